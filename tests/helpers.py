"""Shared helpers of the test-suite: golden fixtures, containers, gauge-invariant comparisons."""
import os
import json
from collections import OrderedDict
import numpy as np
import torch
import ctm_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
C_KEYS = [(-1, -1), (1, -1), (1, 1), (-1, 1)]
T_KEYS = [(0, -1), (-1, 0), (0, 1), (1, 0)]


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + '.npz'))
    meta = json.loads(str(z['meta'])) if 'meta' in z.files else {}
    return z, meta


def golden_sites(z):
    sites = OrderedDict()
    for k in z.files:
        if k.startswith('site_'):
            sites[(int(k[5]), int(k[6]))] = torch.from_numpy(z[k])
    # np.savez keeps insertion order, which is the reference's dict order
    return sites


def golden_env(z, prefix):
    C, T = {}, {}
    for k in z.files:
        if k.startswith(prefix + 'C_') or k.startswith(prefix + 'T_'):
            body = k[len(prefix) + 2:]
            c, vx, vy = body.split('_')
            key = ((int(c[0]), int(c[1])), (int(vx), int(vy)))
            (C if k[len(prefix)] == 'C' else T)[key] = torch.from_numpy(z[k])
    return C, T


def v2s_for(sites):
    n = len(sites)
    if n == 4:
        return orc.v2s_4site, 2, 2
    if n == 2:
        return orc.v2s_2site, 2, 1
    return orc.v2s_1site, 1, 1


class State:
    """Minimal stand-in with the attributes the move reads from the reference's IPEPS
    (ipeps/ipeps.py:89-250): sites, vertexToSite, lX, lY, site()."""

    def __init__(self, sites, vertexToSite, lX, lY):
        self.sites, self.vertexToSite, self.lX, self.lY = sites, vertexToSite, lX, lY
        t = next(iter(sites.values()))
        self.dtype, self.device = t.dtype, t.device

    def site(self, coord):
        return self.sites[self.vertexToSite(coord)]


class Env:
    """Stand-in for ctm.generic.env.ENV (env.py:14-109): chi, C, T dicts."""

    def __init__(self, chi, C, T):
        self.chi, self.C, self.T = chi, C, T


def to_dev(d, device):
    return type(d)((k, v.to(device)) for k, v in d.items())


def maxrel(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-300))


def env_abs_diff(C1, T1, C2, T2):
    w = 0.
    for k in C2:
        w = max(w, maxrel(C1[k].abs().cpu(), C2[k].abs().cpu()))
    for k in T2:
        w = max(w, maxrel(T1[k].abs().cpu(), T2[k].abs().cpu()))
    return w


def spectra_diff(C1, C2):
    s1 = orc.corner_spectra({k: v.cpu() for k, v in C1.items()})
    s2 = orc.corner_spectra({k: v.cpu() for k, v in C2.items()})
    return max(float((s1[k] - s2[k]).abs().max()) for k in s2)


class OracleEngine:
    """HOST-LOGIC TESTS ONLY: an object with the three CtmEngine methods the drop-in modules call, computing with the
    oracle on CPU tensors, so that the Python control flow of run / run_overlap / run_dl (double-layer construction,
    warm-up, option mapping, conv_check protocol) is exercised without a GPU.  Never importable from the package."""
    device = 'cpu'

    def __init__(self):
        self.calls = []

    @staticmethod
    def _args(opt):
        return orc.OracleArgs(projector_method={0: '4X4', 1: '4X2', None: '4X4'}[opt.get('projector_method')],
                              ctm_absorb_normalization='inf' if not opt.get('norm_type') else 'fro',
                              projector_svd_reltol=opt.get('svd_reltol') or 1e-8,
                              projector_eps_multiplet=opt.get('eps_multiplet') or 1e-8,
                              projector_multiplet_abstol=opt.get('multiplet_abstol') or 1e-14)

    def einsum2(self, spec, A, B, conjA=False, conjB=False):
        return torch.einsum(spec, A.conj() if conjA else A, B.conj() if conjB else B).contiguous()

    def move_generic(self, direction, state, env, **opt):
        self.calls.append(('generic', direction, next(iter(state.sites.values())).dim()))
        orc.ctm_move(direction, state.sites, state.vertexToSite, env.C, env.T, env.chi, self._args(opt))

    def move_c4v(self, a, C_, T, chi, **opt):
        self.calls.append(('c4v', a.dim()))
        nC, nT = orc.ctm_move_c4v(a, C_, T, chi, self._args(opt))
        return nC, nT, None

    def rdm2x2(self, coord, state, env, open_sites=(0, 1, 2, 3), sym_pos_def=False, raw=False):
        return orc.rdm2x2(coord, state.sites, state.vertexToSite, env.C, env.T, raw=raw, open_sites=tuple(open_sites),
                          sym_pos_def=sym_pos_def)

    def rdm2x2_sites(self, tensors4, chi, open_sites=(0, 1, 2, 3), sym_pos_def=False, raw=False):
        from peps_torch_b200.engine import C_KEYS, T_KEYS
        coords = [(0, 0), (1, 0), (0, 1), (1, 1)]
        sites = OrderedDict((c, t[0]) for c, t in zip(coords, tensors4))
        C = {(c, k): t[1][i] for c, t in zip(coords, tensors4) for i, k in enumerate(C_KEYS)}
        T = {(c, k): t[2][i] for c, t in zip(coords, tensors4) for i, k in enumerate(T_KEYS)}
        return orc.rdm2x2((0, 0), sites, orc.v2s_4site, C, T, raw=raw, open_sites=tuple(open_sites), sym_pos_def=sym_pos_def)

    def rdm_small(self, kind, coord, state, env, sym_pos_def=False, raw=False):
        f = {'1x1': orc.rdm1x1, '2x1': orc.rdm2x1, '1x2': orc.rdm1x2}[kind]
        return f(coord, state.sites, state.vertexToSite, env.C, env.T, raw=raw, sym_pos_def=sym_pos_def)

    def rdm_small_sites(self, kind, tensors2, chi, sym_pos_def=False, raw=False):
        from peps_torch_b200.engine import C_KEYS, T_KEYS
        coords = [(0, 0), (1, 0)] if kind != '1x2' else [(0, 0), (0, 1)]
        t2 = list(tensors2) if len(tensors2) > 1 else [tensors2[0], tensors2[0]]
        sites = OrderedDict((c, t[0]) for c, t in zip(coords, t2))
        C = {(c, k): t[1][i] for c, t in zip(coords, t2) for i, k in enumerate(C_KEYS)}
        T = {(c, k): t[2][i] for c, t in zip(coords, t2) for i, k in enumerate(T_KEYS)}
        v2s = (lambda c: (c[0] % 2, 0)) if kind != '1x2' else (lambda c: (0, c[1] % 2))
        f = {'1x1': orc.rdm1x1, '2x1': orc.rdm2x1, '1x2': orc.rdm1x2}[kind]
        return f((0, 0), sites, v2s, C, T, raw=raw, sym_pos_def=sym_pos_def)
