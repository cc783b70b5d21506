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
        if not list(open_sites):
            raise ValueError("open_sites must be a non-empty subset of [0,1,2,3]")      # as CtmEngine.rdm2x2_sites
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

    # ---- the piecewise entry points of CtmEngine, for the CPU dry run of the GPU suite (tests/test_gpu_dryrun_cpu.py) ----
    def debug_set_matrix_free(self, mode):
        pass

    def debug_set_m_noise(self, amp):
        pass

    def halves(self, direction, coord, state, env):
        return orc.halves(direction, coord, state.sites, state.vertexToSite, env.C, env.T)

    def move_generic_projectors(self, direction, state, env, jobs, **opt):
        coords = list(state.sites.keys())
        out = [orc.projectors_from_matrices(*orc.halves(direction, coords[j], state.sites, state.vertexToSite, env.C, env.T),
                                            env.chi, self._args(opt)) for j in jobs]
        return [p.contiguous() for p, _ in out], [pt.contiguous() for _, pt in out]

    def move_generic_absorb(self, direction, state, env, jobs, P_all, Pt_all, **opt):
        coords = list(state.sites.keys())
        P = {coords[i]: p for i, p in enumerate(P_all)}
        Pt = {coords[i]: p for i, p in enumerate(Pt_all)}
        out = []
        for j in jobs:
            c = coords[j]
            nC1, nC2, nT = orc.absorb(direction, c, state.sites, state.vertexToSite, env.C, env.T, P, Pt, self._args(opt))
            out.append((state.vertexToSite((c[0] - direction[0], c[1] - direction[1])), nC1, nC2, nT))
        return out

    def c2x2(self, kind, C_, T1, T2, a, chi):
        if a.dim() == 4:
            t = orc.sl_einsum(orc.CORNERS[kind][3], (C_, T1, T2), a)
            return t.reshape(t.shape[0] * t.shape[1], t.shape[2] * t.shape[3])
        return orc.c2x2(kind, C_, T1, T2, a)

    def projectors(self, R, Rt, chi, **opt):
        P, Pt, (M, U, S, V) = orc.projectors_from_matrices(R, Rt, chi, self._args(opt), return_svd=True)
        return P, Pt, S

    def truncated_svd(self, M, chi, **opt):
        return orc.truncated_svd(M, chi, opt.get('eps_multiplet') or 1e-8, opt.get('multiplet_abstol') or 1e-14)

    def truncated_eig_sym(self, M, chi, **opt):
        return orc.truncated_eig_sym(M, chi)

    def sym_pos_def(self, rdm, sym_pos_def=False):
        return orc._sym_pos_def(rdm, sym_pos_def)

    def qr(self, M):
        return torch.linalg.qr(M)


# ----------------------------------------------------------------------------------------------
# tests/golden/variants_*.npz (oracle/gen_golden_variants.py): outputs of the UNMODIFIED reference for the variants of
# the move and for the density matrices.  The same checks run on CPU with OracleEngine patched into the drop-in modules
# (pins the oracle and the host logic without the reference being present) and on the GPU box with libctmb.
# ----------------------------------------------------------------------------------------------
def check_generic_variants(name, dev, tol_move=1e-8, tol_rdm=1e-12, tol_rdm_spd=1e-10):
    from peps_torch_b200.ctm.generic import ctmrg, rdm
    from peps_torch_b200.config import CTMARGS
    z, meta = load_golden(name)
    v, vmeta = load_golden('variants_' + name)
    assert vmeta['source'] == name
    chi = meta['chi']
    sites = golden_sites(z)
    v2s, lX, lY = v2s_for(sites)
    C0, T0 = golden_env(z, 'mid_')
    sites_dev = to_dev(sites, dev)
    dl_dev = type(sites)((c, ctmrg.double_layer(ctmrg._engine(), a)) for c, a in sites_dev.items())
    n_checked = 0
    for tag, kw, ss in (('4x2', dict(projector_method='4X2'), sites_dev), ('dl', dict(ctm_force_dl=True), dl_dev),
                        ('fro', dict(ctm_absorb_normalization='fro'), sites_dev)):
        args = CTMARGS()
        for k, val in kw.items():
            setattr(args, k, val)
        st = State(ss, v2s, lX, lY)
        for d in orc.DIRECTIONS:
            env = Env(chi, to_dev(C0, dev), to_dev(T0, dev))
            ctmrg.ctm_MOVE(d, st, env, ctm_args=args)
            prefix = f'move_{tag}_{d[0]}_{d[1]}_'
            keys = [k for k in v.files if k.startswith(prefix)]
            assert len(keys) == 3 * len(sites), (prefix, len(keys))
            for k in keys:
                kind, body = k[len(prefix)], k[len(prefix) + 2:]
                c, vx, vy = body.split('_')
                key = ((int(c[0]), int(c[1])), (int(vx), int(vy)))
                got = (env.C if kind == 'C' else env.T)[key].cpu()
                want = torch.from_numpy(v[k])
                assert got.shape == want.shape, k
                assert maxrel(got.abs(), want.abs()) < tol_move, (k, maxrel(got.abs(), want.abs()))
                n_checked += 1
    st = State(sites_dev, v2s, lX, lY)
    env = Env(chi, to_dev(C0, dev), to_dev(T0, dev))
    for coord in sites:
        for spd in (False, True):
            for fname, f in (('rdm2x2', rdm.rdm2x2), ('rdm1x1', rdm.rdm1x1), ('rdm2x1', rdm.rdm2x1), ('rdm1x2', rdm.rdm1x2)):
                want = torch.from_numpy(v[f'{fname}_{coord[0]}{coord[1]}_{int(spd)}'])
                got = f(coord, st, env, sym_pos_def=spd).cpu()
                assert got.shape == want.shape, fname
                err = float((got - want).abs().max())
                assert err < (tol_rdm_spd if spd else tol_rdm), (fname, coord, spd, err)
                n_checked += 1
    return n_checked


def check_c4v_variants(name, dev, tol_C=1e-10, tol_T=1e-8, tol_rdm=1e-12, tol_rdm_spd=1e-10):
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v, rdm_c4v
    from peps_torch_b200.ipeps import IPEPS_C4V
    from peps_torch_b200.env import ENV_C4V
    z, meta = load_golden(name)
    v, vmeta = load_golden('variants_' + name)
    assert vmeta['source'] == name
    chi = meta['chi']
    a = torch.from_numpy(z['site']).to(dev)
    stc = IPEPS_C4V(a)
    env = ENV_C4V(chi, stc)
    env.C[env.keyC], env.T[env.keyT] = torch.from_numpy(z['init_C']).to(dev), torch.from_numpy(z['init_T']).to(dev)
    for _ in range(vmeta['n_moves_dl']):
        ctmrg_c4v.ctm_MOVE_dl(a, env, None)
    assert maxrel(env.get_C().cpu(), torch.from_numpy(v['dl3_C'])) < tol_C
    assert maxrel(env.get_T().abs().cpu(), torch.from_numpy(v['dl3_T']).abs()) < tol_T
    env.C[env.keyC], env.T[env.keyT] = torch.from_numpy(z['final_C']).to(dev), torch.from_numpy(z['final_T']).to(dev)
    n_checked = 2
    for spd in (False, True):
        for fname, f in (('rdm2x2_NN', rdm_c4v.rdm2x2_NN_lowmem_sl), ('rdm2x2_NNN', rdm_c4v.rdm2x2_NNN_lowmem_sl),
                         ('rdm2x2', rdm_c4v.rdm2x2), ('rdm1x1', rdm_c4v.rdm1x1_sl), ('rdm2x1', rdm_c4v.rdm2x1_sl)):
            want = torch.from_numpy(v[f'{fname}_{int(spd)}'])
            got = f(stc, env, sym_pos_def=spd).cpu()
            assert got.shape == want.shape, fname
            err = float((got - want).abs().max())
            assert err < (tol_rdm_spd if spd else tol_rdm), (fname, spd, err)
            n_checked += 1
    return n_checked


# ----------------------------------------------------------------------------------------------
# tests/golden/grad_*.npz (oracle/gen_golden_grad.py): gradients of the J1-J2 energy written by the UNMODIFIED reference
# (loss.backward() through its CTM moves).  The same checker runs on CPU with OracleEngine standing in for libctmb (this
# validates peps_torch_b200/ad.py's adjoints and chains) and on the GPU box with libctmb.
# ----------------------------------------------------------------------------------------------
def c4v_symm(A):
    if A.is_complex():
        return orc.make_c4v_symm_A1(A.real) + 1j * orc.make_c4v_symm_A2(A.imag)
    return orc.make_c4v_symm_A1(A)


def check_grad_fixture(name, eng, dev, through_api=True):
    """Returns (|energy - reference|, max |grad - reference grad|, max |reference grad|)."""
    import json
    from peps_torch_b200 import ad
    from peps_torch_b200.config import CTMARGS
    z = np.load(os.path.join(GOLD, name + '.npz'))
    meta = json.loads(str(z['meta']))
    args = CTMARGS()
    args.ad_decomp_reg = meta['ad_decomp_reg']
    args.fwd_checkpoint_move = name.endswith('_ckpt')
    if meta['kind'] == 'c4v':
        A = torch.from_numpy(z['site']).to(dev).requires_grad_(True)
        a = c4v_symm(A)
        C, T = torch.from_numpy(z['C0']).to(dev), torch.from_numpy(z['T0']).to(dev)
        if through_api:      # the drop-in entry point: ctm_MOVE_sl dispatches to the AD path because `a` requires grad
            from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
            from peps_torch_b200.env import ENV_C4V
            from peps_torch_b200.ipeps import IPEPS_C4V
            env = ENV_C4V(meta['chi'], IPEPS_C4V(a.detach()))
            env.C[env.keyC], env.T[env.keyT] = C, T
            for _ in range(meta['moves']):
                ctmrg_c4v.ctm_MOVE_sl(a, env, None, ctm_args=args)
            C, T = env.C[env.keyC], env.T[env.keyT]
        else:
            for _ in range(meta['moves']):
                C, T = ad.ctm_move_c4v(eng, a, C, T, meta['chi'], args)
        loss = orc.energy_j1j2_c4v(a.cpu(), C.cpu(), T.cpu(), 1.0, meta['j2'], as_tensor=True)
        loss.backward()
        g, g_ref = A.grad.cpu(), torch.from_numpy(z['grad'])
        return abs(float(loss.detach()) - float(z['energy'][0])), float((g - g_ref).abs().max()), float(g_ref.abs().max())
    coords = [(0, 0), (1, 0), (0, 1), (1, 1)]
    sites = OrderedDict((c, torch.from_numpy(z[f'site_{c[0]}{c[1]}']).to(dev).requires_grad_(True)) for c in coords)
    C, T = {}, {}
    for key in z.files:
        if key[:3] in ('C0_', 'T0_'):
            c, vx, vy = key[3:].split('_', 1)[0], *key[3:].split('_')[1:]
            k = ((int(c[0]), int(c[1])), (int(vx), int(vy)))
            (C if key[0] == 'C' else T)[k] = torch.from_numpy(z[key]).to(dev)
    st = State(sites, orc.v2s_4site, 2, 2)
    env = Env(meta['chi'], C, T)
    from peps_torch_b200.ctm.generic import ctmrg
    for _ in range(meta['iters']):
        for d in orc.DIRECTIONS:
            for _r in range(2):
                if through_api:
                    ctmrg.ctm_MOVE(d, st, env, ctm_args=args)
                else:
                    ad.ctm_move_generic(eng, d, st, env, args)
    loss = orc.energy_j1j2(OrderedDict((c, t.cpu()) for c, t in sites.items()), orc.v2s_4site,
                           {k: v.cpu() for k, v in env.C.items()}, {k: v.cpu() for k, v in env.T.items()}, 1.0, meta['j2'], as_tensor=True)
    loss.backward()
    d_ = max(float((sites[c].grad.cpu() - torch.from_numpy(z[f'grad_{c[0]}{c[1]}'])).abs().max()) for c in coords)
    s_ = max(float(np.abs(z[f'grad_{c[0]}{c[1]}']).max()) for c in coords)
    return abs(float(loss.detach()) - float(z['energy'][0])), d_, s_
