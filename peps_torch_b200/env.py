"""Environment containers and 'CTMRG' initialisation with the reference's layout
(ctm/generic/env.py:14-109,367-536; ctm/one_site_c4v/env_c4v.py:7-76,257-311).
Initialisation runs once per CTM run on a handful of D^2-sized tensors; it uses torch ops
on the state's device and is not part of the timed hot path."""
import torch

_INIT_C = {(-1, -1): ('mijef,mijab->eafb', (3, 4)), (1, -1): ('miefj,miabj->eafb', (2, 3)),
           (1, 1): ('mefij,mabij->eafb', (1, 2)), (-1, 1): ('meijf,maijb->eafb', (1, 4))}
_INIT_T = {(0, -1): ('miefg,miabc->eafbgc', (2, 3, 4), (True, False, True)),
           (-1, 0): ('meifg,maibc->eafbgc', (1, 3, 4), (True, True, False)),
           (0, 1): ('mefig,mabic->eafbgc', (1, 2, 4), (False, True, True)),
           (1, 0): ('mefgi,mabci->eafbgc', (1, 2, 3), (True, False, True))}


class ENV:
    def __init__(self, chi, state=None):
        self.chi = chi
        self.C, self.T = dict(), dict()
        if state is not None:
            self.dtype, self.device = state.dtype, state.device
            for coord, site in state.sites.items():
                D = site.shape[1:]
                self.T[(coord, (0, -1))] = torch.empty((chi, D[0] ** 2, chi), dtype=self.dtype, device=self.device)
                self.T[(coord, (-1, 0))] = torch.empty((chi, chi, D[1] ** 2), dtype=self.dtype, device=self.device)
                self.T[(coord, (0, 1))] = torch.empty((D[2] ** 2, chi, chi), dtype=self.dtype, device=self.device)
                self.T[(coord, (1, 0))] = torch.empty((chi, D[3] ** 2, chi), dtype=self.dtype, device=self.device)
                for vec in [(-1, -1), (-1, 1), (1, -1), (1, 1)]:
                    self.C[(coord, vec)] = torch.empty((chi, chi), dtype=self.dtype, device=self.device)

    def clone(self):
        e = ENV(self.chi)
        e.C = {k: v.clone() for k, v in self.C.items()}
        e.T = {k: v.clone() for k, v in self.T.items()}
        return e

    def get_spectra(self):
        out = {}
        for k, c in self.C.items():
            s = torch.linalg.svdvals(c)
            out[k] = s / s[0]
        return out


def init_env(state, env):
    """'CTMRG' initialisation: partial traces of a (x) a*, /max|.|, zero-padded to chi."""
    chi = env.chi
    for coord in state.sites.keys():
        for vec, (ein, legs) in _INIT_C.items():
            A = state.site((coord[0] + vec[0], coord[1] + vec[1]))
            d = A.shape
            c = torch.einsum(ein, A, A.conj()).contiguous().view(d[legs[0]] ** 2, d[legs[1]] ** 2)
            c = c / c.abs().max()
            out = torch.zeros(chi, chi, dtype=A.dtype, device=A.device)
            r, q = min(chi, c.shape[0]), min(chi, c.shape[1])
            out[:r, :q] = c[:r, :q]
            env.C[(coord, vec)] = out
        for vec, (ein, legs, is_chi) in _INIT_T.items():
            A = state.site((coord[0] + vec[0], coord[1] + vec[1]))
            d = A.shape
            t = torch.einsum(ein, A, A.conj()).contiguous().view(*[d[l] ** 2 for l in legs])
            t = t / t.abs().max()
            shape = [chi if f else t.shape[i] for i, f in enumerate(is_chi)]
            out = torch.zeros(shape, dtype=A.dtype, device=A.device)
            sl = tuple(slice(0, min(chi, t.shape[i])) if f else slice(None) for i, f in enumerate(is_chi))
            out[sl] = t[sl]
            env.T[(coord, vec)] = out


class ENV_C4V:
    def __init__(self, chi, state=None, bond_dim=None):
        assert state is not None or bond_dim, "either state or bond_dim must be supplied"
        self.chi = chi
        self.keyC = ((0, 0), (-1, -1))
        self.keyT = ((0, 0), (-1, 0))
        self.C, self.T = dict(), dict()
        if state is not None:
            site = next(iter(state.sites.values()))
            bond_dim = site.shape[-1]
            self.dtype, self.device = site.dtype, site.device
            self.C[self.keyC] = torch.zeros((chi, chi), dtype=self.dtype, device=self.device)
            self.T[self.keyT] = torch.zeros((chi, chi, bond_dim ** 2), dtype=self.dtype, device=self.device)
        self.bond_dim = bond_dim

    def get_C(self):
        return self.C[self.keyC]

    def get_T(self):
        return self.T[self.keyT]


def init_env_c4v(state, env):
    """env_c4v.py:257-311: C = diag(eig of the D^2 x D^2 corner), T rotated into its eigenbasis."""
    a = next(iter(state.sites.values()))
    chi = env.chi
    d = a.shape
    dk = [d[i + 1] ** 2 for i in range(4)]
    c = torch.einsum('mijef,mijab->eafb', a, a.conj()).contiguous().view(dk[2], dk[3])
    c = c / c.abs().max()
    Dv, U = torch.linalg.eigh(c)
    p = torch.sort(torch.abs(Dv), descending=True)[1]
    Dv, U = Dv[p], U[:, p]
    C = torch.zeros(chi, chi, dtype=a.dtype, device=a.device)
    r = min(chi, dk[2])
    C[:r, :r] = torch.diag(Dv).to(a.dtype)[:r, :r]
    t = torch.einsum('meifg,maibc->eafbgc', a, a.conj()).contiguous().view(dk[0], dk[2], dk[3])
    t = t / t.abs().max()
    t = torch.einsum('ai,abs,bj->ijs', U, t, U.conj())
    T = torch.zeros(chi, chi, dk[3], dtype=a.dtype, device=a.device)
    T[:r, :r, :] = t[:r, :r, :]
    env.C[env.keyC], env.T[env.keyT] = C, T
