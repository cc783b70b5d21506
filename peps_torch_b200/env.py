"""Environment containers with the reference's layout and API (ctm/generic/env.py:14-233: ENV with clone / detach /
extend / min_chi / get_spectra / get_site_env_t; ctm/one_site_c4v/env_c4v.py:7-164: ENV_C4V), the 'CTMRG' and 'RANDOM'
initialisations (env.py:235-272,367-536; env_c4v.py:166-311) and the corner-spectrum convergence criterion
(env.py:816-875).  Initialisation runs once per CTM run on a handful of D^2-sized tensors; it uses torch ops on the
state's device and is not part of the timed hot path."""
from math import sqrt
import torch
from . import config as cfg


class EnvError(RuntimeError):
    def __init__(self, message="Environment error"):
        super().__init__(message)

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
            # single-layer sites a[s,u,l,d,r]: the environment legs are (ket, bra) pairs; double-layer A[u,l,d,r]: the legs themselves
            numl = 2 if next(iter(state.sites.values())).dim() > 4 else 1
            for coord, site in state.sites.items():
                D = site.shape[-4:]
                self.T[(coord, (0, -1))] = torch.empty((chi, D[0] ** numl, chi), dtype=self.dtype, device=self.device)
                self.T[(coord, (-1, 0))] = torch.empty((chi, chi, D[1] ** numl), dtype=self.dtype, device=self.device)
                self.T[(coord, (0, 1))] = torch.empty((D[2] ** numl, chi, chi), dtype=self.dtype, device=self.device)
                self.T[(coord, (1, 0))] = torch.empty((chi, D[3] ** numl, chi), dtype=self.dtype, device=self.device)
                for vec in [(-1, -1), (-1, 1), (1, -1), (1, 1)]:
                    self.C[(coord, vec)] = torch.empty((chi, chi), dtype=self.dtype, device=self.device)

    def _like(self, chi=None):
        e = ENV(self.chi if chi is None else chi)
        if hasattr(self, 'dtype'):
            e.dtype, e.device = self.dtype, self.device
        return e

    def __str__(self):
        lines = [f"ENV chi={self.chi}"]
        lines += [f"C({k[0]} {k[1]}): {tuple(t.shape)}" for k, t in self.C.items()]
        lines += [f"T({k[0]} {k[1]}): {tuple(t.shape)}" for k, t in self.T.items()]
        return "\n".join(lines)

    def clone(self, ctm_args=None, global_args=None):
        """Copy of the environment (env.py:118-133; keeps gradient tracking)."""
        e = self._like()
        e.C = {k: v.clone() for k, v in self.C.items()}
        e.T = {k: v.clone() for k, v in self.T.items()}
        return e

    def detach(self, ctm_args=None, global_args=None):
        """Detached view of the environment (env.py:135-151)."""
        e = self._like()
        e.C = {k: v.detach() for k, v in self.C.items()}
        e.T = {k: v.detach() for k, v in self.T.items()}
        return e

    def detach_(self):
        for t in list(self.C.values()) + list(self.T.values()):
            t.detach_()

    def min_chi(self):
        """Smallest environment bond dimension over all corners (env.py:157-162)."""
        return min(min(c.shape) for c in self.C.values())

    def extend(self, new_chi, ctm_args=None, global_args=None):
        """New environment of dimension ``new_chi``: the leading min(chi, new_chi) block of every environment leg is
        kept, the rest is zero (env.py:164-202).  Environment legs: both of C; legs (0,2) of T(0,-1) and T(1,0),
        (0,1) of T(-1,0), (1,2) of T(0,1)."""
        e = self._like(new_chi)
        x = min(self.chi, new_chi)
        env_legs = {(0, -1): (0, 2), (1, 0): (0, 2), (-1, 0): (0, 1), (0, 1): (1, 2)}
        for k, c in self.C.items():
            out = c.new_zeros((new_chi, new_chi))
            out[:x, :x] = c[:x, :x].detach()
            e.C[k] = out
        for k, t in self.T.items():
            if k[1] not in env_legs:
                raise Exception(f"Unexpected direction {k[1]}")
            shape = [new_chi if i in env_legs[k[1]] else t.shape[i] for i in range(3)]
            sl = tuple(slice(0, x) if i in env_legs[k[1]] else slice(None) for i in range(3))
            out = t.new_zeros(shape)
            out[sl] = t[sl].detach()
            e.T[k] = out
        return e

    def get_site_env_t(self, coord, state):
        """(C1, C2, C3, C4, T1, T2, T3, T4) of the site at ``coord``: corners clockwise from the upper left, half-row /
        column tensors clockwise from the top (env.py:211-233)."""
        s = state.vertexToSite(coord)
        return tuple(self.C[(s, v)] for v in ((-1, -1), (1, -1), (1, 1), (-1, 1))) \
            + tuple(self.T[(s, v)] for v in ((0, -1), (1, 0), (0, 1), (-1, 0)))

    def get_spectra(self):
        out = {}
        for k, c in self.C.items():
            s = torch.linalg.svdvals(c)
            out[k] = s / s[0]
        return out


def init_random(env, verbosity=0):
    """All C and T uniform in [0,1) (env.py:268-272)."""
    for key, t in env.C.items():
        env.C[key] = torch.rand(t.shape, dtype=t.dtype, device=t.device)
    for key, t in env.T.items():
        env.T[key] = torch.rand(t.shape, dtype=t.dtype, device=t.device)


def init_env(state, env, ctm_args=None):
    """Initialise ``env`` as CTMARGS.ctm_env_init_type says (env.py:235-265): 'CTMRG' (default) or 'RANDOM'; the
    product-state and open-boundary variants 'PROD' (env.py:274-365) and 'CTMRG_OBC' (:538-715) as well."""
    kind = getattr(ctm_args if ctm_args is not None else cfg.ctm_args, 'ctm_env_init_type', 'CTMRG')
    if next(iter(state.sites.values())).dim() == 4 and kind not in ('PROD', 'CTMRG_OBC', 'RANDOM'):
        raise RuntimeError("Incompatible ENV initialization")
    if kind == 'RANDOM':
        return init_random(env)
    if kind == 'PROD':
        return init_prod(state, env)
    if kind == 'CTMRG_OBC':
        return init_from_ipeps_obc(state, env)
    if kind != 'CTMRG':
        raise ValueError("Invalid environment initialization: " + str(kind))
    init_from_ipeps_pbc(state, env)


# open-boundary / product initialisations: (vec, leg kept of a double-layer A[u,l,d,r], single-layer einsum over (A, .), kept legs)
_OBC_C = {(-1, -1): ('ijef->ef', 'mijef,mklab->eafb', (3, 4)), (1, -1): ('iefj->ef', 'miefj,mkabl->eafb', (2, 3)),
          (1, 1): ('efij->ef', 'mefij,mabkl->eafb', (1, 2)), (-1, 1): ('eijf->ef', 'meijf,maklb->eafb', (1, 4))}
_OBC_T = {(0, -1): ('iefg->efg', 'miefg,mkabc->eafbgc', (2, 3, 4), (True, False, True)),
          (-1, 0): ('eifg->efg', 'meifg,makbc->eafbgc', (1, 3, 4), (True, True, False)),
          (0, 1): ('efig->efg', 'mefig,mabkc->eafbgc', (1, 2, 4), (False, True, True)),
          (1, 0): ('efgi->efg', 'mefgi,mabck->eafbgc', (1, 2, 3), (True, False, True))}
# 'PROD': (vec, double-layer einsum, single-layer einsum, where the vector sits, normalised?)  -- the (0,-1) vector is NOT
# divided by its largest magnitude in the reference (env.py:282-292 vs :303, :322, :341); kept as it is
_PROD_T = {(0, -1): ('uldr->d', 'miefg,miebg->fb', (0, slice(None), 0), False),
           (-1, 0): ('uldr->r', 'meifg,meifc->gc', (0, 0, slice(None)), True),
           (0, 1): ('uldr->u', 'mefig,mafig->ea', (slice(None), 0, 0), True),
           (1, 0): ('uldr->l', 'mefgi,mebgi->fb', (0, slice(None), 0), True)}


def init_prod(state, env, verbosity=0):
    """'PROD' (env.py:274-365): C = e_0 e_0^T, T = the on-site double-layer tensor traced over the three legs that do not
    face the site, placed in the (0, 0) slot of the environment legs."""
    chi = env.chi
    any_site = next(iter(state.sites.values()))
    for coord in state.sites.keys():
        for vec in _INIT_C:
            c = torch.zeros(chi, chi, dtype=any_site.dtype, device=any_site.device)
            c[0, 0] = 1.0
            env.C[(coord, vec)] = c
        for vec, (dl, sl, where, normalise) in _PROD_T.items():
            A = state.site((coord[0] + vec[0], coord[1] + vec[1]))
            if A.dim() == 4:
                a = torch.einsum(dl, A).contiguous()
            else:
                a = torch.einsum(sl, A, A.conj()).contiguous()
                a = a.view(a.size(0) ** 2)
            if normalise:
                a = a / a.abs().max()
            shape = [a.size(0) if isinstance(w, slice) else chi for w in where]
            t = torch.zeros(shape, dtype=A.dtype, device=A.device)
            t[where] = a
            env.T[(coord, vec)] = t


def init_from_ipeps_obc(state, env, verbosity=0):
    """'CTMRG_OBC' (env.py:538-715): as 'CTMRG', but the legs that point away from the site are summed over separately in
    the two layers (open boundary) instead of being traced; the reference contracts (A, A) here, not (A, conj A)."""
    chi = env.chi
    for coord in state.sites.keys():
        for table, store in ((_OBC_C, env.C), (_OBC_T, env.T)):
            for vec, spec in table.items():
                dl, sl, legs = spec[0], spec[1], spec[2]
                is_chi = spec[3] if len(spec) > 3 else (True, True)
                A = state.site((coord[0] + vec[0], coord[1] + vec[1]))
                d = A.shape
                if A.dim() == 4:
                    t = torch.einsum(dl, A)
                else:
                    t = torch.einsum(sl, A, A).contiguous().view(*[d[l] ** 2 for l in legs])
                t = t / t.abs().max()
                shape = [chi if f else t.shape[i] for i, f in enumerate(is_chi)]
                out = torch.zeros(shape, dtype=A.dtype, device=A.device)
                sl_ = tuple(slice(0, min(chi, t.shape[i])) if f else slice(None) for i, f in enumerate(is_chi))
                out[sl_] = t[sl_]
                store[(coord, vec)] = out


def init_from_ipeps_pbc(state, env, verbosity=0):
    """'CTMRG' initialisation: partial traces of a (x) a*, /max|.|, zero-padded to chi (env.py:367-536)."""
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


@torch.no_grad()
def ctmrg_conv_specC(state, env, history, p='inf', ctm_args=None):
    """Convergence criterion on the normalised singular spectra of all corners (env.py:816-875): conv_crit = sqrt of the
    largest (p='inf') or of the summed (p='fro' / 2) squared 2-norm difference between the spectra of consecutive
    iterations.  Same history dict ('spec', 'diffs', 'conv_crit') and return value as the reference; the differences are
    reduced on the device and read back with ONE host synchronisation per call (the reference synchronises once per
    corner)."""
    ctm_args = ctm_args if ctm_args is not None else cfg.ctm_args
    if not history:
        history = {'spec': [], 'diffs': [], 'conv_crit': []}
    spec = {k: s.sort(descending=True)[0] for k, s in env.get_spectra().items()}
    conv_crit, diffs = float('inf'), None
    if history['spec']:
        old = history['spec'][-1]
        per_corner = []
        for k, s in spec.items():
            o = old[k]
            m = min(s.shape[0], o.shape[0])
            # spectra of different length: the missing tail counts as zeros
            per_corner.append(((s[:m] - o[:m]) ** 2).sum() + (s[m:] ** 2).sum() + (o[m:] ** 2).sum())
        diffs = torch.stack(per_corner).tolist()                 # the one synchronisation
        if p in ('fro', 2):
            conv_crit = sqrt(sum(diffs))
        elif p in (float('inf'), 'inf'):
            conv_crit = sqrt(max(diffs))
    history['spec'].append(spec)
    history['diffs'].append(diffs)
    history['conv_crit'].append(conv_crit)
    done = (len(history['diffs']) > 1 and conv_crit < ctm_args.ctm_conv_tol) or len(history['diffs']) >= ctm_args.ctm_max_iter
    return bool(done), history


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

    def _like(self, chi=None):
        e = ENV_C4V(self.chi if chi is None else chi, bond_dim=self.bond_dim)
        if hasattr(self, 'dtype'):
            e.dtype, e.device = self.dtype, self.device
        return e

    def clone(self, ctm_args=None, global_args=None):
        """env_c4v.py:92-108"""
        e = self._like()
        e.C[e.keyC], e.T[e.keyT] = self.get_C().clone(), self.get_T().clone()
        return e

    def detach(self, ctm_args=None, global_args=None):
        """env_c4v.py:110-127"""
        e = self._like()
        e.C[e.keyC], e.T[e.keyT] = self.get_C().detach(), self.get_T().detach()
        return e

    def detach_(self):
        self.get_C().detach_()
        self.get_T().detach_()

    def extend(self, new_chi, ctm_args=None, global_args=None):
        """Zero-padded (or cut) copy with environment dimension ``new_chi`` (env_c4v.py:133-153)."""
        e = self._like(new_chi)
        x = min(self.chi, new_chi)
        C_, T = self.get_C(), self.get_T()
        nC, nT = C_.new_zeros((new_chi, new_chi)), T.new_zeros((new_chi, new_chi, T.shape[2]))
        nC[:x, :x] = C_[:x, :x]
        nT[:x, :x, :] = T[:x, :x, :]
        e.C[e.keyC], e.T[e.keyT] = nC, nT
        return e


def compute_multiplets(env, eps_multiplet_gap=1.0e-10):
    """Sizes of the groups of (numerically) degenerate |eigenvalues| of the C4v corner, largest first
    (env_c4v.py:401-418): a group ends where the gap to the next magnitude exceeds ``eps_multiplet_gap``."""
    D = torch.linalg.eigvalsh(env.C[env.keyC]).abs().sort(descending=True)[0]
    D = torch.cat([D, D.new_zeros(1)])
    gaps = (D[:-1] - D[1:] > eps_multiplet_gap).tolist()
    sizes, run = [], 0
    for big in gaps:
        run += 1
        if big:
            sizes.append(run)
            run = 0
    return sizes


def init_random_c4v(env, verbosity=0):
    """Hermitian random C, random T (env_c4v.py:249-255)."""
    C_ = torch.rand(env.get_C().shape, dtype=env.get_C().dtype, device=env.get_C().device)
    env.C[env.keyC] = 0.5 * (C_ + C_.conj().t())
    env.T[env.keyT] = torch.rand(env.get_T().shape, dtype=env.get_T().dtype, device=env.get_T().device)


def init_env_c4v(state, env, C_and_T=None, ctm_args=None):
    """init_env of env_c4v.py:166-213: custom (C, T), 'RANDOM', or 'CTMRG' (default; :257-311: C = diag(eig of the
    D^2 x D^2 corner), T rotated into its eigenbasis)."""
    if C_and_T:
        assert len(C_and_T) == 2 and all(isinstance(t, torch.Tensor) for t in C_and_T), "Invalid C and T. Expects tuple (C, T)."
        C0, T0 = C_and_T
        x = C0.shape[0]
        nC = env.get_C().new_zeros((env.chi, env.chi)) if env.keyC in env.C else C0.new_zeros((env.chi, env.chi))
        nT = C0.new_zeros((env.chi, env.chi, T0.shape[2]))
        nC[:x, :x] = C0
        nT[:x, :x, :] = T0
        env.C[env.keyC], env.T[env.keyT] = nC, nT
        return
    kind = getattr(ctm_args if ctm_args is not None else cfg.ctm_args, 'ctm_env_init_type', 'CTMRG')
    a = next(iter(state.sites.values()))
    if a.dim() == 4 and kind in ('CTMRG', 'CTMRG_OBC'):
        raise RuntimeError("Incompatible ENV_C4V initialization")
    if kind == 'RANDOM':
        return init_random_c4v(env)
    if kind in ('PROD', 'CTMRG_OBC', 'CTMRG_OBC_SL'):
        raise NotImplementedError(f"ctm_env_init_type='{kind}' is not built in peps_torch_b200; use the reference's init_env")
    if kind != 'CTMRG':
        raise ValueError("Invalid environment initialization: " + str(kind))
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


@torch.no_grad()
def ctmrg_conv_rdm2x1(state, env, history, ctm_args=None, min_history=1, lag=0):
    """Convergence criterion of the C4v scripts (examples/j1j2/ctmrg_j1j2_c4v.py:101-129; optim_j1j2_c4v.py:72-86 with
    min_history=0): 2-norm distance between the nearest-neighbour density matrices (rdm2x1_sl, on libctmb) of consecutive
    iterations; history = {'log': [dist...], 'rdm': last rdm}; converged when dist < ctm_conv_tol, stops at ctm_max_iter.

    lag=0 reproduces the reference decision for decision, with ONE host synchronisation per call (the .item() of the
    reference).  lag=1 never stalls the stream: the distance is reduced on the device, copied to pinned memory asynchronously
    and the decision of THIS call is taken on the distance of the PREVIOUS one, which has long arrived while the next move
    was being enqueued; the run then makes one move more than the reference."""
    from .ctm.one_site_c4v.rdm_c4v import rdm2x1_sl
    ctm_args = ctm_args if ctm_args is not None else cfg.ctm_args
    if not history:
        history = {'log': [], 'pending': []}
    rdm = rdm2x1_sl(state, env)
    n = len(history['log']) + len(history.get('pending', []))
    dist = float('inf')
    if lag == 0:
        if n > min_history:
            dist = torch.dist(rdm, history['rdm'], p=2).item()           # the one synchronisation
        history['log'].append(dist)
    else:
        pend = history.setdefault('pending', [])
        if n > min_history:
            host = torch.empty((), dtype=torch.float64, pin_memory=rdm.is_cuda)
            host.copy_(torch.dist(rdm, history['rdm'], p=2).to(torch.float64), non_blocking=True)
            ev = torch.cuda.Event() if rdm.is_cuda else None
            if ev is not None:
                ev.record()
            pend.append((host, ev))
        else:
            pend.append((None, None))
        while len(pend) > lag:                                               # distances old enough to have arrived
            host, ev = pend.pop(0)
            if ev is not None:
                ev.synchronize()                                             # completed long ago: returns at once
            history['log'].append(float(host) if host is not None else float('inf'))
        dist = history['log'][-1] if history['log'] else float('inf')
    history['rdm'] = rdm
    converged = dist < ctm_args.ctm_conv_tol
    if converged or n + 1 >= ctm_args.ctm_max_iter:
        return bool(converged), history
    return False, history
