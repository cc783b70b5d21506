"""Host-side driver of libctmb: turns (state, env) objects into pointer tables and calls the
C ABI.  torch is used for device memory and streams only."""
import ctypes as C
import torch
from . import _lib
from ._lib import lib, check

DIRECTIONS = {(0, -1): _lib.UP, (-1, 0): _lib.LEFT, (0, 1): _lib.DOWN, (1, 0): _lib.RIGHT}
C_KEYS = [(-1, -1), (1, -1), (1, 1), (-1, 1)]          # order of ctmb_site.C
T_KEYS = [(0, -1), (-1, 0), (0, 1), (1, 0)]            # order of ctmb_site.T
# coordinates of the four enlarged corners of the 2x2 patch of a projector job, in the order
# the reference lists them (ctm/generic/ctm_components.py:37-38,105-106,168-169,231-232)
PATCH = {
    (0, -1): [(0, 0), (0, 1), (-1, 0), (-1, 1)],
    (-1, 0): [(0, 0), (1, 0), (0, 1), (1, 1)],
    (0, 1): [(0, 0), (0, -1), (1, 0), (1, -1)],
    (1, 0): [(0, 0), (-1, 0), (0, -1), (-1, -1)],
}
# neighbour whose projectors P1,Pt1 the absorption uses (ctm/generic/ctmrg.py:326-334,442-450,568-576,684-692)
SHIFT = {(0, -1): (1, 0), (-1, 0): (0, -1), (0, 1): (-1, 0), (1, 0): (0, 1)}
# env tensors written by a move: (nC1 key, nC2 key, nT key)  (ctmrg.py:296-307)
OUT_KEYS = {
    (0, -1): ((1, -1), (-1, -1), (0, -1)),
    (-1, 0): ((-1, -1), (-1, 1), (-1, 0)),
    (0, 1): ((-1, 1), (1, 1), (0, 1)),
    (1, 0): ((1, 1), (1, -1), (1, 0)),
}


def _dt(t):
    if t.dtype == torch.float64:
        return _lib.F64
    if t.dtype == torch.complex128:
        return _lib.C128
    raise TypeError(f"libctmb computes in float64/complex128 (as the reference CLI, config.py:113-118); got {t.dtype}")


# float32 / complex64 (reachable in the reference through the library API only: the generic ENV takes its dtype from the
# state, ctm/generic/env.py:78-83; the CLI accepts float64 / complex128, config.py:113-118).  The moves and the density
# matrices accept them: operands are widened on entry, the kernels run in float64 / complex128 (DMMA), results are rounded
# to the input precision on exit -- inside the 1e-4 gate of the north star by construction.  There is no native
# single-precision kernel (tcgen05.mma kind::tf32 + 3xTF32 compensation would be the Blackwell route).
_WIDE = {torch.float32: torch.float64, torch.complex64: torch.complex128}


def _widen(t):
    return t.to(_WIDE[t.dtype]) if (t is not None and t.dtype in _WIDE) else t


class _Shadow:
    """state / environment stand-in holding widened copies (sites, vertexToSite, lX, lY / chi, C, T)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def refuse_autograd(what, tensors):
    """libctmb's kernels are invisible to autograd: a result computed from tensors that require grad would come back
    detached and an optimiser would silently lose the CTM contribution to the gradient.  Fail loudly instead (reverse-mode
    AD through the move is SURVEY 8f row 2; use torch.no_grad() / detached tensors for forward-only evaluation)."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError(f"{what}: an input requires grad, but the libctmb move is forward-only (no autograd graph is "
                           "recorded through it); call it under torch.no_grad() or detach the state and the environment")


def aux2(a):
    """Extents of the four environment legs facing the on-site tensor: D^2 per leg for a single-layer
    a[s,u,l,d,r], the leg itself for a double-layer A[u,l,d,r] (ctm_force_dl / run_overlap / ctm_MOVE_dl)."""
    if a.dim() == 5:
        return [d * d for d in a.shape[1:]]
    if a.dim() == 4:
        return list(a.shape)
    raise ValueError(f"on-site tensor must have rank 5 (single-layer) or 4 (double-layer), got rank {a.dim()}")


class CtmEngine:
    """One engine (= one libctmb handle, one workspace) per process / GPU."""

    def __init__(self, device=None):
        self._h = C.c_void_p()
        if device == 'plan':
            # planning-only handle (include/ctmb.h: device -1): the workspace queries validate and size every call on a
            # machine without a GPU; every compute call raises CtmbError.  Used by the CPU test-suite.
            self.device = torch.device('cpu')
            check(lib.ctmb_create(C.byref(self._h), -1))
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("peps_torch_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
            check(lib.ctmb_create(C.byref(self._h), self.device.index))
        self._ws = None
        self._tables = {}
        self._rsvd_missed_seen = 0
        self.options = _lib.default_options()

    def close(self):
        if self._h:
            lib.ctmb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ----------------------------------------------------------------------------------
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def _stream(self):
        if self.device.type != 'cuda':          # planning-only engine: the C call refuses to compute
            return C.c_void_p(0)
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ----------------------------------------------------------------------------------
    # intra-site split over a group of GPUs (include/ctmb.h: ctmb_set_group)
    # ----------------------------------------------------------------------------------
    def set_group(self, group=None, rank=0, nranks=1):
        """Make this engine one member of a group that shares every projector job: libctmb splits the operator
        applications of the range finder by sketch columns and calls back for an in-place all-gather over `group`
        (a torch.distributed process group on NCCL).  nranks=1 switches it off."""
        import torch.distributed as dist

        class _Raw:                       # zero-copy view of device memory owned by libctmb's caller
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}

        def _allgather(ctx, buf, bytes_per_rank, stream):
            try:
                whole = torch.as_tensor(_Raw(buf, bytes_per_rank * nranks), device=self.device)
                mine = whole[rank * bytes_per_rank:(rank + 1) * bytes_per_rank]
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                # libctmb launches on torch's current stream (self._stream()); c10d orders the NCCL kernel after the work
                # queued there and makes the stream wait for it on return (async_op=False)
                dist.all_gather_into_tensor(whole, mine, group=group)
                b.record()
                self.group_bytes += bytes_per_rank * (nranks - 1)
                self._group_recs.append(('allgather_sketch_slabs', bytes_per_rank * (nranks - 1), a, b))
                return 0
            except Exception as ex:       # never let an exception cross the C boundary
                self._group_error = ex
                return -1

        self.group_bytes = 0
        self._group_recs = []
        self._group_error = None
        if nranks <= 1:
            self._group_cb = _lib.ALLGATHER_FN()      # NULL function pointer
            check(lib.ctmb_set_group(self._h, 0, 1, self._group_cb, None))
            return
        self._group_cb = _lib.ALLGATHER_FN(_allgather)    # keep a reference: ctypes callbacks die with their object
        check(lib.ctmb_set_group(self._h, rank, nranks, self._group_cb, None))

    def debug_set_matrix_free(self, mode):
        """Tests only: 0 = never, 1 = by size (default), 2 = always use the matrix-free projector path; the cached
        workspace sizes depend on it."""
        lib.ctmb_debug_set_matrix_free(int(mode))
        self._tables = {k: v for k, v in self._tables.items() if not (isinstance(k, tuple) and k and k[0] in ('ws', 'wsc4v'))}

    def debug_set_rdm_block_rows(self, rows):
        """tests: rdm2x2 processes the rows of its halves in blocks of `rows` (0: only where 32-bit offsets demand it)."""
        lib.ctmb_debug_set_rdm_block_rows.argtypes = [C.c_int]
        lib.ctmb_debug_set_rdm_block_rows.restype = None
        lib.ctmb_debug_set_rdm_block_rows(int(rows))
        self._tables = {k: v for k, v in self._tables.items() if not (isinstance(k, tuple) and k and k[0] in ('ws', 'wsc4v'))}

    def debug_set_m_noise(self, amp):
        """tests: Gaussian noise of relative amplitude `amp` (x max|M|) on the explicit M = R^T Rt before its decomposition."""
        lib.ctmb_debug_set_m_noise.argtypes = [C.c_double]
        lib.ctmb_debug_set_m_noise.restype = None
        lib.ctmb_debug_set_m_noise(float(amp))

    def group_reset(self):
        self._group_recs = []

    def group_totals(self):
        return list(getattr(self, '_group_recs', []))

    def rsvd_status(self, reset=False):
        """(residual checks, decompositions returned although they missed the residual bound, worst residual / bound)."""
        a, b, w = C.c_longlong(), C.c_longlong(), C.c_double()
        check(lib.ctmb_get_rsvd_status(self._h, C.byref(a), C.byref(b), C.byref(w), None, None, int(bool(reset))))
        return a.value, b.value, w.value

    def rsvd_iterations(self):
        """(iterative decompositions run, power iterations they took in total) since the last reset."""
        c, i = C.c_longlong(), C.c_longlong()
        check(lib.ctmb_get_rsvd_status(self._h, None, None, None, C.byref(c), C.byref(i), 0))
        return c.value, i.value

    def _warn_rsvd(self):
        _, missed, worst = self.rsvd_status()
        if missed > self._rsvd_missed_seen:
            import warnings
            self._rsvd_missed_seen = missed
            warnings.warn(f"libctmb: a truncated decomposition was returned with a residual {worst:.1f}x above the bound "
                          "rsvd_tol*sqrt(n) (rounding floor or rsvd_max_rounds reached); the projectors amplify this by S0/Sj",
                          RuntimeWarning, stacklevel=3)

    def counters(self):
        n, f = C.c_longlong(), C.c_double()
        check(lib.ctmb_get_counters(self._h, C.byref(n), C.byref(f)))
        return n.value, f.value

    def reset_counters(self):
        check(lib.ctmb_reset_counters(self._h))

    PROFILE_CLASSES = ('tc_gemm', 'qr', 'jacobi', 'misc')

    def profile(self, on):
        """Per-kernel-class CUDA-event timing inside libctmb (bench.py's roofline entry)."""
        check(lib.ctmb_profile_enable(self._h, int(bool(on))))

    def profile_totals(self):
        ms, fl, by = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_double * 4)()
        n = (C.c_longlong * 4)()
        check(lib.ctmb_profile_get(self._h, ms, fl, by, n))
        return {k: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=n[i]) for i, k in enumerate(self.PROFILE_CLASSES)}

    def _opts(self, **kw):
        o = _lib.Options()
        C.memmove(C.byref(o), C.byref(self.options), C.sizeof(o))
        for k, v in kw.items():
            if v is not None:
                setattr(o, k, v)
        return o

    @staticmethod
    def _prep(t, device):
        if t.device != device:
            raise ValueError(f"tensor on {t.device}, engine on {device}")
        return t if t.is_contiguous() else t.contiguous()

    def _site(self, a, Cs, Ts, keep):
        s = _lib.Site()
        a = self._prep(a, self.device); keep.append(a)
        s.a = a.data_ptr()
        dims = list(a.shape) if a.dim() == 5 else [0] + aux2(a)      # dims[0] == 0: double-layer tensor (include/ctmb.h)
        for i in range(5):
            s.dims[i] = dims[i]
        for i, t in enumerate(Cs):
            if t is not None:
                t = self._prep(t, self.device); keep.append(t); s.C[i] = t.data_ptr()
        for i, t in enumerate(Ts):
            if t is not None:
                t = self._prep(t, self.device); keep.append(t); s.T[i] = t.data_ptr()
        return s

    # ----------------------------------------------------------------------------------
    # piecewise entry points (parity tests mirror the reference's *_c functions)
    # ----------------------------------------------------------------------------------
    def einsum2(self, spec, A, B, conjA=False, conjB=False):
        if A.dtype in _WIDE and B.dtype == A.dtype:
            return self.einsum2(spec, _widen(A), _widen(B), conjA, conjB).to(A.dtype)
        lhs, out = spec.split('->')
        la, lb = lhs.split(',')
        A, B = self._prep(A, self.device), self._prep(B, self.device)
        ext = dict(zip(la, A.shape)); ext.update(zip(lb, B.shape))
        Cc = torch.empty([ext[c] for c in out], dtype=A.dtype, device=self.device)
        da = (C.c_longlong * len(la))(*A.shape)
        db = (C.c_longlong * len(lb))(*B.shape)
        check(lib.ctmb_einsum2(self._h, _dt(A), spec.encode(), _ptr(A), da, int(conjA), _ptr(B), db, int(conjB),
                               _ptr(Cc), self._stream()))
        return Cc

    def c2x2(self, kind, C_, T1, T2, a, chi):
        """kind: 'LU','RU','RD','LD'; C_,T1,T2 as in SURVEY Appendix A for that corner."""
        k = dict(LU=0, RU=1, RD=2, LD=3)[kind]
        cslot, t1slot, t2slot = [(0, 0, 1), (1, 3, 0), (2, 2, 3), (3, 1, 2)][k]
        Cs, Ts = [None] * 4, [None] * 4
        Cs[cslot], Ts[t1slot], Ts[t2slot] = C_, T1, T2
        keep = []
        s = self._site(a, Cs, Ts, keep)
        D = aux2(a)
        rows = chi * [D[2], D[1], D[0], D[0]][k]
        cols = chi * [D[3], D[2], D[1], D[3]][k]
        out = torch.empty((rows, cols), dtype=a.dtype, device=self.device)
        nb = lib.ctmb_c2x2_workspace(self._h, _dt(a), k, chi, C.byref(s))
        ws = self._workspace(nb)
        check(lib.ctmb_c2x2(self._h, _dt(a), k, chi, C.byref(s), _ptr(out), _ptr(ws), ws.numel(), self._stream()))
        return out

    def halves(self, direction, coord, state, env):
        """halves_of_4x4_CTM_MOVE_{UP,LEFT,DOWN,RIGHT} (ctm/generic/ctm_components.py:10-265): the pair (R, Rt) of the
        projector job at `coord`, built from the four enlarged corners of its 2x2 patch."""
        if direction not in DIRECTIONS:
            raise ValueError("Invalid direction: " + str(direction))
        keep, structs = [], []
        for dx, dy in PATCH[direction]:
            c = state.vertexToSite((coord[0] + dx, coord[1] + dy))
            structs.append(self._site(state.sites[c], [env.C[(c, k)] for k in C_KEYS], [env.T[(c, k)] for k in T_KEYS], keep))
        arr = (C.POINTER(_lib.Site) * 4)(*[C.pointer(s) for s in structs])
        a0 = state.sites[state.vertexToSite(coord)]
        dt = _dt(a0)
        # rows: the bond being truncated, at the first corner's site; columns: the same leg at the second corner's site
        leg = {(0, -1): 1, (-1, 0): 2, (0, 1): 3, (1, 0): 0}[direction]
        c1 = state.vertexToSite((coord[0] + PATCH[direction][1][0], coord[1] + PATCH[direction][1][1]))
        n0, n1 = env.chi * aux2(a0)[leg], env.chi * aux2(state.sites[c1])[leg]
        R = torch.empty((n0, n1), dtype=a0.dtype, device=self.device)
        Rt = torch.empty((n0, n1), dtype=a0.dtype, device=self.device)
        d = DIRECTIONS[direction]
        nb = lib.ctmb_halves_workspace(self._h, dt, d, env.chi, arr)
        if nb == 0:
            raise _lib.CtmbError(lib.ctmb_last_error().decode())
        ws = self._workspace(nb)
        check(lib.ctmb_halves(self._h, dt, d, env.chi, arr, _ptr(R), _ptr(Rt), _ptr(ws), ws.numel(), self._stream()))
        return R, Rt

    def projectors(self, R, Rt, chi, **opt):
        R, Rt = self._prep(R, self.device), self._prep(Rt, self.device)
        n0, n1 = R.shape
        o = self._opts(**opt)
        P = torch.empty((n0, chi), dtype=R.dtype, device=self.device)
        Pt = torch.empty_like(P)
        S = torch.empty(chi, dtype=torch.float64, device=self.device)
        nb = lib.ctmb_projectors_workspace(self._h, _dt(R), n0, n1, chi, C.byref(o))
        ws = self._workspace(nb)
        check(lib.ctmb_projectors(self._h, _dt(R), _ptr(R), _ptr(Rt), n0, n1, chi, C.byref(o), _ptr(P), _ptr(Pt),
                                  _ptr(S), _ptr(ws), ws.numel(), self._stream()))
        return P, Pt, S

    def truncated_svd(self, M, chi, **opt):
        """Returns U (m x chi), S (chi), V (n x chi) with M ~ U S V^H (sign-fixed, multiplet-aware)."""
        M = self._prep(M, self.device)
        m, n = M.shape
        o = self._opts(**opt)
        U = torch.empty((chi, m), dtype=M.dtype, device=self.device)
        V = torch.empty((chi, n), dtype=M.dtype, device=self.device)
        S = torch.empty(chi, dtype=torch.float64, device=self.device)
        nb = lib.ctmb_truncated_svd_workspace(self._h, _dt(M), m, n, chi, C.byref(o))
        ws = self._workspace(nb)
        check(lib.ctmb_truncated_svd(self._h, _dt(M), _ptr(M), m, n, chi, C.byref(o), _ptr(U), _ptr(S), _ptr(V),
                                     _ptr(ws), ws.numel(), self._stream()))
        return U.t(), S, V.t()

    def qr(self, M):
        """Thin QR of a rows x k matrix, rows >= k (torch.linalg.qr(M), LAPACK sign convention) -> (Q, R)."""
        M = self._prep(M, self.device)
        rows, k = M.shape
        A = M.t().contiguous()                      # row-major k x rows = column-major rows x k
        R = torch.empty((k, k), dtype=M.dtype, device=self.device)
        nb = lib.ctmb_qr_workspace(self._h, _dt(M), rows, k)
        if nb == 0:
            raise _lib.CtmbError(lib.ctmb_last_error().decode())
        ws = self._workspace(nb)
        check(lib.ctmb_qr(self._h, _dt(M), _ptr(A), rows, k, _ptr(R), _ptr(ws), ws.numel(), self._stream()))
        return A.t(), R.t()

    def truncated_eig_sym(self, M, chi, **opt):
        M = self._prep(M, self.device)
        n = M.shape[0]
        opt.setdefault('eps_multiplet', 1.0e-12)
        o = self._opts(**opt)
        U = torch.empty((chi, n), dtype=M.dtype, device=self.device)
        D = torch.empty(chi, dtype=torch.float64, device=self.device)
        nb = lib.ctmb_truncated_eig_sym_workspace(self._h, _dt(M), n, chi, C.byref(o))
        ws = self._workspace(nb)
        check(lib.ctmb_truncated_eig_sym(self._h, _dt(M), _ptr(M), n, chi, C.byref(o), _ptr(D), _ptr(U),
                                         _ptr(ws), ws.numel(), self._stream()))
        return D, U.t()

    # ----------------------------------------------------------------------------------
    # the moves
    # ----------------------------------------------------------------------------------
    def _move_tables(self, state, direction):
        coords = tuple(state.sites.keys())
        key = (id(state.vertexToSite), coords, direction)
        tab = self._tables.get(key)
        if tab is None:
            index = {c: i for i, c in enumerate(coords)}
            v2s = state.vertexToSite
            corner = []
            for c in coords:
                for dx, dy in PATCH[direction]:
                    corner.append(index[v2s((c[0] + dx, c[1] + dy))])
            sh = SHIFT[direction]
            nb = [index[v2s((c[0] + sh[0], c[1] + sh[1]))] for c in coords]
            dest = [v2s((c[0] - direction[0], c[1] - direction[1])) for c in coords]
            tab = ((C.c_int * len(corner))(*corner), (C.c_int * len(nb))(*nb), dest, state.vertexToSite)
            self._tables[key] = tab
        return tab

    def _nT_shape(self, direction, a, chi):
        D = aux2(a)
        if direction == (0, -1):
            return (chi, D[2], chi)
        if direction == (-1, 0):
            return (chi, chi, D[3])
        if direction == (0, 1):
            return (D[0], chi, chi)
        return (chi, D[1], chi)

    def _wide_state(self, state):
        key = ('wide', id(state), tuple(id(t) for t in state.sites.values()))
        st2 = self._tables.get(key)
        if st2 is None:
            self._tables = {k: v for k, v in self._tables.items() if not (isinstance(k, tuple) and k and k[0] == 'wide')}
            st2 = _Shadow(sites=type(state.sites)((c, _widen(t)) for c, t in state.sites.items()),
                          vertexToSite=state.vertexToSite, lX=getattr(state, 'lX', None), lY=getattr(state, 'lY', None))
            self._tables[key] = st2
        return st2

    def move_generic(self, direction, state, env, **opt):
        """One ctm_MOVE (ctm/generic/ctmrg.py:179-319): replaces the entries of env.C / env.T
        at coord-direction by freshly allocated tensors; inputs are never modified."""
        if direction not in DIRECTIONS:
            raise ValueError("Invalid direction: " + str(direction))
        coords = list(state.sites.keys())
        n = len(coords)
        chi = env.chi
        refuse_autograd('ctm_MOVE', list(state.sites.values()) + list(env.C.values()) + list(env.T.values()))
        low = next(iter(state.sites.values())).dtype
        if low in _WIDE:
            st2 = self._wide_state(state)
            env2 = _Shadow(chi=chi, C={k: _widen(v) for k, v in env.C.items()}, T={k: _widen(v) for k, v in env.T.items()})
            before = {('C',) + k: v for k, v in env2.C.items()}
            before.update({('T',) + k: v for k, v in env2.T.items()})
            self.move_generic(direction, st2, env2, **opt)
            for k, v in env2.C.items():
                if v is not before[('C',) + k]:
                    env.C[k] = v.to(low)
            for k, v in env2.T.items():
                if v is not before[('T',) + k]:
                    env.T[k] = v.to(low)
            return
        corner, nb, dest, _ = self._move_tables(state, direction)
        keep = []
        sites = (_lib.Site * n)()
        a0 = None
        for i, c in enumerate(coords):
            a = state.sites[c]
            a0 = a if a0 is None else a0
            sites[i] = self._site(a, [env.C[(c, k)] for k in C_KEYS], [env.T[(c, k)] for k in T_KEYS], keep)
        dt = _dt(a0)
        nC1 = [torch.empty((chi, chi), dtype=a0.dtype, device=self.device) for _ in range(n)]
        nC2 = [torch.empty((chi, chi), dtype=a0.dtype, device=self.device) for _ in range(n)]
        nT = [torch.empty(self._nT_shape(direction, state.sites[c], chi), dtype=a0.dtype, device=self.device)
              for c in coords]
        p1 = (C.c_void_p * n)(*[t.data_ptr() for t in nC1])
        p2 = (C.c_void_p * n)(*[t.data_ptr() for t in nC2])
        p3 = (C.c_void_p * n)(*[t.data_ptr() for t in nT])
        o = self._opts(**opt)
        d = DIRECTIONS[direction]
        wkey = ('ws', dt, d, n, chi, tuple(tuple(state.sites[c].shape) for c in coords), o.rsvd_rank_factor,
                o.projector_method)
        nbytes = self._tables.get(wkey)
        if nbytes is None:
            nbytes = lib.ctmb_move_generic_workspace(self._h, dt, d, n, chi, sites, corner, nb, C.byref(o))
            if nbytes == 0:
                raise _lib.CtmbError(lib.ctmb_last_error().decode())
            self._tables[wkey] = nbytes
        ws = self._workspace(nbytes)
        check(lib.ctmb_move_generic(self._h, dt, d, n, chi, sites, corner, nb, C.byref(o), p1, p2, p3,
                                    _ptr(ws), ws.numel(), self._stream()))
        self._warn_rsvd()
        kC1, kC2, kT = OUT_KEYS[direction]
        for i in range(n):
            env.C[(dest[i], kC1)] = nC1[i]
            env.C[(dest[i], kC2)] = nC2[i]
            env.T[(dest[i], kT)] = nT[i]

    # ----------------------------------------------------------------------------------
    # the two halves of a move, restricted to a subset of the site jobs (multi-GPU shard)
    # ----------------------------------------------------------------------------------
    def _sites_array(self, state, env, keep):
        coords = list(state.sites.keys())
        sites = (_lib.Site * len(coords))()
        for i, c in enumerate(coords):
            sites[i] = self._site(state.sites[c], [env.C[(c, k)] for k in C_KEYS], [env.T[(c, k)] for k in T_KEYS], keep)
        return coords, sites

    def projector_shape(self, direction, state, env):
        """(n0, chi) of the projectors of `direction` (uniform bond dimensions)."""
        D = aux2(next(iter(state.sites.values())))
        leg = {(0, -1): D[1], (-1, 0): D[2], (0, 1): D[3], (1, 0): D[0]}[direction]   # bond being truncated (row index of R)
        return env.chi * leg, env.chi

    def move_generic_projectors(self, direction, state, env, jobs, **opt):
        """Projector pairs (P, Pt), each n0 x chi, of the listed site jobs (ctm_get_projectors_4x4)."""
        keep = []
        coords, sites = self._sites_array(state, env, keep)
        n, chi = len(coords), env.chi
        corner, nb, dest, _ = self._move_tables(state, direction)
        a0 = state.sites[coords[0]]
        dt = _dt(a0)
        n0, _ = self.projector_shape(direction, state, env)
        P = [torch.empty((n0, chi), dtype=a0.dtype, device=self.device) for _ in jobs]
        Pt = [torch.empty((n0, chi), dtype=a0.dtype, device=self.device) for _ in jobs]
        o = self._opts(**opt)
        d = DIRECTIONS[direction]
        nbytes = lib.ctmb_move_generic_workspace(self._h, dt, d, n, chi, sites, corner, nb, C.byref(o))
        ws = self._workspace(nbytes)
        jl = (C.c_int * len(jobs))(*jobs)
        pp = (C.c_void_p * len(jobs))(*[t.data_ptr() for t in P])
        ppt = (C.c_void_p * len(jobs))(*[t.data_ptr() for t in Pt])
        check(lib.ctmb_move_generic_projectors(self._h, dt, d, n, chi, sites, corner, len(jobs), jl, C.byref(o), pp, ppt,
                                               _ptr(ws), ws.numel(), self._stream()))
        return P, Pt

    def move_generic_absorb(self, direction, state, env, jobs, P_all, Pt_all, **opt):
        """Absorb + truncate + normalise the listed jobs given the projectors of ALL sites;
        returns [(dest_coord, nC1, nC2, nT)] without touching env."""
        keep = []
        coords, sites = self._sites_array(state, env, keep)
        n, chi = len(coords), env.chi
        corner, nb, dest, _ = self._move_tables(state, direction)
        a0 = state.sites[coords[0]]
        dt = _dt(a0)
        nC1 = [torch.empty((chi, chi), dtype=a0.dtype, device=self.device) for _ in jobs]
        nC2 = [torch.empty((chi, chi), dtype=a0.dtype, device=self.device) for _ in jobs]
        nT = [torch.empty(self._nT_shape(direction, state.sites[coords[j]], chi), dtype=a0.dtype, device=self.device)
              for j in jobs]
        d = DIRECTIONS[direction]
        o = self._opts(**opt)
        nbytes = lib.ctmb_move_generic_workspace(self._h, dt, d, n, chi, sites, corner, nb, C.byref(o))
        ws = self._workspace(nbytes)
        jl = (C.c_int * len(jobs))(*jobs)
        pa = (C.c_void_p * n)(*[t.data_ptr() for t in P_all])
        pta = (C.c_void_p * n)(*[t.data_ptr() for t in Pt_all])
        p1 = (C.c_void_p * len(jobs))(*[t.data_ptr() for t in nC1])
        p2 = (C.c_void_p * len(jobs))(*[t.data_ptr() for t in nC2])
        p3 = (C.c_void_p * len(jobs))(*[t.data_ptr() for t in nT])
        check(lib.ctmb_move_generic_absorb(self._h, dt, d, n, chi, sites, nb, len(jobs), jl, C.byref(o), pa, pta, p1, p2, p3,
                                           _ptr(ws), ws.numel(), self._stream()))
        return [(dest[j], nC1[i], nC2[i], nT[i]) for i, j in enumerate(jobs)]

    def move_c4v(self, a, C_, T, chi, **opt):
        """One ctm_MOVE_sl (ctm/one_site_c4v/ctmrg_c4v.py:325-463) or, with a double-layer A[u,l,d,r],
        ctm_MOVE_dl (:200-322) -> (C', T', D)."""
        refuse_autograd('ctm_MOVE_sl / ctm_MOVE_dl', (a, C_, T))
        if a.dtype in _WIDE:
            Co, To, Dv = self.move_c4v(_widen(a), _widen(C_), _widen(T), chi, **opt)
            return Co.to(C_.dtype), To.to(T.dtype), Dv
        a, C_, T = self._prep(a, self.device), self._prep(C_, self.device), self._prep(T, self.device)
        dt = _dt(a)
        opt.setdefault('eps_multiplet', 1.0e-12)      # truncated_eig_sym default (custom_eig.py:7-8)
        o = self._opts(**opt)
        dims = (C.c_int * 5)(*(list(a.shape) if a.dim() == 5 else [0] + aux2(a)))
        Co = torch.empty_like(C_)
        To = torch.empty_like(T)
        Dv = torch.empty(chi, dtype=torch.float64, device=self.device)
        wkey = ('wsc4v', dt, tuple(a.shape), chi, o.rsvd_rank_factor)
        nbytes = self._tables.get(wkey)
        if nbytes is None:
            nbytes = lib.ctmb_move_c4v_workspace(self._h, dt, dims, chi, C.byref(o))
            if nbytes == 0:
                raise _lib.CtmbError(lib.ctmb_last_error().decode())
            self._tables[wkey] = nbytes
        ws = self._workspace(nbytes)
        check(lib.ctmb_move_c4v(self._h, dt, _ptr(a), dims, _ptr(C_), _ptr(T), chi, C.byref(o), _ptr(Co), _ptr(To),
                                _ptr(Dv), _ptr(ws), ws.numel(), self._stream()))
        self._warn_rsvd()
        return Co, To, Dv


    # ----------------------------------------------------------------------------------
    # observables on the converged environment
    # ----------------------------------------------------------------------------------
    def sym_pos_def(self, rdm, sym_pos_def=False):
        """_sym_pos_def_rdm (ctm/generic/rdm.py:38-68): hermitise, optionally project on the positive part, normalise the trace."""
        if rdm.dtype in _WIDE:
            return self.sym_pos_def(_widen(rdm), sym_pos_def).to(rdm.dtype)
        shape = rdm.shape
        assert len(shape) % 2 == 0, "invalid rank of RDM"
        n = 1
        for d in shape[:len(shape) // 2]:
            n *= d
        raw = self._prep(rdm, self.device)
        out = torch.empty((n, n), dtype=raw.dtype, device=self.device)
        dt = _dt(raw)
        nbytes = lib.ctmb_sym_pos_def_workspace(self._h, dt, n, int(bool(sym_pos_def)))
        if nbytes == 0:
            raise _lib.CtmbError(lib.ctmb_last_error().decode())
        ws = self._workspace(nbytes)
        check(lib.ctmb_sym_pos_def(self._h, dt, _ptr(raw), n, int(bool(sym_pos_def)), _ptr(out), _ptr(ws), ws.numel(),
                                   self._stream()))
        return out.view(shape)

    def rdm2x2_sites(self, tensors4, chi, open_sites=(0, 1, 2, 3), sym_pos_def=False, raw=False):
        """2x2 reduced density matrix from explicit per-site data: tensors4[q] = (a, [C x4 in C_KEYS order], [T x4 in
        T_KEYS order]) for the sites s0 s1 / s2 s3 of the plaquette."""
        open_sites = sorted(set(int(q) for q in open_sites))
        if not open_sites or open_sites[0] < 0 or open_sites[-1] > 3:
            raise ValueError("open_sites must be a non-empty subset of [0,1,2,3]")
        mask = sum(1 << q for q in open_sites)
        refuse_autograd('rdm2x2', [t for (a, Cs, Ts) in tensors4 for t in [a] + list(Cs) + list(Ts)])
        if tensors4[0][0].dtype in _WIDE:
            wide = [(_widen(a), [_widen(t) for t in Cs], [_widen(t) for t in Ts]) for (a, Cs, Ts) in tensors4]
            return self.rdm2x2_sites(wide, chi, open_sites, sym_pos_def, raw).to(tensors4[0][0].dtype)
        keep = []
        structs = [self._site(a, Cs, Ts, keep) for (a, Cs, Ts) in tensors4]
        arr = (C.POINTER(_lib.Site) * 4)(*[C.pointer(s) for s in structs])
        a0 = tensors4[0][0]
        dt = _dt(a0)
        dims = [tensors4[q][0].shape[0] for q in open_sites]
        rho = torch.empty(dims + dims, dtype=a0.dtype, device=self.device)
        nbytes = lib.ctmb_rdm2x2_workspace(self._h, dt, chi, arr, mask)
        if nbytes == 0:
            raise _lib.CtmbError(lib.ctmb_last_error().decode())
        ws = self._workspace(nbytes)
        check(lib.ctmb_rdm2x2(self._h, dt, chi, arr, mask, _ptr(rho), _ptr(ws), ws.numel(), self._stream()))
        return rho if raw else self.sym_pos_def(rho, sym_pos_def)

    def rdm_small_sites(self, kind, tensors2, chi, sym_pos_def=False, raw=False):
        """One- / two-site density matrix from explicit per-site data (see rdm2x2_sites); kind '1x1', '2x1' (second site
        to the right) or '1x2' (second site below)."""
        k = {'1x1': 0, '2x1': 1, '1x2': 2}[kind]
        refuse_autograd('rdm' + kind, [t for (a, Cs, Ts) in tensors2 for t in [a] + list(Cs) + list(Ts)])
        if tensors2[0][0].dtype in _WIDE:
            wide = [(_widen(a), [_widen(t) for t in Cs], [_widen(t) for t in Ts]) for (a, Cs, Ts) in tensors2]
            return self.rdm_small_sites(kind, wide, chi, sym_pos_def, raw).to(tensors2[0][0].dtype)
        keep = []
        structs = [self._site(a, Cs, Ts, keep) for (a, Cs, Ts) in tensors2]
        if len(structs) == 1:
            structs = structs * 2
        arr = (C.POINTER(_lib.Site) * 2)(*[C.pointer(s) for s in structs[:2]])
        a0 = tensors2[0][0]
        dt = _dt(a0)
        dims = [t[0].shape[0] for t in tensors2[:1 if k == 0 else 2]]
        rho = torch.empty(dims + dims, dtype=a0.dtype, device=self.device)
        nbytes = lib.ctmb_rdm_small_workspace(self._h, dt, k, chi, arr)
        if nbytes == 0:
            raise _lib.CtmbError(lib.ctmb_last_error().decode())
        ws = self._workspace(nbytes)
        check(lib.ctmb_rdm_small(self._h, dt, k, chi, arr, _ptr(rho), _ptr(ws), ws.numel(), self._stream()))
        return rho if raw else self.sym_pos_def(rho, sym_pos_def)

    def rdm_small(self, kind, coord, state, env, sym_pos_def=False, raw=False):
        """rdm1x1 / rdm2x1 / rdm1x2 (ctm/generic/rdm.py:71-112, 304-350, 622-670) at vertex `coord`."""
        shifts = {'1x1': [(0, 0)], '2x1': [(0, 0), (1, 0)], '1x2': [(0, 0), (0, 1)]}[kind]
        t2 = []
        for dx, dy in shifts:
            c = state.vertexToSite((coord[0] + dx, coord[1] + dy))
            t2.append((state.sites[c], [env.C[(c, k)] for k in C_KEYS], [env.T[(c, k)] for k in T_KEYS]))
        return self.rdm_small_sites(kind, t2, env.chi, sym_pos_def, raw)

    def rdm2x2(self, coord, state, env, open_sites=(0, 1, 2, 3), sym_pos_def=False, raw=False):
        """rdm2x2 (ctm/generic/rdm.py:1306-1592): rho[s0,s1,s2,s3; s0',s1',s2',s3'] of the plaquette with upper-left
        vertex `coord` (s0 = coord, s1 = coord+(1,0), s2 = coord+(0,1), s3 = coord+(1,1))."""
        t4 = []
        for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):
            c = state.vertexToSite((coord[0] + dx, coord[1] + dy))
            t4.append((state.sites[c], [env.C[(c, k)] for k in C_KEYS], [env.T[(c, k)] for k in T_KEYS]))
        return self.rdm2x2_sites(t4, env.chi, open_sites, sym_pos_def, raw)


_default = None


def default_engine():
    """Process-wide engine on the current CUDA device (created on first use)."""
    global _default
    if _default is None:
        _default = CtmEngine()
    return _default
