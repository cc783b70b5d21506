"""CPU oracle for the CTMRG hot path of peps-torch  --  TEST INFRASTRUCTURE ONLY.

This file is a plain restatement, on torch-CPU float64/complex128 tensors, of the
algorithm the reference implements in

  * ctm/generic/ctmrg.py            (ctm_MOVE, absorb_truncate_CTM_MOVE_*, run)
  * ctm/generic/ctm_components.py   (c2x2_*_sl_c, halves_of_4x4_CTM_MOVE_*_c)
  * ctm/generic/ctm_projectors.py   (ctm_get_projectors_from_matrices)
  * linalg/custom_svd.py, linalg/svd_gesdd.py   (truncated_svd_gesdd, fix_svd_signs)
  * linalg/custom_eig.py, linalg/eig_sym.py     (truncated_eig_sym)
  * ctm/one_site_c4v/ctmrg_c4v.py, ctm_components_c4v.py (ctm_MOVE_sl, c2x2_sl)
  * ctm/generic/env.py, ctm/one_site_c4v/env_c4v.py (init_from_ipeps_pbc)
  * ctm/generic/rdm.py (rdm2x2 incl. open_sites, rdm1x1, rdm2x1, rdm1x2, _sym_pos_def_matrix)
    + models/j1j2.py (get_hp, energy_per_site, energy_1x1_lowmem)
  * ctm/one_site_c4v/rdm_c4v.py (rdm2x2_NN/NNN_lowmem_sl, rdm2x2, rdm1x1_sl, rdm2x1_sl), ctmrg_c4v.py (ctm_MOVE_dl)
  * the non-default variants of the move: projector_method='4X2' (ctm_projectors.py:66-136), double-layer on-site
    tensors (ctm_force_dl, run_overlap), 2-norm normalisation

It exists to CHECK the CUDA path (tests/, __graft_entry__.smoke(), and the
cpu_baseline / --impl reference legs of bench.py).  Nothing under
peps_torch_b200/ may import it; the product path has no CPU fallback.

Parity pin: every function here is compared against the reference itself
(imported from /root/reference in the build container) by oracle/gen_golden.py,
which also writes the fixtures under tests/golden/ that the CPU test-suite
re-checks without the reference being present.  The restatements added after the
fixtures were generated (variants of the move, density matrices) are compared with
the unmodified reference by tests/test_oracle_vs_reference_cpu.py -- element-wise for
the deterministic parts -- every time the CPU suite runs where /root/reference exists.

Conventions (ctm/generic/env.py:57-77): on-site a[s,u,l,d,r]; C(-1,-1)[down,right],
C(1,-1)[left,down], C(1,1)[up,left], C(-1,1)[up,right]; T(0,-1)[l,d,r], T(-1,0)[u,d,r],
T(0,1)[u,l,r], T(1,0)[u,l,d]; fused double-layer legs are (ket,bra), ket-major.
"""
import time
import numpy as np
import torch

UP, LEFT, DOWN, RIGHT = (0, -1), (-1, 0), (0, 1), (1, 0)
DIRECTIONS = [UP, LEFT, DOWN, RIGHT]  # config.py:392 default ctm_move_sequence


class OracleArgs:
    """Hot-path knobs of CTMARGS (config.py:369-409) with the reference defaults."""
    ctm_max_iter = 50
    ctm_absorb_normalization = 'inf'
    projector_svd_reltol = 1.0e-8
    projector_eps_multiplet = 1.0e-8
    projector_multiplet_abstol = 1.0e-14
    svd_driver = None            # LAPACK driver for torch.linalg.svd (None = gesdd)
    projector_method = '4X4'     # or '4X2' (ctm_projectors.py:66-136)
    ctm_force_dl = False

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


# --------------------------------------------------------------------------------------
# single-layer evaluation of a double-layer einsum
# --------------------------------------------------------------------------------------
def _pairwise(subs, ops, out):
    """Left-to-right pairwise evaluation (what torch.einsum does without opt_einsum,
    and what the chains of tensordots in ctm_components.py do)."""
    cur_idx, cur = subs[0], ops[0]
    for i in range(1, len(ops)):
        later = set(''.join(subs[i + 1:]) + out)
        new_idx = ''.join(ch for ch in dict.fromkeys(cur_idx + subs[i]) if ch in later)
        cur = torch.einsum(f'{cur_idx},{subs[i]}->{new_idx}', cur, ops[i])
        cur_idx = new_idx
    if cur_idx != out:
        cur = torch.einsum(f'{cur_idx}->{out}', cur)
    return cur


def sl_einsum(spec, ops, a, open_phys=False):
    """Evaluate the double-layer einsum `spec`, in which the operand written as '@' stands
    for the double-layer tensor A[u,l,d,r] = sum_s a[s,u,l,d,r] conj(a)[s,U,L,D,R] (fused
    (ket,bra)), WITHOUT forming A: every fused letter that belongs to A is split into a
    (ket, bra) pair on every operand that carries it; `a` and conj(a) enter as two separate
    operands, ket layer first (ctm_components.py:372-434).

    spec example: 'ab,buc,ael,@ulfg->efcg'.  Lower-case letters only; 's'/'S' reserved.
    With open_phys the physical indices (s ket, S bra) are appended to the output.
    Returns the tensor with the (ket,bra) pairs of the output fused again."""
    lhs, out = spec.split('->')
    terms = lhs.split(',')
    a_pos = [i for i, t in enumerate(terms) if t.startswith('@')]
    assert len(a_pos) == 1
    a_pos = a_pos[0]
    a_idx = terms[a_pos][1:]
    if a.dim() == 4:
        # `a` IS the double-layer tensor A[u,l,d,r] (ctm_force_dl, ctmrg.py:51-61; run_overlap :137-147;
        # ctm_MOVE_dl ctmrg_c4v.py:229-233): the reference's *_c functions contract it as one operand
        # (ctm_components.py:321-370 "dl" branches)
        assert not open_phys
        xs = list(ops[:a_pos]) + [a] + list(ops[a_pos:])
        subs = [t[1:] if t.startswith('@') else t for t in terms]
        return _pairwise(subs, xs, out)
    assert len(a_idx) == 4 and 's' not in lhs
    D = {ch: a.shape[1 + i] for i, ch in enumerate(a_idx)}
    subs, xs = [], []
    k = 0
    for i, t in enumerate(terms):
        if i == a_pos:
            if open_phys:
                subs += ['s' + a_idx, 'S' + a_idx.upper()]
            else:
                subs += ['s' + a_idx, 's' + a_idx.upper()]
            xs += [a, a.conj()]
            continue
        x = ops[k]; k += 1
        shape, idx = [], ''
        for ch, n in zip(t, x.shape):
            if ch in D:
                assert n == D[ch] ** 2, (spec, t, ch, n)
                shape += [D[ch], D[ch]]; idx += ch + ch.upper()
            else:
                shape += [n]; idx += ch
        subs.append(idx); xs.append(x.reshape(shape))
    out_x = ''.join(ch + ch.upper() if ch in D else ch for ch in out)
    if open_phys:
        out_x += 'sS'
    res = _pairwise(subs, xs, out_x)
    shape, i = [], 0
    for ch in out:
        if ch in D:
            shape.append(D[ch] ** 2); i += 2
        else:
            shape.append(res.shape[i]); i += 1
    shape += list(res.shape[i:])
    return res.reshape(shape)


def double_layer(a, bra=None):
    """A[u,l,d,r] with fused (ket,bra) legs (ctm/generic/ctmrg.py:51-61); with `bra` the overlap tensor of
    run_overlap (ctmrg.py:137-147: ket = state1, bra = state2)."""
    bra = a if bra is None else bra
    d = a.shape
    A = torch.einsum('mefgh,mabcd->eafbgchd', a, bra.conj()).contiguous()
    return A.view(d[1] ** 2, d[2] ** 2, d[3] ** 2, d[4] ** 2)


# --------------------------------------------------------------------------------------
# enlarged corners  (ctm_components.py:314-885)
# --------------------------------------------------------------------------------------
# kind -> (C key, T1 key, T2 key, einsum); output rows x cols = (chi_T2 d)(chi_T1 d)
CORNERS = {
    'LU': ((-1, -1), (0, -1), (-1, 0), 'ab,buc,ael,@ulfg->efcg'),
    'RU': ((1, -1), (1, 0), (0, -1), 'ab,brc,eua,@ulfr->elcf'),
    'RD': ((1, 1), (0, 1), (1, 0), 'ab,feb,cra,@ulfr->cuel'),
    'LD': ((-1, 1), (-1, 0), (0, 1), 'ab,cal,fbe,@ulfr->cuer'),
}


def c2x2(kind, C, T1, T2, a, open_phys=False):
    """Enlarged corner as an n x n matrix (n = chi * D^2), optionally with open (s,S)."""
    spec = CORNERS[kind][3]
    t = sl_einsum(spec, (C, T1, T2), a, open_phys=open_phys)
    n0, n1 = t.shape[0] * t.shape[1], t.shape[2] * t.shape[3]
    return t.reshape((n0, n1) + tuple(t.shape[4:]))


def corner_at(kind, coord, sites, v2s, C, T, open_phys=False):
    s = v2s(coord)
    kc, k1, k2, _ = CORNERS[kind]
    return c2x2(kind, C[(s, kc)], T[(s, k1)], T[(s, k2)], sites[s], open_phys=open_phys)


# --------------------------------------------------------------------------------------
# halves (ctm_components.py:55-75,123-139,186-201,249-265)
# each entry: (kind, dx, dy, transpose?) for R = X1.X2 and Rt = X3.X4
# --------------------------------------------------------------------------------------
HALVES = {
    UP: ((('RU', 0, 0, False), ('RD', 0, 1, False)), (('LU', -1, 0, True), ('LD', -1, 1, False))),
    LEFT: ((('LU', 0, 0, False), ('RU', 1, 0, False)), (('LD', 0, 1, False), ('RD', 1, 1, True))),
    DOWN: ((('LD', 0, 0, True), ('LU', 0, -1, False)), (('RD', 1, 0, True), ('RU', 1, -1, True))),
    RIGHT: ((('RD', 0, 0, False), ('LD', -1, 0, True)), (('RU', 0, -1, True), ('LU', -1, -1, True))),
}


def corners_4x2(direction, coord, sites, v2s, C, T):
    """ctm_get_projectors_4x2 (ctm_projectors.py:118-136): R and Rt are the first enlarged corner of each half
    (same corner kinds, offsets and transpositions as the first factors of halves_of_4x4_CTM_MOVE_*)."""
    res = []
    for pair in HALVES[direction]:
        kind, dx, dy, tr = pair[0]
        m = corner_at(kind, (coord[0] + dx, coord[1] + dy), sites, v2s, C, T)
        res.append(m.t() if tr else m)
    return res[0], res[1]


def halves(direction, coord, sites, v2s, C, T):
    res = []
    for pair in HALVES[direction]:
        ms = []
        for kind, dx, dy, tr in pair:
            m = corner_at(kind, (coord[0] + dx, coord[1] + dy), sites, v2s, C, T)
            ms.append(m.t() if tr else m)
        res.append(ms[0] @ ms[1])
    return res[0], res[1]


# --------------------------------------------------------------------------------------
# truncated SVD / EIG front-ends
# --------------------------------------------------------------------------------------
def fix_svd_signs(U, V):
    """linalg/svd_gesdd.py:18-26: make the first max-|.| entry of each U column real positive."""
    Uamp = (U.abs() * (2 ** 40)).to(dtype=torch.int64)
    ii = torch.argmax(Uamp, dim=0, keepdim=True)
    phase = torch.take_along_dim(U, ii, dim=0)
    phase = phase / phase.abs()
    return U * phase.conj().reshape(1, -1), V * phase.conj().reshape(1, -1)


def multiplet_chi(s_abs, chi, eps_multiplet, abs_tol):
    """custom_svd.py:70-88 / custom_eig.py:39-57.  s_abs: at least chi+1 leading values
    (descending). Returns chi_new; columns > chi_new are zeroed by the caller."""
    chi_new = chi
    gaps = s_abs[:chi + 1].clone()
    gaps[gaps < abs_tol] = 0.
    gaps = (gaps[:chi] - s_abs[1:chi + 1]) / (gaps[:chi] + 1.0e-16)
    gaps[gaps > 1.0] = 0.
    if gaps[chi - 1] < eps_multiplet:
        for i in range(chi - 1, -1, -1):
            if gaps[i] > eps_multiplet:
                chi_new = i
                break
    return chi_new


def truncated_svd(M, chi, eps_multiplet, abs_tol, driver=None):
    """custom_svd.py:38-101 with keep_multiplets=True; M = U S V^dag."""
    if driver == 'scipy_gesvd':
        import scipy.linalg
        u, s, vh = scipy.linalg.svd(M.numpy(), lapack_driver='gesvd')
        U, S, Vh = torch.from_numpy(u), torch.from_numpy(s), torch.from_numpy(vh)
    else:
        U, S, Vh = torch.linalg.svd(M)
    V = Vh.transpose(-2, -1).conj()
    U, V = fix_svd_signs(U, V)
    if chi < S.shape[0]:
        chi_new = multiplet_chi(S, chi, eps_multiplet, abs_tol)
        St = S[:chi].clone(); St[chi_new + 1:] = 0.
        Ut = U[:, :chi].clone(); Ut[:, chi_new + 1:] = 0.
        Vt = V[:, :chi].clone(); Vt[:, chi_new + 1:] = 0.
        return Ut, St, Vt
    k = min(chi, S.shape[0])
    return U[:, :k], S[:k], V[:, :k]


def truncated_eig_sym(M, chi, eps_multiplet=1.0e-12, abs_tol=1.0e-14):
    """custom_eig.py:7-67 + eig_sym.py:14-34 (eigh, sort by |lambda| descending)."""
    D, U = torch.linalg.eigh(M)
    _, p = torch.sort(torch.abs(D), descending=True)
    D, U = D[p], U[:, p]
    if chi < D.shape[0]:
        chi_new = multiplet_chi(torch.abs(D), chi, eps_multiplet, abs_tol)
        Dt = D[:chi].clone(); Dt[chi_new + 1:] = 0.
        Ut = U[:, :chi].clone(); Ut[:, chi_new + 1:] = 0.
        return Dt, Ut
    k = min(chi, D.shape[0])
    return D[:k], U[:, :k]


def projectors_from_matrices(R, Rt, chi, args, return_svd=False):
    """ctm_projectors.py:142-293 (projector_full_matrices=True)."""
    M = R.t() @ Rt
    U, S, V = truncated_svd(M, chi, args.projector_eps_multiplet,
                            args.projector_multiplet_abstol, driver=args.svd_driver)
    nz = S / S[0] > args.projector_svd_reltol
    S_nz = S[nz]
    S_sqrt = S * 0
    S_sqrt[:S_nz.shape[0]] = torch.rsqrt(S_nz)
    P = (R @ U.conj()) * S_sqrt[None, :]
    Pt = (Rt @ V) * S_sqrt[None, :]
    if return_svd:
        return P, Pt, (M, U, S, V)
    return P, Pt


# --------------------------------------------------------------------------------------
# absorption (ctmrg.py:324-804)
# --------------------------------------------------------------------------------------
# direction -> dict(keys of C1,T1,T,T2,C2; shift of the neighbour whose P1,Pt1 are used;
#   einsums for nC1 (C1,T1,Pt1), nC2 (C2,T2,P2), nT (4 operands incl. '@');
#   operand order of nT; which new env tensors are written)
ABSORB = {
    UP: dict(C1=(1, -1), T1=(1, 0), T=(0, -1), T2=(-1, 0), C2=(-1, -1), shift=(1, 0),
             nC1='ab,brc,arx->xc', nC2='ab,ael,blx->ex',
             nT=('auc,alx,@uldr,cry->xdy', ('T', 'Pt2', 'P1')),
             out=((1, -1), (-1, -1), (0, -1))),
    LEFT: dict(C1=(-1, -1), T1=(0, -1), T=(-1, 0), T2=(0, 1), C2=(-1, 1), shift=(0, -1),
               nC1='ab,buc,aux->xc', nC2='ab,fbe,afx->xe',
               nT=('acl,aux,@uldr,cdy->xyr', ('T', 'P1', 'Pt2')),
               out=((-1, -1), (-1, 1), (-1, 0))),
    DOWN: dict(C1=(-1, 1), T1=(-1, 0), T=(0, 1), T2=(1, 0), C2=(1, 1), shift=(-1, 0),
               nC1='ab,cal,blx->cx', nC2='ab,cra,brx->cx',
               nT=('fbe,blx,@ulfr,ery->uxy', ('T', 'P1', 'Pt2')),
               out=((-1, 1), (1, 1), (0, 1))),
    RIGHT: dict(C1=(1, 1), T1=(0, 1), T=(1, 0), T2=(0, -1), C2=(1, -1), shift=(0, 1),
                nC1='ab,feb,afx->xe', nC2='ab,eua,bux->ex',
                nT=('arc,aux,@uldr,cdy->xly', ('T', 'Pt2', 'P1')),
                out=((1, 1), (1, -1), (1, 0))),
}


def _normalize(t, norm_type):
    if norm_type == 'inf':
        return t / t.abs().max()
    return t / torch.linalg.vector_norm(t, ord=2)


def absorb(direction, coord, sites, v2s, C, T, P, Pt, args):
    sp = ABSORB[direction]
    s = v2s(coord)
    sn = v2s((coord[0] + sp['shift'][0], coord[1] + sp['shift'][1]))
    a = sites[s]
    chi_new = P[s].shape[1]
    # view the projectors as (chi, d, chi_new); d is the bond dim squared of the leg they act on
    C1, T1, Tm, T2, C2 = (C[(s, sp['C1'])], T[(s, sp['T1'])], T[(s, sp['T'])],
                          T[(s, sp['T2'])], C[(s, sp['C2'])])
    def _pview(p, chi_env):
        return p.view(chi_env, p.shape[0] // chi_env, chi_new)
    chi_env = C1.shape[0]
    P2, Pt2 = _pview(P[s], chi_env), _pview(Pt[s], chi_env)
    P1, Pt1 = _pview(P[sn], chi_env), _pview(Pt[sn], chi_env)
    nC1 = torch.einsum(sp['nC1'], C1, T1, Pt1)
    nC2 = torch.einsum(sp['nC2'], C2, T2, P2)
    spec, order = sp['nT']
    pool = dict(T=Tm, P1=P1, Pt1=Pt1, P2=P2, Pt2=Pt2)
    nT = sl_einsum(spec, tuple(pool[k] for k in order), a)
    nt = args.ctm_absorb_normalization
    return _normalize(nC1, nt), _normalize(nC2, nt), _normalize(nT, nt)


def ctm_move(direction, sites, v2s, C, T, chi, args=None, return_proj=False):
    """One directional move over all sites (ctmrg.py:179-319). C, T are updated in place
    (dict entries replaced). `sites`: OrderedDict coord -> a[s,u,l,d,r]."""
    args = args or OracleArgs()
    P, Pt = {}, {}
    for coord in sites.keys():
        if args.projector_method == '4X4':
            R, Rt = halves(direction, coord, sites, v2s, C, T)
        elif args.projector_method == '4X2':
            R, Rt = corners_4x2(direction, coord, sites, v2s, C, T)
        else:
            raise ValueError("Invalid Projector method: " + str(args.projector_method))
        P[coord], Pt[coord] = projectors_from_matrices(R, Rt, chi, args)
    new = {}
    for coord in sites.keys():
        new[coord] = absorb(direction, coord, sites, v2s, C, T, P, Pt, args)
    kC1, kC2, kT = ABSORB[direction]['out']
    for coord in sites.keys():
        nc = v2s((coord[0] - direction[0], coord[1] - direction[1]))
        C[(nc, kC1)], C[(nc, kC2)], T[(nc, kT)] = new[coord]
    if return_proj:
        return P, Pt


def ctm_iteration(sites, v2s, lX, lY, C, T, chi, args=None):
    """ctmrg.py:63-69: for each direction, lX (left/right) or lY (up/down) moves."""
    n = 0
    for direction in DIRECTIONS:
        reps = lX if direction in (LEFT, RIGHT) else lY
        for _ in range(reps):
            ctm_move(direction, sites, v2s, C, T, chi, args)
            n += 1
    return n


def run(sites, v2s, lX, lY, C, T, chi, n_iter, args=None):
    """ctmrg.py:18-110 with conv_check=None. Returns (#moves, t_ctm seconds)."""
    t0 = time.perf_counter()
    moves = 0
    for _ in range(n_iter):
        moves += ctm_iteration(sites, v2s, lX, lY, C, T, chi, args)
    return moves, time.perf_counter() - t0


# --------------------------------------------------------------------------------------
# environment initialisation (ctm/generic/env.py:367-536)
# --------------------------------------------------------------------------------------
_INIT_C = {(-1, -1): ('mijef,mijab->eafb', (3, 4)), (1, -1): ('miefj,miabj->eafb', (2, 3)),
           (1, 1): ('mefij,mabij->eafb', (1, 2)), (-1, 1): ('meijf,maijb->eafb', (1, 4))}
_INIT_T = {(0, -1): ('miefg,miabc->eafbgc', (2, 3, 4), (True, False, True)),
           (-1, 0): ('meifg,maibc->eafbgc', (1, 3, 4), (True, True, False)),
           (0, 1): ('mefig,mabic->eafbgc', (1, 2, 4), (False, True, True)),
           (1, 0): ('mefgi,mabci->eafbgc', (1, 2, 3), (True, False, True))}


def init_env(sites, v2s, chi):
    """'CTMRG' initialisation: partial traces of a (x) a*, /max|.|, zero-padded to chi."""
    C, T = {}, {}
    for coord, site in sites.items():
        for vec, (ein, legs) in _INIT_C.items():
            A = sites[v2s((coord[0] + vec[0], coord[1] + vec[1]))]
            d = A.shape
            c = torch.einsum(ein, A, A.conj()).contiguous().view(d[legs[0]] ** 2, d[legs[1]] ** 2)
            c = c / c.abs().max()
            out = torch.zeros(chi, chi, dtype=A.dtype)
            r, q = min(chi, c.shape[0]), min(chi, c.shape[1])
            out[:r, :q] = c[:r, :q]
            C[(coord, vec)] = out
        for vec, (ein, legs, is_chi) in _INIT_T.items():
            A = sites[v2s((coord[0] + vec[0], coord[1] + vec[1]))]
            d = A.shape
            t = torch.einsum(ein, A, A.conj()).contiguous().view(*[d[l] ** 2 for l in legs])
            t = t / t.abs().max()
            shape = [chi if f else t.shape[i] for i, f in enumerate(is_chi)]
            out = torch.zeros(shape, dtype=A.dtype)
            sl = tuple(slice(0, min(chi, t.shape[i])) if f else slice(None) for i, f in enumerate(is_chi))
            out[sl] = t[sl]
            T[(coord, vec)] = out
    return C, T


# --------------------------------------------------------------------------------------
# C4v one-site move (ctm/one_site_c4v/ctmrg_c4v.py:325-463, ctm_components_c4v.py:52-130)
# --------------------------------------------------------------------------------------
def c2x2_c4v(a, C, T):
    t = sl_einsum('ab,xbu,ael,@uldr->edxr', (C, T, T), a)
    return t.reshape(t.shape[0] * t.shape[1], t.shape[2] * t.shape[3])


def ctm_move_c4v(a, C, T, chi, args=None, return_decomp=False):
    """Returns the new (C, T). eps_multiplet/abs_tol are the *function defaults* of
    truncated_eig_sym (custom_eig.py:7-8), as used by ctmrg_c4v.py:50-52.  With a rank-4 `a` (the double-layer
    tensor) this is ctm_MOVE_dl (ctmrg_c4v.py:200-322)."""
    args = args or OracleArgs()
    C2X2 = c2x2_c4v(a, C, T)
    D, U = truncated_eig_sym(C2X2, chi)
    nC = torch.diag((1. + 0.j) * D) if C2X2.is_complex() else torch.diag(D)
    P = U.view(C.shape[0], T.shape[2], U.shape[1])
    nT = sl_einsum('acl,aux,@uldr,cdy->xyr', (T, P, P.conj()), a)
    nT = 0.5 * (nT + nT.conj().permute(1, 0, 2))
    nC = nC / torch.abs(nC[0, 0])
    nT = _normalize(nT, args.ctm_absorb_normalization)
    if return_decomp:
        return nC, nT.contiguous(), (C2X2, D, U)
    return nC, nT.contiguous()


def ctm_move_qr_c4v(a, C, T, chi, args=None):
    """ctm_MOVE_QR_sl (ctmrg_c4v.py:465-602): projector = thin Q of (C.T) reshaped to (chi d) x chi."""
    args = args or OracleArgs()
    C2X2 = c2x2_c4v(a, C, T)
    C1x2 = torch.tensordot(C, T, ([1], [1])).permute(0, 2, 1).reshape(-1, chi)
    P, _ = torch.linalg.qr(C1x2)
    nC = P.t() @ C2X2 @ P
    Pv = P.view(C.shape[0], T.shape[2], P.shape[1])
    nT = sl_einsum('acl,aux,@uldr,cdy->xyr', (T, Pv, Pv.conj()), a)
    nT = 0.5 * (nT + nT.conj().permute(1, 0, 2))
    nC = nC / torch.abs(nC[0, 0])
    nT = _normalize(nT, args.ctm_absorb_normalization)
    return nC, nT.contiguous()


def init_env_c4v(a, chi):
    """env_c4v.py:257-311 (_init_from_ipeps_pbc with a_bra = conj(a))."""
    d = a.shape
    dk = [d[i + 1] ** 2 for i in range(4)]
    c = torch.einsum('mijef,mijab->eafb', a, a.conj()).contiguous().view(dk[2], dk[3])
    c = c / c.abs().max()
    D, U = truncated_eig_sym(c, c.shape[0])
    c = torch.diag(D).to(a.dtype)
    C = torch.zeros(chi, chi, dtype=a.dtype)
    r = min(chi, dk[2])
    C[:r, :r] = c[:r, :r]
    t = torch.einsum('meifg,maibc->eafbgc', a, a.conj()).contiguous().view(dk[0], dk[2], dk[3])
    t = t / t.abs().max()
    t = torch.einsum('ai,abs,bj->ijs', U, t, U.conj())
    T = torch.zeros(chi, chi, dk[3], dtype=a.dtype)
    T[:r, :r, :] = t[:r, :r, :]
    return C, T


# --------------------------------------------------------------------------------------
# observables used for parity: corner spectra, 2x2 RDM, J1-J2 energy
# --------------------------------------------------------------------------------------
def corner_spectra(C):
    """env.py:204-209 get_spectra: normalised singular values of every corner."""
    out = {}
    for k, c in C.items():
        s = torch.linalg.svdvals(c)
        out[k] = s / s[0]
    return out


def _sym_pos_def(rho, sym_pos_def=False):
    """ctm/generic/rdm.py:38-68: hermitise (always), clamp negative eigenvalues (optional),
    trace-normalise."""
    shape = rho.shape
    n = int(np.prod(shape[:len(shape) // 2]))
    m = rho.reshape(n, n)
    m = 0.5 * (m + m.conj().t())
    if sym_pos_def:
        D, U = torch.linalg.eigh(m)
        if D.min() < 0:
            m = U @ torch.diag(torch.clamp(D, min=0).to(m.dtype)) @ U.conj().t()
    m = m / m.diagonal().sum().real
    return m.reshape(shape)


def rdm2x2(coord, sites, v2s, C, T, raw=False, open_sites=(0, 1, 2, 3), sym_pos_def=False):
    """ctm/generic/rdm.py:1306-1675: rho(s0,s1,s2,s3; s0',s1',s2',s3') of the plaquette
    s0=coord, s1=coord+(1,0), s2=coord+(0,1), s3=coord+(1,1); hermitised, trace-normalised.
    First index group = ket (the index of `a`), second = bra (conj(a)).  Sites not in `open_sites` are traced
    (their enlarged corner is the closed one); the remaining indices keep the order above (rdm.py:1352-1353)."""
    x, y = coord
    op = [q in open_sites for q in range(4)]
    assert any(op)
    LU = corner_at('LU', (x, y), sites, v2s, C, T, open_phys=op[0])          # [down,right,(s,S)]
    RU = corner_at('RU', (x + 1, y), sites, v2s, C, T, open_phys=op[1])      # [left,down,(s,S)]
    RD = corner_at('RD', (x + 1, y + 1), sites, v2s, C, T, open_phys=op[3])  # [up,left,(s,S)]
    LD = corner_at('LD', (x, y + 1), sites, v2s, C, T, open_phys=op[2])      # [up,right,(s,S)]
    p = ['iI' if op[0] else '', 'jJ' if op[1] else '', 'kK' if op[2] else '', 'lL' if op[3] else '']
    upper = torch.einsum(f'ab{p[0]},bc{p[1]}->ac{p[0]}{p[1]}', LU, RU)
    lower = torch.einsum(f'ab{p[2]},cb{p[3]}->ac{p[2]}{p[3]}', LD, RD)
    kets = ''.join(s[0] for s in p if s)
    bras = ''.join(s[1] for s in p if s)
    rho = torch.einsum(f'ac{p[0]}{p[1]},ac{p[2]}{p[3]}->{kets}{bras}', upper, lower)
    return rho if raw else _sym_pos_def(rho, sym_pos_def)


def spin_half_ops(dtype=torch.float64):
    sz = torch.tensor([[0.5, 0.], [0., -0.5]], dtype=dtype)
    sp = torch.tensor([[0., 1.], [0., 0.]], dtype=dtype)
    sm = torch.tensor([[0., 0.], [1., 0.]], dtype=dtype)
    return sz, sp, sm, torch.eye(2, dtype=dtype)


def j1j2_hp(j1=1.0, j2=0.0, dtype=torch.float64):
    """models/j1j2.py:120-140 get_hp with hz_stag=0, h_uni=0, delta_zz=1:
    0.5*j1*(SS_01 + SS_02 + SS_23 + SS_13) + j2*(SS_03 + SS_12), sites ordered s0 s1 / s2 s3."""
    sz, sp, sm, I = spin_half_ops(dtype)
    SS = (torch.einsum('ij,ab->iajb', sz, sz) + 0.5 * (torch.einsum('ij,ab->iajb', sp, sm)
                                                       + torch.einsum('ij,ab->iajb', sm, sp)))
    id2 = torch.einsum('ij,ab->iajb', I, I)
    h = torch.einsum('ijab,klcd->ijklabcd', SS, id2)  # acts on (s0,s1)
    nn = h + h.permute(0, 2, 1, 3, 4, 6, 5, 7) + h.permute(2, 3, 0, 1, 6, 7, 4, 5) \
        + h.permute(3, 1, 2, 0, 7, 5, 6, 4)
    nnn = h.permute(0, 3, 2, 1, 4, 7, 6, 5) + h.permute(2, 1, 0, 3, 6, 5, 4, 7)
    return 0.5 * j1 * nn + j2 * nnn


def energy_j1j2(sites, v2s, C, T, j1=1.0, j2=0.0, as_tensor=False):
    """models/j1j2.py:223-247 energy_per_site (= energy_2x2_4site / _2site).  as_tensor: the 0-dim torch tensor (the
    gradient tests differentiate through it)."""
    any_site = next(iter(sites.values()))
    hp = j1j2_hp(j1, j2, dtype=any_site.dtype)
    e = 0.
    for coord in sites.keys():
        rho = rdm2x2(coord, sites, v2s, C, T)
        e = e + torch.einsum('ijklabcd,ijklabcd', rho, hp)
    e = e / len(sites)
    if as_tensor:
        return e.real if e.is_complex() else e
    return float(e.real if e.is_complex() else e)


def c4v_to_generic_env(C, T):
    """Rotate the single C4v pair (C[0,1], T[up,down,d] = left T) into the eight generic
    environment tensors of a 1x1 cell (env_c4v.py:25-45 vs env.py:57-77)."""
    s = (0, 0)
    Cg = {(s, (-1, -1)): C, (s, (1, -1)): C, (s, (1, 1)): C, (s, (-1, 1)): C.t()}
    Tg = {(s, (-1, 0)): T, (s, (0, 1)): T.permute(2, 0, 1), (s, (1, 0)): T.permute(1, 2, 0),
          (s, (0, -1)): T.permute(1, 2, 0)}
    return Cg, Tg


def rdm1x1(coord, sites, v2s, C, T, raw=False, sym_pos_def=False):
    """ctm/generic/rdm.py:114-258 (rdm1x1_dl): rho(s;s') of the site at `coord`: open enlarged LU corner closed by
    C(1,-1) T(1,0) C(1,1) T(0,1) C(-1,1) of the same site."""
    s = v2s(coord)
    LU = corner_at('LU', coord, sites, v2s, C, T, open_phys=True)            # [(l,D),(x,R),s,S]
    E = torch.einsum('xa,aRb,bc,Dec,le->lDxR', C[(s, (1, -1))], T[(s, (1, 0))], C[(s, (1, 1))], T[(s, (0, 1))], C[(s, (-1, 1))])
    rho = torch.einsum('yzsS,yz->sS', LU, E.reshape(LU.shape[0], LU.shape[1]))
    return rho if raw else _sym_pos_def(rho, sym_pos_def)


def rdm2x1(coord, sites, v2s, C, T, raw=False, sym_pos_def=False):
    """ctm/generic/rdm.py:352-500 (rdm2x1_dl): rho(s0,s1;s0',s1'), s0 = coord, s1 = coord+(1,0)."""
    c0, c1 = v2s(coord), v2s((coord[0] + 1, coord[1]))
    LU = corner_at('LU', coord, sites, v2s, C, T, open_phys=True)                        # [(l,D),(x,R),s,S]
    RU = corner_at('RU', (coord[0] + 1, coord[1]), sites, v2s, C, T, open_phys=True)     # [(e,l),(c,f),j,J]
    B0 = torch.einsum('le,Dec->lDc', C[(c0, (-1, 1))], T[(c0, (0, 1))])
    B1 = torch.einsum('fEb,cb->cfE', T[(c1, (0, 1))], C[(c1, (1, 1))])
    L = torch.einsum('yzsS,yc->czsS', LU, B0.reshape(LU.shape[0], -1))
    R = torch.einsum('tqjJ,qE->tEjJ', RU, B1.reshape(RU.shape[1], -1))
    rho = torch.einsum('czsS,zcjJ->sjSJ', L, R)
    return rho if raw else _sym_pos_def(rho, sym_pos_def)


def rdm1x2(coord, sites, v2s, C, T, raw=False, sym_pos_def=False):
    """ctm/generic/rdm.py:672-826 (rdm1x2_dl): rho(s0,s1;s0',s1'), s0 = coord, s1 = coord+(0,1)."""
    c0, c1 = v2s(coord), v2s((coord[0], coord[1] + 1))
    LU = corner_at('LU', coord, sites, v2s, C, T, open_phys=True)                        # [(l,D),(x,R),s,S]
    LD = corner_at('LD', (coord[0], coord[1] + 1), sites, v2s, C, T, open_phys=True)     # [(c,u),(e,r),k,K]
    Rr = torch.einsum('xa,aRb->xRb', C[(c0, (1, -1))], T[(c0, (1, 0))])
    Rb = torch.einsum('Brc,ce->erB', T[(c1, (1, 0))], C[(c1, (1, 1))])
    Tp = torch.einsum('yzsS,zb->ybsS', LU, Rr.reshape(LU.shape[1], -1))
    Bt = torch.einsum('tqkK,qB->tBkK', LD, Rb.reshape(LD.shape[1], -1))
    rho = torch.einsum('ybsS,ybkK->skSK', Tp, Bt)
    return rho if raw else _sym_pos_def(rho, sym_pos_def)


def rdm_small_c4v(kind, a, C, T, sym_pos_def=False):
    """rdm1x1_sl / rdm2x1_sl of ctm/one_site_c4v/rdm_c4v.py:266-392,530-665 through the generic construction on the
    rotated environment (kind '1x1' or '2x1')."""
    from collections import OrderedDict
    Cg, Tg = c4v_to_generic_env(C, T)
    f = {'1x1': rdm1x1, '2x1': rdm2x1, '1x2': rdm1x2}[kind]
    return f((0, 0), OrderedDict({(0, 0): a}), v2s_1site, Cg, Tg, sym_pos_def=sym_pos_def)


def rdm3x1_c4v(a, C, T, sym_pos_def=False):
    """rdm3x1 / rdm3x1_sl of ctm/one_site_c4v/rdm_c4v.py:667-1011: the two END sites of a row of three (centre traced), indices
    s0 s1 ; s0' s1' (ket ; bra).  Restated column by column with plain einsums: left boundary C-T-C, three columns T-(a a*)-T
    (the first and the last with open physical indices), right boundary."""
    D = a.shape[1]
    chi = C.shape[0]
    Tk = T.reshape(chi, chi, D, D)                                    # [chi, chi, ket, bra]
    E = torch.einsum('ab,acd,ce->bde', C, T, C).reshape(chi, D, D, chi)
    col_open = 'axuU,xyYz,puydr,qUYDR,zedD->arRepq'
    col = 'axuU,xyYz...,suydr,sUYDR,zedD->arRe...'
    E = torch.einsum(col_open, Tk, E, a, a.conj(), Tk)
    E = torch.einsum(col, Tk, E, a, a.conj(), Tk)
    E = torch.einsum('axuU,xyYzij,puydr,qUYDR,zedD->arReijpq', Tk, E, a, a.conj(), Tk)
    R = torch.einsum('xc,tcy,tz->xyz', C, T, C).reshape(chi, D, D, chi)
    rho = torch.einsum('arReijpq,arRe->ipjq', E, R)
    return _sym_pos_def(rho, sym_pos_def)


def rdm2x2_c4v(a, C, T, open_sites=(0, 1, 2, 3), sym_pos_def=False):
    """The 2x2 plaquette of the one-site C4v state through the generic construction: open_sites=(0,1) is
    rdm2x2_NN_lowmem_sl, (0,3) rdm2x2_NNN_lowmem_sl, all four rdm2x2 (ctm/one_site_c4v/rdm_c4v.py:1160-1202,
    1329-1371, 1446-1546; those rotate ONE enlarged corner, which C4v symmetry makes equivalent)."""
    from collections import OrderedDict
    Cg, Tg = c4v_to_generic_env(C, T)
    return rdm2x2((0, 0), OrderedDict({(0, 0): a}), v2s_1site, Cg, Tg, open_sites=open_sites, sym_pos_def=sym_pos_def)


def energy_j1j2_c4v(a, C, T, j1=1.0, j2=0.0, as_tensor=False):
    """models/j1j2.py:641-679 energy_1x1_lowmem (j3=hz_stag=h_uni=0): bipartite rotation
    on one sublattice, e = 2 j1 <SS_rot>_NN + 2 j2 <SS>_NNN with sym_pos_def RDMs."""
    from collections import OrderedDict
    sites = OrderedDict({(0, 0): a})
    Cg, Tg = c4v_to_generic_env(C, T)
    rho = rdm2x2((0, 0), sites, v2s_1site, Cg, Tg, raw=True)
    sz, sp, sm, I = spin_half_ops(a.dtype)
    SS = (torch.einsum('ij,ab->iajb', sz, sz) + 0.5 * (torch.einsum('ij,ab->iajb', sp, sm)
                                                       + torch.einsum('ij,ab->iajb', sm, sp)))
    rot = torch.tensor([[0., 1.], [-1., 0.]], dtype=a.dtype)       # groups/su2.py:172-176
    SS_rot = torch.einsum('ki,kjcb,ca->ijab', rot, SS, rot)
    nn = _sym_pos_def(torch.einsum('ijklabkl->ijab', rho), True)
    e = 2.0 * j1 * torch.einsum('ijab,ijab', nn, SS_rot)
    if abs(j2) > 0:
        nnn = _sym_pos_def(torch.einsum('ijklajkd->ilad', rho), True)
        e = e + 2.0 * j2 * torch.einsum('ijab,ijab', nnn, SS)
    if as_tensor:
        return e.real if e.is_complex() else e
    return float(e.real if e.is_complex() else e)


# --------------------------------------------------------------------------------------
# synthetic inputs (examples/j1j2/ctmrg_j1j2.py:56-110, ctmrg_j1j2_c4v.py:61-66)
# --------------------------------------------------------------------------------------
def v2s_4site(coord):
    """4SITE tiling of examples/j1j2/ctmrg_j1j2.py:56-60."""
    return ((coord[0] + abs(coord[0]) * 2) % 2, (coord[1] + abs(coord[1]) * 2) % 2)


def v2s_2site(coord):
    """2SITE tiling (2x1 cell) of examples/j1j2/ctmrg_j1j2.py:51-55."""
    return ((coord[0] + abs(coord[0]) * 2) % 2, 0)


def v2s_1site(coord):
    return (0, 0)


def random_state_4site(D, p=2, seed=123, family='A', dtype=torch.float64):
    """Family A: rand in [0,1) (script-exact); family B: rand-0.5. Draw order (0,0),(1,0),(0,1),(1,1)."""
    from collections import OrderedDict
    torch.manual_seed(seed)
    sites = OrderedDict()
    for c in [(0, 0), (1, 0), (0, 1), (1, 1)]:
        A = torch.rand((p, D, D, D, D), dtype=dtype)
        if family == 'B':
            A = A - 0.5
        sites[c] = A / A.abs().max()
    return sites


def _symmetrise(A, steps):
    for sign, perm in steps:
        A = 0.5 * (A + sign * A.permute(perm))
    return A


def make_c4v_symm_A1(A):
    """groups/pg.py:44-56: projection on the A1 irrep of C4v acting on a[s,u,l,d,r]:
    left-right and up-down reflections, then the two quarter turns."""
    return _symmetrise(A, [(1, (0, 1, 4, 3, 2)), (1, (0, 3, 2, 1, 4)),
                           (1, (0, 4, 1, 2, 3)), (1, (0, 2, 3, 4, 1))])


def make_c4v_symm_A2(A):
    """groups/pg.py:57-69: projection on A2 (odd under reflections, even under rotations)."""
    return _symmetrise(A, [(-1, (0, 1, 4, 3, 2)), (-1, (0, 4, 3, 2, 1)),
                           (1, (0, 4, 1, 2, 3)), (1, (0, 3, 4, 1, 2))])


def random_state_c4v(D, p=2, seed=123, family='A', dtype=torch.float64):
    """Real: make_c4v_symm(rand)/max (ctmrg_j1j2_c4v.py:61-66). Complex: A1(Re)+i*A2(Im)
    (ipeps/ipeps_c4v.py:90-91), which keeps the enlarged corner Hermitian."""
    torch.manual_seed(seed)
    if dtype.is_complex:
        X = torch.rand((p, D, D, D, D), dtype=dtype)
        if family == 'B':
            X = X - (0.5 + 0.5j)
        A = make_c4v_symm_A1(X.real) + 1j * make_c4v_symm_A2(X.imag)
    else:
        X = torch.rand((p, D, D, D, D), dtype=dtype)
        if family == 'B':
            X = X - 0.5
        A = make_c4v_symm_A1(X)
    return A / A.abs().max()


def random_state_kagome(D, seed=123, family='A', dtype=torch.float64):
    """examples/kagome/ctmrg_spin_half_kagome.py:130-147 + ipeps/ipess_kagome.py:62-82:
    T_u,T_d = rand(D,D,D)-1; B_c,B_a,B_b = rand(2,D,D)-1 contracted into one on-site
    a[8,D,D,D,D], /max|a|.  (family B: rand-0.5.)"""
    torch.manual_seed(seed)
    off = 1.0 if family == 'A' else 0.5
    T_u = torch.rand(D, D, D, dtype=dtype) - off
    T_d = torch.rand(D, D, D, dtype=dtype) - off
    B_c = torch.rand(2, D, D, dtype=dtype) - off
    B_a = torch.rand(2, D, D, dtype=dtype) - off
    B_b = torch.rand(2, D, D, dtype=dtype) - off
    return kagome_onsite(T_u, B_c, T_d, B_b, B_a)


def kagome_onsite(T_u, B_c, T_d, B_b, B_a):
    """ipeps/ipess_kagome.py:62-82 build_onsite_tensors: the three physical spins of the
    up-triangle are fused into one index of dimension 8."""
    A = torch.einsum('iab,uji,jkl,vkc,wld->uvwabcd', T_u, B_c, T_d, B_b, B_a)
    A = A.reshape(B_a.shape[0] * B_b.shape[0] * B_c.shape[0], T_u.shape[1], T_u.shape[2],
                  B_b.shape[2], B_a.shape[2])
    return A / A.abs().max()
