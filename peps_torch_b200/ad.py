"""Reverse-mode AD through the CTM move: what makes optim_j1j2*.py of the reference drop in.

The fused C calls (ctmb_move_generic, ctmb_move_c4v) are forward-only.  When autograd is recording and an input of the move
requires grad, the drop-in modules (ctm/generic/ctmrg.py, ctm/one_site_c4v/ctmrg_c4v.py) build the SAME move from
differentiable pieces, each of which is one libctmb call wrapped in a torch.autograd.Function:

  * `contract`      -- one pairwise tensor contraction (ctmb_einsum2); its adjoint is two more contractions of the same kind;
  * `svd_full`      -- complete SVD of the (small) projector matrix M by libctmb's Householder QR + one-sided Jacobi path
                       (ctmb_truncated_svd with chi = n), adjoint = the reference's regularised formula
                       (linalg/svd_gesdd.py:209-328, SVDGESDD.backward: F = 1/(S_i - S_j), G = 1/(S_i + S_j) through
                       safe_inverse with epsilon = S_0 * ad_decomp_reg);
  * `eig_sym_full`  -- complete Hermitian eigendecomposition (ctmb_truncated_eig_sym with chi = n), adjoint of
                       linalg/eig_sym.py:56-78 (F = safe_inverse(D_j - D_i, ad_decomp_reg)).

The reference differentiates a FULL decomposition and truncates by slicing (custom_svd.py:38-101, custom_eig.py:7-67), so
its gradient couples the kept triplets to the discarded ones through F; reproducing `loss.backward()` therefore needs the
complete spectrum, which is what the two functions above compute.  Truncation (multiplet rule included), S^-1/2, masks,
views, permutes and normalisation are element-wise torch operations on the device, recorded by autograd as in the
reference.  Activation checkpointing (CTMARGS.fwd_checkpoint_move, config.py:402-407) works as in the reference: the
functions are pure.

The engine is duck-typed (einsum2 / truncated_svd / truncated_eig_sym): the CPU tests validate every adjoint against
gradients written by the unmodified reference (tests/golden/grad_*.npz, oracle/gen_golden_grad.py) with the oracle as the
engine; the GPU tests run the same checks with libctmb.
"""
import torch
from torch.utils.checkpoint import checkpoint

UP, LEFT, DOWN, RIGHT = (0, -1), (-1, 0), (0, 1), (1, 0)


def needs_grad(tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# ----------------------------------------------------------------------------------------------
# contraction
# ----------------------------------------------------------------------------------------------
class _Contract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, spec, conjA, conjB, A, B):
        ctx.eng, ctx.spec, ctx.conjA, ctx.conjB = eng, spec, conjA, conjB
        ctx.save_for_backward(A, B)
        return eng.einsum2(spec, A.detach(), B.detach(), conjA, conjB)

    @staticmethod
    def backward(ctx, g):
        A, B = ctx.saved_tensors
        eng = ctx.eng
        lhs, lc = ctx.spec.split('->')
        la, lb = lhs.split(',')
        g = g.contiguous()
        gA = gB = None
        # C = op(A) . op(B): with torch's convention (gradients are conjugate Wirtinger derivatives)
        # g_op(A) = g . conj(op(B)) over the labels of B that are not in A;  g_A = conj(g_op(A)) if op = conj
        if ctx.needs_input_grad[4]:
            assert all(ch in lc or ch in lb for ch in la), ctx.spec
            gA = eng.einsum2(f'{lc},{lb}->{la}', g, B.detach(), False, not ctx.conjB)
            if ctx.conjA and gA.is_complex():
                gA = gA.conj().resolve_conj()
        if ctx.needs_input_grad[5]:
            assert all(ch in lc or ch in la for ch in lb), ctx.spec
            gB = eng.einsum2(f'{lc},{la}->{lb}', g, A.detach(), False, not ctx.conjA)
            if ctx.conjB and gB.is_complex():
                gB = gB.conj().resolve_conj()
        return None, None, None, None, gA, gB


def contract(eng, spec, A, B, conjA=False, conjB=False):
    """einsum of two operands through the engine; differentiable."""
    return _Contract.apply(eng, spec, bool(conjA), bool(conjB), A.contiguous(), B.contiguous())


def _mm(eng, A, B, hA=False, hB=False):
    """op(A) @ op(B) with op = conjugate transpose where flagged (plain, non-recorded engine call)."""
    sa = 'ki' if hA else 'ik'
    sb = 'jk' if hB else 'kj'
    return eng.einsum2(f'{sa},{sb}->ij', A.contiguous(), B.contiguous(), hA, hB)


def _chain(eng, subs, ops, conj, out):
    """Left-to-right pairwise evaluation (what the chains of tensordots of ctm_components.py do): a label survives a step
    iff a later operand or the output carries it."""
    cur_idx, cur = subs[0], ops[0]
    assert not conj[0]
    for i in range(1, len(ops)):
        later = set(''.join(subs[i + 1:]) + out)
        new_idx = ''.join(ch for ch in dict.fromkeys(cur_idx + subs[i]) if ch in later)
        cur = contract(eng, f'{cur_idx},{subs[i]}->{new_idx}', cur, ops[i], False, conj[i])
        cur_idx = new_idx
    if cur_idx != out:
        cur = cur.permute([cur_idx.index(ch) for ch in out])
    return cur


def sl_chain(eng, spec, ops, a, conj_last=False, a_ket=None, ket_extra='', phys='ss'):
    """The double-layer einsum `spec` ('@' = the on-site double-layer tensor) evaluated layer by layer without forming
    a (x) a*: every fused label of '@' is split into (ket, bra) on the operands that carry it (SURVEY Appendix A;
    ctm_components.py:372-434).  A rank-4 `a` is the double-layer tensor itself and enters as one operand.  `a_ket`
    replaces the ket layer (an operator applied to the physical leg, corrf.py:415-419); the bra layer stays conj(a).
    `ket_extra` labels trailing indices of `a_ket` beyond [s,u,l,d,r] (the bond of a two-site operator split over two
    sites, corrf_c4v.py:501-503).  `phys` = the labels of the physical index on the ket and on the bra layer: equal (default)
    contracts it, two different labels that also appear in the output keep both open (density matrices)."""
    lhs, out = spec.split('->')
    terms = lhs.split(',')
    a_pos = [i for i, t in enumerate(terms) if t.startswith('@')][0]
    a_idx = terms[a_pos][1:]
    assert phys != 'ss' or 's' not in lhs.replace(',', '') + out, "sl_chain: 's' is the physical index"
    nops = len(ops)
    if a.dim() == 4:
        xs = list(ops[:a_pos]) + [a] + list(ops[a_pos:])
        subs = [t[1:] if t.startswith('@') else t for t in terms]
        cj = [False] * len(xs)
        if conj_last:
            cj[-1] = True
        return _chain(eng, subs, xs, cj, out)
    D = {ch: a.shape[1 + i] for i, ch in enumerate(a_idx)}
    subs, xs, cj = [], [], []
    k = 0
    for i, t in enumerate(terms):
        if i == a_pos:
            subs += [phys[0] + a_idx + ket_extra, phys[1] + a_idx.upper()]
            xs += [a if a_ket is None else a_ket, a]
            cj += [False, True]
            continue
        x = ops[k]; k += 1
        shape, idx = [], ''
        for ch, n in zip(t, x.shape):
            if ch in D:
                shape += [D[ch], D[ch]]; idx += ch + ch.upper()
            else:
                shape += [n]; idx += ch
        subs.append(idx); xs.append(x.reshape(shape)); cj.append(conj_last and k == nops)
    out_x = ''.join(ch + ch.upper() if ch in D else ch for ch in out)
    res = _chain(eng, subs, xs, cj, out_x)
    shape, i = [], 0
    for ch in out:
        if ch in D:
            shape.append(D[ch] ** 2); i += 2
        else:
            shape.append(res.shape[i]); i += 1
    return res.reshape(shape)


# ----------------------------------------------------------------------------------------------
# decompositions
# ----------------------------------------------------------------------------------------------
def _safe_inverse(x, epsilon):
    return x / (x ** 2 + epsilon)


def _safe_inverse_2(x, epsilon):
    x = x.clone()
    x[x.abs() < epsilon] = float('inf')
    return x.pow(-1)


class _SvdFull(torch.autograd.Function):
    """M = U S V^H, complete (k = min(m, n) triplets), U's columns phase-fixed as linalg/svd_gesdd.py:18-26.  The move only
    decomposes square matrices (M = R^T Rt); the m > k / n > k terms of the reference's backward are carried along for
    completeness and are not exercised by the tests."""

    @staticmethod
    def forward(ctx, eng, M, reg):
        k = min(M.shape)
        U, S, V = eng.truncated_svd(M.detach().contiguous(), k, eps_multiplet=0.0)
        U, S, V = U.contiguous(), S.contiguous(), V.contiguous()
        ctx.eng, ctx.reg = eng, float(reg)
        ctx.save_for_backward(U, S, V)
        return U, S, V

    @staticmethod
    def backward(ctx, gu, gsigma, gv):
        # linalg/svd_gesdd.py:209-328 (SVDGESDD.backward), the matrix products through the engine
        eng = ctx.eng
        u, sigma, v = ctx.saved_tensors
        m, n, k = u.shape[0], v.shape[0], sigma.shape[0]
        eps = sigma[0] * ctx.reg
        if gsigma is not None:
            sigma_term = _mm(eng, u * gsigma.to(u.dtype).unsqueeze(-2), v, False, True)
        else:
            sigma_term = torch.zeros(m, n, dtype=u.dtype, device=u.device)
        if gu is None and gv is None:
            return None, sigma_term, None
        sigma_inv = _safe_inverse_2(sigma, eps)
        F = sigma.unsqueeze(-2) - sigma.unsqueeze(-1)
        F = _safe_inverse(F, eps)
        F.diagonal(0, -2, -1).fill_(0)
        G = sigma.unsqueeze(-2) + sigma.unsqueeze(-1)
        G = _safe_inverse(G, eps)
        G.diagonal(0, -2, -1).fill_(0)
        dA = sigma_term
        if gu is not None:
            uhgu = _mm(eng, u, gu, True, False)
            inner = ((F + G) * (uhgu - uhgu.conj().transpose(-2, -1))) * 0.5
            u_term = _mm(eng, u, inner.to(u.dtype))
            if m > k:
                w = gu * sigma_inv.unsqueeze(-2)
                u_term = u_term + w - _mm(eng, u, _mm(eng, u, w, True, False))
            dA = dA + _mm(eng, u_term, v, False, True)
        if gv is not None:
            vhgv = _mm(eng, v, gv, True, False)
            inner = ((F - G) * (vhgv - vhgv.conj().transpose(-2, -1))) * 0.5
            v_term = _mm(eng, inner.to(v.dtype), v, False, True)
            if n > k:
                gvh = gv.conj().transpose(-2, -1)
                w = sigma_inv.unsqueeze(-1) * gvh
                v_term = v_term + w - _mm(eng, _mm(eng, w, v), v, False, True)
            dA = dA + _mm(eng, u, v_term)
        if u.is_complex() and gu is not None:
            L = _mm(eng, u, gu, True, False).diagonal(0, -2, -1).clone()
            L = 1j * L.imag * sigma_inv
            dA = dA + _mm(eng, u * L.unsqueeze(-2), v, False, True)
        return None, dA, None


class _EigSymFull(torch.autograd.Function):
    """A = U D U^H, complete, eigenpairs ordered by |D| descending (linalg/eig_sym.py:14-34)."""

    @staticmethod
    def forward(ctx, eng, A, reg):
        n = A.shape[0]
        D, U = eng.truncated_eig_sym(A.detach().contiguous(), n, eps_multiplet=0.0)
        D, U = D.contiguous(), U.contiguous()
        ctx.eng, ctx.reg = eng, float(reg)
        ctx.save_for_backward(D, U)
        return D, U

    @staticmethod
    def backward(ctx, dD, dU):
        # linalg/eig_sym.py:56-78
        eng = ctx.eng
        D, U = ctx.saved_tensors
        inner = torch.zeros(U.shape[1], U.shape[1], dtype=U.dtype, device=U.device)
        if dD is not None:
            inner = inner + torch.diag(dD).to(U.dtype)
        if dU is not None:
            F = D - D[:, None]
            F = _safe_inverse(F, ctx.reg)
            F.diagonal().fill_(0)
            inner = inner + F * _mm(eng, U, dU, True, False)
        dA = _mm(eng, _mm(eng, U, inner), U, False, True)
        return None, dA, None


def _multiplet_chi(s_abs, chi, eps_multiplet, abs_tol):
    """custom_svd.py:70-88 / custom_eig.py:39-57 on detached |spectrum| values (at least chi + 1 of them)."""
    gaps = s_abs[:chi + 1].clone().detach()
    gaps[gaps < abs_tol] = 0.
    gaps = (gaps[:chi] - s_abs[1:chi + 1].detach()) / (gaps[:chi] + 1.0e-16)
    gaps[gaps > 1.0] = 0.
    chi_new = chi
    g = gaps.cpu()
    if g[chi - 1] < eps_multiplet:
        for i in range(chi - 1, -1, -1):
            if g[i] > eps_multiplet:
                chi_new = i
                break
    return chi_new


def truncated_svd(eng, M, chi, abs_tol=1.0e-14, ad_decomp_reg=1.0e-12, eps_multiplet=1.0e-12):
    """truncated_svd_gesdd(M, chi, keep_multiplets=True, ...) (custom_svd.py:38-101)."""
    U, S, V = _SvdFull.apply(eng, M, ad_decomp_reg)
    if chi < S.shape[0]:
        chi_new = _multiplet_chi(S, chi, eps_multiplet, abs_tol)
        St = S[:chi].clone(); St[chi_new + 1:] = 0.
        Ut = U[:, :chi].clone(); Ut[:, chi_new + 1:] = 0.
        Vt = V[:, :chi].clone(); Vt[:, chi_new + 1:] = 0.
        return Ut, St, Vt
    k = min(chi, S.shape[0])
    return U[:, :k], S[:k], V[:, :k]


def truncated_eig_sym(eng, M, chi, abs_tol=1.0e-14, ad_decomp_reg=1.0e-12, eps_multiplet=1.0e-12):
    """truncated_eig_sym(M, chi, keep_multiplets=True, ...) (custom_eig.py:7-67)."""
    D, U = _EigSymFull.apply(eng, M, ad_decomp_reg)
    if chi < D.shape[0]:
        chi_new = _multiplet_chi(D.abs(), chi, eps_multiplet, abs_tol)
        Dt = D[:chi].clone(); Dt[chi_new + 1:] = 0.
        Ut = U[:, :chi].clone(); Ut[:, chi_new + 1:] = 0.
        return Dt, Ut
    k = min(chi, D.shape[0])
    return D[:k], U[:, :k]


# ----------------------------------------------------------------------------------------------
# generic move (ctm/generic/ctmrg.py:179-319)
# ----------------------------------------------------------------------------------------------
# kind -> (C key, T1 key, T2 key, einsum); rows x cols = (chi_T2 d)(chi_T1 d)   
CORNERS = {
    'LU': ((-1, -1), (0, -1), (-1, 0), 'ab,buc,ael,@ulfg->efcg'),
    'RU': ((1, -1), (1, 0), (0, -1), 'ab,brc,eua,@ulfr->elcf'),
    'RD': ((1, 1), (0, 1), (1, 0), 'ab,feb,cra,@ulfr->cuel'),
    'LD': ((-1, 1), (-1, 0), (0, 1), 'ab,cal,fbe,@ulfr->cuer'),
}
# direction -> R = X1 X2, Rt = X3 X4 with X = enlarged corner (kind, dx, dy, transposed)   (ctm_components.py:55-265)
HALVES = {
    UP: ((('RU', 0, 0, False), ('RD', 0, 1, False)), (('LU', -1, 0, True), ('LD', -1, 1, False))),
    LEFT: ((('LU', 0, 0, False), ('RU', 1, 0, False)), (('LD', 0, 1, False), ('RD', 1, 1, True))),
    DOWN: ((('LD', 0, 0, True), ('LU', 0, -1, False)), (('RD', 1, 0, True), ('RU', 1, -1, True))),
    RIGHT: ((('RD', 0, 0, False), ('LD', -1, 0, True)), (('RU', 0, -1, True), ('LU', -1, -1, True))),
}
# absorption (ctmrg.py:324-804)
ABSORB = {
    UP: dict(C1=(1, -1), T1=(1, 0), T=(0, -1), T2=(-1, 0), C2=(-1, -1), shift=(1, 0),
             nC1='ab,brc,arx->xc', nC2='ab,ael,blx->ex', nT=('auc,alx,@uldr,cry->xdy', ('T', 'Pt2', 'P1')),
             out=((1, -1), (-1, -1), (0, -1))),
    LEFT: dict(C1=(-1, -1), T1=(0, -1), T=(-1, 0), T2=(0, 1), C2=(-1, 1), shift=(0, -1),
               nC1='ab,buc,aux->xc', nC2='ab,fbe,afx->xe', nT=('acl,aux,@uldr,cdy->xyr', ('T', 'P1', 'Pt2')),
               out=((-1, -1), (-1, 1), (-1, 0))),
    DOWN: dict(C1=(-1, 1), T1=(-1, 0), T=(0, 1), T2=(1, 0), C2=(1, 1), shift=(-1, 0),
               nC1='ab,cal,blx->cx', nC2='ab,cra,brx->cx', nT=('fbe,blx,@ulfr,ery->uxy', ('T', 'P1', 'Pt2')),
               out=((-1, 1), (1, 1), (0, 1))),
    RIGHT: dict(C1=(1, 1), T1=(0, 1), T=(1, 0), T2=(0, -1), C2=(1, -1), shift=(0, 1),
                nC1='ab,feb,afx->xe', nC2='ab,eua,bux->ex', nT=('arc,aux,@uldr,cdy->xly', ('T', 'Pt2', 'P1')),
                out=((1, 1), (1, -1), (1, 0))),
}


def _corner(eng, kind, coord, sites, v2s, C, T):
    s = v2s(coord)
    kc, k1, k2, spec = CORNERS[kind]
    t = sl_chain(eng, spec, (C[(s, kc)], T[(s, k1)], T[(s, k2)]), sites[s])
    return t.reshape(t.shape[0] * t.shape[1], t.shape[2] * t.shape[3])


def _halves(eng, direction, coord, sites, v2s, C, T, method):
    res = []
    for pair in HALVES[direction]:
        ms = []
        for kind, dx, dy, tr in (pair if method == '4X4' else pair[:1]):
            m = _corner(eng, kind, (coord[0] + dx, coord[1] + dy), sites, v2s, C, T)
            ms.append(m.t() if tr else m)
        res.append(ms[0] if len(ms) == 1 else contract(eng, 'ik,kj->ij', ms[0], ms[1]))
    return res[0], res[1]


def _projectors(eng, R, Rt, chi, ctm_args):
    """ctm_get_projectors_from_matrices (ctm_projectors.py:142-293), projector_full_matrices = True."""
    M = contract(eng, 'ki,kj->ij', R, Rt)
    U, S, V = truncated_svd(eng, M, chi, abs_tol=ctm_args.projector_multiplet_abstol,
                            ad_decomp_reg=getattr(ctm_args, 'ad_decomp_reg', 1.0e-12),
                            eps_multiplet=ctm_args.projector_eps_multiplet)
    nz_mask = S / S[0] > ctm_args.projector_svd_reltol
    S_nz = S[nz_mask]
    S_sqrt = S * 0
    S_sqrt[:S_nz.size(0)] = torch.rsqrt(S_nz)
    P = contract(eng, 'ik,kj->ij', R, U, False, True) * S_sqrt[None, :]
    Pt = contract(eng, 'ik,kj->ij', Rt, V) * S_sqrt[None, :]
    return P, Pt


def _normalize(t, norm_type):
    with torch.no_grad():
        scale = torch.linalg.vector_norm(t, ord=float('inf') if norm_type == 'inf' else 2)
    return t / scale


def _absorb(eng, direction, coord, sites, v2s, C, T, P, Pt, norm_type):
    sp = ABSORB[direction]
    s = v2s(coord)
    sn = v2s((coord[0] + sp['shift'][0], coord[1] + sp['shift'][1]))
    C1, T1, Tm, T2, C2 = C[(s, sp['C1'])], T[(s, sp['T1'])], T[(s, sp['T'])], T[(s, sp['T2'])], C[(s, sp['C2'])]
    chi_env = C1.shape[0]

    def pview(p):
        return p.reshape(chi_env, p.shape[0] // chi_env, p.shape[1])
    pool = dict(T=Tm, P1=pview(P[sn]), Pt1=pview(Pt[sn]), P2=pview(P[s]), Pt2=pview(Pt[s]))
    l1, l2, l3 = sp['nC1'].split('->')[0].split(',')
    nC1 = _chain(eng, [l1, l2, l3], [C1, T1, pool['Pt1']], [False] * 3, sp['nC1'].split('->')[1])
    l1, l2, l3 = sp['nC2'].split('->')[0].split(',')
    nC2 = _chain(eng, [l1, l2, l3], [C2, T2, pool['P2']], [False] * 3, sp['nC2'].split('->')[1])
    spec, order = sp['nT']
    nT = sl_chain(eng, spec, tuple(pool[k] for k in order), sites[s])
    return _normalize(nC1, norm_type), _normalize(nC2, norm_type), _normalize(nT, norm_type)


def ctm_move_generic(eng, direction, state, env, ctm_args):
    """Differentiable ctm_MOVE: replaces the entries of env.C / env.T at coord - direction (ctmrg.py:302-319)."""
    direction = tuple(direction)
    if direction not in ABSORB:
        raise ValueError("Invalid direction: " + str(direction))
    method = getattr(ctm_args, 'projector_method', '4X4')
    if method not in ('4X4', '4X2'):
        raise ValueError("Invalid Projector method: " + str(method))
    sites, v2s = state.sites, state.vertexToSite
    coords = list(sites.keys())
    ckeys, tkeys = list(env.C.keys()), list(env.T.keys())
    norm_type = getattr(ctm_args, 'ctm_absorb_normalization', 'inf')

    def core(*tensors):
        st = dict(zip(coords, tensors[:len(coords)]))
        C = dict(zip(ckeys, tensors[len(coords):len(coords) + len(ckeys)]))
        T = dict(zip(tkeys, tensors[len(coords) + len(ckeys):]))
        P, Pt = {}, {}
        for c in coords:
            R, Rt = _halves(eng, direction, c, st, v2s, C, T, method)
            P[c], Pt[c] = _projectors(eng, R, Rt, env.chi, ctm_args)
        out = []
        for c in coords:
            out += list(_absorb(eng, direction, c, st, v2s, C, T, P, Pt, norm_type))
        return tuple(out)

    tensors = tuple(sites[c] for c in coords) + tuple(env.C[k] for k in ckeys) + tuple(env.T[k] for k in tkeys)
    if getattr(ctm_args, 'fwd_checkpoint_move', False):
        new = checkpoint(core, *tensors, use_reentrant=True)
    else:
        new = core(*tensors)
    kC1, kC2, kT = ABSORB[direction]['out']
    for i, c in enumerate(coords):
        nc = v2s((c[0] - direction[0], c[1] - direction[1]))
        env.C[(nc, kC1)], env.C[(nc, kC2)], env.T[(nc, kT)] = new[3 * i], new[3 * i + 1], new[3 * i + 2]


# ----------------------------------------------------------------------------------------------
# C4v move (ctm/one_site_c4v/ctmrg_c4v.py:200-463)
# ----------------------------------------------------------------------------------------------
def ctm_move_c4v(eng, a, C, T, chi, ctm_args):
    """Differentiable ctm_MOVE_sl (rank-5 a) / ctm_MOVE_dl (rank-4 double-layer a) -> (C', T')."""
    norm_type = getattr(ctm_args, 'ctm_absorb_normalization', 'inf')
    reg = getattr(ctm_args, 'ad_decomp_reg', 1.0e-12)

    def core(a, C, T):
        t = sl_chain(eng, 'ab,xbu,ael,@uldr->edxr', (C, T, T), a)
        C2X2 = t.reshape(t.shape[0] * t.shape[1], t.shape[2] * t.shape[3])
        # ctmrg_c4v.py:49-52: truncated_eig_sym(M, chi, keep_multiplets=True, ad_decomp_reg=...), function defaults otherwise
        D, U = truncated_eig_sym(eng, C2X2, chi, ad_decomp_reg=reg)
        nC = torch.diag((1. + 0.j) * D) if C2X2.is_complex() else torch.diag(D)
        P = U.reshape(C.shape[0], T.shape[2], U.shape[1])
        nT = sl_chain(eng, 'acl,aux,@uldr,cdy->xyr', (T, P, P), a, conj_last=True)
        nT = 0.5 * (nT + nT.conj().permute(1, 0, 2))
        with torch.no_grad():
            scale_nC = torch.abs(nC[0, 0])
        nC = nC / scale_nC
        return nC, _normalize(nT, norm_type).contiguous()

    if getattr(ctm_args, 'fwd_checkpoint_move', False):
        return checkpoint(core, a, C, T, use_reentrant=True)
    return core(a, C, T)
