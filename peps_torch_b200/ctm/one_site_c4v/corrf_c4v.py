"""Two-point functions of the C4v one-site ansatz, ctm/one_site_c4v/corrf_c4v.py of peps-torch: get_edge :5-41, apply_edge
:85-143, get_edge_L :43-83, apply_edge_L :145-176, apply_TM_1sO :178-271, apply_TM_1sO_2 :273-433, apply_TM_2sO :435-591,
corrf_1sO1sO :593-664, corrf_2sOH2sOH_E1 :666-737, corrf_2sOV2sOV_E2 :739-829 -- what eval_corrf_SS / eval_corrf_DD_H /
eval_corrf_DD_V of models/j1j2.py:826-925 call at the tail of ctmrg_j1j2_c4v.py.  Every transfer-matrix
application is one contraction chain through libctmb; the double-layer tensor (with or without an operator) is never formed,
the operator is applied to the ket layer and the two layers are contracted one after the other."""
import torch
from ... import ad


def _engine():
    from ...engine import default_engine
    return default_engine()


def _parts(state, env):
    return next(iter(state.sites.values())), env.C[env.keyC], env.T[env.keyT]


def _dot(x, y):
    return ad.contract(_engine(), 'ka,kb->ab', x.reshape(-1, 1), y.reshape(-1, 1)).reshape(())


def get_edge(state, env, verbosity=0):
    r"""The left boundary C--T--C as a :math:`\chi \times D^2 \times \chi` tensor (corrf_c4v.py:5-41)."""
    eng = _engine()
    _, C, T = _parts(state, env)
    return ad.contract(eng, 'bcd,ce->bde', ad.contract(eng, 'ab,acd->bcd', C, T), C)


def apply_edge(state, env, vec, verbosity=0):
    r"""``vec`` closed with the right boundary C--T--C, a scalar (corrf_c4v.py:85-143)."""
    eng = _engine()
    _, C, T = _parts(state, env)
    E = ad.contract(eng, 'xty,tz->xyz', ad.contract(eng, 'xc,tcy->xty', C, T), C)
    return _dot(vec, E)


_TM = 'axu,xyz,@uydr,zed->are'       # T (upper), edge, on-site, T (lower)                        (corrf_c4v.py:206-271)


def apply_TM_1sO(state, env, edge, op=None, verbosity=0):
    r"""One transfer matrix, with ``op`` (rank 2) or the identity on the physical leg, applied to ``edge`` (corrf_c4v.py:178-271)."""
    if op is not None and op.dim() != 2:
        raise ValueError("apply_TM_1sO: op must be a one-site operator")
    eng = _engine()
    a, _, T = _parts(state, env)
    a_ket = None if op is None else ad.contract(eng, 'mefgh,mn->nefgh', a, op.to(dtype=a.dtype, device=a.device))
    return ad.sl_chain(eng, _TM, (T, edge, T), a, a_ket=a_ket).contiguous()


def _split_two_site(op):
    """op[s0,s1;s0',s1'] = sum_k op_l[s0,s0',k] op_r[s1,s1',k] by an SVD of the (s0 s0') x (s1 s1') matrix, exactly as
    corrf_c4v.py:463-472 does it (including its plain transpose of V, so a complex ``op`` is split the way the reference
    splits it)."""
    p = op.size(0)
    U, S, V = torch.svd(op.permute(0, 2, 1, 3).contiguous().reshape(p * p, p * p))
    op_l = U.reshape(p, p, S.size(0))
    op_r = (S[:, None] * V.t()).reshape(S.size(0), p, p).permute(1, 2, 0).contiguous()
    return op_l, op_r


def apply_TM_2sO(state, env, edge, op=None, verbosity=0):
    r"""Two transfer matrices carrying the two-site operator ``op`` (rank 4; None = identity) applied to ``edge``
    (corrf_c4v.py:435-591): the operator is split over the two sites and its bond travels with the edge in between."""
    if op is None:
        return apply_TM_1sO(state, env, apply_TM_1sO(state, env, edge))
    if op.dim() != 4:
        raise ValueError("apply_TM_2sO: op must be a two-site operator")
    eng = _engine()
    a, _, T = _parts(state, env)
    op_l, op_r = _split_two_site(op.to(dtype=a.dtype, device=a.device))
    a_l = ad.contract(eng, 'mefgh,mnk->nefghk', a, op_l)
    a_r = ad.contract(eng, 'mefgh,mnk->nefghk', a, op_r)
    E = ad.sl_chain(eng, 'axu,xyz,@uydr,zed->arke', (T, edge, T), a, a_ket=a_l, ket_extra='k')
    return ad.sl_chain(eng, 'axu,xykz,@uydr,zed->are', (T, E.contiguous(), T), a, a_ket=a_r, ket_extra='k').contiguous()


def corrf_1sO1sO(state, env, op1, get_op2, dist, rl_0=None, verbosity=0):
    r""":math:`\langle O_1(0)\, O_2(r) \rangle`, r = 1 .. dist+1 (corrf_c4v.py:593-664); ``rl_0`` = (left, right) edges that
    replace the environment's."""
    def close(E):
        return apply_edge(state, env, E) if rl_0 is None else _dot(E, rl_0[1])

    E0 = get_edge(state, env) if rl_0 is None else rl_0[0]
    E1 = apply_TM_1sO(state, env, E0, op=op1)
    E0 = apply_TM_1sO(state, env, E0)
    out = torch.empty(dist + 1, dtype=state.dtype, device=state.device)
    for r in range(dist + 1):
        E12 = apply_TM_1sO(state, env, E1, op=get_op2(r))
        E0 = apply_TM_1sO(state, env, E0)
        E1 = apply_TM_1sO(state, env, E1)
        out[r] = close(E12) / close(E0)
        scale = E0.abs().max()
        E0, E1 = E0 / scale, E1 / scale
    return out


def corrf_2sOH2sOH_E1(state, env, op1, get_op2, dist, verbosity=0):
    r""":math:`\langle O_1(0)\, O_2(r) \rangle` of two horizontal two-site operators, r = 2 .. dist+2 (corrf_c4v.py:666-737)."""
    E0 = get_edge(state, env)
    E1 = apply_TM_2sO(state, env, E0, op=op1)
    E0 = apply_TM_2sO(state, env, E0)
    out = torch.empty(dist + 1, dtype=state.dtype, device=state.device)
    for r in range(dist + 1):
        E12 = apply_TM_2sO(state, env, E1, op=get_op2(r))
        E0 = apply_TM_1sO(state, env, E0)
        E1 = apply_TM_1sO(state, env, E1)
        out[r] = apply_edge(state, env, E12) / apply_edge(state, env, apply_TM_1sO(state, env, E0))
        scale = E0.abs().max()
        E0, E1 = E0 / scale, E1 / scale
    return out


def get_edge_L(state, env, l=1, verbosity=0):
    r"""The boundary C--T--...--T--C with ``l`` T tensors, indices from the first C to the last (corrf_c4v.py:43-83)."""
    eng = _engine()
    _, C, T = _parts(state, env)
    E, idx = C, 'ab'
    for i in range(l):
        # E[..., b] T[d, b, c] -> E[..., c, d]
        nxt = idx[:-1] + 'yz'
        E = ad.contract(eng, f'{idx[:-1]}x,zxy->{nxt}', E, T)
        idx = ''.join(chr(ord('a') + j) for j in range(len(nxt)))
    return ad.contract(eng, f'{idx[:-1]}x,xz->{idx[:-1]}z', E, C)


def apply_edge_L(state, env, vec, verbosity=0):
    r"""``vec`` closed with the opposite boundary of its own width, a scalar (corrf_c4v.py:145-176): the indices of the edge
    run the other way round."""
    E = get_edge_L(state, env, l=vec.dim() - 2)
    return _dot(vec, E.permute(*reversed(range(E.dim()))).contiguous())


def apply_TM_1sO_2(state, env, edge, op=None, verbosity=0):
    r"""One width-2 transfer matrix (two sites on top of each other) applied to ``edge`` (:math:`\chi \times D^2 \times D^2
    \times \chi`), with the two-site operator ``op`` (rank 4) split over the two sites or the identity (corrf_c4v.py:273-433)."""
    eng = _engine()
    a, _, T = _parts(state, env)
    if op is None:
        E = ad.sl_chain(eng, 'axu,xyvz,@uydr->avzdr', (T, edge), a)
        return ad.sl_chain(eng, 'avzdr,@dvfq,zef->arqe', (E.contiguous(), T), a).contiguous()
    if op.dim() != 4:
        raise ValueError(f"Invalid op: rank {op.size()}")
    op_1, op_2 = _split_two_site(op.to(dtype=a.dtype, device=a.device))
    a_1 = ad.contract(eng, 'mefgh,mnk->nefghk', a, op_1)
    a_2 = ad.contract(eng, 'mefgh,mnk->nefghk', a, op_2)
    E = ad.sl_chain(eng, 'axu,xyvz,@uydr->avzdrk', (T, edge), a, a_ket=a_1, ket_extra='k')
    return ad.sl_chain(eng, 'avzdrk,@dvfq,zef->arqe', (E.contiguous(), T), a, a_ket=a_2, ket_extra='k').contiguous()


def corrf_2sOV2sOV_E2(state, env, op1, get_op2, dist, verbosity=0):
    r""":math:`\langle O_1(0)\, O_2(r) \rangle` of two vertical two-site operators, r = 1 .. dist+1, through the width-2
    transfer matrix (corrf_c4v.py:739-829)."""
    E0 = get_edge_L(state, env, l=2)
    E1 = apply_TM_1sO_2(state, env, E0, op=op1)
    E0 = apply_TM_1sO_2(state, env, E0)
    out = torch.empty(dist + 1, dtype=state.dtype, device=state.device)
    for r in range(dist + 1):
        E12 = apply_TM_1sO_2(state, env, E1, op=get_op2(r))
        E0 = apply_TM_1sO_2(state, env, E0)
        E1 = apply_TM_1sO_2(state, env, E1)
        out[r] = apply_edge_L(state, env, E12) / apply_edge_L(state, env, E0)
        scale = E0.abs().max()
        E0, E1 = E0 / scale, E1 / scale
    return out
