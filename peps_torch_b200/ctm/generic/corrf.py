"""Transfer-matrix mat-vecs and two-point functions of ctm/generic/corrf.py of peps-torch (get_edge :10-104, apply_edge
:234-277, apply_TM_0sO :278-362, apply_TM_1sO :364-650 with identity, one-site operators and MPO pieces,
get_edge_2 :105-232, apply_TM_2sO_2sChannel :671-912, apply_TM_2sO_1sChannel :914-978, corrf_1sO1sO :980-1067,
corrf_2sOH2sOH_E1 :1069-1156, corrf_2sOV2sOV_E2 :1158-1248).  One
application of the width-0 / width-1 transfer matrix of a site to an edge tensor is one contraction chain through libctmb
(the double-layer tensor a (x) a* is never formed: the chain contracts the two layers one after the other)."""
import torch
from ... import ad


def _engine():
    from ...engine import default_engine
    return default_engine()


# direction -> (T1 key, T2 key, einsum with '@' = the on-site double-layer tensor A[u,l,d,r])       (corrf.py:427-648)
_TM1 = {
    (0, -1): ((-1, 0), (1, 0), 'abl,bdc,@uldr,erc->aue'),
    (-1, 0): ((0, -1), (0, 1), 'aub,brc,@uldr,dec->ale'),
    (0, 1): ((-1, 0), (1, 0), 'bal,buc,@uldr,cre->ade'),
    (1, 0): ((0, -1), (0, 1), 'bua,blc,@uldr,dce->are'),
}


def apply_TM_1sO(coord, direction, state, env, edge, op=None, verbosity=0):
    r"""``edge`` (:math:`\chi \times D^2 \times \chi`, optionally with a trailing MPO index) with one transfer matrix of site
    ``coord`` applied (corrf.py:364-650).  ``op``: None (identity on the physical leg), a one-site operator (rank 2, applied to
    the ket layer as corrf.py:415-419), or a piece of an MPO (corrf.py:423-450): rank 3 ``[s, s', l]`` -- ``l`` closes the
    MPO index of the edge if it has one, else becomes the trailing index of the result -- or rank 4 ``[s, s', l, k]`` (``l``
    contracted with the edge's MPO index, ``k`` the trailing index of the result)."""
    if direction not in _TM1:
        raise ValueError("Invalid direction: " + str(direction))
    mpo_in = edge.dim() == 4
    rank = 0 if op is None else op.dim()
    if edge.dim() not in (3, 4) or rank not in ((3, 4) if mpo_in else (0, 2, 3)):
        raise ValueError(f"apply_TM_1sO: edge of rank {edge.dim()} with an operator of rank {rank}")
    eng = _engine()
    c = state.vertexToSite(coord)
    k1, k2, spec = _TM1[direction]
    a = state.site(c)
    if op is None:
        return ad.sl_chain(eng, spec, (env.T[(c, k1)], edge, env.T[(c, k2)]), a).contiguous()
    if a.dim() != 5:
        raise NotImplementedError("apply_TM_1sO: an operator insertion needs the single-layer on-site tensor")
    op = op.to(dtype=a.dtype, device=a.device)
    ket_idx = {2: '', 3: 'm', 4: 'mn'}[rank]
    a_ket = ad.contract(eng, f'qefgh,qp{ket_idx}->pefgh{ket_idx}', a, op)
    lhs, out = spec.split('->')
    t1, e, at, t2 = lhs.split(',')
    if mpo_in:
        e += 'm'                                      # the edge's MPO index meets the operator's
    if (rank == 3 and not mpo_in) or rank == 4:
        out += ket_idx[-1]                            # the operator's open MPO index trails the result
    return ad.sl_chain(eng, f'{t1},{e},{at},{t2}->{out}', (env.T[(c, k1)], edge, env.T[(c, k2)]), a, a_ket=a_ket,
                       ket_extra=ket_idx).contiguous()


def _split_two_site(op):
    """op[s0,s1;s0',s1'] = sum_k op_l[s0,s0',k] op_r[s1,s1',k] through the SVD of the (s0 s0') x (s1 s1') matrix
    (corrf.py:952-959: op_r = S V^H)."""
    p = op.size(0)
    U, S, Vh = torch.linalg.svd(op.permute(0, 2, 1, 3).contiguous().reshape(p * p, p * p))
    op_l = U.reshape(p, p, S.size(0))
    op_r = (Vh * S[:, None].to(Vh.dtype)).reshape(S.size(0), p, p).permute(1, 2, 0).contiguous()
    return op_l, op_r


# direction -> (T1 key, T2 key, step to the second site, chain of the first site, chain of the second)   (corrf.py:779-910)
_TM2 = {
    (0, 1): ((-1, 0), (1, 0), (1, 0), 'xal,xuvz,@uldr->avzdr', 'avzdr,@vrfq,zqe->adfe'),
    (1, 0): ((0, -1), (0, 1), (0, 1), 'xua,xlvz,@uldr->avzdr', 'avzdr,@dvfq,fze->arqe'),
}


def apply_TM_2sO_2sChannel(coord, direction, state, env, edge, op=None, verbosity=0):
    r"""One width-2 transfer matrix -- site ``coord`` and its right (growing down) or lower (growing right) neighbour -- applied
    to the rank-4 ``edge``, with the two-site operator ``op`` (rank 4; None = identity) split over the two sites
    (corrf.py:671-912; as there, only the directions down and right exist).  The reference splits ``op`` = U S V^H into U and
    V^H and drops S (corrf.py:717-722); so does this function: the results are the reference's."""
    if direction in ((0, -1), (-1, 0)):
        raise ValueError("Direction: " + str(direction) + "not implemented")
    if direction not in _TM2:
        raise ValueError("Invalid direction: " + str(direction))
    eng = _engine()
    k1, k2, step, s1, s2 = _TM2[direction]
    c = state.vertexToSite(coord)
    c2 = state.vertexToSite((coord[0] + step[0], coord[1] + step[1]))
    a1, a2 = state.site(c), state.site(c2)
    T1, T2 = env.T[(c, k1)], env.T[(c2, k2)]
    if op is None:
        E = ad.sl_chain(eng, s1, (T1, edge), a1)
        return ad.sl_chain(eng, s2, (E.contiguous(), T2), a2).contiguous()
    if op.dim() != 4:
        raise ValueError(f"Invalid op: rank {op.size()}")
    op = op.to(dtype=a1.dtype, device=a1.device)
    p = op.size(0)
    U, S, Vh = torch.linalg.svd(op.permute(0, 2, 1, 3).contiguous().reshape(p * p, p * p))
    op_1 = U.reshape(p, p, S.size(0))
    op_2 = Vh.reshape(S.size(0), p, p).permute(1, 2, 0).contiguous()
    k_1 = ad.contract(eng, 'qefgh,qpk->pefghk', a1, op_1)
    k_2 = ad.contract(eng, 'qefgh,qpk->pefghk', a2, op_2)
    l1, o1 = s1.split('->')
    E = ad.sl_chain(eng, f'{l1}->{o1}k', (T1, edge), a1, a_ket=k_1, ket_extra='k')
    return ad.sl_chain(eng, s2.replace('avzdr,', 'avzdrk,'), (E.contiguous(), T2), a2, a_ket=k_2, ket_extra='k').contiguous()


def apply_TM_2sO_1sChannel(coord, direction, state, env, edge, op=None, verbosity=0):
    r"""Two consecutive transfer matrices (sites ``coord`` and ``coord + direction``) carrying the two-site operator ``op``
    (rank 4; None = identity), split over the two sites with its bond travelling on the edge in between (corrf.py:914-978)."""
    op_l, op_r = (None, None) if op is None else _split_two_site(op.to(dtype=state.dtype, device=state.device))
    E = apply_TM_1sO(coord, direction, state, env, edge, op=op_l)
    return apply_TM_1sO((coord[0] + direction[0], coord[1] + direction[1]), direction, state, env, E, op=op_r)


# direction -> (C1 key, C2 key, the two contractions T (x) C1 and C2 (x) that)                        (corrf.py:48-101)
_EDGE = {
    (0, -1): ((1, -1), (-1, -1), ('T', 'abc,cd->abd'), ('C2first', 'xa,abd->xbd')),
    (-1, 0): ((-1, -1), (-1, 1), ('C', 'ab,acd->bcd'), ('C2last', 'bcd,ce->bde')),
    (0, 1): ((-1, 1), (1, 1), ('C', 'ab,cbd->acd'), ('C2last', 'acd,ed->ace')),
    (1, 0): ((1, 1), (1, -1), ('T', 'abc,cd->abd'), ('C2first', 'xa,abd->xbd')),
}


def _dot(x, y):
    """sum_k x_k y_k (no conjugation) as a 1 x 1 contraction."""
    return ad.contract(_engine(), 'ka,kb->ab', x.reshape(-1, 1), y.reshape(-1, 1)).reshape(())


def get_edge(coord, direction, state, env, verbosity=0):
    r"""The boundary C--T--C of site ``coord`` on the side ``direction`` as a :math:`\chi \times D^2 \times \chi` tensor, indices
    ordered left to right / top to bottom (corrf.py:10-104)."""
    if direction not in _EDGE:
        raise ValueError("Invalid direction: " + str(direction))
    eng = _engine()
    c = state.vertexToSite(coord)
    k1, k2, (first, s1), (second, s2) = _EDGE[direction]
    C1, C2, T = env.C[(c, k1)], env.C[(c, k2)], env.T[(c, direction)]
    E = ad.contract(eng, s1, T, C1) if first == 'T' else ad.contract(eng, s1, C1, T)
    return ad.contract(eng, s2, C2, E) if second == 'C2first' else ad.contract(eng, s2, E, C2)


# direction -> (first C, last C, step to the second site, C (x) T, (x) T', (x) C)                      (corrf.py:138-229)
_EDGE2 = {
    (0, -1): ((-1, -1), (1, -1), (1, 0), 'ab,bcd->acd', 'acd,def->acef', 'acef,fg->aceg'),
    (-1, 0): ((-1, -1), (-1, 1), (0, 1), 'ab,acd->bcd', 'bcd,cef->bdef', 'bdef,eg->bdfg'),
    (0, 1): ((-1, 1), (1, 1), (1, 0), 'ab,cbd->acd', 'acd,edf->acef', 'acef,gf->aceg'),
    (1, 0): ((1, -1), (1, 1), (0, 1), 'ab,bcd->acd', 'acd,def->acef', 'acef,fg->aceg'),
}


def get_edge_2(coord, direction, state, env, verbosity=0):
    r"""The boundary C--T--T--C of the two sites ``coord`` and its right (up / down edges) or lower (left / right edges)
    neighbour, indices ordered left to right / top to bottom (corrf.py:105-232)."""
    if direction not in _EDGE2:
        raise ValueError("Invalid direction: " + str(direction))
    eng = _engine()
    kc1, kc2, step, s1, s2, s3 = _EDGE2[direction]
    c = state.vertexToSite(coord)
    c2 = state.vertexToSite((coord[0] + step[0], coord[1] + step[1]))
    E = ad.contract(eng, s1, env.C[(c, kc1)], env.T[(c, direction)])
    E = ad.contract(eng, s2, E, env.T[(c2, direction)])
    return ad.contract(eng, s3, E, env.C[(c2, kc2)])


def apply_edge(coord, direction, state, env, vec, verbosity=0):
    r"""The scalar ``vec`` . get_edge(coord, direction) (rank-3 ``vec``) or ``vec`` . get_edge_2(coord, direction) (rank 4) over
    all indices, no conjugation (corrf.py:234-277)."""
    if vec.dim() == 3:
        E = get_edge(coord, direction, state, env, verbosity=verbosity)
    elif vec.dim() == 4:
        E = get_edge_2(coord, direction, state, env, verbosity=verbosity)
    else:
        raise ValueError("Unsupported edge: " + str(tuple(vec.shape)))
    return _dot(vec, E)


def corrf_1sO1sO(coord, direction, state, env, op1, get_op2, dist, rl_0=None, verbosity=0):
    r""":math:`\langle O_1(0)\, O_2(r) \rangle` for r = 1 .. dist+1 along ``direction`` from site ``coord`` (corrf.py:980-1067):
    three edges are carried from site to site -- the norm network, the network with ``op1`` at the origin, and the latter
    closed with ``get_op2(r)`` -- and each distance is one ratio of two scalars.  ``rl_0`` = (left, right) callables of a
    site that replace the environment's edges (leading transfer-matrix eigenvectors), as in the reference.  The edges are
    rescaled by the largest element of the norm edge after every site, which the ratio does not see."""
    def shift(c):
        return (c[0] + direction[0], c[1] + direction[1])

    def close(E, c):
        if rl_0 is None:
            return apply_edge(c, direction, state, env, E, verbosity=verbosity)
        return _dot(E, rl_0[1](c))

    c0 = coord
    rev = (-direction[0], -direction[1])
    E0 = get_edge(c0, rev, state, env, verbosity=verbosity) if rl_0 is None else rl_0[0](c0)
    E1 = apply_TM_1sO(c0, direction, state, env, E0, op=op1, verbosity=verbosity)
    E0 = apply_TM_1sO(c0, direction, state, env, E0, verbosity=verbosity)
    out = torch.empty(dist + 1, dtype=state.dtype, device=state.device)
    for r in range(dist + 1):
        c0 = shift(c0)
        E12 = apply_TM_1sO(c0, direction, state, env, E1, op=get_op2(r), verbosity=verbosity)
        E0 = apply_TM_1sO(c0, direction, state, env, E0, verbosity=verbosity)
        E1 = apply_TM_1sO(c0, direction, state, env, E1, verbosity=verbosity)
        out[r] = close(E12, c0) / close(E0, c0)
        scale = E0.abs().max()
        E0, E1 = E0 / scale, E1 / scale
    return out


def apply_TM_0sO(coord, direction, state, env, edge, verbosity=0):
    r"""``edge`` (:math:`\chi \times \chi`) with one width-0 transfer matrix applied (corrf.py:278-362; as there, right is
    evaluated with the formula of left and down with the formula of up)."""
    if direction not in _TM1:
        raise ValueError("Invalid direction: " + str(direction))
    eng = _engine()
    if direction == (1, 0):
        direction = (-1, 0)
    if direction == (0, 1):
        direction = (0, -1)
    c = state.vertexToSite(coord)
    if direction == (0, -1):
        T1 = env.T[(c, (-1, 0))]
        T2 = env.T[(state.vertexToSite((c[0] - 1, c[1])), (1, 0))]
        E = ad.contract(eng, 'abl,bc->alc', T1, edge)
        return ad.contract(eng, 'alc,elc->ae', E, T2)
    T1 = env.T[(c, (0, -1))]
    T2 = env.T[(state.vertexToSite((c[0], c[1] - 1)), (0, 1))]
    E = ad.contract(eng, 'aub,bc->auc', T1, edge)
    return ad.contract(eng, 'auc,uec->ae', E, T2)


def corrf_2sOH2sOH_E1(coord, direction, state, env, op1, get_op2, dist, verbosity=0):
    r""":math:`\langle O_1(0,1)\, O_2(r, r+1) \rangle` of two two-site operators laid along ``direction``, r = 2 .. dist+2
    (corrf.py:1069-1156)."""
    def shift(c, r=1):
        return (c[0] + r * direction[0], c[1] + r * direction[1])

    c0 = coord
    rev = (-direction[0], -direction[1])
    E0 = get_edge(c0, rev, state, env)
    E1 = apply_TM_2sO_1sChannel(c0, direction, state, env, E0, op=op1)
    E0 = apply_TM_2sO_1sChannel(c0, direction, state, env, E0)
    c0 = shift(c0, 2)
    out = torch.empty(dist + 1, dtype=state.dtype, device=state.device)
    for r in range(dist + 1):
        E12 = apply_TM_2sO_1sChannel(c0, direction, state, env, E1, op=get_op2(r))
        E0 = apply_TM_1sO(c0, direction, state, env, E0)
        E1 = apply_TM_1sO(c0, direction, state, env, E1)
        c0 = shift(c0)
        E00 = apply_TM_1sO(c0, direction, state, env, E0)
        out[r] = apply_edge(c0, direction, state, env, E12) / apply_edge(c0, direction, state, env, E00)
        scale = E0.abs().max()
        E0, E1 = E0 / scale, E1 / scale
    return out


def corrf_2sOV2sOV_E2(coord, direction, state, env, op1, get_op2, dist, verbosity=0):
    r""":math:`\langle O_1(0)\, O_2(r) \rangle` of two two-site operators laid ACROSS ``direction`` (down or right), r = 1 ..
    dist+1, through the width-2 transfer matrix (corrf.py:1158-1248)."""
    c0 = coord
    rev = (-direction[0], -direction[1])
    E0 = get_edge_2(c0, rev, state, env)
    E1 = apply_TM_2sO_2sChannel(c0, direction, state, env, E0, op=op1)
    E0 = apply_TM_2sO_2sChannel(c0, direction, state, env, E0)
    out = torch.empty(dist + 1, dtype=state.dtype, device=state.device)
    for r in range(dist + 1):
        c0 = (c0[0] + direction[0], c0[1] + direction[1])
        E12 = apply_TM_2sO_2sChannel(c0, direction, state, env, E1, op=get_op2(r))
        E0 = apply_TM_2sO_2sChannel(c0, direction, state, env, E0)
        E1 = apply_TM_2sO_2sChannel(c0, direction, state, env, E1)
        out[r] = apply_edge(c0, direction, state, env, E12) / apply_edge(c0, direction, state, env, E0)
        scale = E0.abs().max()
        E0, E1 = E0 / scale, E1 / scale
    return out
