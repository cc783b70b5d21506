"""Transfer-matrix mat-vecs and two-point functions of ctm/generic/corrf.py of peps-torch (get_edge :10-104, apply_edge
:234-277, apply_TM_0sO :278-362, apply_TM_1sO :364-650 with op=None or a one-site operator, corrf_1sO1sO :980-1067).  One
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
    r"""``edge`` (:math:`\chi \times D^2 \times \chi`) with one transfer matrix of site ``coord`` applied (corrf.py:364-650).
    ``op`` is None (identity on the physical leg) or a one-site operator (rank 2, applied to the ket layer as corrf.py:415-419);
    the MPO variants (rank-3 ``op``, rank-4 ``edge``) are not built."""
    if edge.dim() != 3 or (op is not None and op.dim() != 2):
        raise NotImplementedError("apply_TM_1sO: MPO operators / MPO legs are not built in peps_torch_b200")
    if direction not in _TM1:
        raise ValueError("Invalid direction: " + str(direction))
    c = state.vertexToSite(coord)
    k1, k2, spec = _TM1[direction]
    a = state.site(c)
    a_ket = None
    if op is not None:
        if a.dim() != 5:
            raise NotImplementedError("apply_TM_1sO: an operator insertion needs the single-layer on-site tensor")
        a_ket = ad.contract(_engine(), 'mefgh,mn->nefgh', a, op.to(dtype=a.dtype, device=a.device))
    return ad.sl_chain(_engine(), spec, (env.T[(c, k1)], edge, env.T[(c, k2)]), a, a_ket=a_ket).contiguous()


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


def apply_edge(coord, direction, state, env, vec, verbosity=0):
    r"""The scalar ``vec`` . get_edge(coord, direction) over all three indices, no conjugation (corrf.py:234-277)."""
    if vec.dim() != 3:
        raise NotImplementedError("apply_edge: width-2 edges are not built in peps_torch_b200")
    E = get_edge(coord, direction, state, env, verbosity=verbosity)
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
