"""Transfer-matrix mat-vecs of ctm/generic/corrf.py of peps-torch (apply_TM_0sO :278-362, apply_TM_1sO :364-650 with op=None):
one application of the width-0 / width-1 transfer matrix of a site to an edge tensor.  Each is one contraction chain through
libctmb (the double-layer tensor a (x) a* is never formed: the chain contracts the two layers one after the other)."""
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
    Only the identity insertion (``op=None``) is built."""
    if op is not None or edge.dim() != 3:
        raise NotImplementedError("apply_TM_1sO: operator insertions / MPO legs are not built in peps_torch_b200")
    if direction not in _TM1:
        raise ValueError("Invalid direction: " + str(direction))
    c = state.vertexToSite(coord)
    k1, k2, spec = _TM1[direction]
    return ad.sl_chain(_engine(), spec, (env.T[(c, k1)], edge, env.T[(c, k2)]), state.site(c)).contiguous()


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
