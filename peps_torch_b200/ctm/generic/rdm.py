"""Drop-in for the reduced density matrices of ctm/generic/rdm.py that the J1-J2 scripts evaluate: rdm2x2 (:1306-1360,
rdm2x2_legacy :1362-1592) -- the energy is tr(rho_2x2 h_p) over the plaquettes of the unit cell (models/j1j2.py:223-247)
-- and rdm1x1 / rdm2x1 / rdm1x2 (:71-112, 304-350, 622-670) behind the magnetisations and nearest-neighbour correlations
(models/j1j2.py eval_obs).  The networks -- enlarged corners with open physical legs closed by the remaining edge
tensors -- are contracted by libctmb."""
from ... import config as cfg


def _engine():
    from ...engine import default_engine
    return default_engine()


def rdm2x2(coord, state, env, open_sites=[0, 1, 2, 3], unroll=[], checkpoint_unrolled=False, checkpoint_on_device=False,
           sym_pos_def=False, force_cpu=False, verbosity=0, global_args=cfg.global_args):
    r"""
    :param coord: vertex (x,y) of the upper-left site of the 2x2 plaquette
    :param open_sites: sites left open (``s0 s1 / s2 s3``); the others are traced, the order of the remaining indices is kept
    :param sym_pos_def: enforce hermiticity (always) and positive definiteness if ``True``
    :return: reduced density matrix with indices :math:`s_0s_1s_2s_3;s'_0s'_1s'_2s'_3`, trace-normalised
    ``unroll`` / ``checkpoint_*`` / ``force_cpu`` of the reference steer memory use of its einsum path and are ignored.
    """
    return _engine().rdm2x2(coord, state, env, open_sites=open_sites, sym_pos_def=sym_pos_def)


def rdm2x2_legacy(coord, state, env, sym_pos_def=False, verbosity=0):
    return _engine().rdm2x2(coord, state, env, sym_pos_def=sym_pos_def)


def rdm2x2_oe(coord, state, env, open_sites=[0, 1, 2, 3], unroll=False, checkpoint_unrolled=False, checkpoint_on_device=False,
              sym_pos_def=False, force_cpu=False, verbosity=0, global_args=cfg.global_args):
    r"""The opt_einsum variant the reference's ``rdm2x2`` dispatches to (rdm.py:1594-1675): the same density matrix."""
    return _engine().rdm2x2(coord, state, env, open_sites=open_sites, sym_pos_def=sym_pos_def)


def _rdm1x1(coord, state, env, operator, sym_pos_def):
    """With ``operator`` the reference returns the scalar sum_{s s'} rho[s;s'] operator[s';s] of the UNNORMALISED network
    (rdm.py:89-90,175-181,275-299): the raw 1-site network comes from libctmb, the p x p trace against the operator is taken here."""
    if operator is None:
        return _engine().rdm_small('1x1', coord, state, env, sym_pos_def=sym_pos_def)
    raw = _engine().rdm_small('1x1', coord, state, env, raw=True)
    return (raw * operator.to(dtype=raw.dtype, device=raw.device).t()).sum()


def rdm1x1(coord, state, env, mode='sl', operator=None, sym_pos_def=False, force_cpu=False, verbosity=0):
    r""":return: 1-site reduced density matrix with indices :math:`s;s'` (ctm/generic/rdm.py:71-112); ``mode`` selects
    between equivalent contraction orders in the reference and is ignored"""
    return _rdm1x1(coord, state, env, operator, sym_pos_def)


def rdm2x1(coord, state, env, mode='sl', sym_pos_def=False, force_cpu=False, unroll=False, checkpoint_unrolled=False,
           checkpoint_on_device=False, verbosity=0):
    r""":return: 2-site reduced density matrix :math:`s_0s_1;s'_0s'_1`, s1 at ``coord+(1,0)`` (rdm.py:304-350)"""
    return _engine().rdm_small('2x1', coord, state, env, sym_pos_def=sym_pos_def)


def rdm1x2(coord, state, env, mode='sl', sym_pos_def=False, force_cpu=False, unroll=False, checkpoint_unrolled=False,
           checkpoint_on_device=False, verbosity=0):
    r""":return: 2-site reduced density matrix :math:`s_0s_1;s'_0s'_1`, s1 at ``coord+(0,1)`` (rdm.py:622-670)"""
    return _engine().rdm_small('1x2', coord, state, env, sym_pos_def=sym_pos_def)


def rdm1x1_dl(coord, state, env, operator=None, sym_pos_def=False, force_cpu=False, verbosity=0):
    return _rdm1x1(coord, state, env, operator, sym_pos_def)


def rdm2x1_dl(coord, state, env, sym_pos_def=False, force_cpu=False, verbosity=0):
    return _engine().rdm_small('2x1', coord, state, env, sym_pos_def=sym_pos_def)


def rdm1x2_dl(coord, state, env, sym_pos_def=False, force_cpu=False, verbosity=0):
    return _engine().rdm_small('1x2', coord, state, env, sym_pos_def=sym_pos_def)


rdm1x1_sl, rdm2x1_sl, rdm1x2_sl = rdm1x1_dl, rdm2x1_dl, rdm1x2_dl
