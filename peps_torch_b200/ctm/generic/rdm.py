"""Drop-in for the 2x2 reduced density matrix of ctm/generic/rdm.py (rdm2x2 :1306-1360, rdm2x2_legacy :1362-1592):
the energy of the J1-J2 scripts is tr(rho_2x2 h_p) over the plaquettes of the unit cell (models/j1j2.py:223-247).
The network -- four enlarged corners with open physical legs, two halves, one trace -- is contracted by libctmb."""
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
