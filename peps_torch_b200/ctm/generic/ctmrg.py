"""Drop-in for ctm/generic/ctmrg.py of peps-torch: same public names and signatures
(run :18-110, ctm_MOVE :179-319); the body of a move is one call into libctmb."""
import time
import logging
import torch
from ... import config as cfg

log = logging.getLogger(__name__)
_SUPPORTED_SVD = ('DEFAULT', 'GESDD', 'RSVD', 'RSVD_CUSTOM')


def _engine():
    from ...engine import default_engine      # raises if there is no CUDA device / library
    return default_engine()


def _options(ctm_args):
    method = getattr(ctm_args, 'projector_svd_method', 'DEFAULT')
    if method not in _SUPPORTED_SVD:
        # the reference raises a bare string here (ctm_projectors.py:257), i.e. a TypeError
        raise TypeError(f'Projector svd method "{method}" not implemented')
    norm = getattr(ctm_args, 'ctm_absorb_normalization', 'inf')
    if norm != 'inf':
        raise ValueError("libctmb implements ctm_absorb_normalization='inf' only, got " + str(norm))
    pm = getattr(ctm_args, 'projector_method', '4X4')
    if pm not in ('4X4', '4X2'):
        raise ValueError("Invalid Projector method: " + str(pm))
    return dict(projector_method={'4X4': 0, '4X2': 1}[pm],
                svd_reltol=ctm_args.projector_svd_reltol,
                eps_multiplet=ctm_args.projector_eps_multiplet,
                multiplet_abstol=ctm_args.projector_multiplet_abstol,
                rsvd_niter=getattr(ctm_args, 'b200_rsvd_niter', None),
                rsvd_rank_factor=getattr(ctm_args, 'b200_rsvd_rank_factor', None),
                rsvd_tol=getattr(ctm_args, 'b200_rsvd_tol', None))


def ctm_MOVE(direction, state, env, ctm_args=cfg.ctm_args, global_args=cfg.global_args,
             verbosity=0, diagnostics=None):
    r"""
    :param direction: one of Up=(0,-1), Left=(-1,0), Down=(0,1), Right=(1,0)
    :param state: wavefunction (``sites``, ``vertexToSite``)
    :param env: environment (``chi``, ``C``, ``T``); entries at coord-direction are replaced
    Executes a single directional CTM move (projectors for all sites, then absorption,
    truncation and normalisation for all sites) on the GPU.
    """
    eng = _engine()
    if direction not in ((0, -1), (-1, 0), (0, 1), (1, 0)):
        raise ValueError("Invalid direction: " + str(direction))
    eng.move_generic(direction, state, env, **_options(ctm_args))


def run(state, env, conv_check=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args):
    r"""
    Executes directional CTM for a generic iPEPS starting from ``env``; same contract as the
    reference: returns ``(env, history, t_ctm, t_obs)`` and mutates ``env`` in place.
    ``t_ctm`` is bracketed by stream synchronisation.
    """
    eng = _engine()
    dev = eng.device

    def _ctmrg_iter():
        for direction in ctm_args.ctm_move_sequence:
            reps = state.lX if direction in [(-1, 0), (1, 0)] else state.lY
            for _ in range(reps):
                ctm_MOVE(direction, state, env, ctm_args=ctm_args, global_args=global_args)

    t_obs = t_ctm = 0.
    history = None
    for i in range(ctm_args.ctm_max_iter):
        torch.cuda.synchronize(dev)
        t0_ctm = time.perf_counter()
        _ctmrg_iter()
        torch.cuda.synchronize(dev)
        t1_ctm = time.perf_counter()

        t0_obs = time.perf_counter()
        if conv_check is not None:
            converged, history = conv_check(state, env, history, ctm_args=ctm_args)
            if converged:
                if getattr(ctm_args, 'verbosity_ctm_convergence', 0) > 0:
                    print(f"CTMRG  converged at iter= {i}, history= {history['conv_crit'][-1]}")
                break
        t1_obs = time.perf_counter()
        t_ctm += t1_ctm - t0_ctm
        t_obs += t1_obs - t0_obs
    return env, history, t_ctm, t_obs
