"""Drop-in for ctm/one_site_c4v/ctmrg_c4v.py of peps-torch (run :16-108, ctm_MOVE_sl :325-463)."""
import time
import torch
from ... import config as cfg


def _engine():
    from ...engine import default_engine
    return default_engine()


def ctm_MOVE_sl(a, env, f_c2x2_decomp=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args,
                past_steps_data=None):
    r"""
    :param a: on-site C4v symmetric tensor a[s,u,l,d,r]
    :param env: C4v environment (``chi``, ``C[keyC]``, ``T[keyT]``)
    :param f_c2x2_decomp: ignored (the truncated Hermitian decomposition runs inside libctmb);
                          kept for signature compatibility
    One C4v move: enlarged corner -> leading-chi eigenpairs -> C' = diag(D), T' = P T a a* P*,
    symmetrised and normalised; the two env entries are replaced by fresh tensors.
    """
    eng = _engine()
    if getattr(ctm_args, 'ctm_absorb_normalization', 'inf') != 'inf':
        raise ValueError("libctmb implements ctm_absorb_normalization='inf' only")
    nC, nT, _ = eng.move_c4v(a, env.C[env.keyC], env.T[env.keyT], env.chi,
                             rsvd_niter=getattr(ctm_args, 'b200_rsvd_niter', None),
                             rsvd_rank_factor=getattr(ctm_args, 'b200_rsvd_rank_factor', None),
                             rsvd_tol=getattr(ctm_args, 'b200_rsvd_tol_c4v', None))
    env.C[env.keyC] = nC
    env.T[env.keyT] = nT


def run(state, env, conv_check=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args):
    r"""Same contract as the reference: ``(env, history, t_ctm, t_obs)``; one move per iteration."""
    if ctm_args.projector_svd_method not in ('DEFAULT', 'SYMEIG'):
        raise Exception(f"Projector eig/svd method \"{ctm_args.projector_svd_method}\" not implemented")
    eng = _engine()
    a = next(iter(state.sites.values()))
    t_obs = t_ctm = 0.
    history = None
    for i in range(ctm_args.ctm_max_iter):
        torch.cuda.synchronize(eng.device)
        t0_ctm = time.perf_counter()
        ctm_MOVE_sl(a, env, None, ctm_args=ctm_args, global_args=global_args)
        torch.cuda.synchronize(eng.device)
        t1_ctm = time.perf_counter()
        t0_obs = time.perf_counter()
        if conv_check is not None:
            converged, history = conv_check(state, env, history, ctm_args=ctm_args)
            if converged:
                break
        t1_obs = time.perf_counter()
        t_ctm += t1_ctm - t0_ctm
        t_obs += t1_obs - t0_obs
    return env, history, t_ctm, t_obs
