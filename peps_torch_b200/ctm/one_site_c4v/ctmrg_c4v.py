"""Drop-in for ctm/one_site_c4v/ctmrg_c4v.py of peps-torch (run :16-108, run_dl :110-176,
ctm_MOVE_dl :200-322, ctm_MOVE_sl :325-463)."""
import time
import torch
from ... import config as cfg

# all of these ask for the chi eigenpairs of largest magnitude of the Hermitian enlarged corner
# (ctmrg_c4v.py:49-67,127-140); here every one of them runs libctmb's residual-checked subspace iteration
_SUPPORTED_EIG = ('DEFAULT', 'SYMEIG', 'SYMARP', 'SYMLOBPCG', 'QR')


def _engine():
    from ...engine import default_engine
    return default_engine()


def _sync(dev):
    if isinstance(dev, torch.device) and dev.type == 'cuda':
        torch.cuda.synchronize(dev)


def _move(a, env, ctm_args):
    eng = _engine()
    from ... import ad
    if ad.needs_grad((a, env.C[env.keyC], env.T[env.keyT])):
        # reverse-mode AD (optim_j1j2_c4v.py): the same move from differentiable libctmb calls (peps_torch_b200/ad.py)
        env.C[env.keyC], env.T[env.keyT] = ad.ctm_move_c4v(eng, a, env.C[env.keyC], env.T[env.keyT], env.chi, ctm_args)
        return
    # ctmrg_c4v.py:182-197: C is always scaled by |C[0,0]|; T by the infinity norm ('inf') or the 2-norm (anything else)
    norm = 0 if getattr(ctm_args, 'ctm_absorb_normalization', 'inf') == 'inf' else 1
    nC, nT, _ = eng.move_c4v(a, env.C[env.keyC], env.T[env.keyT], env.chi, norm_type=norm,
                             rsvd_niter=getattr(ctm_args, 'b200_rsvd_niter', None),
                             rsvd_rank_factor=getattr(ctm_args, 'b200_rsvd_rank_factor', None),
                             rsvd_tol=getattr(ctm_args, 'b200_rsvd_tol_c4v', None))
    env.C[env.keyC] = nC
    env.T[env.keyT] = nT


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
    _engine()                           # no CUDA device / library: fail before anything else
    if a.dim() != 5:
        raise ValueError("ctm_MOVE_sl contracts the single-layer tensor a[s,u,l,d,r]; use ctm_MOVE_dl for a double-layer one")
    _move(a, env, ctm_args)


def ctm_MOVE_QR_sl(a, env, ctm_args=cfg.ctm_args, global_args=cfg.global_args, past_steps_data=None):
    r"""
    The QR variant of the C4v move (ctmrg_c4v.py:465-602, ``projector_svd_method='QR'``): the projector is the thin Q of the
    (chi D^2) x chi matrix C.T instead of the leading eigenvectors of the enlarged corner; C' = P^T C2x2 P, T' as in
    ctm_MOVE_sl.  Composed from libctmb calls: contraction chains, the Householder QR (ctmb_qr), element-wise
    symmetrisation / normalisation on the device.  Forward only (no QR adjoint is built).
    """
    eng = _engine()
    from ... import ad
    C, T = env.C[env.keyC], env.T[env.keyT]
    if a.dim() != 5:
        raise ValueError("ctm_MOVE_QR_sl contracts the single-layer tensor a[s,u,l,d,r]")
    if ad.needs_grad((a, C, T)):
        raise NotImplementedError("ctm_MOVE_QR_sl: reverse-mode AD through the QR projector is not built")
    with torch.no_grad():
        t = ad.sl_chain(eng, 'ab,xbu,ael,@uldr->edxr', (C, T, T), a)
        C2X2 = t.reshape(t.shape[0] * t.shape[1], t.shape[2] * t.shape[3])
        C1x2 = eng.einsum2('ab,cbd->adc', C, T).reshape(-1, env.chi)
        P, _ = eng.qr(C1x2)
        P = P.contiguous()
        nC = eng.einsum2('ki,kj->ij', P, eng.einsum2('ik,kj->ij', C2X2, P))          # P^T C2X2 P (plain transpose)
        Pv = P.reshape(C.shape[0], T.shape[2], P.shape[1])
        nT = ad.sl_chain(eng, 'acl,aux,@uldr,cdy->xyr', (T, Pv, Pv), a, conj_last=True)
        nT = 0.5 * (nT + nT.conj().permute(1, 0, 2))
        nC = nC / torch.abs(nC[0, 0])
        norm = getattr(ctm_args, 'ctm_absorb_normalization', 'inf')
        nT = nT / torch.linalg.vector_norm(nT, ord=float('inf') if norm == 'inf' else 2)
    env.C[env.keyC] = nC
    env.T[env.keyT] = nT.contiguous()


def ctm_MOVE_dl(a, env, f_c2x2_decomp=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args):
    r"""
    :param a: on-site C4v symmetric tensor a[s,u,l,d,r], or the double-layer tensor A[D^2,D^2,D^2,D^2]
    The double-layer variant of the move (ctmrg_c4v.py:200-322): A = a (x) a* is formed once (O(D^8) memory, by
    libctmb's contraction kernel) and the corner / nT chains contract it as one operand.
    """
    eng = _engine()
    if a.dim() == 5:
        from ..generic.ctmrg import double_layer
        a = double_layer(eng, a)
    elif a.dim() != 4:
        raise ValueError(f"on-site tensor of rank {a.dim()}")
    _move(a, env, ctm_args)


def _run(state, env, conv_check, ctm_args, global_args, move):
    if ctm_args.projector_svd_method not in _SUPPORTED_EIG:
        raise Exception(f"Projector eig/svd method \"{ctm_args.projector_svd_method}\" not implemented")
    if getattr(ctm_args, 'fpcm_freq', -1) > 0:
        raise Exception("fixed-point corner acceleration (fpcm_MOVE_sl, CTMARGS.fpcm_freq > 0) is not implemented in libctmb")
    eng = _engine()
    a = next(iter(state.sites.values()))
    t_obs = t_ctm = 0.
    history = None
    for i in range(ctm_args.ctm_max_iter):
        _sync(eng.device)
        t0_ctm = time.perf_counter()
        if ctm_args.projector_svd_method == 'QR' and i > getattr(ctm_args, 'fpcm_init_iter', 1) and move is ctm_MOVE_sl:
            # ctmrg_c4v.py:87-89: the first moves use the eigenvalue projector, then the QR projector takes over
            ctm_MOVE_QR_sl(a, env, ctm_args=ctm_args, global_args=global_args)
        else:
            move(a, env, None, ctm_args=ctm_args, global_args=global_args)
        _sync(eng.device)
        t1_ctm = time.perf_counter()
        t0_obs = time.perf_counter()
        if conv_check is not None:
            converged, history = conv_check(state, env, history, ctm_args=ctm_args)
            if converged:
                if getattr(ctm_args, 'verbosity_ctm_convergence', 0) > 0:
                    print(f"CTMRG converged at iter= {i}")
                break
        t1_obs = time.perf_counter()
        t_ctm += t1_ctm - t0_ctm
        t_obs += t1_obs - t0_obs
    return env, history, t_ctm, t_obs


def run(state, env, conv_check=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args):
    r"""Same contract as the reference: ``(env, history, t_ctm, t_obs)``; one move per iteration."""
    return _run(state, env, conv_check, ctm_args, global_args, ctm_MOVE_sl)


def run_dl(state, env, conv_check=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args):
    r"""Double-layer variant of :func:`run` (ctmrg_c4v.py:110-176): every move is ``ctm_MOVE_dl``."""
    a = next(iter(state.sites.values()))
    if a.dim() == 5:                    # build A once instead of once per move
        from ..generic.ctmrg import double_layer
        A = double_layer(_engine(), a)

        def move(_a, env, f, ctm_args=ctm_args, global_args=global_args):
            return ctm_MOVE_dl(A, env, f, ctm_args=ctm_args, global_args=global_args)
        return _run(state, env, conv_check, ctm_args, global_args, move)
    return _run(state, env, conv_check, ctm_args, global_args, ctm_MOVE_dl)
