"""Drop-in for ctm/generic/ctmrg.py of peps-torch: same public names and signatures
(run :18-110, run_overlap :112-175, ctm_MOVE :179-319); the body of a move is one call into libctmb."""
import copy
import time
import logging
from math import ceil
import torch
from ... import config as cfg
from ...ipeps import IPEPS

log = logging.getLogger(__name__)
# every value the reference dispatches on (ctm_projectors.py:213-257) asks for the same object -- the leading chi
# singular triplets of M -- from a different LAPACK / ARPACK / randomised driver; here all of them run the
# residual-checked randomised decomposition of libctmb
_announced = set()
_SUPPORTED_SVD = ('DEFAULT', 'GESDD', 'GESDD_CPU', 'AF', 'ARP', 'PROPACK', 'RSVD', 'RSVD_CUSTOM')


def _engine():
    from ...engine import default_engine      # raises if there is no CUDA device / library
    return default_engine()


def _options(ctm_args):
    method = getattr(ctm_args, 'projector_svd_method', 'DEFAULT')
    if method not in _SUPPORTED_SVD:
        # the reference raises a bare string here (ctm_projectors.py:257), i.e. a TypeError
        raise TypeError(f'Projector svd method "{method}" not implemented')
    if method not in ('DEFAULT', 'RSVD', 'RSVD_CUSTOM') and method not in _announced \
            and getattr(ctm_args, 'verbosity_projectors', 0) + getattr(ctm_args, 'verbosity_ctm_move', 0) > 0:
        # no multi-backend dispatch here: say so once instead of silently serving a different driver
        _announced.add(method)
        log.info(f"projector_svd_method={method}: libctmb serves every method with its residual-checked randomised "
                 "decomposition (leading chi triplets to LAPACK-grade residuals); the named driver is not called")
    # ctmrg.py:212-214: 'inf' is the infinity norm, every other value the vector 2-norm
    norm = 0 if getattr(ctm_args, 'ctm_absorb_normalization', 'inf') == 'inf' else 1
    pm = getattr(ctm_args, 'projector_method', '4X4')
    if pm not in ('4X4', '4X2'):
        raise ValueError("Invalid Projector method: " + str(pm))
    return dict(projector_method={'4X4': 0, '4X2': 1}[pm], norm_type=norm,
                svd_reltol=ctm_args.projector_svd_reltol,
                eps_multiplet=ctm_args.projector_eps_multiplet,
                multiplet_abstol=ctm_args.projector_multiplet_abstol,
                rsvd_niter=getattr(ctm_args, 'b200_rsvd_niter', None),
                rsvd_rank_factor=getattr(ctm_args, 'b200_rsvd_rank_factor', None),
                rsvd_tol=getattr(ctm_args, 'b200_rsvd_tol', None))


def double_layer(eng, ket, bra=None):
    """A[(e,a),(f,b),(g,c),(h,d)] = sum_m ket[m,e,f,g,h] conj(bra)[m,a,b,c,d]  (ctmrg.py:51-61, :137-147),
    contracted by libctmb straight into the interleaved layout (the reference materialises einsum + contiguous)."""
    bra = ket if bra is None else bra
    d = ket.shape
    A = eng.einsum2('mefgh,mabcd->eafbgchd', ket, bra, conjB=True)
    return A.view(d[1] * bra.shape[1], d[2] * bra.shape[2], d[3] * bra.shape[3], d[4] * bra.shape[4])


def ctm_MOVE(direction, state, env, ctm_args=cfg.ctm_args, global_args=cfg.global_args,
             verbosity=0, diagnostics=None):
    r"""
    :param direction: one of Up=(0,-1), Left=(-1,0), Down=(0,1), Right=(1,0)
    :param state: wavefunction (``sites``, ``vertexToSite``); on-site tensors of rank 5 (single-layer
                  a[s,u,l,d,r]) or rank 4 (double-layer A[u,l,d,r], as built under ``ctm_force_dl``)
    :param env: environment (``chi``, ``C``, ``T``); entries at coord-direction are replaced
    Executes a single directional CTM move (projectors for all sites, then absorption,
    truncation and normalisation for all sites) on the GPU.
    """
    eng = _engine()
    if direction not in ((0, -1), (-1, 0), (0, 1), (1, 0)):
        raise ValueError("Invalid direction: " + str(direction))
    opts = _options(ctm_args)
    from ... import ad
    if ad.needs_grad(list(state.sites.values()) + list(env.C.values()) + list(env.T.values())):
        # reverse-mode AD (optim_*.py): the same move from differentiable libctmb calls (peps_torch_b200/ad.py)
        ad.ctm_move_generic(eng, direction, state, env, ctm_args)
        return
    eng.move_generic(direction, state, env, **opts)


def _sync(dev):
    if isinstance(dev, torch.device) and dev.type == 'cuda':
        torch.cuda.synchronize(dev)


def run(state, env, conv_check=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args):
    r"""
    Executes directional CTM for a generic iPEPS starting from ``env``; same contract as the
    reference: returns ``(env, history, t_ctm, t_obs)`` and mutates ``env`` in place.
    ``t_ctm`` is bracketed by stream synchronisation.
    """
    eng = _engine()
    dev = eng.device

    # 0) double-layer tensors on request (ctmrg.py:51-61)
    first = next(iter(state.sites.values()))
    if not getattr(ctm_args, 'ctm_force_dl', False) or first.dim() == 4:
        stateDL = state
    else:
        stateDL = IPEPS({coord: double_layer(eng, A) for coord, A in state.sites.items()},
                        vertexToSite=state.vertexToSite, lX=state.lX, lY=state.lY)

    def _ctmrg_iter(loc_ctm_args=ctm_args):
        for direction in loc_ctm_args.ctm_move_sequence:
            reps = stateDL.lX if direction in [(-1, 0), (1, 0)] else stateDL.lY
            for _ in range(reps):
                ctm_MOVE(direction, stateDL, env, ctm_args=loc_ctm_args, global_args=global_args)

    t_obs = t_ctm = 0.
    history = None

    # 1.1) warm-up (ctmrg.py:76-86)
    if getattr(ctm_args, 'ctm_warmup_iter', -1) >= 0:
        warmup_ctm_args = copy.deepcopy(ctm_args)
        warmup_ctm_args.projector_svd_method = getattr(ctm_args, 'warmup_projector_svd_method',
                                                       ctm_args.projector_svd_method)
        if hasattr(state, 'get_aux_bond_dims'):
            maxD = max(state.get_aux_bond_dims())       # as the reference (ctmrg.py:80); for rank-4 sites these are D^2
        else:
            maxD = max(max(t.shape[1:]) if t.dim() == 5 else max(t.shape) for t in state.sites.values())
        for i in range(max(ctm_args.ctm_warmup_iter, ceil(env.chi / maxD ** 2))):
            _sync(dev)
            t0_ctm = time.perf_counter()
            _ctmrg_iter(loc_ctm_args=warmup_ctm_args)
            _sync(dev)
            t_ctm += time.perf_counter() - t0_ctm

    # 1.2) main loop
    for i in range(ctm_args.ctm_max_iter):
        _sync(dev)
        t0_ctm = time.perf_counter()
        _ctmrg_iter()
        _sync(dev)
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


def run_overlap(state1, state2, env, conv_check=None, ctm_args=cfg.ctm_args, global_args=cfg.global_args):
    r"""
    CTM of the overlap network <state2|state1> (ctmrg.py:112-175): the double-layer tensor of every site is
    ket = state1, bra = state2; one move per direction and iteration, as in the reference.
    ``conv_check(state1, state2, env, history, ctm_args=...)``.
    """
    assert ctm_args.ctm_force_dl, 'ctmrg for wavefunction overlap requires use of double-layer routines'
    eng = _engine()
    dev = eng.device
    sitesDL = {coord: double_layer(eng, state1.site(coord), state2.site(coord)) for coord in state1.sites.keys()}
    stateDL = IPEPS(sitesDL, vertexToSite=state1.vertexToSite, lX=getattr(state1, 'lX', None), lY=getattr(state1, 'lY', None))

    t_obs = t_ctm = 0.
    history = None
    for i in range(ctm_args.ctm_max_iter):
        _sync(dev)
        t0_ctm = time.perf_counter()
        for direction in ctm_args.ctm_move_sequence:
            ctm_MOVE(direction, stateDL, env, ctm_args=ctm_args, global_args=global_args)
        _sync(dev)
        t1_ctm = time.perf_counter()

        t0_obs = time.perf_counter()
        if conv_check is not None:
            converged, history = conv_check(state1, state2, env, history, ctm_args=ctm_args)
            if converged:
                if getattr(ctm_args, 'verbosity_ctm_convergence', 0) > 0:
                    print(f"CTMRG  converged at iter= {i}, history= {history[-1]}")
                break
        t1_obs = time.perf_counter()
        t_ctm += t1_ctm - t0_ctm
        t_obs += t1_obs - t0_obs
    return env, history, t_ctm, t_obs
