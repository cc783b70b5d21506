"""Mirror of the hot-path part of the reference's config singletons (config.py:369-409,
507-511).  When the reference's own `config` module is importable (drop-in use through
peps_torch_b200.run), its singletons are used instead, so that --CTMARGS_* flags keep working."""
import torch


class CTMARGS:
    def __init__(self):
        self.ctm_max_iter = 50
        self.ctm_warmup_iter = -1
        self.ctm_conv_tol = 1.0e-8
        self.ctm_env_init_type = 'CTMRG'
        self.ctm_absorb_normalization = 'inf'
        self.projector_method = '4X4'
        self.projector_svd_method = 'DEFAULT'
        self.warmup_projector_svd_method = self.projector_svd_method
        self.projector_svd_reltol = 1.0e-8
        self.projector_eps_multiplet = 1.0e-8
        self.projector_multiplet_abstol = 1.0e-14
        self.projector_rsvd_niter = 2
        self.ctm_move_sequence = [(0, -1), (-1, 0), (0, 1), (1, 0)]
        self.ctm_force_dl = False
        self.verbosity_ctm_convergence = 0
        self.fpcm_init_iter = 1
        self.fpcm_freq = -1
        # reverse-mode AD through the move (peps_torch_b200/ad.py; config.py:391,402-407 of the reference)
        self.ad_decomp_reg = 1.0e-12
        self.fwd_checkpoint_move = False
        # engine-specific (no counterpart in the reference)
        self.b200_rsvd_niter = None       # None: library default (ctmb_default_options, include/ctmb.h)
        self.b200_rsvd_rank_factor = None  # None: library default; sketch width k = ceil(rank_factor * chi)
        self.b200_rsvd_tol = None         # None: library default 2e-15 (x sqrt(n)) residual bound; 0: fixed iteration count
        self.b200_rsvd_tol_c4v = None     # C4v eigen path: same default


class GLOBALARGS:
    def __init__(self):
        self.dtype = 'float64'
        self.torch_dtype = torch.float64
        self.device = 'cuda:0'
        self.offload_to_gpu = 'None'


try:                                    # the reference is on sys.path: share its singletons
    import config as _ref_cfg
    ctm_args, global_args = _ref_cfg.ctm_args, _ref_cfg.global_args
except Exception:                       # stand-alone use (tests, bench on the GPU box)
    ctm_args, global_args = CTMARGS(), GLOBALARGS()
