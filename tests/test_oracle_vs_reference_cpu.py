"""Oracle restatements that have no committed fixture are pinned here against the reference itself.
Needs the reference tree (build container only): skipped on the GPU box."""
import os
import sys
import pytest
import torch

REF = os.environ.get('PEPS_TORCH_REF', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'ctm', 'generic')), reason='reference tree not present')


def test_oracle_4x2_projector_move_matches_reference(tmp_path):
    import ctm_oracle as orc
    import helpers as H
    cwd = os.getcwd()
    os.chdir(tmp_path)                       # config.configure may write log files into cwd
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    try:
        import config as cfg
        from ipeps.ipeps import IPEPS
        from ctm.generic.env import ENV
        from ctm.generic import ctmrg
        z, meta = H.load_golden('generic_4site_D3_chi12_B')
        chi = meta['chi']
        sites = H.golden_sites(z)
        v2s, lX, lY = H.v2s_for(sites)
        C0, T0 = H.golden_env(z, 'mid_')
        cfg.global_args.dtype, cfg.global_args.device = 'float64', 'cpu'
        state = IPEPS(sites={c: t.clone() for c, t in sites.items()}, vertexToSite=v2s, lX=lX, lY=lY)
        old = cfg.ctm_args.projector_method
        cfg.ctm_args.projector_method = '4X2'
        try:
            for d in orc.DIRECTIONS:
                env = ENV(chi, state)
                env.C = {k: v.clone() for k, v in C0.items()}
                env.T = {k: v.clone() for k, v in T0.items()}
                ctmrg.ctm_MOVE(d, state, env, ctm_args=cfg.ctm_args, global_args=cfg.global_args)
                C, T = dict(C0), dict(T0)
                orc.ctm_move(d, sites, v2s, C, T, chi, orc.OracleArgs(projector_method='4X2'))
                # |.|: fix_svd_signs can flip a column between two evaluations that differ by 1e-14
                worst = max([float((env.C[k].abs() - C[k].abs()).abs().max()) for k in C]
                            + [float((env.T[k].abs() - T[k].abs()).abs().max()) for k in T])
                assert worst < 1e-11, (d, worst)
        finally:
            cfg.ctm_args.projector_method = old
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
