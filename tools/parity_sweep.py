"""Parity of the range-finder parameters at config-c2 size (4SITE D=3 chi=48): for each (rank_factor, niter)
run a few CTMRG iterations on the GPU and compare with the CPU oracle on the same inputs.
Usage: python tools/parity_sweep.py [iters]"""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import ctm_oracle as orc
import helpers as H
from peps_torch_b200.engine import default_engine
from peps_torch_b200.ipeps import IPEPS
from peps_torch_b200.env import ENV, init_env

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda:0')
eng = default_engine()
if 'SWEEP_TOL' in os.environ:                 # default: the library's residual bound (2e-15 * sqrt(n)); 0 = fixed iteration count
    eng.options.rsvd_tol = float(os.environ['SWEEP_TOL'])
COMBOS = ((2.0, 4), (1.75, 4), (1.5, 4)) if eng.options.rsvd_tol > 0 else ((2.0, 4), (2.0, 3), (1.75, 4), (1.75, 3), (1.625, 4), (1.5, 5))
for D, chi, family in ((3, 48, 'B'), (3, 48, 'A'), (4, 40, 'B'), (4, 64, 'B')):
    sites = orc.random_state_4site(D, family=family)
    C, T = orc.init_env(sites, orc.v2s_4site, chi)
    orc.run(sites, orc.v2s_4site, 2, 2, C, T, chi, iters)
    e_cpu = orc.energy_j1j2(sites, orc.v2s_4site, C, T, 1.0, 0.3)
    st = IPEPS(H.to_dev(sites, dev), orc.v2s_4site, 2, 2)
    for rf, q in COMBOS:
        eng.options.rsvd_rank_factor = rf
        eng.options.rsvd_niter = q
        env = ENV(chi, st)
        init_env(st, env)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            for d in orc.DIRECTIONS:
                for _r in range(2):
                    eng.move_generic(d, st, env)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (8 * iters)
        e_gpu = orc.energy_j1j2(sites, orc.v2s_4site, {k: v.cpu() for k, v in env.C.items()}, {k: v.cpu() for k, v in env.T.items()}, 1.0, 0.3)
        print(json.dumps({'D': D, 'chi': chi, 'family': family, 'rank_factor': rf, 'niter': q, 'ms_per_move': 1e3 * dt,
                          'spectra_diff': H.spectra_diff(env.C, C), 'absCT_diff': H.env_abs_diff(env.C, env.T, C, T),
                          'energy_rel': abs(e_gpu - e_cpu) / abs(e_cpu)}), flush=True)
