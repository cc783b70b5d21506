"""Run a few CTMRG iterations of one bench config on the GPU and print the per-kernel-class device
times (CUDA events inside libctmb) and the Jacobi sweep statistics.  Short enough to sit under ncu.
Usage: python tools/run_cfg.py [c2] [iters] [key=value engine options ...]"""
import sys, os, json, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import bench
from peps_torch_b200.engine import default_engine
from peps_torch_b200 import _lib
from peps_torch_b200.ipeps import IPEPS, IPEPS_C4V
from peps_torch_b200.env import ENV, init_env, ENV_C4V, init_env_c4v

cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opts = {}
for kv in sys.argv[3:]:
    k, v = kv.split('=')
    opts[k] = float(v) if '.' in v or 'e' in v else int(v)
dev = torch.device('cuda:0')
eng = default_engine()
onemove = bool(opts.pop('onemove', 0))      # time single ctm_MOVE calls (config c5: seconds per move)
for k, v in opts.items():
    setattr(eng.options, k, v)
kind, sites, v2s, lX, lY, chi = bench.make_state(cfg)
if kind == 'c4v':
    st = IPEPS_C4V(sites.to(dev)); env = ENV_C4V(chi, st); init_env_c4v(st, env)
else:
    st = IPEPS({c: t.to(dev) for c, t in sites.items()}, v2s, lX, lY); env = ENV(chi, st); init_env(st, env)
DIRS = [(0, -1), (-1, 0), (0, 1), (1, 0)]


_dir = [0]


def iteration():
    if onemove and kind != 'c4v':
        eng.move_generic(DIRS[_dir[0] % 4], st, env); _dir[0] += 1
        return 1
    if kind == 'c4v':
        nC, nT, _ = eng.move_c4v(st.site(), env.C[env.keyC], env.T[env.keyT], chi)
        env.C[env.keyC], env.T[env.keyT] = nC, nT
        return 1
    n = 0
    for d in DIRS:
        for _ in range(lX if d in ((-1, 0), (1, 0)) else lY):
            eng.move_generic(d, st, env); n += 1
    return n


iteration(); torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
plain_moves = sum(iteration() for _ in range(iters))
t1.record(); torch.cuda.synchronize()
plain_ms = t0.elapsed_time(t1) / plain_moves
stats = (ctypes.c_ulonglong * 2)()
_lib.lib.ctmb_debug_jacobi_stats(stats)
eng.profile(True); eng.reset_counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
moves = sum(iteration() for _ in range(iters))
e1.record(); torch.cuda.synchronize()
prof = eng.profile_totals()
_lib.lib.ctmb_debug_jacobi_stats(stats)
out = {'config': cfg, 'moves': moves, 'ms_per_move': plain_ms, 'ms_per_move_profiled': e0.elapsed_time(e1) / moves, 'launches_per_move': eng.counters()[0] / moves,
       'per_move_ms_by_class': {k: round(v['ms'] / moves, 4) for k, v in prof.items()},
       'launches_by_class': {k: v['launches'] / moves for k, v in prof.items()},
       'gemm_tflops': prof['tc_gemm']['flops'] / max(prof['tc_gemm']['ms'], 1e-9) / 1e9,
       'jacobi_sweeps_avg': stats[0] / max(1, stats[1]), 'jacobi_matrices': int(stats[1]), 'options': opts}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
open(os.path.join(ROOT, 'gpurun_out', 'run_cfg.jsonl'), 'a').write(json.dumps(out) + '\n')
