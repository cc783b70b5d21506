"""Average number of Jacobi sweeps per decomposition over a CTM run (diagnostic): python tools/jacobi_sweeps.py c1|c3|c2"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import torch
import bench
from peps_torch_b200 import _lib
from peps_torch_b200.engine import default_engine
from peps_torch_b200.ipeps import IPEPS, IPEPS_C4V
from peps_torch_b200.env import ENV, init_env, ENV_C4V, init_env_c4v
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c3'
kind, sites, v2s, lX, lY, chi = bench.make_state(cfg)
dev = torch.device('cuda:0')
eng = default_engine()
out = (C.c_ulonglong * 2)()
_lib.lib.ctmb_debug_jacobi_stats.argtypes = [C.POINTER(C.c_ulonglong)]
_lib.lib.ctmb_debug_jacobi_stats.restype = None
if kind == 'c4v':
    st = IPEPS_C4V(sites.to(dev)); env = ENV_C4V(chi, st); init_env_c4v(st, env)
    Cc, Tc = env.C[env.keyC], env.T[env.keyT]
    for it in range(16):
        _lib.lib.ctmb_debug_jacobi_stats(out)
        Cc, Tc, _ = eng.move_c4v(st.site(), Cc, Tc, chi)
        torch.cuda.synchronize()
        _lib.lib.ctmb_debug_jacobi_stats(out)
        print(it, 'sweeps', out[0], 'matrices', out[1])
else:
    st = IPEPS({c: t.to(dev) for c, t in sites.items()}, v2s, lX, lY); env = ENV(chi, st); init_env(st, env)
    for it in range(6):
        for d in [(0, -1), (-1, 0), (0, 1), (1, 0)]:
            _lib.lib.ctmb_debug_jacobi_stats(out)
            eng.move_generic(d, st, env); torch.cuda.synchronize()
            _lib.lib.ctmb_debug_jacobi_stats(out)
            print(it, d, 'sweeps', out[0], 'matrices', out[1])
