"""Ad-hoc edge-shape sweep on a GPU box: a few CTM iterations of unusual unit cells / dimensions through the drop-in move against
the oracle; one line per case, never raises.  python tools/edge_shapes.py [cpu]  (cpu: oracle as engine, checks the script)."""
import os
import sys
import traceback
from collections import OrderedDict
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests'), ROOT):
    sys.path.insert(0, p)
import ctm_oracle as orc          # noqa: E402
import helpers as H               # noqa: E402
from peps_torch_b200.ipeps import IPEPS                 # noqa: E402
from peps_torch_b200.env import ENV, init_env           # noqa: E402
from peps_torch_b200.ctm.generic import ctmrg           # noqa: E402

cpu = len(sys.argv) > 1 and sys.argv[1] == 'cpu'
dev = 'cpu' if cpu else 'cuda:0'
if cpu:
    e = H.OracleEngine()
    ctmrg._engine = lambda: e

CASES = [  # name, lX, lY, p, (Du, Dl, Dd, Dr), chi, dtype, iterations
    ('D1_product', 1, 1, 2, (1, 1, 1, 1), 4, torch.float64, 2),
    ('D1_2x2', 2, 2, 2, (1, 1, 1, 1), 3, torch.complex128, 2),
    ('p1', 2, 2, 1, (2, 2, 2, 2), 8, torch.float64, 2),
    ('p3_1x2', 1, 2, 3, (2, 2, 2, 2), 9, torch.float64, 3),
    ('2x1_c128', 2, 1, 2, (2, 3, 2, 3), 10, torch.complex128, 3),
    ('chi_odd_5x1', 5, 1, 2, (2, 2, 2, 2), 7, torch.float64, 2),
    ('chi_gt_rank', 1, 1, 2, (2, 2, 2, 2), 40, torch.float64, 3),
    ('c128_D2_chi64_k112', 2, 2, 2, (2, 2, 2, 2), 64, torch.complex128, 4),
    ('c128_D2_chi61_k106', 1, 1, 2, (2, 2, 2, 2), 61, torch.complex128, 4),
    ('f64_D2_chi73_k127', 1, 1, 2, (2, 2, 2, 2), 73, torch.float64, 4),
    ('f64_D3_chi74_k129', 1, 1, 2, (3, 3, 3, 3), 74, torch.float64, 3),
    ('D5_chi20', 1, 1, 2, (5, 5, 5, 5), 20, torch.float64, 2),
    ('D6_chi12_c128', 1, 1, 2, (6, 6, 6, 6), 12, torch.complex128, 2),
    ('D7_chi10', 2, 1, 2, (7, 7, 7, 7), 10, torch.float64, 1),
]

for name, lX, lY, p, (Du, Dl, Dd, Dr), chi, dt, iters in CASES:
    try:
        g = torch.Generator().manual_seed(len(name) * 7 + chi)
        sites = OrderedDict()
        for y in range(lY):
            for x in range(lX):
                a = torch.randn(p, Du, Dl, Dd, Dr, dtype=dt, generator=g)
                sites[(x, y)] = a / a.abs().max()
        v2s = (lambda lX, lY: (lambda c: (c[0] % lX, c[1] % lY)))(lX, lY)
        C, T = orc.init_env(sites, v2s, chi)
        orc.run(sites, v2s, lX, lY, C, T, chi, iters)
        st = IPEPS(H.to_dev(sites, dev), v2s, lX, lY)
        env = ENV(chi, st)
        init_env(st, env)
        for _ in range(iters):
            for d in orc.DIRECTIONS:
                for _r in range(lX if d in (orc.LEFT, orc.RIGHT) else lY):
                    ctmrg.ctm_MOVE(d, st, env)
        shapes_ok = all(tuple(env.T[k].shape) == tuple(T[k].shape) for k in T) and all(tuple(env.C[k].shape) == tuple(C[k].shape) for k in C)
        print(f'{name:24s} spectra {H.spectra_diff(env.C, C):.2e}  absCT {H.env_abs_diff(env.C, env.T, C, T):.2e}  shapes {shapes_ok}', flush=True)
    except Exception as ex:          # noqa: BLE001
        print(f'{name:24s} FAILED {type(ex).__name__}: {str(ex)[:300]}', flush=True)
        if cpu:
            traceback.print_exc()
