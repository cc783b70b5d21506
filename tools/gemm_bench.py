"""Times libctmb's contraction GEMM on plain matrices (CUDA events, best of 5) for the four fast-direction pairs.
CTMB_GEMM_TMA=0 disables the TMA-fed kernel (tc_gemm_tma.cu) and falls back to the offset-table gather kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from peps_torch_b200.engine import default_engine
eng = default_engine()
dev = torch.device('cuda:0')
out = {}
for (M, N, K) in ((8192, 8192, 8192), (16384, 512, 16384), (16384, 256, 16384), (4096, 4096, 4096)):
    A = torch.randn(M, K, dtype=torch.float64, device=dev); B = torch.randn(K, N, dtype=torch.float64, device=dev)
    At, Bt = A.t().contiguous(), B.t().contiguous()
    for spec, X, Y in (('ik,kj->ij', A, B), ('ki,kj->ij', At, B), ('ik,jk->ij', A, Bt), ('ki,jk->ij', At, Bt)):
        eng.einsum2(spec, X, Y); torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.einsum2(spec, X, Y); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f'{M}x{N}x{K} {spec}'] = round(2.0 * M * N * K / best / 1e9, 2)
    if M == 8192:
        torch.matmul(A, B); torch.cuda.synchronize(); best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(A, B); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out['cublas 8192^3'] = round(2.0 * M * N * K / best / 1e9, 2)
    del A, B, At, Bt
print(json.dumps({'tma': os.environ.get('CTMB_GEMM_TMA', '1'), 'tflops': out}))
