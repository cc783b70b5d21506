"""Kernel-level timings on the GPU box (CUDA events, warm, best of N): the tensor-contraction GEMM
against cuBLAS DGEMM, and the enlarged-corner chain at the sizes of BASELINE.json's configs.
Usage: python tools/bench_kernels.py [gemm] [corner] [c5]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
from peps_torch_b200.engine import CtmEngine

dev = torch.device('cuda:0')
eng = CtmEngine()
what = set(sys.argv[1:]) or {'gemm', 'corner'}


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = []
if 'gemm1' in what:      # one large GEMM, short enough for an `ncu --set full` capture
    n = 4096
    A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: eng.einsum2('ik,kj->ij', A, B), reps=2, warm=1)
    print(json.dumps(dict(kind='gemm1', n=n, ms=ms, tflops=2.0 * n ** 3 / ms / 1e9)), flush=True)
if 'gemmbig' in what:
    for n in (2048, 4096, 8192):
        A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
        for spec in ('ik,kj->ij', 'ki,kj->ij', 'ik,jk->ij'):
            ms = timeit(lambda: eng.einsum2(spec, A, B), reps=3, warm=1)
            err = float((eng.einsum2(spec, A, B) - torch.einsum(spec, A, B)).abs().max())
            r = dict(kind='gemmbig', n=n, spec=spec, ms=ms, tflops=2.0 * n ** 3 / ms / 1e9, maxerr=err)
            out.append(r); print(json.dumps(r), flush=True)
if 'gemm' in what:
    for dt in (torch.float64, torch.complex128):
        for n in ((432, 1536, 4096, 8192) if dt == torch.float64 else (432, 1536, 4096)):
            A = torch.randn(n, n, dtype=dt, device=dev); B = torch.randn(n, n, dtype=dt, device=dev)
            f = 2.0 * n ** 3 * (4 if dt.is_complex else 1)
            for spec in ('ik,kj->ij', 'ki,kj->ij', 'ik,jk->ij'):
                ms = timeit(lambda: eng.einsum2(spec, A, B))
                ref = timeit(lambda: torch.einsum(spec, A, B))
                r = dict(kind='gemm', dtype=str(dt), n=n, spec=spec, ms=ms, tflops=f / ms / 1e9, cublas_ms=ref, cublas_tflops=f / ref / 1e9)
                out.append(r); print(json.dumps(r), flush=True)
            # skinny: n x n times n x 2chi
            k = max(16, n // 8)
            Bk = torch.randn(k, n, dtype=dt, device=dev)
            f = 2.0 * n * n * k * (4 if dt.is_complex else 1)
            ms = timeit(lambda: eng.einsum2('ij,sj->si', A, Bk))
            ref = timeit(lambda: torch.einsum('ij,sj->si', A, Bk))
            r = dict(kind='skinny', dtype=str(dt), n=n, k=k, ms=ms, tflops=f / ms / 1e9, cublas_ms=ref, cublas_tflops=f / ref / 1e9)
            out.append(r); print(json.dumps(r), flush=True)

if 'corner' in what or 'c5' in what:
    cases = []
    if 'corner' in what:
        cases += [(3, 48, 2, torch.float64), (4, 96, 2, torch.complex128), (3, 64, 8, torch.float64), (6, 144, 2, torch.float64)]
    if 'c5' in what:
        cases += [(8, 256, 2, torch.float64)]
    for D, chi, p, dt in cases:
        d = D * D
        a = torch.randn(p, D, D, D, D, dtype=dt, device=dev)
        C = torch.randn(chi, chi, dtype=dt, device=dev)
        T1 = torch.randn(chi, d, chi, dtype=dt, device=dev)
        T2 = torch.randn(chi, chi, d, dtype=dt, device=dev)
        Fc = (2 * chi ** 3 * d + 2 * chi ** 3 * d ** 2 + 4 * p * chi ** 2 * D ** 6) * (4 if dt.is_complex else 1)
        ms = timeit(lambda: eng.c2x2('LU', C, T1, T2, a, chi), reps=3, warm=1)
        r = dict(kind='c2x2_LU', D=D, chi=chi, p=p, dtype=str(dt), ms=ms, tflops=Fc / ms / 1e9, out_GB=(chi * d) ** 2 * a.element_size() / 1e9)
        out.append(r); print(json.dumps(r), flush=True)

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', 'bench_kernels.json'), 'a') as f:
    for r in out:
        f.write(json.dumps(r) + '\n')
