#!/usr/bin/env python
"""bench.py -- CTM moves/sec of the B200-native engine (and of the reference CPU path).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]

Workload (BASELINE.json configs[1], named in config.workload): J1-J2 generic 4SITE iPEPS,
D=3, chi=48, float64, synthetic random state (family B: rand-0.5, seed 123),
environment from the 'CTMRG' initialisation.  One STEP = one full CTMRG iteration
= 2(lX+lY) = 8 calls of ctm_MOVE (ctm/generic/ctmrg.py:63-69); metric = ctm_MOVE calls / s,
timed on the device with CUDA events (conv_check excluded, as the reference's t_ctm).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CONFIGS = {
    # name: (kind, D, chi, dtype, family, description)
    'c1': ('c4v', 2, 16, 'float64', 'A', 'J1-J2 one-site C4v D=2 chi=16 float64'),
    'c2': ('4site', 3, 48, 'float64', 'B', 'J1-J2 generic 4SITE D=3 chi=48 float64'),
    'c3': ('c4v', 4, 96, 'complex128', 'B', 'J1-J2 one-site C4v complex128 D=4 chi=96'),
    'c4': ('kagome', 3, 64, 'float64', 'A', 'Kagome spin-1/2 iPESS D=3 chi=64 float64 (p=8, generic engine)'),
    'c5': ('4site', 8, 256, 'float64', 'B', 'J1-J2 generic 4SITE D=8 chi=256 float64'),
    # one site of c5 (1x1 cell): the unit of the intra-site group split (G = 2N) that two GPUs can measure
    'c5s': ('1site', 8, 256, 'float64', 'B', 'generic 1SITE D=8 chi=256 float64 (one site job of config 5)'),
}
BIG = ('c5', 'c5s')             # a move takes seconds: a step is ONE ctm_MOVE, no CPU arm


def algorithmic_flops_per_move(kind, D, chi, p, cplx):
    """reference-algorithm real FLOPs of one move, decomposition excluded."""
    d = D * D
    n = chi * d
    Fc = 2 * chi ** 3 * d + 2 * chi ** 3 * d ** 2 + 4 * p * chi ** 2 * D ** 6
    f = 4.0 if cplx else 1.0
    if kind == 'c4v':
        return f * (Fc + 4 * chi ** 3 * d ** 2 + 4 * p * chi ** 2 * D ** 6)
    nsites = 4 if kind == '4site' else 1
    Fsite = 4 * Fc + 4 * n ** 3 + 2 * n ** 3 + 4 * n ** 2 * chi + (8 * chi ** 3 * d + 4 * chi ** 3 * d ** 2 + 4 * p * chi ** 2 * D ** 6)
    return f * nsites * Fsite


def make_state(cfg_name):
    import ctm_oracle as orc
    from collections import OrderedDict
    kind, D, chi, dt, fam, _ = CONFIGS[cfg_name]
    dtype = torch.complex128 if dt == 'complex128' else torch.float64
    if kind == '4site':
        return kind, orc.random_state_4site(D, family=fam, dtype=dtype), orc.v2s_4site, 2, 2, chi
    if kind == '1site':
        a = orc.random_state_4site(D, family=fam, dtype=dtype)[(0, 0)]
        return kind, OrderedDict({(0, 0): a}), orc.v2s_1site, 1, 1, chi
    if kind == 'kagome':
        return kind, OrderedDict({(0, 0): orc.random_state_kagome(D, family=fam, dtype=dtype)}), orc.v2s_1site, 1, 1, chi
    return kind, orc.random_state_c4v(D, family=fam, dtype=dtype), None, 1, 1, chi


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path on host cores
# ------------------------------------------------------------------------------------------
def cpu_moves_per_s(cfg_name, steps, warmup, threads):
    import ctm_oracle as orc
    kind, sites, v2s, lX, lY, chi = make_state(cfg_name)
    torch.set_num_threads(threads)
    if kind == 'c4v':
        a = sites
        C, T = orc.init_env_c4v(a, chi)
        for _ in range(warmup):
            C, T = orc.ctm_move_c4v(a, C, T, chi)
        t0 = time.perf_counter()
        for _ in range(steps):
            C, T = orc.ctm_move_c4v(a, C, T, chi)
        return steps / (time.perf_counter() - t0), 1
    C, T = orc.init_env(sites, v2s, chi)
    for _ in range(warmup):
        orc.ctm_iteration(sites, v2s, lX, lY, C, T, chi)
    t0 = time.perf_counter()
    moves = 0
    for _ in range(steps):
        moves += orc.ctm_iteration(sites, v2s, lX, lY, C, T, chi)
    return moves / (time.perf_counter() - t0), moves // steps


def cpu_baseline(cfg_name, steps=3, warmup=1):
    """Thread sweep {1, all host cores} (MKL threading can cost 20x on small problems,
    BASELINE.md section 5): report the best."""
    ncpu = os.cpu_count() or 1
    best = None
    for th in sorted({1, ncpu}):
        v, mps = cpu_moves_per_s(cfg_name, steps, warmup, th)
        if best is None or v > best[0]:
            best = (v, th, mps)
    return {'value': best[0], 'unit': 'ctm_MOVE/s', 'cores': best[1], 'kind': 'port',
            'sample': f'{warmup} warm-up + {steps} timed CTMRG iterations ({best[2]} ctm_MOVE each) of the same workload; '
                      f'oracle/ctm_oracle.py (torch CPU, LAPACK gesdd/syevd) on {ncpu} host cores, best of thread counts {{1,{ncpu}}}'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    kind, D, chi, dt, fam, desc = CONFIGS[args.config]
    ncpu = os.cpu_count() or 1
    if args.config in BIG:
        print(json.dumps({'impl': 'reference', 'unavailable': 'config c5 on CPU is four full 16384^2 LAPACK SVDs per ctm_MOVE '
                          '(~56 min per move, SURVEY.md section 6); the default config c2 has a live reference arm'}))
        return
    best = None
    for th in sorted({1, ncpu}):
        v, mps = cpu_moves_per_s(args.config, args.steps, min(args.warmup, 1), th)
        if best is None or v > best[0]:
            best = (v, th, mps)
    v, th, mps = best
    line = {'impl': 'reference', 'metric': 'CTM moves/sec', 'value': v, 'unit': 'ctm_MOVE/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': min(args.warmup, 1), 'ms_per_step': 1e3 * mps / v, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c128' if dt == 'complex128' else 'f64', 'data': 'synthetic',
            'config': {'workload': desc + f' ({"1 ctm_MOVE_sl" if kind == "c4v" else str(mps) + " ctm_MOVE"} per step)',
                       'family': fam, 'seed': 123, 'moves_per_step': mps},
            'cpu_baseline': {'value': v, 'unit': 'ctm_MOVE/s', 'cores': th, 'kind': 'port',
                             'sample': f'{args.steps} timed CTMRG iterations, oracle port (torch CPU) of the reference path'},
            'e2e': {'value': v, 'unit': 'ctm_MOVE/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def measure_fp64_peak(dev):
    """cuBLAS DGEMM 8192^3 (2N^3 flops), best of 5: the FP64 roofline denominator.
    MEASURED_PEAKS.json holds only bf16/HBM figures."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def kernel_level(eng, dev, fp64_peak):
    """FLOP-bound evidence next to the (latency-bound) c2 step: the enlarged corner at config-5 size
    (D=8, chi=256, float64: the north star's named kernel, F_c = 2.77e11 FLOP, 2.1 GB output) and the
    n x n x n contraction that carries 94 % of a config-5 move, timed alone with CUDA events (best of 3,
    warm; inputs 2-4 GB >> L2)."""
    out = {}

    def best_ms(fn, reps=3):
        fn(); torch.cuda.synchronize(dev)
        b = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
            b = min(b, e0.elapsed_time(e1))
        return b
    try:
        D, chi, p = 8, 256, 2
        d = D * D
        g = torch.Generator(device='cpu').manual_seed(1)
        a = torch.randn(p, D, D, D, D, dtype=torch.float64, generator=g).to(dev)
        C = torch.randn(chi, chi, dtype=torch.float64, generator=g).to(dev)
        T1 = torch.randn(chi, d, chi, dtype=torch.float64, generator=g).to(dev)
        T2 = torch.randn(chi, chi, d, dtype=torch.float64, generator=g).to(dev)
        Fc = 2 * chi ** 3 * d + 2 * chi ** 3 * d ** 2 + 4 * p * chi ** 2 * D ** 6
        ms = best_ms(lambda: eng.c2x2('LU', C, T1, T2, a, chi))
        out['enlarged_corner_D8_chi256'] = {'ms': ms, 'flops': Fc, 'tflops': Fc / ms / 1e9, 'frac_of_fp64_dgemm_peak': Fc / ms / 1e9 / fp64_peak}
        del a, C, T1, T2
        n = 8192
        A = torch.randn(n, n, dtype=torch.float64, device=dev)
        B = torch.randn(n, n, dtype=torch.float64, device=dev)
        ms = best_ms(lambda: eng.einsum2('ki,kj->ij', A, B))
        out['gemm_RtR_8192'] = {'ms': ms, 'flops': 2.0 * n ** 3, 'tflops': 2.0 * n ** 3 / ms / 1e9, 'frac_of_fp64_dgemm_peak': 2.0 * n ** 3 / ms / 1e9 / fp64_peak}
        del A, B
        eng._ws = None
        torch.cuda.empty_cache()
    except Exception as ex:          # never lose the bench line over the side measurement
        out['error'] = str(ex)[:200]
    return out


def run_ours(args):
    import torch.distributed as dist
    import ctm_oracle as orc
    from peps_torch_b200.engine import default_engine
    from peps_torch_b200.ipeps import IPEPS, IPEPS_C4V
    from peps_torch_b200.env import ENV, init_env, ENV_C4V, init_env_c4v
    from peps_torch_b200.ctm.generic import ctmrg
    from peps_torch_b200.ctm.one_site_c4v import ctmrg_c4v
    from peps_torch_b200 import config as pcfg

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if os.environ.get('CTMB_BENCH_SINGLE_DEVICE'):     # functional dry-run of the N>1 control flow on one GPU
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        backend = os.environ.get('CTMB_BENCH_BACKEND', 'nccl')
        if backend == 'nccl':
            dist.init_process_group('nccl', device_id=dev)
        else:
            dist.init_process_group(backend)
    eng = default_engine()
    kind, D, chi, dt, fam, desc = CONFIGS[args.config]
    shard = world > 1 and args.parallel == 'shard' and kind != 'c4v'
    sharded = None
    if shard:
        from peps_torch_b200.dist import ShardedCtm
        sharded = ShardedCtm(eng)
    kind, sites_cpu, v2s, lX, lY, chi = make_state(args.config)
    cplx = dt == 'complex128'
    ctm_args = pcfg.CTMARGS()
    ctm_args.ctm_max_iter = 1
    if args.rank_factor is not None:            # experiments: sketch width k = ceil(rank_factor * chi)
        eng.options.rsvd_rank_factor = args.rank_factor
    ctm_args.b200_rsvd_rank_factor = eng.options.rsvd_rank_factor if args.rank_factor is not None else None

    # pinned host copies (e2e) and device-resident state (value)
    if kind == 'c4v':
        a_host = sites_cpu.pin_memory()
        st = IPEPS_C4V(a_host.to(dev))
        env = ENV_C4V(chi, st)
        init_env_c4v(st, env)
        p_phys = a_host.shape[0]

        def one_step(state, e):
            ctmrg_c4v.run(state, e, ctm_args=ctm_args)
        moves_per_step = 1
    else:
        host_sites = {c: t.pin_memory() for c, t in sites_cpu.items()}
        st = IPEPS({c: t.to(dev) for c, t in host_sites.items()}, v2s, lX, lY)
        env = ENV(chi, st)
        init_env(st, env)
        p_phys = next(iter(sites_cpu.values())).shape[0]

        # config c5 (n = 16384): a move takes seconds, so a STEP is ONE ctm_MOVE (directions cycle through the
        # reference's sequence U,U,L,L,D,D,R,R) instead of a full iteration of eight
        per_move = args.config in BIG
        seq = [d for d in ctm_args.ctm_move_sequence for _ in range(lX if d in [(-1, 0), (1, 0)] else lY)]
        cursor = [0]

        def next_direction():
            d = seq[cursor[0] % len(seq)]
            cursor[0] += 1
            return d

        def one_step(state, e):
            if per_move:
                if shard:
                    sharded.ctm_MOVE(next_direction(), state, e)
                else:
                    ctmrg.ctm_MOVE(next_direction(), state, e, ctm_args=ctm_args)
            elif shard:
                sharded.iteration(state, e, ctm_args.ctm_move_sequence)
            else:
                ctmrg.run(state, e, ctm_args=ctm_args)
        moves_per_step = 1 if per_move else 2 * (lX + lY)
    # a few iterations so that the timed environment is not the zero-padded initial one
    for _ in range(1 if args.config in BIG else max(args.warmup, 3)):
        one_step(st, env)
    torch.cuda.synchronize(dev)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_region(step_fn, K):
        """K steps, each bracketed by CUDA events on the launching stream; the L2 is flushed
        between steps outside the event brackets. Returns total device milliseconds."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        for i in range(K):
            flush.fill_(i & 0xff)
            ev[i][0].record()
            step_fn()
            ev[i][1].record()
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev)

    # ------------------------------------------------------------------ value (inputs resident)
    def resident_step():
        if kind == 'c4v':
            ctmrg_c4v.ctm_MOVE_sl(st.site(), env, ctm_args=ctm_args)
        elif per_move:
            one_step(st, env)
        elif shard:
            sharded.iteration(st, env, ctm_args.ctm_move_sequence)
        else:
            for direction in ctm_args.ctm_move_sequence:
                for _ in range(lX if direction in [(-1, 0), (1, 0)] else lY):
                    ctmrg.ctm_MOVE(direction, st, env, ctm_args=ctm_args)

    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.reset_counters()
    ms = timed_region(resident_step, args.steps)
    launches, flops_exec = eng.counters()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # shard: ONE CTM run spread over the ranks; replicas: every rank runs its own CTM run
    mult = 1 if (shard or world == 1) else world
    value = mult * moves_per_step * args.steps / (ms * 1e-3)

    # ------------------------------------------------------------------ e2e (host buffers)
    if kind == 'c4v':
        hostC = {k: v.cpu().pin_memory() for k, v in env.C.items()}
        hostT = {k: v.cpu().pin_memory() for k, v in env.T.items()}
    else:
        hostC = {k: v.cpu().pin_memory() for k, v in env.C.items()}
        hostT = {k: v.cpu().pin_memory() for k, v in env.T.items()}
    h2d = sum(t_.numel() * t_.element_size() for t_ in list(hostC.values()) + list(hostT.values()))
    h2d += (a_host.numel() * a_host.element_size()) if kind == 'c4v' else sum(t_.numel() * t_.element_size() for t_ in host_sites.values())
    d2h = sum(t_.numel() * t_.element_size() for t_ in list(hostC.values()) + list(hostT.values()))

    def e2e_step():
        # host -> device: the state and its environment from pinned host memory
        if kind == 'c4v':
            s2 = IPEPS_C4V(a_host.to(dev, non_blocking=True))
            e2 = ENV_C4V(chi, s2)
        else:
            s2 = IPEPS({c: t_.to(dev, non_blocking=True) for c, t_ in host_sites.items()}, v2s, lX, lY)
            e2 = ENV(chi)
        e2.C = {k: v.to(dev, non_blocking=True) for k, v in hostC.items()}
        e2.T = {k: v.to(dev, non_blocking=True) for k, v in hostT.items()}
        one_step(s2, e2)                       # the public drop-in API: ctmrg.run(state, env)
        for k, v in e2.C.items():
            hostC[k].copy_(v, non_blocking=True)
        for k, v in e2.T.items():
            hostT[k].copy_(v, non_blocking=True)

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    ms_e2e = timed_region(e2e_step, args.steps)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = mult * moves_per_step * args.steps / (float(t.item()) * 1e-3)

    # ------------------------------------------------------------------ roofline of the dominant kernel
    # (all ranks take part: in shard mode a step contains collectives)
    eng.profile(True)
    eng.reset_counters()
    timed_region(resident_step, args.steps)
    prof = eng.profile_totals()
    eng.profile(False)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    fp64_peak = measure_fp64_peak(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    total_ms = sum(v['ms'] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k_: prof[k_]['ms'])
    g = prof['tc_gemm']
    share = {k_: round(v['ms'] / total_ms, 4) for k_, v in prof.items()}
    if dom == 'tc_gemm':
        ach = g['flops'] / (g['ms'] * 1e-3) / 1e12
        roof = {'kernel': 'tc_kernel / tc_kernel_ws (DMMA tensor-contraction GEMM)', 'bound': 'tensor', 'achieved': ach, 'peak': fp64_peak,
                'unit': 'TFLOP/s', 'frac': ach / fp64_peak, 'traffic': None}
    else:
        v = prof[dom]
        ach = v['bytes'] / (v['ms'] * 1e-3) / 1e9
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
        # (profiles/r1_c2_qr_tsolve_metrics.csv: qr_reg_kernel reads 1.37 MB = the four 432x96 sketches, writes stay in L2)
        traffic = {'qr': 1.370624e6 if args.config == 'c2' else None, 'jacobi': None, 'misc': None}[dom]
        roof = {'kernel': {'qr': 'qr_reg_kernel + wy_tsolve_kernel (Householder QR of the range-finder sketches, register-resident over a cluster of 8 CTAs per matrix; latency-bound: one cluster barrier per column)',
                           'jacobi': 'jacobi_kernel (one-sided Jacobi SVD of the k x k factor in shared memory; shared-memory-bandwidth bound)',
                           'misc': 'misc kernels'}[dom], 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': ach / hbm_peak, 'traffic': traffic,
                'note': 'the step is latency-bound (2.8 GFLOP per move); the FLOP-bound kernels of the path are reported under kernel_level; '
                        'traffic is the ncu capture of the k = 96 build (profiles/r1_c2_qr_tsolve_metrics.csv), the sketches are 84 columns wide since'}
    roof['peak_source'] = ('cuBLAS DGEMM 8192^3 measured in this run' if roof['bound'] == 'tensor'
                           else ('MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s'))
    roof['avg_launch_us'] = 1e3 * prof[dom]['ms'] / max(1, prof[dom]['launches'])
    roof['time_share_by_kernel_class'] = share
    roof['gemm_tflops_in_step'] = g['flops'] / max(g['ms'], 1e-9) / 1e9
    roof['fp64_dgemm_peak_tflops'] = fp64_peak

    # CTMB_BENCH_FAST=1 (profiling runs under ncu): skip the side measurements that are not part of the timed step
    fast = bool(os.environ.get('CTMB_BENCH_FAST'))
    roof['kernel_level'] = None if fast else kernel_level(eng, dev, fp64_peak)
    F_move = algorithmic_flops_per_move(kind, D, chi, p_phys, cplx)
    if args.config in BIG:
        base = {'value': None, 'unit': 'ctm_MOVE/s', 'cores': os.cpu_count(), 'kind': 'port',
                'sample': 'not run: one reference move at n = 16384 is four full 16384^2 LAPACK SVDs (~56 min per ctm_MOVE '
                          'extrapolated from DGEMM / gesdd timings, SURVEY.md section 6)'}
    elif fast:
        base = {'value': None, 'unit': 'ctm_MOVE/s', 'cores': os.cpu_count(), 'kind': 'port', 'sample': 'skipped (CTMB_BENCH_FAST)'}
    else:
        base = cpu_baseline(args.config)
    if world == 1:
        parallelism = 'single GPU'
    elif not shard:
        parallelism = f'{world} independent replicas (one CTM run per GPU)'
    elif sharded._layout is not None and sharded._layout[1] > 1:
        parallelism = (f'{sharded._layout[0]} site job(s) x groups of {sharded._layout[1]} GPUs: the n x n x k operator applications of the '
                       f'range finder are split by sketch columns inside a group (in-place NCCL all-gather of the slabs, '
                       f'{getattr(eng, "group_bytes", 0) / 1e6:.0f} MB received per rank in this run); NCCL all-gather of P/Pt and of '
                       f'the new C/T per move')
    else:
        parallelism = f'per-site shard over {world} GPUs, NCCL all-gather of P/Pt and of the new C/T per move'
    line = {'metric': 'CTM moves/sec', 'value': value, 'unit': 'ctm_MOVE/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if shard else 'weak', 'vs_baseline': None, 'dtype': 'c128' if cplx else 'f64', 'data': 'synthetic',
            'config': {'workload': desc + f' ({"1 ctm_MOVE_sl" if kind == "c4v" else str(moves_per_step) + " ctm_MOVE"} per step)',
                       'family': fam, 'seed': 123, 'moves_per_step': moves_per_step, 'l2': 'flushed between timed steps (256 MiB write)',
                       'parallelism': parallelism,
                       'rsvd': {'rank_factor': eng.options.rsvd_rank_factor or 'auto (1.75 for n <= 1024, else 2.0)',
                                'niter_first_call': eng.options.rsvd_niter, 'iterations': 'residual-checked (rsvd_tol 2e-15 sqrt(n))'}},
            'e2e': {'value': e2e_value, 'unit': 'ctm_MOVE/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': base,
            'flops': {'reference_algorithm_per_move': F_move, 'executed_per_move': flops_exec / (moves_per_step * args.steps),
                      'move_level_frac_of_fp64_peak': F_move * value / world / (fp64_peak * 1e12)}}
    if world > 1 and base.get('value') is not None:
        line['cpu_baseline']['note'] = 'measured on rank 0 host cores while the other ranks idle'
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--rank-factor', type=float, default=None, dest='rank_factor',
                    help='sketch width of the range finder in units of chi (default: the library default)')
    ap.add_argument('--parallel', default='auto', choices=['auto', 'shard', 'replicas'],
                    help='N>1: "shard" = per-site shard of ONE CTM run with NCCL all-gathers of P/Pt and C/T (strong scaling; '
                         'default for c5, whose move is FLOP-bound); "replicas" = one independent CTM run per GPU (weak scaling; '
                         'default for c1-c4, whose move time is set by the serial column steps of the decomposition and does not '
                         'drop when a rank holds one site instead of four)')
    args = ap.parse_args()
    if args.parallel == 'auto':
        args.parallel = 'shard' if args.config in BIG else 'replicas'
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
