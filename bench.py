#!/usr/bin/env python
"""bench.py -- CTM moves/sec of the B200-native engine (and of the reference CPU path).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload for EVERY N (BASELINE.json: "CTM moves/sec (4SITE, D,chi named) at 1/2/4/8 B200", configs[4]):
J1-J2 generic 4SITE iPEPS, D=8, chi=256, float64 (n = chi D^2 = 16384), synthetic random state (family B: rand-0.5,
seed 123), environment from the 'CTMRG' initialisation.  One STEP = ONE ctm_MOVE (ctm/generic/ctmrg.py:179-319;
the directions cycle through the reference's sequence U,U,L,L,D,D,R,R): a move is 1.13e14 reference-algorithm FLOP and
takes seconds.  N > 1: ONE CTM run sharded per site over the ranks (N = 2, 4) and, on 8 GPUs, 4 site jobs x groups of 2
(peps_torch_b200/dist.py), `scaling: strong`.  `--config c1..c4` select the other BASELINE configs (a step is then one
full CTMRG iteration, N > 1 runs replicas); the c2 line with its live reference arm is recorded in profiles/.
metric = ctm_MOVE calls / s, timed on the device with CUDA events (conv_check excluded, as the reference's t_ctm).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
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
BIG = ('c5', 'c5s')             # a move takes seconds: a step is ONE ctm_MOVE


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
# reference arm / cpu baseline on the host cores: the UNMODIFIED reference through its own library API when a copy
# of it travelled with the repo (baseline/_ref, written by __graft_entry__.build() where /root/reference exists:
# kind "reference"), else the oracle port (kind "port").  Never reads /root/reference at run time.
# ------------------------------------------------------------------------------------------
REF_ROOT = os.path.join(ROOT, 'baseline', '_ref')


class RefArm:
    def __init__(self):
        self.kind = 'port'
        self.ref = None
        if os.path.isfile(os.path.join(REF_ROOT, 'ctm', 'generic', 'ctmrg.py')):
            try:
                self.ref = self._import_reference()
                self.kind = 'reference'
            except Exception as ex:                # an unimportable copy must not take the bench line down
                self.import_error = repr(ex)[:200]

    @staticmethod
    def _import_reference():
        sys.dont_write_bytecode = True
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        import contextlib
        import importlib
        cwd = os.getcwd()
        os.chdir('/tmp')                           # config.configure / logging write into cwd (config.py:129)
        try:
            with contextlib.redirect_stdout(sys.stderr):       # the reference prints warnings about optional back-ends at import
                m = {name: importlib.import_module(name) for name in
                     ('config', 'ipeps.ipeps', 'ipeps.ipeps_c4v', 'ctm.generic.env', 'ctm.generic.ctmrg',
                      'ctm.generic.ctm_components', 'ctm.one_site_c4v.env_c4v', 'ctm.one_site_c4v.ctmrg_c4v',
                      'linalg.custom_svd')}
        finally:
            os.chdir(cwd)
        return m

    def describe(self):
        return ('the unmodified reference (baseline/_ref) through its library API: ENV, init_env, ctmrg.run(conv_check=None), t_ctm'
                if self.kind == 'reference' else 'oracle/ctm_oracle.py, the port of the reference CPU path (torch CPU, LAPACK gesdd/syevd)')

    # ---- configs c1-c4: whole iterations -------------------------------------------------
    def _moves_per_s_once(self, cfg_name, steps, warmup):
        import ctm_oracle as orc
        kind, sites, v2s, lX, lY, chi = make_state(cfg_name)
        if self.kind == 'reference':
            m = self.ref
            cfg = m['config']
            dt = sites.dtype if kind == 'c4v' else next(iter(sites.values())).dtype
            cfg.global_args.torch_dtype = dt
            cfg.global_args.dtype = 'complex128' if dt.is_complex else 'float64'
            cfg.global_args.device = 'cpu'
            if kind == 'c4v':
                state = m['ipeps.ipeps_c4v'].IPEPS_C4V(sites)
                env = m['ctm.one_site_c4v.env_c4v'].ENV_C4V(chi, state)
                m['ctm.one_site_c4v.env_c4v'].init_env(state, env)
                run, mps = m['ctm.one_site_c4v.ctmrg_c4v'].run, 1
            else:
                state = m['ipeps.ipeps'].IPEPS(sites, vertexToSite=v2s, lX=lX, lY=lY)
                env = m['ctm.generic.env'].ENV(chi, state)
                m['ctm.generic.env'].init_env(state, env)
                run, mps = m['ctm.generic.ctmrg'].run, 2 * (lX + lY)
            if warmup > 0:
                cfg.ctm_args.ctm_max_iter = warmup
                run(state, env, conv_check=None)
            cfg.ctm_args.ctm_max_iter = steps
            _, _, t_ctm, _ = run(state, env, conv_check=None)
            return steps * mps / t_ctm, mps
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

    def moves_per_s(self, cfg_name, steps, warmup, repeats=3):
        """thread sweep {1, all host cores} (MKL threading can cost 20x on small problems), median
        of `repeats` repeats per thread count, best median reported."""
        ncpu = os.cpu_count() or 1
        best = None
        for th in sorted({1, ncpu}):
            torch.set_num_threads(th)
            vals, mps = [], 1
            for _ in range(repeats):
                v, mps = self._moves_per_s_once(cfg_name, steps, warmup)
                vals.append(v)
            med = statistics.median(vals)
            if best is None or med > best[0]:
                best = (med, th, mps, vals)
        v, th, mps, vals = best
        return {'value': v, 'unit': 'ctm_MOVE/s', 'cores': th, 'kind': self.kind, 'moves_per_step': mps,
                'sample': f'{warmup} warm-up + {steps} timed CTMRG iterations ({mps} ctm_MOVE each) of the same workload, median of '
                          f'{repeats} repeats ({", ".join(f"{x:.3g}" for x in vals)}), best of thread counts {{1,{ncpu}}} on {ncpu} host cores; '
                          + self.describe()}

    # ---- config c5: phases of one site, SVD extrapolated --------
    def c5_sample(self, cfg_name, repeats=1):
        """One reference ctm_MOVE at n = 16384 is four full 16384^2 LAPACK SVDs (tens of minutes per move).  Bounded
        sample, timed with the reference's own functions on the workload's own tensors: ONE enlarged corner at full size
        (c2x2_LU, ctm_components.py:314-434), ONE n x n x n product at full size (the halves and M = R^T Rt are three of
        these per site, ctm_components.py:55-75, ctm_projectors.py:260-263), and ONE truncated_svd_gesdd
        (custom_svd.py:38-101) at n/4 = 4096 scaled by 4^3 (gesdd is O(n^3)) -- the only extrapolated phase, stated as
        such.  Projector products and the absorption (< 1 % of the FLOPs) are left out, which favours the CPU."""
        import ctm_oracle as orc
        kind, sites, v2s, lX, lY, chi = make_state(cfg_name)
        nsites = len(sites)
        ncpu = os.cpu_count() or 1
        torch.set_num_threads(ncpu)
        D = next(iter(sites.values())).shape[1]
        n = chi * D * D
        if self.kind == 'reference':
            m = self.ref
            cfg = m['config']
            cfg.global_args.torch_dtype, cfg.global_args.dtype, cfg.global_args.device = torch.float64, 'float64', 'cpu'
            state = m['ipeps.ipeps'].IPEPS(sites, vertexToSite=v2s, lX=lX, lY=lY)
            env = m['ctm.generic.env'].ENV(chi, state)
            m['ctm.generic.env'].init_env(state, env)
            corner = lambda: m['ctm.generic.ctm_components'].c2x2_LU((0, 0), state, env, mode='sl')     # noqa: E731
            svd = lambda M: m['linalg.custom_svd'].truncated_svd_gesdd(M, chi // 4)                       # noqa: E731
        else:
            C, T = orc.init_env(sites, v2s, chi)
            corner = lambda: orc.corner_at('LU', (0, 0), sites, v2s, C, T)                                # noqa: E731
            svd = lambda M: orc.truncated_svd(M, chi // 4, 1e-8, 1e-14)                                   # noqa: E731
        g = torch.Generator().manual_seed(123)
        A = torch.rand(n, n, dtype=torch.float64, generator=g) - 0.5
        ns = n // 4
        Ms = (torch.rand(ns, ns, dtype=torch.float64, generator=g) - 0.5) * torch.logspace(0, -12, ns, dtype=torch.float64)
        tc, tg, ts = [], [], []
        for _ in range(repeats):
            t0 = time.perf_counter(); X = corner(); tc.append(time.perf_counter() - t0)
            assert X.shape == (n, n)
            t0 = time.perf_counter(); torch.mm(A.t(), X); tg.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); svd(Ms); ts.append(time.perf_counter() - t0)
        t_corner, t_gemm, t_svd_s = statistics.median(tc), statistics.median(tg), statistics.median(ts)
        t_svd = t_svd_s * (n / ns) ** 3
        per_site = 4 * t_corner + 3 * t_gemm + t_svd
        return {'value': 1.0 / (nsites * per_site), 'unit': 'ctm_MOVE/s', 'cores': ncpu, 'kind': self.kind, 'extrapolated': True,
                'phases_s': {'enlarged_corner_full_size': t_corner, 'gemm_n3_full_size': t_gemm,
                             f'svd_gesdd_n{ns}': t_svd_s, f'svd_gesdd_n{n}_extrapolated_x{int((n / ns) ** 3)}': t_svd},
                'seconds_per_move_extrapolated': nsites * per_site,
                'sample': f'phases of ONE site job timed once each on {ncpu} host cores (median of {repeats}): one enlarged corner at full '
                          f'size ({t_corner:.2f} s), one {n}^3 DGEMM ({t_gemm:.2f} s), one truncated_svd_gesdd at n = {ns} ({t_svd_s:.2f} s) '
                          f'scaled by {int((n / ns) ** 3)} to n = {n} (EXTRAPOLATED, O(n^3)); move = {nsites} sites x (4 corners + 3 GEMMs + 1 SVD), '
                          'projector products and absorption left out (BASELINE.md section 4.3); ' + self.describe()}

    def baseline(self, cfg_name, steps=3, warmup=1, repeats=3):
        if cfg_name in BIG:
            return self.c5_sample(cfg_name, repeats=1)
        return self.moves_per_s(cfg_name, steps, warmup, repeats)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    kind, D, chi, dt, fam, desc = CONFIGS[args.config]
    arm = RefArm()
    if args.config in BIG:
        base = arm.c5_sample(args.config, repeats=min(max(args.steps, 1), 3))
        mps = 1
        steps_done = min(max(args.steps, 1), 3)
    else:
        steps_done = max(1, min(args.steps, 20))
        base = arm.moves_per_s(args.config, steps_done, min(args.warmup, 1), repeats=3)
        mps = base['moves_per_step']
    v = base['value']
    per_step = f'{"1 ctm_MOVE_sl" if kind == "c4v" else str(mps) + " ctm_MOVE"} per step'
    line = {'impl': 'reference', 'metric': 'CTM moves/sec', 'value': v, 'unit': 'ctm_MOVE/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * mps / v, 'higher_is_better': True,
            'scaling': 'strong' if args.config in BIG else 'weak', 'vs_baseline': None,
            'dtype': 'c128' if dt == 'complex128' else 'f64', 'data': 'synthetic',
            'config': {'workload': desc + f' ({per_step})', 'family': fam, 'seed': 123, 'moves_per_step': mps},
            'cpu_baseline': base,
            'e2e': {'value': v, 'unit': 'ctm_MOVE/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
            'note': (f'the CPU arm timed {steps_done} bounded sample(s), not {args.steps} whole moves: one reference ctm_MOVE at n = 16384 takes '
                     'tens of minutes (four full 16384^2 gesdd); value is an explicit extrapolation from timed phases, see cpu_baseline.sample'
                     if args.config in BIG else f'{steps_done} timed iterations per repeat')}
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
    """Kernel-level FLOP-bound evidence: the enlarged corner at config-5 size (D=8, chi=256, float64: the north star's
    named kernel, F_c = 2.77e11 FLOP, 2.1 GB output) and an n x n x n contraction, timed alone with CUDA events (best of
    3, warm; inputs 2-4 GB >> L2)."""
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
        torch.cuda.empty_cache()
    except Exception as ex:          # never lose the bench line over the side measurement
        out['error'] = str(ex)[:200]
    return out


def run_ours(args):
    import torch.distributed as dist
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
    per_move = args.config in BIG

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
    # a few moves so that the timed environment is not the zero-padded initial one; for the per-move workloads one full
    # iteration (2(lX+lY) moves: every direction once per row / column), during which the residual-checked range finder also
    # settles its iteration count -- a failed downward probe costs an extra round, and with 10-20 timed moves on a rank that
    # holds ONE site job two of those are a 10 % swing of the reported value (measured: 739 vs 812 ms at N = 4)
    for _ in range(2 * (lX + lY) if per_move else max(args.warmup, 3)):
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
    if sharded is not None:
        sharded.comm_reset()
    ms = timed_region(resident_step, args.steps)
    launches, flops_exec = eng.counters()
    comm = sharded.comm_totals() if sharded is not None else None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # shard: ONE CTM run spread over the ranks; replicas: every rank runs its own CTM run
    mult = 1 if (shard or world == 1) else world
    value = mult * moves_per_step * args.steps / (ms * 1e-3)

    # ------------------------------------------------------------------ e2e (host buffers)
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
        one_step(s2, e2)                       # the public drop-in API: ctmrg.run(state, env) / ctm_MOVE(direction, state, env)
        for k, v in e2.C.items():
            hostC[k].copy_(v, non_blocking=True)
        for k, v in e2.T.items():
            hostT[k].copy_(v, non_blocking=True)

    for _ in range(min(args.warmup, 1 if per_move else 3)):
        e2e_step()
    ms_e2e = timed_region(e2e_step, args.steps)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = mult * moves_per_step * args.steps / (float(t.item()) * 1e-3)

    # ------------------------------------------------------------------ roofline of the dominant kernel
    # (all ranks take part: in shard mode a step contains collectives)
    prof_steps = max(1, min(args.steps, 4)) if per_move else args.steps
    eng.profile(True)
    eng.reset_counters()
    timed_region(resident_step, prof_steps)
    prof = eng.profile_totals()
    eng.profile(False)
    rsvd_checks, rsvd_missed, rsvd_worst = eng.rsvd_status()
    rsvd_calls, rsvd_iters = eng.rsvd_iterations()
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
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    # capture of the same config (profiles/traffic.json: {config: {class: {"bytes_per_launch": ..., "source": ...}}})
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:
        traffic_tab = {}
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    total_ms = sum(v['ms'] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k_: prof[k_]['ms'])
    g = prof['tc_gemm']
    share = {k_: round(v['ms'] / total_ms, 4) for k_, v in prof.items()}
    tr = traffic_tab.get(args.config, {}).get(dom, {})
    if dom == 'tc_gemm':
        ach = g['flops'] / (g['ms'] * 1e-3) / 1e12
        roof = {'kernel': 'tc_kernel_tma (TMA-fed) / tc_kernel_ws / tc_kernel (DMMA tensor-contraction GEMM: operator applications of '
                          'the range finder, corners, projector products, absorption)', 'bound': 'tensor', 'achieved': ach, 'peak': fp64_peak,
                'unit': 'TFLOP/s', 'frac': ach / fp64_peak, 'traffic': tr.get('bytes_per_launch'),
                'algorithmic_bytes_per_launch': g['bytes'] / max(1, g['launches'])}
    else:
        v = prof[dom]
        ach = v['bytes'] / (v['ms'] * 1e-3) / 1e9
        roof = {'kernel': {'qr': 'Householder QR of the range-finder sketches (qr.cu)',
                           'jacobi': 'jacobi_kernel (one-sided Jacobi SVD of the k x k factor in shared memory)',
                           'misc': 'misc kernels'}[dom], 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': ach / hbm_peak, 'traffic': tr.get('bytes_per_launch'),
                'note': 'this config is latency-bound (a few GFLOP per move); the FLOP-bound kernels of the path are reported under kernel_level'}
    if tr.get('source'):
        roof['traffic_source'] = tr['source']
    roof['peak_source'] = ('cuBLAS DGEMM 8192^3 measured in this run (FP64 tensor path = DMMA; tcgen05 has no f64 kind)' if roof['bound'] == 'tensor'
                           else ('MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s'))
    roof['avg_launch_us'] = 1e3 * prof[dom]['ms'] / max(1, prof[dom]['launches'])
    roof['time_share_by_kernel_class'] = share
    roof['gemm_tflops_in_step'] = g['flops'] / max(g['ms'], 1e-9) / 1e9
    roof['fp64_dgemm_peak_tflops'] = fp64_peak

    # CTMB_BENCH_FAST=1 (profiling runs under ncu): skip the side measurements that are not part of the timed step
    fast = bool(os.environ.get('CTMB_BENCH_FAST'))
    eng._ws = None
    torch.cuda.empty_cache()
    roof['kernel_level'] = None if fast else kernel_level(eng, dev, fp64_peak)
    F_move = algorithmic_flops_per_move(kind, D, chi, p_phys, cplx)
    if fast or world > 1:
        base = {'value': None, 'unit': 'ctm_MOVE/s', 'cores': os.cpu_count(), 'kind': 'port',
                'sample': 'skipped (CTMB_BENCH_FAST)' if fast else 'reported at N = 1 only (rank 0 host cores)'}
    else:
        base = RefArm().baseline(args.config)
    if world == 1:
        parallelism = 'single GPU'
    elif not shard:
        parallelism = f'{world} independent replicas (one CTM run per GPU)'
    elif sharded._layout is not None and sharded._layout[1] > 1:
        parallelism = (f'shard: {sharded._layout[0]} site job(s) x groups of {sharded._layout[1]} GPUs -- the n x n x k operator applications of the '
                       f'range finder are split by sketch columns inside a group (in-place NCCL all-gather of the slabs); NCCL all-gather of P/Pt '
                       f'and of the new C/T per move')
    else:
        parallelism = f'shard: per-site shard of ONE CTM run over {world} GPUs, NCCL all-gather of P/Pt and of the new C/T per move'
    line = {'metric': 'CTM moves/sec', 'value': value, 'unit': 'ctm_MOVE/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if (shard or (per_move and world == 1)) else 'weak', 'vs_baseline': None, 'dtype': 'c128' if cplx else 'f64',
            'data': 'synthetic',
            'config': {'workload': desc + f' ({"1 ctm_MOVE_sl" if kind == "c4v" else str(moves_per_step) + " ctm_MOVE"} per step)',
                       'family': fam, 'seed': 123, 'moves_per_step': moves_per_step, 'l2': 'flushed between timed steps (256 MiB write)',
                       'parallelism': parallelism,
                       'rsvd': {'rank_factor': eng.options.rsvd_rank_factor or 'auto (1.75 for n <= 1024, else 2.0)',
                                'niter_first_call': eng.options.rsvd_niter, 'iterations': 'residual-checked (rsvd_tol 2e-15 sqrt(n))',
                                'residual_checks': rsvd_checks, 'results_above_bound': rsvd_missed,
                                'power_iterations_per_decomposition': (rsvd_iters / rsvd_calls) if rsvd_calls else None}},
            'e2e': {'value': e2e_value, 'unit': 'ctm_MOVE/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': base,
            'flops': {'reference_algorithm_per_move': F_move, 'executed_per_move': flops_exec / (moves_per_step * prof_steps),
                      'move_level_frac_of_fp64_peak': F_move * value / world / (fp64_peak * 1e12)}}
    if comm is not None:
        moves = moves_per_step * args.steps
        line['collectives'] = {'nccl_bytes_received_per_move': comm['bytes'] / moves, 'ms_in_collectives_per_move': comm['ms'] / moves,
                               'share_of_step': comm['ms'] / ms if ms else None, 'by_kind_ms_per_move': {k_: v / moves for k_, v in comm['by_kind'].items()},
                               'note': 'CUDA events around every NCCL call on rank 0 (includes waiting for the slowest rank)'}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c5', choices=sorted(CONFIGS))
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
