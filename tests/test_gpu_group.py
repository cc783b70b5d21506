"""Intra-site group split (SURVEY 8e, G = 2N) on two GPUs: the range finder of ONE site job shared by a group of two
ranks over NCCL.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import os
import sys
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, ret):
    for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'oracle'), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import ctm_oracle as orc
    import helpers as H
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    import datetime
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    try:
        from collections import OrderedDict
        from peps_torch_b200 import _lib
        from peps_torch_b200.engine import CtmEngine
        from peps_torch_b200.dist import ShardedCtm
        eng = CtmEngine()
        sh = ShardedCtm(eng)
        out = {}
        # 1x1 cell, D=3, chi=24: n = 216 > 160 => randomised path with residual-checked rounds (k = 48 columns, 24 per rank)
        D, chi = 3, 24
        a = orc.random_state_4site(D, family='B')[(0, 0)]
        sites = OrderedDict({(0, 0): a})
        C0, T0 = orc.init_env(sites, orc.v2s_1site, chi)
        for _ in range(2):
            orc.ctm_iteration(sites, orc.v2s_1site, 1, 1, C0, T0, chi)
        st = H.State(H.to_dev(sites, dev), orc.v2s_1site, 1, 1)
        for mode in (1, 2):                       # 1: explicit M (this size), 2: matrix-free factored operator forced
            _lib.lib.ctmb_debug_set_matrix_free(mode)
            worst = 0.0
            for d in orc.DIRECTIONS:
                C, T = dict(C0), dict(T0)
                orc.ctm_move(d, sites, orc.v2s_1site, C, T, chi)
                env = H.Env(chi, H.to_dev(C0, dev), H.to_dev(T0, dev))
                sh.ctm_MOVE(d, st, env)
                worst = max(worst, H.env_abs_diff(env.C, env.T, C, T))
            out[f'worst_mode{mode}'] = worst
        _lib.lib.ctmb_debug_set_matrix_free(1)
        out['group_bytes'] = eng.group_bytes
        out['layout'] = (sh._layout[1], sh._layout[2])
        # every rank holds the same replica bit for bit
        flat = torch.cat([env.C[k].reshape(-1) for k in sorted(env.C)] + [env.T[k].reshape(-1) for k in sorted(env.T)])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        out['same'] = all(torch.equal(other[0], o) for o in other)
        out['err'] = repr(eng._group_error)
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_group_split_two_gpus_matches_oracle():
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        o = ret[r]
        assert o['err'] == 'None', o['err']
        assert o['layout'] == (2, 0)
        assert o['group_bytes'] > 0                 # slabs really travelled
        assert o['worst_mode1'] < 1e-8 and o['worst_mode2'] < 1e-8, o
        assert o['same']
