"""CPU, world_size 2, gloo: the per-site sharding logic of peps_torch_b200.dist (job partition,
the two all-gathers, environment bookkeeping) with the oracle as compute backend."""
import os
import sys
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import ctm_oracle as orc
import helpers as H


class OracleBackend:
    """CPU stand-in for CtmEngine exposing the three methods ShardedCtm uses."""

    def projector_shape(self, direction, state, env):
        a = next(iter(state.sites.values()))
        D = a.shape[1:]
        leg = {(0, -1): D[1], (-1, 0): D[0], (0, 1): D[3], (1, 0): D[2]}[direction]
        return env.chi * leg * leg, env.chi

    def _nT_shape(self, direction, a, chi):
        D = a.shape[1:]
        return {(0, -1): (chi, D[2] ** 2, chi), (-1, 0): (chi, chi, D[3] ** 2),
                (0, 1): (D[0] ** 2, chi, chi), (1, 0): (chi, D[1] ** 2, chi)}[direction]

    grp = None

    def set_group(self, group, rank, nranks):
        """Group mode: libctmb would split the range finder over the members; the oracle backend simply computes the
        whole job on every member (same result on all of them, which is what the sharding logic relies on)."""
        self.grp = (rank, nranks, dist.get_world_size(group))

    def move_generic_projectors(self, direction, state, env, jobs, **opt):
        coords = list(state.sites.keys())
        P, Pt = [], []
        for j in jobs:
            R, Rt = orc.halves(direction, coords[j], state.sites, state.vertexToSite, env.C, env.T)
            p, pt = orc.projectors_from_matrices(R, Rt, env.chi, orc.OracleArgs())
            P.append(p.contiguous()); Pt.append(pt.contiguous())
        return P, Pt

    def move_generic_absorb(self, direction, state, env, jobs, P_all, Pt_all, **opt):
        coords = list(state.sites.keys())
        P = {coords[i]: p for i, p in enumerate(P_all)}
        Pt = {coords[i]: p for i, p in enumerate(Pt_all)}
        out = []
        for j in jobs:
            c = coords[j]
            nC1, nC2, nT = orc.absorb(direction, c, state.sites, state.vertexToSite, env.C, env.T, P, Pt, orc.OracleArgs())
            dest = state.vertexToSite((c[0] - direction[0], c[1] - direction[1]))
            out.append((dest, nC1.contiguous(), nC2.contiguous(), nT.contiguous()))
        return out


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from peps_torch_b200.dist import ShardedCtm
        torch.set_num_threads(2)
        z, meta = H.load_golden('generic_4site_D2_chi8_B')
        sites = H.golden_sites(z)
        C, T = H.golden_env(z, 'mid_')
        st = H.State(sites, orc.v2s_4site, 2, 2)
        env = H.Env(meta['chi'], C, T)
        sh = ShardedCtm(OracleBackend())
        moves = sh.iteration(st, env)
        # single-process oracle on the same snapshot
        C2, T2 = H.golden_env(z, 'mid_')
        orc.ctm_iteration(sites, orc.v2s_4site, 2, 2, C2, T2, meta['chi'])
        worst = max([float((env.C[k] - C2[k]).abs().max()) for k in C2] + [float((env.T[k] - T2[k]).abs().max()) for k in T2])
        # every rank must hold the same replica bit for bit
        flat = torch.cat([env.C[k].reshape(-1) for k in sorted(env.C)] + [env.T[k].reshape(-1) for k in sorted(env.T)])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        same = all(torch.equal(other[0], o) for o in other)
        ret[rank] = (moves, worst, same)
    finally:
        dist.destroy_process_group()


def test_sharded_move_world2_gloo():
    from peps_torch_b200.dist import partition_jobs
    assert partition_jobs(4, 2) == [[0, 2], [1, 3]]
    assert partition_jobs(4, 8)[5] == [] and partition_jobs(1, 2) == [[0], []]
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        moves, worst, same = ret[r]
        assert moves == 8
        assert worst < 1e-12, worst          # same arithmetic as the single-process oracle, job order aside
        assert same


def _worker_group(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from peps_torch_b200.dist import ShardedCtm
        torch.set_num_threads(2)
        z, meta = H.load_golden('kagome_1site_D2_chi8_A')
        sites = H.golden_sites(z)
        v2s, lX, lY = H.v2s_for(sites)
        C, T = H.golden_env(z, 'mid_')
        st = H.State(sites, v2s, lX, lY)
        env = H.Env(meta['chi'], C, T)
        be = OracleBackend()
        sh = ShardedCtm(be)
        moves = sh.iteration(st, env)
        C2, T2 = H.golden_env(z, 'mid_')
        orc.ctm_iteration(sites, v2s, lX, lY, C2, T2, meta['chi'])
        worst = max([float((env.C[k] - C2[k]).abs().max()) for k in C2] + [float((env.T[k] - T2[k]).abs().max()) for k in T2])
        flat = torch.cat([env.C[k].reshape(-1) for k in sorted(env.C)] + [env.T[k].reshape(-1) for k in sorted(env.T)])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        ret[rank] = (moves, worst, all(torch.equal(other[0], o) for o in other), be.grp)
    finally:
        dist.destroy_process_group()


def test_group_mode_world2_one_site_gloo():
    """world >= 2 N: the ranks form groups that share a site job (SURVEY 8e, G = 2N); here N = 1, one group of two."""
    from peps_torch_b200.dist import group_layout
    assert group_layout(4, 8) == (2, [[0, 1], [2, 3], [4, 5], [6, 7]])
    assert group_layout(4, 4) == (1, None) and group_layout(4, 7) == (1, None) and group_layout(1, 2) == (2, [[0, 1]])
    assert group_layout(2, 8)[0] == 4
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_group, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        moves, worst, same, grp = ret[r]
        assert moves == 4 and worst < 1e-12 and same
        assert grp == (r, 2, 2)


def _worker_two_groups(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from peps_torch_b200.dist import ShardedCtm
        torch.set_num_threads(1)
        z, meta = H.load_golden('generic_4site_D2_chi8_B')
        # a 2-site cell (2x1) cut out of the 4-site fixture: sites (0,0) and (1,0), 2SITE tiling
        all_sites = H.golden_sites(z)
        from collections import OrderedDict
        sites = OrderedDict([((0, 0), all_sites[(0, 0)]), ((1, 0), all_sites[(1, 0)])])
        C, T = orc.init_env(sites, orc.v2s_2site, meta['chi'])
        st = H.State(sites, orc.v2s_2site, 2, 1)
        env = H.Env(meta['chi'], dict(C), dict(T))
        be = OracleBackend()
        sh = ShardedCtm(be)
        moves = sh.iteration(st, env)
        C2, T2 = dict(C), dict(T)
        orc.ctm_iteration(sites, orc.v2s_2site, 2, 1, C2, T2, meta['chi'])
        worst = max([float((env.C[k] - C2[k]).abs().max()) for k in C2] + [float((env.T[k] - T2[k]).abs().max()) for k in T2])
        flat = torch.cat([env.C[k].reshape(-1) for k in sorted(env.C)] + [env.T[k].reshape(-1) for k in sorted(env.T)])
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        ret[rank] = (moves, worst, all(torch.equal(other[0], o) for o in other), be.grp, sh._layout[2])
    finally:
        dist.destroy_process_group()


def test_group_mode_world4_two_sites_gloo():
    """Two sites on four ranks: two groups of two (ranks {0,1} share site 0, {2,3} site 1); every rank creates both
    process groups in the same order, only the leaders contribute to the exchanges."""
    world = 4
    port = 33700 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_two_groups, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        moves, worst, same, grp, mine = ret[r]
        assert moves == 6                       # 2(lX + lY) with lX = 2, lY = 1
        assert worst < 1e-12 and same
        assert grp == (r % 2, 2, 2) and mine == r // 2
