"""Sharding of one CTM move over the GPUs of a node (one process per GPU).

Inside ctm_MOVE all N projector pairs are computed from the same environment snapshot, then
all N absorptions from that snapshot plus the projectors, and results are written only
afterwards (ctm/generic/ctmrg.py:246-275,313-319).

* world <= N: rank r owns the site jobs r, r+W, r+2W, ...
* world >= 2N (config 5 on 8 GPUs: N = 4): the ranks form N GROUPS of g = world // N members; a group shares the
  projector job of one site -- libctmb splits the n x n x k operator applications of the range finder by sketch
  columns over the group and all-gathers the slabs in place over NVLink (engine.set_group / ctmb_set_group), so
  every member ends with identical projectors; the group's first rank (leader) does the absorption.  Ranks beyond
  g*N stay idle.

Either way the move has two exchange steps over ALL ranks:

  1. all-gather of (P, Pt) of every job   -- the absorption at `coord` needs the projectors of
     the neighbouring site coord+shift (ctmrg.py:326-334), which another rank may own;
  2. all-gather of the new (nC1, nC2, nT) -- every rank keeps a full replica of the (small)
     environment, as the next move reads all of it.

Collectives go through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
The compute backend is anything with the methods of CtmEngine used below, so the same
code is exercised on CPU with the oracle as backend (tests/test_dist_cpu.py).
"""
import time
import torch
import torch.distributed as dist
from .engine import OUT_KEYS, DIRECTIONS


def partition_jobs(nsites, world):
    """Job (site index) lists per rank: round-robin, ranks beyond nsites stay idle."""
    return [list(range(r, nsites, world)) for r in range(world)]


def group_layout(nsites, world):
    """(members per group g, [ranks of group j]) when the ranks outnumber the sites at least 2:1, else (1, None)."""
    g = world // nsites if nsites > 0 else 1
    if g < 2:
        return 1, None
    return g, [list(range(j * g, (j + 1) * g)) for j in range(nsites)]


class ShardedCtm:
    def __init__(self, backend, group=None):
        self.backend = backend
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._layout = None         # (nsites, g, my group index or None, parts)
        self._comm = []             # (kind, bytes received, start, end): CUDA events on the current stream (CPU: seconds)

    # ---- accounting of the exchange steps (bench.py: NCCL bytes and time in collectives per move) ----
    def comm_reset(self):
        self._comm = []
        if hasattr(self.backend, 'group_reset'):
            self.backend.group_reset()

    def comm_totals(self):
        """Synchronises the recorded events.  {'bytes', 'ms', 'by_kind': {kind: ms}}; the slab gathers of the intra-site
        group split are reported by the backend (CtmEngine.group_totals)."""
        tot = {'bytes': 0, 'ms': 0.0, 'by_kind': {}}
        recs = list(self._comm)
        if hasattr(self.backend, 'group_totals'):
            recs += self.backend.group_totals()
        for kind, nbytes, a, b in recs:
            if isinstance(a, float):
                ms = 1e3 * (b - a)
            else:
                b.synchronize()
                ms = a.elapsed_time(b)
            tot['bytes'] += nbytes
            tot['ms'] += ms
            tot['by_kind'][kind] = tot['by_kind'].get(kind, 0.0) + ms
        return tot

    def _all_gather(self, flat, counts, kind='gather'):
        """all-gather of per-rank 1-D tensors: one in-place NCCL all-gather into a [world, m] buffer (ragged counts are
        padded to the maximum m; with equal counts -- the per-site shard with N | 4 -- nothing is padded or copied twice)."""
        m = max(counts)
        whole = flat.new_empty((self.world, m))
        mine = whole[self.rank]
        mine[:flat.numel()] = flat
        cuda = flat.is_cuda
        if cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        else:
            a = time.perf_counter()
        if cuda:
            dist.all_gather_into_tensor(whole, mine, group=self.group)
        else:                       # gloo (CPU tests) has no all_gather_into_tensor
            out = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(out, mine.clone(), group=self.group)
            for r, o in enumerate(out):
                whole[r] = o
        if cuda:
            b.record()
        else:
            b = time.perf_counter()
        self._comm.append((kind, flat.element_size() * (sum(counts) - counts[self.rank]), a, b))
        return [whole[r, :c] for r, c in enumerate(counts)]

    def _setup(self, nsites):
        """Decide the layout for `nsites` jobs once; in group mode create the process groups (a collective call: every
        rank creates every group, in the same order) and hand this rank's group to the backend."""
        if self._layout is not None and self._layout[0] == nsites:
            return self._layout
        g, groups = group_layout(nsites, self.world)
        if groups is None:
            parts = partition_jobs(nsites, self.world)
            self._layout = (nsites, 1, None, parts, parts)
            return self._layout
        base = list(range(self.world)) if self.group is None else dist.get_process_group_ranks(self.group)
        mine, my_pg = None, None
        for j, members in enumerate(groups):
            pg = dist.new_group([base[r] for r in members])
            if self.rank in members:
                mine, my_pg = j, pg
        if mine is not None:
            self.backend.set_group(my_pg, groups[mine].index(self.rank), g)
        # compute: every member of group j runs job j; exchange: only the leader contributes
        compute = [[r // g] if r < g * nsites else [] for r in range(self.world)]
        contribute = [[r // g] if (r < g * nsites and r % g == 0) else [] for r in range(self.world)]
        self._layout = (nsites, g, mine, compute, contribute)
        return self._layout

    def ctm_MOVE(self, direction, state, env, **opt):
        """Same contract as ctm.generic.ctmrg.ctm_MOVE; every rank ends with the same env."""
        if direction not in DIRECTIONS:
            raise ValueError("Invalid direction: " + str(direction))
        coords = list(state.sites.keys())
        n = len(coords)
        _, g, _, compute, parts = self._setup(n)
        mine = compute[self.rank]
        contributes = bool(parts[self.rank])
        n0, chi = self.backend.projector_shape(direction, state, env)
        a0 = state.sites[coords[0]]
        # 1) projectors of my jobs (in group mode: shared with the other members of my group), then all-gather
        if mine:
            P, Pt = self.backend.move_generic_projectors(direction, state, env, mine, **opt)
        if contributes:
            flat = torch.cat([t.reshape(-1) for pair in zip(P, Pt) for t in pair])
        else:
            flat = a0.new_zeros(0)
        per_job = 2 * n0 * chi
        gathered = self._all_gather(flat, [len(p) * per_job for p in parts], 'allgather_P_Pt')
        P_all, Pt_all = [None] * n, [None] * n
        for r, jobs in enumerate(parts):
            for i, j in enumerate(jobs):
                blk = gathered[r][i * per_job:(i + 1) * per_job]
                P_all[j] = blk[:n0 * chi].view(n0, chi)
                Pt_all[j] = blk[n0 * chi:].view(n0, chi)
        # 2) absorption of my jobs (group mode: by the leader), then all-gather of the new tensors
        todo = parts[self.rank]
        res = self.backend.move_generic_absorb(direction, state, env, todo, P_all, Pt_all, **opt) if todo else []
        flat = torch.cat([t.reshape(-1) for (_, c1, c2, t3) in res for t in (c1, c2, t3)]) if res else a0.new_zeros(0)
        shapes = {j: self.backend._nT_shape(direction, state.sites[coords[j]], chi) for j in range(n)}
        def job_len(j):
            s = shapes[j]
            return 2 * chi * chi + s[0] * s[1] * s[2]
        gathered = self._all_gather(flat, [sum(job_len(j) for j in p) for p in parts], 'allgather_C_T')
        kC1, kC2, kT = OUT_KEYS[direction]
        v2s = state.vertexToSite
        for r, jobs in enumerate(parts):
            off = 0
            for j in jobs:
                c = coords[j]
                dest = v2s((c[0] - direction[0], c[1] - direction[1]))
                blk = gathered[r]
                env.C[(dest, kC1)] = blk[off:off + chi * chi].view(chi, chi); off += chi * chi
                env.C[(dest, kC2)] = blk[off:off + chi * chi].view(chi, chi); off += chi * chi
                s = shapes[j]
                ln = s[0] * s[1] * s[2]
                env.T[(dest, kT)] = blk[off:off + ln].view(*s); off += ln

    def iteration(self, state, env, move_sequence=((0, -1), (-1, 0), (0, 1), (1, 0)), **opt):
        n = 0
        for direction in move_sequence:
            for _ in range(state.lX if direction in [(-1, 0), (1, 0)] else state.lY):
                self.ctm_MOVE(direction, state, env, **opt)
                n += 1
        return n
