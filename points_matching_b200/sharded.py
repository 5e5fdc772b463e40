"""Multi-GPU host logic: one process per GPU, query rows / hypothesis batches sharded per rank.

The reference is single-process (main.cpp calls OpenCV once per stage); BASELINE.json's
north_star shards the same two calls over the 8 GPUs of a B200 box:

  kNN matching   query rows split contiguously, train set replicated, global trainIdx unchanged
                 -> tie-breaks identical to one GPU; no data-path collective; results gathered.
  cross-check    one exchange: each rank's column minima over its query shard, packed as
                 u64 = float_bits(dist) << 32 | queryIdx, all_reduce(MIN) (Nt x 8 bytes), then
                 the local filter bwd[fwd[i]] == i  (BFMatcher(crossCheck=true), main.cpp:43-46).
  RANSAC-F       hypotheses split by batch, index sets identical on every rank; one 8-byte
                 all_reduce(MAX) of key = count << 32 | (0xFFFFFFFF - model_id); every rank then
                 re-solves the winning index set locally (cv::findFundamentalMat, main.cpp:95-98).

All compute goes through an *engine*: `DeviceEngine` drives libpm's device-resident entry points
on torch CUDA tensors.  With `DeviceEngine.init_comm()` the two exchanges run INSIDE the C ABI
(pm_match_cross_sharded_dev / pm_find_fundamental_sharded_dev: ncclAllReduce(min | max, uint64) on the ctx
stream, between the kernels they connect -- what a C++ host calls, include/pm.h "multi-GPU"); without
it they go through torch.distributed, which is also how the protocol is covered on CPU: world-size-2
gloo tests plug the CPU oracle in as the engine (tests/test_sharded_gloo.py) -- the product itself has
no CPU engine.
"""
import numpy as np
import torch
import torch.distributed as dist

from .api import METRIC_SAMPSON, NORM_HAMMING, NORM_L2

I64_MAX = torch.iinfo(torch.int64).max


def shard_bounds(n, world, rank):
    """Contiguous shard [lo, hi) of n items for `rank` of `world` (sizes differ by at most one)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class DeviceEngine:
    """libpm device-resident calls on torch CUDA tensors of one GPU (the product engine)."""

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, torch.device(device)
        stream = torch.cuda.current_stream(self.device)
        # libpm and torch must see each other's work in stream order (the torch.zeros / .to() that make the buffers,
        # the collectives that follow the kernels).  A real torch stream is shared with the ctx as it is; torch's
        # legacy default stream has handle 0, which pm_set_stream reads as "ctx-owned", so in that case the engine
        # runs libpm on a torch stream of its own and orders the two with events on both sides of every call.
        self._shared_stream = stream.cuda_stream != 0
        self._stream = stream if self._shared_stream else torch.cuda.Stream(device=self.device)
        ctx.set_stream(self._stream.cuda_stream)

    native_comm = False

    def init_comm(self, group=None):
        """Gives the ctx its own NCCL communicator over the ranks of `group` (pm_comm_init; the unique id travels through
        torch.distributed).  From then on match_cross / sharded_find_fundamental use the C ABI's sharded entries."""
        self.ctx.comm_init_from_torch(group)
        self.native_comm = self.ctx.comm_info()[0] > 1
        return self.native_comm

    def match_cross_native(self, q, t, norm, base):
        """pm_match_cross_sharded_dev on this rank's query shard: (kNN rows [nq,2,4], mutual matches [k,4])."""
        nq, nt = q.shape[0], t.shape[0]
        knn = torch.zeros((max(nq, 1), 2, 4), dtype=torch.int32, device=self.device)
        col = torch.zeros(max(nt, 1), dtype=torch.int64, device=self.device)
        out = torch.zeros((max(nq, 1), 4), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._before()
        self.ctx.match_cross_sharded_dev(q.data_ptr(), nq, t.data_ptr(), nt, q.shape[1] if nq else t.shape[1], norm, base,
                                         knn.data_ptr(), col.data_ptr(), out.data_ptr(), cnt.data_ptr())
        self._sync_for_collective()
        return knn[:nq], out[: int(cnt[0].item())]

    def ransac_native(self, p1, p2, idx_full, lo, hi, m, metric, thr, refit):
        """pm_find_fundamental_sharded_dev: this rank scores hypotheses [lo, hi) of idx_full; returns the GLOBAL winner
        (key i64[1], F f64[9], mask u8[n], n_inliers i32[1]) -- identical on every rank."""
        n = p1.shape[0]
        F = torch.zeros(16, dtype=torch.float64, device=self.device)
        mask = torch.zeros(n, dtype=torch.uint8, device=self.device)
        ninl = torch.zeros(4, dtype=torch.int32, device=self.device)
        key = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._before()
        self.ctx.find_fundamental_sharded_dev(p1.data_ptr(), p2.data_ptr(), n, idx_full.data_ptr(), idx_full.shape[0], lo, hi - lo, m,
                                              metric, thr, refit, F.data_ptr(), mask.data_ptr(), ninl.data_ptr(), key.data_ptr())
        self._sync_for_collective()
        return key[:1], F[:9], mask, ninl[:1]

    def _before(self):
        """Work enqueued on torch's current stream so far happens before the libpm kernels enqueued next."""
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self._stream.cuda_stream:
            self._stream.wait_stream(cur)

    def _sync_for_collective(self):
        """The libpm kernels enqueued so far happen before whatever torch enqueues next (collectives, reads)."""
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self._stream.cuda_stream:
            cur.wait_stream(self._stream)

    def tensor(self, a):
        return torch.as_tensor(np.ascontiguousarray(a)).to(self.device)

    def knn2(self, q, t, norm, base):
        nq, nt = q.shape[0], t.shape[0]
        out = torch.zeros((nq, 2, 4), dtype=torch.int32, device=self.device)
        if nq == 0:
            return out
        self._before()
        if norm == NORM_HAMMING:
            self.ctx.knn2_hamming_dev(q.data_ptr(), nq, t.data_ptr(), nt, q.shape[1], out.data_ptr(), base)
        elif q.dtype == torch.uint8:
            self.ctx.knn2_l2_u8_dev(q.data_ptr(), nq, t.data_ptr(), nt, q.shape[1], out.data_ptr(), base)
        else:
            self.ctx.knn2_l2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, q.shape[1], out.data_ptr(), base)
        self._sync_for_collective()
        return out

    def col_best(self, q, t, norm, base):
        nq, nt = q.shape[0], t.shape[0]
        col = torch.full((nt,), -1, dtype=torch.int64, device=self.device)       # ~0 = no query in this shard
        self._before()
        if nq and nt:
            if norm == NORM_HAMMING:
                self.ctx.col_best_hamming_dev(q.data_ptr(), nq, t.data_ptr(), nt, q.shape[1], col.data_ptr(), base)
            else:
                self.ctx.col_best_l2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, q.shape[1], col.data_ptr(), base)
        self._sync_for_collective()
        return col

    def cross_check(self, knn, col):
        nq = knn.shape[0]
        out = torch.zeros((max(nq, 1), 4), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._before()
        if nq:
            self.ctx.cross_check_dev(knn.data_ptr(), nq, 2, col.data_ptr(), col.shape[0], out.data_ptr(), cnt.data_ptr())
        self._sync_for_collective()
        return out[: int(cnt[0].item())]

    def ransac(self, p1, p2, idx, m, metric, thr, refit, base):
        """Solve + score + pick the local winner; returns (key i64[1], F f64[9], mask u8[n], n_inliers)."""
        n = p1.shape[0]
        F = torch.zeros(16, dtype=torch.float64, device=self.device)
        mask = torch.zeros(n, dtype=torch.uint8, device=self.device)
        ninl = torch.zeros(4, dtype=torch.int32, device=self.device)
        key = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._before()
        if idx.shape[0]:
            self.ctx.find_fundamental_dev(p1.data_ptr(), p2.data_ptr(), n, idx.data_ptr(), idx.shape[0], m, metric, thr,
                                          refit, F.data_ptr(), mask.data_ptr(), ninl.data_ptr(), key.data_ptr(), base)
        self._sync_for_collective()
        return key[:1], F[:9], mask, ninl[:1]


def _gather_rows(local, n_total, world, group):
    """all_gather of row shards produced by shard_bounds (pads to the largest shard)."""
    if world == 1:
        return local
    width = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        rows.append(parts[r][: hi - lo])
    return torch.cat(rows, dim=0)


class ShardedMatcher:
    """BFMatcher(normType, crossCheck) over the ranks of a process group.  Every rank passes the
    same (replicated) query and train arrays; each matches its own query shard."""

    def __init__(self, engine, normType=NORM_L2, group=None):
        self.engine, self.norm, self.group = engine, normType, group
        self.world, self.rank = _world(group)

    def _shard(self, query):
        lo, hi = shard_bounds(query.shape[0], self.world, self.rank)
        return lo, hi, query[lo:hi].contiguous()

    def knn2_local(self, query, train):
        """This rank's [hi-lo, 2, 4] int32 DMatch rows (queryIdx already global)."""
        lo, hi, q = self._shard(query)
        return self.engine.knn2(q, train, self.norm, lo)

    def knn2(self, query, train):
        """knnMatch(k=2) of the whole query set, identical on every rank: [nq, 2, 4] int32 DMatch rows."""
        return _gather_rows(self.knn2_local(query, train), query.shape[0], self.world, self.group)

    def match_cross(self, query, train):
        """BFMatcher(norm, crossCheck=true).match: mutual nearest neighbours of the whole query set in
        queryIdx order, identical on every rank ([n, 4] int32 DMatch rows)."""
        lo, hi, q = self._shard(query)
        if getattr(self.engine, "native_comm", False) and self.group is None:
            _, mine = self.engine.match_cross_native(q, train, self.norm, lo)     # min-reduce inside the C ABI
            return self._gather_survivors(mine)
        knn = self.engine.knn2(q, train, self.norm, lo)
        col = self.engine.col_best(q, train, self.norm, lo)
        if self.world > 1:
            col = torch.where(col < 0, torch.full_like(col, I64_MAX), col)      # ~0 must lose a signed MIN
            dist.all_reduce(col, op=dist.ReduceOp.MIN, group=self.group)
            col = torch.where(col == I64_MAX, torch.full_like(col, -1), col)
        mine = self.engine.cross_check(knn, col)
        return self._gather_survivors(mine)

    def _gather_survivors(self, mine):
        if self.world == 1:
            return mine
        # survivors per rank differ: exchange counts, then padded rows
        cnt = torch.tensor([mine.shape[0]], dtype=torch.int64, device=mine.device)
        cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(cnts, cnt, group=self.group)
        width = max(1, max(int(c.item()) for c in cnts))
        pad = torch.zeros((width, 4), dtype=torch.int32, device=mine.device)
        pad[: mine.shape[0]] = mine
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        return torch.cat([parts[r][: int(cnts[r].item())] for r in range(self.world)], dim=0)


def sharded_find_fundamental(engine, p1, p2, sample_idx, sample_size=8, metric=METRIC_SAMPSON, threshold=1.0,
                             refit=True, group=None):
    """RANSAC-F with the hypothesis batch `sample_idx` ([n_hyp, m], identical on every rank) split
    across ranks.  Returns (F f64[9], mask u8[n], n_inliers, winner_model_id) -- identical on every
    rank and identical to a single-rank run over the whole batch -- or None when no hypothesis gave a
    model.  model id = hyp for 8-point samples, 3 * hyp + k for 7-point samples."""
    world, rank = _world(group)
    per = 1 if sample_size == 8 else 3
    lo, hi = shard_bounds(sample_idx.shape[0], world, rank)
    if world > 1 and getattr(engine, "native_comm", False) and group is None:
        # one call: solve + score the shard, ncclAllReduce(max) of the key, local re-solve of the winner, mask, refit
        key, F, mask, ninl = engine.ransac_native(p1, p2, sample_idx, lo, hi, sample_size, metric, threshold, refit)
        k = int(key[0].item())
        return None if k == 0 else (F, mask, int(ninl[0].item()), 0xFFFFFFFF - (k & 0xFFFFFFFF))
    key, F, mask, ninl = engine.ransac(p1, p2, sample_idx[lo:hi].contiguous(), sample_size, metric, threshold,
                                       refit if world == 1 else False, lo)
    if world > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)       # 8 bytes: max count, lowest id on ties
    k = int(key[0].item())
    if k == 0:
        return None
    winner = 0xFFFFFFFF - (k & 0xFFFFFFFF)
    if world > 1:
        # every rank knows the winning index set: re-solve it locally instead of broadcasting F and the mask
        h = winner // per
        key1, F, mask, ninl = engine.ransac(p1, p2, sample_idx[h:h + 1].contiguous(), sample_size, metric, threshold,
                                            refit, h)
        assert int(key1[0].item()) == k, "winner re-solve disagrees with the reduced key"
    return F, mask, int(ninl[0].item()), winner
