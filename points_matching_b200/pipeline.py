"""Device-resident end-to-end flow of main.cpp:43-98 for one image pair, and its batched form
(BASELINE config 5: many independent pairs, partitioned across ranks).

    knnMatch(k=2) -> ratio test -> KeyPoint::convert on both sides -> findFundamentalMat(RANSAC)

Descriptors, keypoints, matches and correspondences never leave HBM between the stages
(SURVEY 8 f1): K1-K3 write the kNN rows, K5 compacts the good matches, gather_matches turns them
into the two point lists, K6-K8 estimate F.  The only host round trip per pair is the 4-byte match
count (the RANSAC launch geometry depends on it); F and the inlier count are read once at the end.
"""
import numpy as np
import torch

from .api import METRIC_SAMPSON
from .sharded import _world, shard_bounds


class PairPipeline:
    def __init__(self, ctx, device, max_desc, n_hyp=4096, ratio=0.75, threshold=1.0, metric=METRIC_SAMPSON, refit=True):
        self.ctx, self.dev = ctx, torch.device(device)
        self.n_hyp, self.ratio, self.thr, self.metric, self.refit = n_hyp, ratio, threshold, metric, refit
        # torch copies and libpm kernels share ONE stream, so stream order is the only synchronisation needed.
        # (torch's legacy default stream has handle 0, which pm_set_stream reads as "ctx-owned": use a real one)
        stream = torch.cuda.current_stream(self.dev)
        self.stream = stream if stream.cuda_stream != 0 else torch.cuda.Stream(device=self.dev)
        ctx.set_stream(self.stream.cuda_stream)
        i32 = dict(dtype=torch.int32, device=self.dev)
        self.knn = torch.zeros((max_desc, 2, 4), **i32)
        self.good = torch.zeros((max_desc, 4), **i32)
        self.ngood = torch.zeros(4, **i32)
        self.p1 = torch.zeros((max_desc, 2), dtype=torch.float32, device=self.dev)
        self.p2 = torch.zeros((max_desc, 2), dtype=torch.float32, device=self.dev)
        self.samples = torch.zeros((n_hyp, 8), **i32)
        self.F = torch.zeros(16, dtype=torch.float64, device=self.dev)
        self.mask = torch.zeros(max_desc, dtype=torch.uint8, device=self.dev)
        self.ninl = torch.zeros(4, **i32)
        self.key = torch.zeros(2, dtype=torch.int64, device=self.dev)
        self.h_count = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.res_dev = torch.zeros(16, dtype=torch.float64, device=self.dev)
        self.h_res = torch.zeros(16, dtype=torch.float64).pin_memory()
        self._ev = torch.cuda.Event()

    # The flow is split into three host stages so that several pipelines (one ctx + stream each) can be
    # interleaved: while one waits for its 4-byte match count, the GPU runs another pair's kernels.
    def start(self, desc1, desc2, kp1, kp2, seed=0):
        """Stage 1: enqueue matching + ratio filter + gather and the asynchronous read of the match count."""
        c, n1, n2 = self.ctx, desc1.shape[0], desc2.shape[0]
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))     # inputs produced on the caller's stream
        if desc1.dtype == torch.uint8:
            c.knn2_l2_u8_dev(desc1.data_ptr(), n1, desc2.data_ptr(), n2, desc1.shape[1], self.knn.data_ptr(), 0)
        else:
            c.knn2_l2_f32_dev(desc1.data_ptr(), n1, desc2.data_ptr(), n2, desc1.shape[1], self.knn.data_ptr(), 0)
        c.ratio_filter_dev(self.knn.data_ptr(), n1, self.ratio, self.good.data_ptr(), self.ngood.data_ptr())
        c.gather_matches_dev(self.good.data_ptr(), self.ngood.data_ptr(), n1, kp1.data_ptr(), n1, kp2.data_ptr(), n2,
                             self.p1.data_ptr(), self.p2.data_ptr())
        with torch.cuda.stream(self.stream):
            self.h_count.copy_(self.ngood, non_blocking=True)
            self._ev.record(self.stream)
        return dict(seed=seed, stage=1)

    def estimate(self, res):
        """Stage 2: wait for the match count, enqueue RANSAC-F and the asynchronous read of its result."""
        c = self.ctx
        self._ev.synchronize()
        n = int(self.h_count[0])
        res.update(n_matches=n, stage=2, pending=n >= 8)
        if n < 8:
            return res
        c.make_sample_sets_dev(n, self.n_hyp, 8, res["seed"], self.samples.data_ptr())   # == the host generator's sets
        c.find_fundamental_dev(self.p1.data_ptr(), self.p2.data_ptr(), n, self.samples.data_ptr(), self.n_hyp, 8, self.metric,
                               self.thr, self.refit, self.F.data_ptr(), self.mask.data_ptr(), self.ninl.data_ptr(),
                               self.key.data_ptr(), 0)
        with torch.cuda.stream(self.stream):
            # one packed read: [key, n_inliers, F(9)] as 11 doubles
            self.res_dev[0] = self.key[0].to(torch.float64)          # key != 0 <=> a model exists (exact for the test)
            self.res_dev[1] = self.ninl[0].to(torch.float64)
            self.res_dev[2:11] = self.F[:9]
            self.h_res.copy_(self.res_dev, non_blocking=True)
            self._ev.record(self.stream)
        return res

    def finish(self, res):
        """Stage 3: read F / inlier count of the pair whose RANSAC was enqueued by estimate()."""
        if res.get("stage") == 1:
            res = self.estimate(res)
        if not res.get("pending"):
            return dict(n_matches=res["n_matches"], n_inliers=0, F=None)
        self._ev.synchronize()
        h = self.h_res.numpy()
        ok = h[0] != 0.0
        return dict(n_matches=res["n_matches"], n_inliers=int(h[1]) if ok else 0, F=h[2:11].reshape(3, 3).copy() if ok else None)

    def run(self, desc1, desc2, kp1, kp2, seed=0):
        """desc*/kp*: CUDA tensors ([n,128] f32 or u8, [n,2] f32).  Returns the dict finish() completes."""
        return self.estimate(self.start(desc1, desc2, kp1, kp2, seed))


def match_and_estimate_batch_native(ctx, pairs, n_hyp=4096, ratio=0.75, threshold=1.0, metric=METRIC_SAMPSON, refit=True,
                                    group=None, sync=True):
    """The same flow through the C ABI's batched entry (pm_match_estimate_batched_dev): rank r enqueues its
    contiguous shard of `pairs` in ONE call, nothing returns to the host between the stages of a pair (the
    match count stays on the device), and the per-pair records are read once at the end.  Pair p uses seed p,
    like match_and_estimate_batch.  Returns [(pair_index, result dict)], or (lo, device records) when
    sync=False (records are complete once the ctx stream has drained)."""
    from ._lib import PAIR_RESULT
    world, rank = _world(group)
    lo, hi = shard_bounds(len(pairs), world, rank)
    mine = pairs[lo:hi]
    dev = mine[0][0].device if mine else torch.device("cuda")
    res = torch.zeros((max(len(mine), 1), PAIR_RESULT.itemsize), dtype=torch.uint8, device=dev)
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()  # inputs / records made on torch's stream; libpm uses the ctx stream
    if mine:
        is_u8 = mine[0][0].dtype == torch.uint8
        ctx.match_estimate_batched_dev([p[0].data_ptr() for p in mine], [p[0].shape[0] for p in mine],
                                       [p[1].data_ptr() for p in mine], [p[1].shape[0] for p in mine],
                                       mine[0][0].shape[1], is_u8, [p[2].data_ptr() for p in mine],
                                       [p[3].data_ptr() for p in mine], ratio, res.data_ptr(), n_hyp, 8, metric, threshold,
                                       refit, seed=lo)
    if not sync:
        return lo, res
    ctx.sync()
    rec = res.cpu().numpy().view(PAIR_RESULT).reshape(-1)[: len(mine)]
    return [(lo + i, dict(n_matches=int(r["n_matches"]), n_inliers=int(r["n_inliers"]),
                          F=r["F"].reshape(3, 3).copy() if r["has_model"] else None)) for i, r in enumerate(rec)]


def match_and_estimate_batch(pipelines, pairs, group=None):
    """pairs: list of (desc1, desc2, kp1, kp2) CUDA tensors, identical on every rank.  Rank r processes the
    pairs of its contiguous shard; returns this rank's list of (pair_index, result dict).  No collective is
    needed on the data path (the pairs are independent); gather the small results if every rank wants all.
    `pipelines` is one PairPipeline or a list of them (each with its own ctx): with two or more, the pairs are
    software-pipelined so one pipeline's host round trip hides behind another's kernels."""
    if isinstance(pipelines, PairPipeline):
        pipelines = [pipelines]
    world, rank = _world(group)
    lo, hi = shard_bounds(len(pairs), world, rank)
    k, out = len(pipelines), []
    for base in range(lo, hi, k):                 # rounds of k pairs, one per pipeline (a pipeline holds one pair in flight)
        chunk = list(range(base, min(base + k, hi)))
        st = [pipelines[i].start(*pairs[p], seed=p) for i, p in enumerate(chunk)]
        st = [pipelines[i].estimate(r) for i, r in enumerate(st)]
        out += [(p, pipelines[i].finish(st[i])) for i, p in enumerate(chunk)]
    return out
