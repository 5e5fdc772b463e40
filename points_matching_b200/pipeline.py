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

from .api import METRIC_SAMPSON, make_sample_sets
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
        self._sets = {}                       # host sample sets per match count (deterministic in (n, seed))

    def run(self, desc1, desc2, kp1, kp2, seed=0):
        """desc*/kp*: CUDA tensors ([n,128] f32 or u8, [n,2] f32).  Returns dict(n_matches, n_inliers, F [3,3] or None)."""
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
        self.stream.synchronize()
        n = int(self.h_count[0])
        if n < 8:
            return dict(n_matches=n, n_inliers=0, F=None)
        sets = self._sets.get((n, seed))
        if sets is None:
            sets = torch.from_numpy(make_sample_sets(n, self.n_hyp, 8, seed)).pin_memory()
            if len(self._sets) < 64:
                self._sets[(n, seed)] = sets
        with torch.cuda.stream(self.stream):
            self.samples.copy_(sets, non_blocking=True)
        c.find_fundamental_dev(self.p1.data_ptr(), self.p2.data_ptr(), n, self.samples.data_ptr(), self.n_hyp, 8, self.metric,
                               self.thr, self.refit, self.F.data_ptr(), self.mask.data_ptr(), self.ninl.data_ptr(),
                               self.key.data_ptr(), 0)
        return dict(n_matches=n, pending=True)

    def finish(self, res):
        """Reads F / inlier count of the pair whose RANSAC was enqueued last."""
        if not res.get("pending"):
            return res
        self.stream.synchronize()
        ok = int(self.key[0].item()) != 0
        return dict(n_matches=res["n_matches"], n_inliers=int(self.ninl[0].item()) if ok else 0,
                    F=self.F[:9].cpu().numpy().reshape(3, 3).copy() if ok else None)


def match_and_estimate_batch(pipeline, pairs, group=None):
    """pairs: list of (desc1, desc2, kp1, kp2) CUDA tensors, identical on every rank.  Rank r processes the
    pairs of its contiguous shard; returns this rank's list of (pair_index, result dict).  No collective is
    needed on the data path (the pairs are independent); gather the small results if every rank wants all."""
    world, rank = _world(group)
    lo, hi = shard_bounds(len(pairs), world, rank)
    out = []
    for p in range(lo, hi):
        d1, d2, k1, k2 = pairs[p]
        out.append((p, pipeline.finish(pipeline.run(d1, d2, k1, k2, seed=p))))
    return out
