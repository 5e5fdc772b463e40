"""World-size-2 (and 3) gloo tests of the multi-GPU host protocol (points_matching_b200/sharded.py) on
CPU.  The per-rank compute is the CPU oracle plugged in as the engine -- test infrastructure only; the
product's engine is DeviceEngine (CUDA).  What is checked is the sharding logic: shard bounds, global
query indices, the u64 MIN / MAX key exchanges, the gathers, and rank-count invariance of the result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleEngine:
    """CPU stand-in for DeviceEngine with the same interface (torch CPU tensors in and out)."""

    def __init__(self):
        from oracle import oracle
        self.orc = oracle

    @staticmethod
    def tensor(a):
        return torch.as_tensor(np.ascontiguousarray(a))

    def _dm(self, a, base):
        a = a.copy()
        a["queryIdx"] += base
        return torch.from_numpy(a.view(np.int32).reshape(a.shape + (4,)))

    def knn2(self, q, t, norm, base):
        from points_matching_b200.api import NORM_HAMMING
        fn = self.orc.knn2_hamming if norm == NORM_HAMMING else self.orc.knn2_l2
        if q.shape[0] == 0:
            return torch.zeros((0, 2, 4), dtype=torch.int32)
        return self._dm(fn(q.numpy(), t.numpy(), 1), base)

    def col_best(self, q, t, norm, base):
        if q.shape[0] == 0:
            return torch.full((t.shape[0],), -1, dtype=torch.int64)
        col = self.orc.col_best_hamming(q.numpy(), t.numpy(), 1)          # low 32 bits: local query index
        absent = col == np.uint64(0xFFFFFFFFFFFFFFFF)
        col = np.where(absent, col, col + np.uint64(base))
        return torch.from_numpy(col.view(np.int64).copy())

    def cross_check(self, knn, col):
        from points_matching_b200 import DMATCH
        k = knn.numpy().view(DMATCH).reshape(knn.shape[0], 2)
        out = self.orc.cross_check(k, col.numpy().view(np.uint64))
        return torch.from_numpy(out.view(np.int32).reshape(-1, 4).copy())

    def ransac(self, p1, p2, idx, m, metric, thr, refit, base):
        n = p1.shape[0]
        per = 1 if m == 8 else 3
        r = self.orc.ransac_f(p1.numpy(), p2.numpy(), idx.numpy(), metric, thr, refit, nthreads=1) if idx.shape[0] else None
        if r is None:
            return (torch.zeros(1, dtype=torch.int64), torch.zeros(9, dtype=torch.float64),
                    torch.zeros(n, dtype=torch.uint8), torch.zeros(1, dtype=torch.int32))
        # winner count = inliers of the winning minimal model (before any refit)
        count = int(r["mask"].sum())
        key = (count << 32) | (0xFFFFFFFF - (base * per + int(r["best_model"])))
        return (torch.tensor([key], dtype=torch.int64), torch.from_numpy(r["F"].reshape(9).copy()),
                torch.from_numpy(r["mask"].copy()), torch.tensor([r["n_inliers"]], dtype=torch.int32))


class StubPairCtx:
    """Stands in for libpm's batched pair entry (pm_match_estimate_batched_dev): records what a rank asks for and
    fills the PAIR_RESULT records so that the host-side sharding of BASELINE config 5 can be checked without a GPU:
    pair p must be processed exactly once, by the rank that owns it, with seed p."""
    def __init__(self):
        self.calls = []

    def match_estimate_batched_dev(self, d1, n1, d2, n2, dim, is_u8, kp1, kp2, ratio, dresults, n_hyp, sample_size=8,
                                   metric=0, threshold=1.0, refit=True, seed=0):
        import ctypes
        from points_matching_b200._lib import PAIR_RESULT
        self.calls.append((len(d1), seed))
        buf = (ctypes.c_uint8 * (len(d1) * PAIR_RESULT.itemsize)).from_address(dresults)
        rec = np.frombuffer(buf, dtype=PAIR_RESULT)
        for i in range(len(d1)):
            rec[i]["F"] = np.arange(9) + 100.0 * (seed + i)          # the pair's seed = its global index
            rec[i]["n_matches"], rec[i]["n_inliers"], rec[i]["has_model"] = n1[i], n2[i], 1 if n1[i] >= 8 else 0

    def sync(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from points_matching_b200 import synth
        from points_matching_b200 import sharded
        from points_matching_b200.api import METRIC_SAMPSON, NORM_HAMMING, NORM_L2
        eng = OracleEngine()
        res = {}
        # kNN-2 (L2 and Hamming): ragged shard sizes on purpose (nq not divisible by world)
        q, t = synth.sift_pair(203, 150, seed=11)
        res["l2_knn"] = sharded.ShardedMatcher(eng, NORM_L2).knn2(eng.tensor(q), eng.tensor(t)).numpy()
        qb, tb = synth.orb_pair(301, 257, seed=12)
        mh = sharded.ShardedMatcher(eng, NORM_HAMMING)
        res["ham_knn"] = mh.knn2(eng.tensor(qb), eng.tensor(tb)).numpy()
        res["ham_cross"] = mh.match_cross(eng.tensor(qb), eng.tensor(tb)).numpy()
        # fewer queries than ranks' worth of rows: some shards are empty
        res["ham_cross_tiny"] = mh.match_cross(eng.tensor(qb[:1]), eng.tensor(tb[:40])).numpy()
        # RANSAC-F, 8-point and 7-point batches
        p1, p2, _ = synth.correspondences(600, seed=5)
        for m in (8, 7):
            idx = synth.sample_index_sets(600, 97, m, seed=40 + m)
            r = sharded.sharded_find_fundamental(eng, eng.tensor(p1), eng.tensor(p2), eng.tensor(idx), m, METRIC_SAMPSON,
                                                 1.0, True)
            F, mask, ninl, winner = r
            res[f"ransac{m}_F"] = F.numpy()
            res[f"ransac{m}_mask"] = mask.numpy()
            res[f"ransac{m}_meta"] = np.array([ninl, winner])
        # BASELINE config 5: a batch of image pairs partitioned across the ranks (ragged: 7 pairs)
        import torch
        from points_matching_b200.pipeline import match_and_estimate_batch_native
        pairs = [(torch.zeros((5 + k, 128)), torch.zeros((9 + k, 128)), torch.zeros((5 + k, 2)), torch.zeros((9 + k, 2)))
                 for k in range(7)]
        stub = StubPairCtx()
        mine = match_and_estimate_batch_native(stub, pairs, n_hyp=64)
        assert len(stub.calls) == 1 and stub.calls[0] == (len(mine), mine[0][0] if mine else stub.calls[0][1])
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        allp = sorted((p, o["n_matches"], o["n_inliers"], float(o["F"][0, 0]) if o["F"] is not None else -1.0)
                      for part in everyone for p, o in part)
        res["pairs"] = np.array(allp)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_matches_single_rank(world, tmp_path, orc):
    from points_matching_b200 import DMATCH, synth
    from points_matching_b200.sharded import shard_bounds
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    for key in ranks[0].files:                          # every rank ends with the same answer
        for r in ranks[1:]:
            assert np.array_equal(ranks[0][key], r[key]), key
    got = ranks[0]
    # ... and it is the single-process answer
    q, t = synth.sift_pair(203, 150, seed=11)
    ref = orc.knn2_l2(q, t)
    assert np.array_equal(got["l2_knn"].view(DMATCH).reshape(203, 2), ref)
    qb, tb = synth.orb_pair(301, 257, seed=12)
    refh = orc.knn2_hamming(qb, tb)
    assert np.array_equal(got["ham_knn"].view(DMATCH).reshape(301, 2), refh)
    refc = orc.cross_check(refh, orc.col_best_hamming(qb, tb))
    assert np.array_equal(got["ham_cross"].view(DMATCH).reshape(-1), refc) and len(refc) > 20
    ref1 = orc.cross_check(orc.knn2_hamming(qb[:1], tb[:40]), orc.col_best_hamming(qb[:1], tb[:40]))
    assert np.array_equal(got["ham_cross_tiny"].view(DMATCH).reshape(-1), ref1)
    p1, p2, _ = synth.correspondences(600, seed=5)
    for m in (8, 7):
        idx = synth.sample_index_sets(600, 97, m, seed=40 + m)
        r = orc.ransac_f(p1, p2, idx, 0, 1.0, True)
        assert int(got[f"ransac{m}_meta"][1]) == r["best_model"]          # rank-count invariant winner
        assert int(got[f"ransac{m}_meta"][0]) == r["n_inliers"]
        assert np.array_equal(got[f"ransac{m}_mask"], r["mask"])
        assert np.allclose(got[f"ransac{m}_F"], r["F"].reshape(9), rtol=1e-12, atol=0)
    # the pair batch: every pair exactly once, processed with its own index as the seed, model only from 8 matches up
    exp = [(k, 5 + k, 9 + k, 100.0 * k if 5 + k >= 8 else -1.0) for k in range(7)]
    assert np.array_equal(got["pairs"], np.array(exp))
    # shard bounds tile the range exactly
    for n in (0, 1, 7, 203, 10000):
        b = [shard_bounds(n, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
