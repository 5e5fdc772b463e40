"""Two-rank NCCL run of the multi-GPU protocol with the product engine (libpm on CUDA tensors): results must be
byte-identical to the single-GPU / oracle answer (sharding must not change tie-breaks).  Needs >= 2 GPUs; the
CPU (gloo) twin of this test is tests/test_sharded_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import points_matching_b200 as pm
        from points_matching_b200 import sharded, synth
        ctx = pm.Context(rank)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        eng = sharded.DeviceEngine(ctx, f"cuda:{rank}")
        res = {}
        q, t = synth.sift_pair(2003, 1500, seed=11)
        res["l2_knn"] = sharded.ShardedMatcher(eng, pm.NORM_L2).knn2(eng.tensor(q), eng.tensor(t)).cpu().numpy()
        qb, tb = synth.orb_pair(3001, 2570, seed=12)
        mh = sharded.ShardedMatcher(eng, pm.NORM_HAMMING)
        res["ham_knn"] = mh.knn2(eng.tensor(qb), eng.tensor(tb)).cpu().numpy()
        res["ham_cross"] = mh.match_cross(eng.tensor(qb), eng.tensor(tb)).cpu().numpy()
        p1, p2, _ = synth.correspondences(3000, seed=5)
        idx = synth.sample_index_sets(3000, 1001, 8, seed=48)
        F, mask, ninl, winner = sharded.sharded_find_fundamental(eng, eng.tensor(p1), eng.tensor(p2), eng.tensor(idx), 8,
                                                                 pm.METRIC_SAMPSON, 1.0, True)
        res["ransac_F"], res["ransac_mask"] = F.cpu().numpy(), mask.cpu().numpy()
        res["ransac_meta"] = np.array([ninl, winner])
        # the same two exchanges INSIDE the C ABI (pm_comm_init + pm_match_cross_sharded_dev /
        # pm_find_fundamental_sharded_dev: ncclAllReduce on the ctx stream) -- what a C++ host runs
        assert eng.init_comm() and ctx.comm_info() == (world, rank)
        res["native_ham_cross"] = mh.match_cross(eng.tensor(qb), eng.tensor(tb)).cpu().numpy()
        res["native_l2_cross"] = sharded.ShardedMatcher(eng, pm.NORM_L2).match_cross(eng.tensor(q), eng.tensor(t)).cpu().numpy()
        F2, mask2, ninl2, winner2 = sharded.sharded_find_fundamental(eng, eng.tensor(p1), eng.tensor(p2), eng.tensor(idx), 8,
                                                                     pm.METRIC_SAMPSON, 1.0, True)
        res["native_ransac_F"], res["native_ransac_mask"] = F2.cpu().numpy(), mask2.cpu().numpy()
        res["native_ransac_meta"] = np.array([ninl2, winner2])
        # sample sets generated from the seed on every rank (no index array at all), 7-point samples
        dF = torch.zeros(16, dtype=torch.float64, device=eng.device); dmask = torch.zeros(3000, dtype=torch.uint8, device=eng.device)
        dn = torch.zeros(4, dtype=torch.int32, device=eng.device); dkey = torch.zeros(2, dtype=torch.int64, device=eng.device)
        lo, hi = sharded.shard_bounds(777, world, rank)
        d1, d2 = eng.tensor(p1), eng.tensor(p2)
        eng._before()
        ctx.find_fundamental_sharded_dev(d1.data_ptr(), d2.data_ptr(), 3000, 0, 777, lo, hi - lo, 7, pm.METRIC_SYMEPI, 1.0, False,
                                         dF.data_ptr(), dmask.data_ptr(), dn.data_ptr(), dkey.data_ptr(), seed=321)
        ctx.sync()
        res["seeded_F"], res["seeded_mask"] = dF[:9].cpu().numpy(), dmask.cpu().numpy()
        res["seeded_meta"] = np.array([int(dn[0].item()), int(dkey[0].item())])
        # BASELINE config 5: image pairs partitioned across ranks, each rank's shard through the batched C-ABI entry
        from points_matching_b200.pipeline import match_and_estimate_batch_native
        mine = match_and_estimate_batch_native(ctx, _pairs(torch, f"cuda:{rank}"), n_hyp=512)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        allp = sorted((p, o) for part in everyone for p, o in part)
        res["pairs_idx"] = np.array([p for p, _ in allp])
        res["pairs_meta"] = np.array([[o["n_matches"], o["n_inliers"]] for _, o in allp])
        res["pairs_F"] = np.stack([o["F"] if o["F"] is not None else np.zeros((3, 3)) for _, o in allp])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
        ctx.close()
    finally:
        dist.destroy_process_group()


def _pairs(torch, device):
    from points_matching_b200 import synth
    out = []
    for k, (n1, n2) in enumerate([(1500, 1700), (900, 800), (2048, 2048), (600, 1000), (1200, 1300)]):
        d1, d2, k1, k2, _ = synth.image_pair(n1, n2, seed=40 + k)
        out.append(tuple(torch.from_numpy(a).to(device) for a in (d1, d2, k1, k2)))
    return out


def test_two_rank_nccl_equals_single_gpu(tmp_path, orc):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import points_matching_b200 as pm
    from points_matching_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(2))
    for k in r0.files:
        assert np.array_equal(r0[k], r1[k]), k
    q, t = synth.sift_pair(2003, 1500, seed=11)
    assert np.array_equal(r0["l2_knn"].view(pm.DMATCH).reshape(2003, 2), orc.knn2_l2(q, t))
    qb, tb = synth.orb_pair(3001, 2570, seed=12)
    refh = orc.knn2_hamming(qb, tb)
    assert np.array_equal(r0["ham_knn"].view(pm.DMATCH).reshape(3001, 2), refh)
    assert np.array_equal(r0["ham_cross"].view(pm.DMATCH).reshape(-1), orc.cross_check(refh, orc.col_best_hamming(qb, tb)))
    p1, p2, _ = synth.correspondences(3000, seed=5)
    idx = synth.sample_index_sets(3000, 1001, 8, seed=48)
    r = orc.ransac_f(p1, p2, idx, 0, 1.0, True)
    assert int(r0["ransac_meta"][1]) == r["best_model"] and abs(int(r0["ransac_meta"][0]) - r["n_inliers"]) <= 3
    assert (r0["ransac_mask"] == r["mask"]).mean() > 0.998
    # the exchanges inside the C ABI give the same answers as the torch.distributed protocol, bit for bit
    assert np.array_equal(r0["native_ham_cross"], r0["ham_cross"])
    ctx1 = pm.Context(0)
    assert np.array_equal(r0["native_l2_cross"].view(pm.DMATCH).reshape(-1), ctx1.match_cross(q, t, pm.NORM_L2))
    for k in ("ransac_F", "ransac_mask", "ransac_meta"):
        assert np.array_equal(r0["native_" + k], r0[k]), k
    from points_matching_b200.api import make_sample_sets
    one = ctx1.find_fundamental(p1, p2, sample_size=7, metric=pm.METRIC_SYMEPI, threshold=1.0, refit=False,
                                sample_idx=make_sample_sets(3000, 777, 7, 321))
    ctx1.close()
    assert np.array_equal(r0["seeded_F"].reshape(3, 3), one[0]) and np.array_equal(r0["seeded_mask"], one[1])
    assert int(r0["seeded_meta"][0]) == one[2]
    # the partitioned pair batch equals the same batch on one GPU, pair by pair and bit for bit
    from points_matching_b200.pipeline import match_and_estimate_batch_native
    ctx = pm.Context(0)
    one = match_and_estimate_batch_native(ctx, _pairs(torch, "cuda:0"), n_hyp=512)
    ctx.close()
    assert list(r0["pairs_idx"]) == [p for p, _ in one] == list(range(5))
    for i, (_, o) in enumerate(one):
        assert list(r0["pairs_meta"][i]) == [o["n_matches"], o["n_inliers"]] and o["F"] is not None
        assert np.array_equal(r0["pairs_F"][i], o["F"])


def test_cpp_host_two_gpus_example():
    """examples/sharded_two_gpus: a C++ host, one thread and one pm_ctx per GPU, NCCL inside the C ABI
    (pm_comm_init / pm_match_cross_sharded_dev / pm_allgather_matches_dev / pm_find_fundamental_sharded_dev);
    it checks itself against the same calls on one GPU and exits 0 when every bit agrees."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    exe = os.path.join(ROOT, "examples", "sharded_two_gpus")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])
    r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical to one GPU" in r.stdout
