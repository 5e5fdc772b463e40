#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, libpm.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (OpenCV)

Metric (BASELINE.json): SIFT kNN-2 match pairs/sec.  A "step" is one pass of the hot path over
one batch: 10k x 10k x 128-d float SIFT-like descriptors -> K1 pack/norms -> K2 tcgen05 L2 GEMM
with fused top-k -> K3 FP32 merge/re-rank -> K5 ratio test 0.75 (BASELINE configs[1]); a pair
is one (query, train) distance evaluation.  `value` times it with the inputs resident in HBM,
`e2e` through the host-buffer C-ABI call (pinned host buffers, H2D + D2H inside the timed
region).  For N > 1 the query rows are sharded: every rank matches its own 10k-row query shard
against the replicated train set (weak scaling, no data-path collective).
The second half of the metric, RANSAC-F hypotheses/sec (configs[3]: 100k correspondences, 50%
outliers, 8-point, Sampson 1 px), and the other configurations (cfg3 Hamming cross-check shard, cfg5
batched image pairs, the split mode for general floats, the literal LMedS call) ride in the same line:
their roofline / e2e / CPU-baseline figures sit in the "secondary" members of `roofline`, `e2e` and
`cpu_baseline`, and the compact `summary` object is the LAST key of the line.  For N > 1 every rank
checks its results against the single-GPU answer (`parity_ok`).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NQ, NT, DIM, RATIO = 10000, 10000, 128, 0.75
POOL = 16                                   # rotating input sets: 16 x 10.24 MB = 164 MB > 126 MB L2
R_N, R_HYP, R_THR = 100000, 1 << 20, 1.0    # RANSAC-F config 4
METRIC, UNIT = "sift_knn2_match_pairs_per_sec", "pairs/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def config_dict(n_gpus, lanes=4):
    return {"workload": "cfg2: synthetic SIFT-like 128-d f32 descriptors 10000x10000, L2 kNN-2 + ratio 0.75",
            "nq_per_gpu": NQ, "nt": NT, "dim": DIM, "ratio": RATIO,
            "sharding": "query rows per rank, train replicated" if n_gpus > 1 else "single GPU",
            "l2_policy": f"rotating {POOL} distinct input sets ({POOL * (NQ + NT) * DIM * 4 / 1e6:.0f} MB) > 126 MB L2",
            "gpu_lanes": lanes}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the observed range (idle samples sit at ~120 MHz)
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (OpenCV), host cores only
# ----------------------------------------------------------------------------------------
def cpu_match_fn():
    """Returns (fn(q, t) -> n_good, kind, cores, description)."""
    try:
        import cv2
        cores = os.cpu_count() or 1
        cv2.setNumThreads(cores)

        def run(q, t):
            # array form of BFMatcher(NORM_L2).knnMatch(q, t, k=2) (same batchDistance call, no
            # per-match Python objects), then the ratio test
            dist, idx = cv2.batchDistance(q, t, cv2.CV_32F, None, None, cv2.NORM_L2, 2)
            return int((dist[:, 0] < RATIO * dist[:, 1]).sum())

        desc = (f"OpenCV {cv2.__version__} (cv2) batchDistance K=2 = BFMatcher(NORM_L2).knnMatch + ratio {RATIO}, "
                f"{cv2.getNumThreads()} threads; the reference links OpenCV 2.4.13's same routine (main.cpp:43-46)")
        return run, "reference", cv2.getNumThreads(), desc
    except Exception:
        from oracle import oracle as orc
        cores = os.cpu_count() or 1

        def run(q, t):
            return len(orc.ratio_filter(orc.knn2_l2(q, t, cores), RATIO))

        return run, "port", cores, f"oracle/pm_oracle.c (OpenMP, {cores} threads): f64 L2 kNN-2 + ratio {RATIO}"


def time_cpu_sample(fn, q, t, budget_s, reps=3):
    """Times fn on a bounded query sample sized for ~budget_s seconds per call."""
    n0 = min(512, q.shape[0])
    fn(q[:64], t)                                   # thread-pool warm-up
    t0 = time.perf_counter(); fn(q[:n0], t); dt = time.perf_counter() - t0
    rate = n0 * t.shape[0] / max(dt, 1e-6)
    rows = int(min(q.shape[0], max(64, budget_s * rate / t.shape[0])))
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); fn(q[:rows], t); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return rows, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from points_matching_b200 import synth
    q, t = synth.sift_pair(NQ, NT, seed=1234)
    fn, kind, cores, desc = cpu_match_fn()
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = min(2.0, 150.0 / (steps + warm + 4))     # whole run within a few minutes
    rows, _ = time_cpu_sample(fn, q, t, budget, reps=1)
    for _ in range(warm):
        fn(q[:rows], t)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(q[:rows], t)
    dt = (time.perf_counter() - t0) / steps
    value = rows * NT / dt
    sample = f"{rows} of {NQ} query rows x {NT} train rows per step; {desc}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus, args.lanes),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import points_matching_b200 as pm
    from points_matching_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (the NCCL version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    numa_note = None
    if world > 1 and not args.no_affinity:
        # several ranks on one host: keep this rank's threads (and so the first-touch pages of its pinned staging
        # buffers) on the cores next to its GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1} & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa_note = f"rank threads bound to the {len(cpus)} cores local to GPU {local}"
        except Exception as e:      # noqa: BLE001 -- an optimisation only
            numa_note = f"cpu affinity not set: {e}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    peaks = load_peaks()

    ctx = pm.Context(local)
    # a real (non-default) stream shared by torch and libpm: torch's legacy default stream has
    # handle 0, which pm_set_stream reads as "use the ctx-owned stream", and torch events would
    # then not see libpm's kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: a pool of distinct (query shard, train) sets, resident in HBM ----------
    q0, t0 = synth.sift_pair(NQ, NT, seed=1234 + rank)
    if world > 1:                                    # train set replicated: every rank uses rank 0's
        _, t0 = synth.sift_pair(NQ, NT, seed=1234)
    gq, gt = torch.from_numpy(q0).to(dev), torch.from_numpy(t0).to(dev)
    g = torch.Generator(device=dev); g.manual_seed(7 + rank)
    pool = [(gq, gt)]
    for _ in range(POOL - 1):                        # row permutations: distinct memory, same statistics
        pool.append((gq[torch.randperm(NQ, device=dev, generator=g)].contiguous(),
                     gt[torch.randperm(NT, device=dev, generator=g)].contiguous()))
    knn = torch.zeros((NQ, 2, 4), dtype=torch.int32, device=dev)
    good = torch.zeros((NQ, 4), dtype=torch.int32, device=dev)
    ngood = torch.zeros(4, dtype=torch.int32, device=dev)
    qbase = rank * NQ

    # Lanes: the steps of the timed loop go round-robin over `--lanes` contexts (own stream, own workspaces, own output
    # buffers each) -- every step is still one whole pass K1 -> K2 -> K3 -> K5 over its own input set, but the launch
    # hand-offs and the small kernels of one lane run under the GEMM of another.  One lane: consecutive steps overlap
    # through pm_set_pipelining instead (K1 of step i+1 under K3 / K5 of step i; the inputs are resident and complete
    # before the loop starts, as that mode requires).  Lane 0 is `ctx`: every other leg, and the per-kernel timing, use it alone.
    n_lanes = max(1, args.lanes)
    pipelining = (not args.no_pipelining) and n_lanes == 1
    ctx.set_pipelining(pipelining)
    lane_ctx, lane_stream, lane_out = [ctx], [stream], [(knn, good, ngood)]
    for _ in range(1, n_lanes):
        c = pm.Context(local)
        s_ = torch.cuda.Stream(device=dev)
        c.set_stream(s_.cuda_stream)
        lane_ctx.append(c); lane_stream.append(s_)
        lane_out.append((torch.zeros_like(knn), torch.zeros_like(good), torch.zeros_like(ngood)))

    # one device-resident C-ABI call per step: K1 pack -> K2 GEMM + fused top-2 -> K3 re-rank -> K5 ratio test + compaction.
    # The ctypes arguments are built once (a step is ~34 us of GPU work: per-call Python argument marshalling is not free)
    import ctypes as C
    from points_matching_b200 import _lib as _pmlib
    _fn = _pmlib.lib().pm_knn2_ratio_l2_f32_dev
    _pool_args = [(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr())) for a, b in pool]
    _nq, _nt, _dim, _ratio, _base = C.c_int(NQ), C.c_int(NT), C.c_int(DIM), C.c_float(RATIO), C.c_int(qbase)
    _lane_args = [(c._h, C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()), C.c_void_p(o[2].data_ptr()))
                  for c, o in zip(lane_ctx, lane_out)]

    def step_on(lane, i):
        a, b = _pool_args[i % POOL]
        h, k_, g_, n_ = _lane_args[lane]
        if _fn(h, a, _nq, b, _nt, _dim, _ratio, _base, k_, g_, n_) != 0:
            lane_ctx[lane]._chk(-2)

    def step(i):                                     # lane 0 alone (per-kernel timing, parity, the other legs)
        step_on(0, i)

    def join_lanes():                                # lane 0's stream waits for everything the other lanes have queued
        for s_ in lane_stream[1:]:
            e = torch.cuda.Event()
            e.record(s_)
            stream.wait_event(e)

    def timed_loop(n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for s_ in lane_stream[1:]:                   # no lane starts before the opening event
            s_.wait_event(a)
        for i in range(n):
            step_on(i % n_lanes, i)
        join_lanes()
        b.record(stream)
        barrier()
        return a.elapsed_time(b)

    steps, warm = max(1, args.steps), max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    # clock ramp: an idle B200 sits at ~120 MHz; run the workload ~1 s before anything is timed
    t_end = time.perf_counter() + (0.0 if args.no_ramp else 1.0)
    ramp = 0
    while time.perf_counter() < t_end:
        for i in range(50):
            step_on((ramp + i) % n_lanes, ramp + i)
        ramp += 50
        torch.cuda.synchronize()
    for i in range(max(warm, 3 * n_lanes)):
        step_on(i % n_lanes, i)
    barrier()
    launches0 = sum(c.launch_count() for c in lane_ctx)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_total = timed_loop(steps)
    launches = sum(c.launch_count() for c in lane_ctx) - launches0
    # the same loop on lane 0 alone: what one stream of consecutive steps costs (explanatory)
    ms_single = ms_total
    if n_lanes > 1:
        ctx.set_pipelining(not args.no_pipelining)
        for i in range(warm):
            step(i)
        barrier()
        ev0.record(stream)
        for i in range(steps):
            step(i)
        ev1.record(stream)
        barrier()
        ms_single = ev0.elapsed_time(ev1)
    # per-kernel timing for the roofline: the same K steps replayed with an event pair around every
    # K2 launch (kept out of the region above: the event records would serialise the
    # programmatic-dependent-launch overlap between the kernels of a step)
    ctx.profile_enable(True)
    for i in range(steps):
        step(i)
    k2_ms, k2_n = ctx.profile_read(0)
    ctx.profile_enable(False)
    # the same kernel launched 8 times back to back inside ONE event pair per step (it only re-writes the same candidates):
    # its steady-state duration, the per-launch event gap and launch latency amortised -- how the bf16 peak it is held
    # against was measured (a 0.68 ms GEMM, best of 10)
    K2_REP = 8
    _pmlib.lib().pm_debug_k2_repeat(K2_REP)
    ctx.profile_enable(True)
    for i in range(min(steps, 200)):
        step(i)
    k2b_ms, k2b_n = ctx.profile_read(0)
    ctx.profile_enable(False)
    _pmlib.lib().pm_debug_k2_repeat(1)
    n_good_last = int(ngood[0].item())
    stats = ctx.l2_stats()
    # the same kernel inside the undisturbed chain: %globaltimer stamps written by K2 itself (first CTA past its
    # waits -> last CTA out) for ONE step in the middle of a running loop, 15 repetitions, median
    k2_chain_us = None
    try:
        import ctypes as C0
        from points_matching_b200 import _lib as _l0
        L0 = _l0.lib()
        span = torch.zeros(15 + 17 + 2 * 1024, dtype=torch.int64, device=dev)
        init = np.zeros(span.numel(), dtype=np.int64)
        init[[0, 1, 3, 4, 6, 7, 9, 10, 12, 13]] = np.iinfo(np.int64).max
        init_d = torch.from_numpy(init).to(dev)
        durs = []
        for rep in range(15):
            span.copy_(init_d)
            for i in range(20):
                step(i)
            L0.pm_debug_set_span(C0.c_void_p(span.data_ptr()))
            step(20)
            L0.pm_debug_set_span(C0.c_void_p(0))
            for i in range(3):
                step(21 + i)
            torch.cuda.synchronize()
            r = span.cpu().numpy()
            durs.append((int(r[5]) - int(r[4])) / 1e3)
        k2_chain_us = float(np.median(durs))
    except Exception as e:      # instrumentation only
        print("bench.py: in-chain K2 timing unavailable:", e, file=sys.stderr)
    ctx.set_pipelining(False)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / steps
    ms_step_single = ms_single / steps
    value = world * NQ * NT / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (K2, tcgen05 GEMM + fused top-k) -----------------
    flops = 2.0 * DIM * NQ * NT                       # 256 FLOP per pair
    k2_avg_ms = k2_ms / max(k2_n, 1)
    achieved = flops / (k2_avg_ms * 1e-3) / 1e12 if k2_n else 0.0
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "k2_traffic.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    except Exception:
        pass
    # Denominator: the timed window is steps x ~34 us -- milliseconds at the boost clock, far below the ~1 s it takes
    # the GPU to settle at its sustained tensor clock -- so the kernel is held against the BURST bf16 peak unless the
    # window is long enough to be a sustained measurement.
    window_s = ms_total * 1e-3
    sustained_window = window_s >= 2.0
    peak = peaks["bf16_sustained"] if sustained_window else peaks["bf16_burst"]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "kernel": "l2_tc_kernel (tcgen05.mma cta_group::1 kind::f16 bf16, M128 N128 K16, 256x128 work items, fused top-2 epilogue)",
                "kernel_ms": k2_avg_ms,
                # the event-timed launch against one stream of consecutive steps (with several lanes the step period is
                # shorter than one event-timed K2: its launch latency and prologue run under the other lanes' kernels --
                # the in-chain share below is the one that compares with the lane-interleaved step)
                "kernel_share_of_step": k2_avg_ms / ms_step_single if ms_step_single else None,
                "peak_source": peaks["source"] + (", sustained bf16 (timed window %.1f s)" % window_s if sustained_window else
                                                  ", BURST bf16 (timed window %.1f ms at the boost clock)" % (window_s * 1e3)),
                "frac_of_sustained_peak": achieved / peaks["bf16_sustained"],
                "timing": "CUDA-event pair around every K2 launch on the launching stream (pm_profile_*), same K steps replayed; the "
                          "pair breaks the programmatic-dependent-launch overlap, so kernel_ms includes K2's launch latency and prologue",
                "traffic_source": traffic_src,
                "back_to_back": None if not k2b_n else {
                    "kernel_ms": k2b_ms / (k2b_n * K2_REP), "achieved": flops / (k2b_ms / (k2b_n * K2_REP) * 1e-3) / 1e12,
                    "frac": flops / (k2b_ms / (k2b_n * K2_REP) * 1e-3) / 1e12 / peak, "launches_per_event_pair": K2_REP,
                    "how": "explanatory: CUDA-event pair around 8 consecutive launches of K2 on the same operands (pm_debug_k2_repeat), "
                           "duration / 8 -- the launch latency and prologue of a launch run under its predecessor"},
                "in_chain": None if not k2_chain_us else {
                    "kernel_ms": k2_chain_us * 1e-3, "achieved": flops / (k2_chain_us * 1e-6) / 1e12,
                    "frac": flops / (k2_chain_us * 1e-6) / 1e12 / peak,
                    "share_of_step": k2_chain_us * 1e-3 / ms_step if ms_step else None,
                    "how": "explanatory: K2's own %globaltimer stamps (first CTA past griddepcontrol.wait -> last CTA out) inside the "
                           "undisturbed PDL chain, median of 15 single steps"},
                "algorithmic_flops_per_launch": flops, "mma_k_blocks_per_tile": stats["k_blocks"],
                "exact_integer_mode": stats["exact_mode"], "exact_fallback_rows": stats["fallback_rows"]}

    # ---- e2e: the host-buffer C-ABI call, pinned host memory, H2D + D2H in the timed region ---
    hq = [torch.from_numpy(q0).pin_memory()]
    ht = [torch.from_numpy(t0).pin_memory()]
    for k in range(1, 4):
        hq.append(pool[k][0].cpu().pin_memory()); ht.append(pool[k][1].cpu().pin_memory())
    hknn = torch.zeros((NQ, 2, 4), dtype=torch.int32).pin_memory()
    hgood = torch.zeros((NQ, 4), dtype=torch.int32).pin_memory()
    # one host thread per lane, each with its own context and pinned result buffers: a call is synchronous (it returns with
    # the matches on the host), so the upload of one lane's step runs under the kernels and the download of another's
    from concurrent.futures import ThreadPoolExecutor
    lane_host = [(hknn, hgood)] + [(torch.zeros_like(hknn).pin_memory(), torch.zeros_like(hgood).pin_memory()) for _ in lane_ctx[1:]]
    pool_exec = ThreadPoolExecutor(max_workers=n_lanes)
    e_steps = min(max(steps, 100), 200)          # five blocks of at least 20 calls whatever --steps is (a block of four calls
    n_host_good = 0                              # cannot show what four lanes overlap)

    def host_calls(lane, n_calls, n_active):
        c, (hk, hg) = lane_ctx[lane], lane_host[lane]
        r = 0
        for i in range(lane, n_calls, n_active):
            r = c.knn2_ratio_l2_ptr(hq[i % 4].data_ptr(), NQ, ht[i % 4].data_ptr(), NT, DIM, RATIO, hk.data_ptr(), hg.data_ptr())
        return r

    def host_block(n_calls, n_active):
        if n_active == 1:
            return host_calls(0, n_calls, 1)
        futs = [pool_exec.submit(host_calls, l, n_calls, n_active) for l in range(n_active)]
        return [f.result() for f in futs][0]

    t_warm, i = time.perf_counter() + 0.3, 0    # warm-up by time: the first few hundred calls run up to 1.4x slower (host / PCIe side ramping up)
    while time.perf_counter() < t_warm or i < 3:
        host_block(2 * n_lanes, n_lanes)
        i += 1

    # five blocks of e_steps / 5 calls; the reported figure is the MEDIAN block (host-side interference -- other
    # tenants on the PCIe switch, the nvidia-smi sampler -- moved single blocks by 2x between otherwise equal runs)
    def e2e_blocks(n_active):
        nonlocal n_host_good
        blocks, per = [], max(n_active, -(-max(1, e_steps // 5) // n_active) * n_active)
        for b in range(5):
            barrier()
            ev0.record(stream)
            n_host_good = host_block(per, n_active)
            ev1.record(stream)          # every call has returned: all lanes' work is complete
            barrier()
            emax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(emax, op=dist.ReduceOp.MAX)
            blocks.append(float(emax.item()) / per)
        return blocks, per

    e_blocks, e_per = e2e_blocks(n_lanes)
    e_one, e_one_per = (e_blocks, e_per) if n_lanes == 1 else e2e_blocks(1)
    # the faster of the two host-side arrangements is the figure (the block times are already the max over ranks, so every
    # rank decides alike): on one or two GPUs the lanes hide everything but the upload; with eight ranks on one host the
    # PCIe root / host memory is the limit and four uploading threads per rank only add contention
    e_lanes_used = n_lanes
    if float(np.median(e_one)) < float(np.median(e_blocks)):
        e_blocks, e_one, e_per, e_lanes_used = e_one, e_blocks, e_one_per, 1
    e_steps = 5 * e_per
    e_ms = float(np.median(e_blocks))
    e2e = {"value": world * NQ * NT / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms,
           "h2d_bytes_per_step": (NQ + NT) * DIM * 4, "d2h_bytes_per_step": NQ * 2 * 16 + NQ * 16 + 4,
           "api": "pm_knn2_ratio_l2_f32 (host buffers, pinned)" + (f"; {e_lanes_used} host threads, one context each" if e_lanes_used > 1 else ""),
           "steps": e_steps, "lanes": e_lanes_used, "value_per_gpu": NQ * NT / (e_ms * 1e-3),
           ("ms_per_step_one_lane" if e_lanes_used > 1 else f"ms_per_step_{n_lanes}_lanes"): float(np.median(e_one)),
           "ms_per_step_blocks": e_blocks, "timing": "median of 5 blocks, max over ranks per block"}

    # ---- the same call with SIFT shipped as bytes (pm_knn2_l2_u8: 4x fewer PCIe bytes, identical matches) ----
    def u8_leg():
        import ctypes as C
        from points_matching_b200 import _lib
        hq8 = torch.from_numpy(q0.astype(np.uint8)).pin_memory()
        ht8 = torch.from_numpy(t0.astype(np.uint8)).pin_memory()
        L = _lib.lib()

        lane_h = [c._h for c in lane_ctx]

        def u8_calls(lane, n_calls, n_active):
            out = C.c_void_p(lane_host[lane][0].data_ptr())
            for _ in range(lane, n_calls, n_active):
                st = L.pm_knn2_l2_u8(lane_h[lane], C.c_void_p(hq8.data_ptr()), NQ, C.c_void_p(ht8.data_ptr()), NT, DIM, out)
                assert st == 0, st

        def u8_block(n_calls, n_active):
            if n_active == 1:
                return u8_calls(0, n_calls, 1)
            for f in [pool_exec.submit(u8_calls, l, n_calls, n_active) for l in range(n_active)]:
                f.result()

        def u8_time(n_active):
            t_w = time.perf_counter() + 0.1
            while time.perf_counter() < t_w:
                u8_block(2 * n_active, n_active)
            per = -(-e_steps // n_active) * n_active
            barrier()
            ev0.record(stream)
            u8_block(per, n_active)
            ev1.record(stream)
            barrier()
            return ev0.elapsed_time(ev1) / per

        u8_ms = u8_time(n_lanes)
        u8_one = u8_ms if n_lanes == 1 else u8_time(1)
        e2e["u8_wire_format"] = {"value": world * NQ * NT / (u8_ms * 1e-3), "unit": UNIT, "ms_per_step": u8_ms,
                                 "h2d_bytes_per_step": (NQ + NT) * DIM, "d2h_bytes_per_step": NQ * 2 * 16,
                                 "lanes": n_lanes, "ms_per_step_one_lane": u8_one,
                                 "api": "pm_knn2_l2_u8 (kNN-2 only, host buffers, pinned)"}
        return None

    guarded("e2e.u8_wire_format", world, u8_leg)

    # ---- the other legs.  Measured pipe peaks first (FP32 FFMA for K7, POPC for K4a): BASELINE.md asks for measured numbers
    measured = guarded("measured_peaks", world, lambda: {"fp32_ffma_tflops": ctx.measure_peak(0), "popc_tera_per_s": ctx.measure_peak(1),
                                                          "how": "pm_measure_peak: independent-chain microkernels (8 chains x 2048 threads per SM), "
                                                                 "best of 5 launches, CUDA events"})
    if world > 1:                # the ctx gets its own NCCL communicator: the sharded C-ABI entries run their exchange on the ctx stream
        ctx.comm_init_from_torch()
    secondary = None
    if not args.no_ransac:
        secondary = guarded("ransac", world, lambda: bench_ransac(ctx, torch, dist, dev, world, rank, stream, barrier, peaks, args, measured))
    extra = {}
    if not args.no_hamming:
        extra["hamming"] = guarded("hamming", world, lambda: bench_hamming(ctx, torch, dist, dev, world, rank, stream, barrier, peaks, measured, args))
        extra["l2_general_floats"] = guarded("l2_general_floats", world, lambda: bench_split_mode(ctx, torch, dev, rank, stream, barrier, peaks))
        extra["lmeds"] = guarded("lmeds", world, lambda: bench_lmeds(ctx, rank, world, args))
    if not args.no_cfg5:
        extra["cfg5"] = guarded("cfg5", world, lambda: bench_cfg5(ctx, torch, dev, world, rank, barrier))
    extra = extra or None

    # ---- parity: the results of the timed steps against an independent answer --------------------------------------
    parity = guarded("parity", world, lambda: parity_l2(ctx, torch, dist, dev, world, rank, step, knn, good, ngood, pool, q0, t0))

    clocks = sampler.stop() if sampler else None

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        def cpu_leg():
            fn, kind, cores, desc = cpu_match_fn()
            rows, best = time_cpu_sample(fn, q0, t0, budget_s=4.0, reps=3)
            return {"value": rows * NT / best, "unit": UNIT, "cores": cores, "kind": kind,
                    "sample": f"{rows} of {NQ} query rows x {NT} train rows, best of 3; {desc}"}
        cpu_baseline = guarded("cpu_baseline", world, cpu_leg)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    def get(d, *path):
        for k in path:
            if not isinstance(d, dict) or d.get(k) is None:
                return None
            d = d[k]
        return d

    # the other configurations' figures where the driver keeps them: inside roofline / e2e / cpu_baseline
    ham = (extra or {}).get("hamming") or {}
    roofline["secondary"] = {
        "ransac_score_kernel": get(secondary, "roofline"),
        "hamming_popc_kernel": get(ham, "popc", "roofline"),
        "hamming_tensor_kernel": get(ham, "tensor", "roofline"),
        "l2_split_mode_kernel": get(extra, "l2_general_floats", "roofline"),
        "measured_pipe_peaks": measured}
    e2e["secondary"] = {"ransac_f": get(secondary, "e2e"), "ransac_f_seeded": get(secondary, "e2e_seeded"),
                        "cfg5_image_pairs": get(extra, "cfg5", "e2e"), "hamming_cross_check": get(ham, "e2e")}
    if cpu_baseline is not None and isinstance(cpu_baseline, dict):
        cpu_baseline["secondary"] = {"ransac_f": get(secondary, "cpu_baseline"), "hamming": get(ham, "cpu_baseline"),
                                     "lmeds": get(extra, "lmeds", "cpu_baseline")}
    parity_all = {"l2_knn": parity, "cfg3_cross_check": get(ham, "parity"), "cfg4_ransac": get(secondary, "parity"),
                  "cfg5_pairs": get(extra, "cfg5", "parity")}
    flags = [v.get("ok") for v in parity_all.values() if isinstance(v, dict) and "ok" in v]
    parity_ok = bool(flags) and all(bool(x) for x in flags)
    summary = {
        "l2_pairs_per_s": value, "l2_step_us": ms_step * 1e3, "l2_step_us_one_lane": ms_step_single * 1e3, "lanes": n_lanes,
        "k2_us_event_timed": k2_avg_ms * 1e3,
        "k2_frac_of_burst_peak": achieved / peaks["bf16_burst"], "k2_frac_of_sustained_peak": achieved / peaks["bf16_sustained"],
        "k2_us_in_chain": k2_chain_us, "k2_us_back_to_back": get(roofline, "back_to_back", "kernel_ms") and get(roofline, "back_to_back", "kernel_ms") * 1e3,
        "k2_frac_back_to_back": get(roofline, "back_to_back", "frac"), "l2_e2e_pairs_per_s": e2e["value"], "l2_e2e_u8_pairs_per_s": get(e2e, "u8_wire_format", "value"),
        "ransac_hyp_per_s": get(secondary, "value"), "ransac_k7_frac_of_measured_fp32": get(secondary, "roofline", "frac"),
        "ransac_e2e_hyp_per_s": get(secondary, "e2e", "value"), "ransac_cpu_iters_per_s": get(secondary, "cpu_baseline", "value"),
        "hamming_tensor_pairs_per_s": get(ham, "tensor", "pairs_per_s_knn_kernel"), "hamming_tensor_frac": get(ham, "tensor", "roofline", "frac"),
        "hamming_popc_pairs_per_s": get(ham, "popc", "pairs_per_s_knn_kernel"), "hamming_popc_frac_of_measured_popc": get(ham, "popc", "roofline", "frac"),
        "cfg3_cross_check_step_ms": get(ham, "ms_per_step"),
        "split_mode_step_us": (get(extra, "l2_general_floats", "ms_per_step") or 0) * 1e3 or None,
        "split_mode_frac": get(extra, "l2_general_floats", "roofline", "frac"),
        "cfg5_image_pairs_per_s": get(extra, "cfg5", "image_pairs_per_s"), "cfg5_e2e_image_pairs_per_s": get(extra, "cfg5", "e2e", "value"),
        "lmeds_models_per_s": get(extra, "lmeds", "models_per_s"),
        "fp32_peak_measured_tflops": get(measured, "fp32_ffma_tflops"), "popc_peak_measured_tera": get(measured, "popc_tera_per_s"),
        "parity_ok": parity_ok, "parity": {k: (v.get("ok") if isinstance(v, dict) else None) for k, v in parity_all.items()}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config_dict(world, n_lanes),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "parity_ok": parity_ok, "secondary": secondary, "extra": extra,
            "notes": {"good_matches_last_step": n_good_last, "clock_ramp_steps": ramp,
                      "step_overlap": (f"{n_lanes} lanes: steps round-robin over {n_lanes} contexts (own stream + workspaces), each step one whole "
                                       "K1-K2-K3-K5 pass over its own input set") if n_lanes > 1 else
                                      ("none" if args.no_pipelining else "pm_set_pipelining: K1 of step i+1 under K3/K5 of step i"),
                      "ms_per_step_one_lane": ms_step_single, "host_affinity": numa_note,
                      "dtype_detail": "bf16 operands (exact for 0..255 integers), fp32 accumulate in TMEM, fp32 norms / selection / output"},
            "summary": summary}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    return 0


def parity_l2(ctx, torch, dist, dev, world, rank, step, knn, good, ngood, pool, q0, t0):
    """The kNN rows / good matches the timed loop produces against an independent answer.  N = 1: OpenCV (cv2.batchDistance,
    else the C oracle) on 256 sampled query rows.  N > 1: rank 0 gathers 128 sampled rows of every rank's shard (queryIdx is
    global: rank * 10000 + i) and recomputes those rows itself, single-GPU, from the regenerated shard of that rank."""
    from points_matching_b200 import DMATCH, synth
    step(0)                                             # pool[0] = (q0, t0 / rank 0's train set), unpermuted
    torch.cuda.synchronize()
    rng = np.random.default_rng(3)
    rows = np.sort(rng.choice(NQ, 128 if world > 1 else 256, replace=False))
    mine = knn[torch.from_numpy(rows).to(dev)].contiguous()                   # [r, 2, 4] int32
    n_good = int(ngood[0].item())
    if world == 1:
        k = mine.cpu().numpy().view(DMATCH).reshape(len(rows), 2)
        try:
            import cv2
            d, i = cv2.batchDistance(q0[rows], t0, cv2.CV_32F, None, None, cv2.NORM_L2, 2)
            ref_idx, ref_d, who = i, d, f"cv2 {cv2.__version__} batchDistance"
        except Exception:
            from oracle import oracle as orc
            r = orc.knn2_l2(q0[rows], t0)
            ref_idx, ref_d, who = r["trainIdx"], r["distance"], "oracle/pm_oracle.c"
        idx_ok = bool(np.array_equal(k["trainIdx"], ref_idx))
        dist_ok = bool(np.array_equal(k["distance"], ref_d))                   # integer data: sqrtf of an exact integer on both sides
        kn_all = knn.cpu().numpy().view(DMATCH).reshape(NQ, 2)
        cnt_ok = n_good == int((kn_all["distance"][:, 0] < np.float32(RATIO) * kn_all["distance"][:, 1]).sum())
        return {"ok": idx_ok and dist_ok and cnt_ok, "rows": len(rows), "indices_equal": idx_ok, "distances_bit_equal": dist_ok,
                "good_count_consistent": cnt_ok, "against": who}
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    ok = True
    if rank == 0:
        out = torch.zeros((len(rows), 2, 4), dtype=torch.int32, device=dev)
        dt_ = pool[0][1]
        for r in range(world):
            qr, _ = synth.sift_pair(NQ, NT, seed=1234 + r)
            dq = torch.from_numpy(np.ascontiguousarray(qr[rows])).to(dev)
            torch.cuda.synchronize()
            ctx.knn2_l2_f32_dev(dq.data_ptr(), len(rows), dt_.data_ptr(), NT, DIM, out.data_ptr(), 0)
            ctx.sync()
            a = parts[r].cpu().numpy().view(DMATCH).reshape(len(rows), 2)
            b = out.cpu().numpy().view(DMATCH).reshape(len(rows), 2)
            ok = ok and np.array_equal(a["trainIdx"], b["trainIdx"]) and np.array_equal(a["distance"], b["distance"]) \
                and np.array_equal(a["queryIdx"][:, 0], r * NQ + rows)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"ok": bool(flag.item()), "rows_per_rank": len(rows),
            "against": "the same rows recomputed on rank 0 alone from the regenerated shard of every rank (global queryIdx checked)"}


def guarded(name, world, fn):
    print(f"bench.py: leg '{name}' ...", file=sys.stderr, flush=True)
    return _guarded(name, world, fn)


def _guarded(name, world, fn):
    """The legs after the headline (RANSAC, Hamming, general floats, cfg5) must not cost the JSON line: on a single GPU a
    failing leg is reported in its place.  With several ranks an exception on one of them would leave the others in a
    barrier, so there it propagates and the run fails fast."""
    if world > 1:
        return fn()
    try:
        return fn()
    except Exception as e:      # noqa: BLE001
        print(f"bench.py: leg '{name}' failed: {e!r}", file=sys.stderr)
        return {"error": repr(e)}


def measure_fp8_peak(torch, dev):
    """Dense e4m3 GEMM throughput of this GPU through the library path (torch._scaled_mm -> cuBLASLt), 8192^3, best of
    10: the burst figure, which is what a 0.2 ms kernel timed alone should be held against (MEASURED_PEAKS.json
    carries bf16 only)."""
    try:
        n = 8192
        a = torch.randn((n, n), device=dev).to(torch.float8_e4m3fn)
        b = torch.randn((n, n), device=dev).to(torch.float8_e4m3fn).t()      # column-major operand
        one = torch.ones((), device=dev, dtype=torch.float32)
        for _ in range(3):
            torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
        best = 1e9
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            e0.record(); torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "measured here: torch._scaled_mm e4m3 8192^3, best of 10 (burst, kernel timed alone)"
    except Exception as e:
        print("bench.py: fp8 peak measurement unavailable:", e, file=sys.stderr)
        return None, None


def bench_hamming(ctx, torch, dist, dev, world, rank, stream, barrier, peaks, measured, args):
    """cfg3's per-GPU shard (12.5k of 100k ORB queries per rank x 100k train rows, 256-bit): BFMatcher(NORM_HAMMING,
    crossCheck=true).match through pm_match_cross_sharded_dev -- kNN-2 of the shard, column minima of the shard,
    (N > 1) ncclAllReduce(min, u64) of the 800 KB packed column minima on the ctx stream, cross-check filter.  Both kNN
    kernels are timed: the POPC kernel the north_star names, and the tensor-core kernel (default for large problems)
    that runs K2's GEMM on bits expanded to E4M3 operands."""
    import points_matching_b200 as pm
    from points_matching_b200 import DMATCH, _lib, synth
    nq, nt = 12500, 100000
    qall, t = synth.orb_pair(nq * world, nt, seed=4321)          # train set replicated, query rows sharded
    q = np.ascontiguousarray(qall[rank * nq:(rank + 1) * nq])
    dq, dt_ = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
    knn = torch.zeros((nq, 2, 4), dtype=torch.int32, device=dev)
    col = torch.zeros(nt, dtype=torch.int64, device=dev)
    out = torch.zeros((nq, 4), dtype=torch.int32, device=dev)
    cnt = torch.zeros(4, dtype=torch.int32, device=dev)

    def step():
        ctx.match_cross_sharded_dev(dq.data_ptr(), nq, dt_.data_ptr(), nt, 32, pm.NORM_HAMMING, rank * nq, knn.data_ptr(),
                                    col.data_ptr(), out.data_ptr(), cnt.data_ptr())

    pairs = float(nq) * nt
    res = {"workload": "cfg3 shard: ORB-like 256-bit, 12500 query rows per rank x 100000 train rows, pm_match_cross_sharded_dev = kNN-2 + "
                       "reverse pass over the train rows that are some query's best match (N > 1: ncclAllReduce(max, u8) of the 100 KB "
                       "mark bytes first) + (N > 1: ncclAllReduce(min, u64) of the 800 KB packed column minima) + cross-check, all inside "
                       "the C ABI; pairs_per_s_step counts the nq x nt query-train pairs of the step once"}
    for name, path in (("popc", 1), ("tensor", 2)):
        _lib.lib().pm_debug_hamming_path(path)
        for _ in range(2):
            step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        ev0.record(stream)
        for _ in range(steps):
            step()
        ev1.record(stream)
        barrier()
        tmax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item()) / steps
        n_mutual = int(cnt[0].item())
        # the matching kernel's own time: the forward kNN-2 of the shard alone (nq x nt pairs per launch), event pair around
        # every launch of the kernel (the step's reverse pass runs the same kernel on the marked train rows only)
        for _ in range(2):
            ctx.knn2_hamming_dev(dq.data_ptr(), nq, dt_.data_ptr(), nt, 32, knn.data_ptr(), rank * nq)
        ctx.sync()
        ctx.profile_enable(True)
        for _ in range(steps):
            ctx.knn2_hamming_dev(dq.data_ptr(), nq, dt_.data_ptr(), nt, 32, knn.data_ptr(), rank * nq)
        ctx.sync()
        k4_ms, k4_n = ctx.profile_read(1)
        ctx.profile_enable(False)
        k4_avg = k4_ms / max(k4_n, 1)
        r = {"ms_per_step": ms, "kernel_ms": k4_avg, "pairs_per_s_knn_kernel": pairs / (k4_avg * 1e-3) if k4_n else None,
             "pairs_per_s_step": world * pairs / (ms * 1e-3), "mutual_matches": n_mutual}
        if name == "popc":
            popc_meas = (measured or {}).get("popc_tera_per_s") if isinstance(measured, dict) else None
            popc_rate = popc_meas * 1e12 if popc_meas else 148 * 16 * 1.965e9
            popc_peak = popc_rate / 8                    # pairs/s at 8 POPC.32 per 256-bit pair (the algorithmic count)
            r["roofline"] = {"bound": "popc_issue", "peak_pairs_per_s": popc_peak, "peak_popc_per_s": popc_rate,
                             "frac": (pairs / (k4_avg * 1e-3)) / popc_peak if k4_n else None,
                             "frac_as_issued": (pairs * 5 / (k4_avg * 1e-3)) / popc_rate if k4_n else None,
                             "peak_source": ("measured here (pm_measure_peak: POPC.32 issue rate)" if popc_meas else "nominal 148 SM x 16 POPC/clk x 1965 MHz") +
                                            "; 8 POPC per pair is the algorithmic count -- the kernel issues 5 after its carry-save adders "
                                            "(frac_as_issued counts those), so `frac` may exceed 1",
                             "hbm_equivalent": {"bytes_per_pair": 64, "achieved_gbs": pairs * 64 / (k4_avg * 1e-3) / 1e9 if k4_n else None,
                                                "peak_gbs": peaks["hbm"]}}
        else:
            tf = pairs * 512 / (k4_avg * 1e-3) / 1e12 if k4_n else None     # 2 x 256 FLOP per pair
            fp8_peak, fp8_src = measure_fp8_peak(torch, dev)
            if fp8_peak is None:
                fp8_peak = 2 * peaks["bf16_burst"]
                fp8_src = ("derived: 2 x the measured burst bf16 peak (fp8 e4m3 is twice the bf16 rate; MEASURED_PEAKS.json has "
                           "no fp8 entry and torch._scaled_mm was not usable here)")
            r["roofline"] = {"bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "peak": fp8_peak,
                             "frac": tf / fp8_peak if tf else None, "peak_source": fp8_src,
                             "kernel": "l2_tc_kernel<FP8> (tcgen05.mma kind::f8f6f4 E4M3, M128 N128 K32, same fused top-2 epilogue)"}
        res[name] = r
    _lib.lib().pm_debug_hamming_path(0)
    # parity: the sharded list, gathered (pm_allgather_matches_dev), against the whole problem on rank 0 alone / the C oracle
    width = nq
    allm = torch.zeros((world, width, 4), dtype=torch.int32, device=dev)
    counts = torch.zeros(world, dtype=torch.int32, device=dev)
    step()
    ctx.allgather_matches_dev(out.data_ptr(), cnt.data_ptr(), width, allm.data_ptr(), counts.data_ptr())
    ctx.sync()
    ok, against = True, None
    if rank == 0:
        c = counts.cpu().numpy()
        got = np.concatenate([allm[r, : int(c[r])].cpu().numpy() for r in range(world)]).view(DMATCH).reshape(-1)
        if world > 1:
            one = pm.Context(dev.index)
            dqa = torch.from_numpy(qall).to(dev)
            k1 = torch.zeros((nq * world, 2, 4), dtype=torch.int32, device=dev)
            c1 = torch.zeros(nt, dtype=torch.int64, device=dev)
            o1 = torch.zeros((nq * world, 4), dtype=torch.int32, device=dev)
            n1 = torch.zeros(4, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            one.match_cross_sharded_dev(dqa.data_ptr(), nq * world, dt_.data_ptr(), nt, 32, pm.NORM_HAMMING, 0, k1.data_ptr(),
                                        c1.data_ptr(), o1.data_ptr(), n1.data_ptr())
            one.sync()
            ref = o1[: int(n1[0].item())].cpu().numpy().view(DMATCH).reshape(-1)
            one.close()
            against = "the whole 100k x 100k problem on rank 0 alone (single-GPU pm_match_cross_sharded_dev)"
            ok = bool(np.array_equal(got, ref))
        else:
            # one GPU: 2000 sampled query rows of the kNN against the C oracle, and the mutual list against the oracle's rule
            from oracle import oracle as orc
            rows = np.sort(np.random.default_rng(4).choice(nq, 1000, replace=False))
            refk = orc.knn2_hamming(q[rows], t)
            kk = knn.cpu().numpy().view(DMATCH).reshape(nq, 2)[rows]
            ok = bool(np.array_equal(kk["trainIdx"], refk["trainIdx"]) and np.array_equal(kk["distance"], refk["distance"]))
            colh = col.cpu().numpy().view(np.uint64)
            fwd = knn.cpu().numpy().view(DMATCH).reshape(nq, 2)[:, 0]
            mutual = (colh[fwd["trainIdx"]] & np.uint64(0xFFFFFFFF)) == fwd["queryIdx"].astype(np.uint64)
            ok = ok and bool(np.array_equal(got["queryIdx"], fwd["queryIdx"][mutual])) and int(mutual.sum()) == len(got)
            against = "oracle/pm_oracle.c on 1000 sampled query rows (bit-exact) + the mutual-nearest rule re-evaluated on the host"
    if world > 1:
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    res["parity"] = {"ok": ok, "against": against, "mutual_matches_total": int(counts.sum().item())}
    # e2e: the host-buffer call (pinned), H2D of the shard + train set and D2H of the matches inside the timed region
    hq, ht = torch.from_numpy(q).pin_memory(), torch.from_numpy(t).pin_memory()
    hout = torch.zeros((nq, 4), dtype=torch.int32).pin_memory()
    import ctypes as C
    L = _lib.lib()
    nout = C.c_int(0)

    def host_call():
        st = L.pm_match_cross_hamming(ctx._h, C.c_void_p(hq.data_ptr()), nq, C.c_void_p(ht.data_ptr()), nt, 32, C.c_void_p(hout.data_ptr()), C.byref(nout))
        assert st == 0, st

    for _ in range(2):
        host_call()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(5):
        host_call()
    ev1.record(stream)
    barrier()
    e_ms = ev0.elapsed_time(ev1) / 5
    res["e2e"] = {"value": pairs / (e_ms * 1e-3), "unit": "pairs/s (query-train pairs cross-checked)", "ms_per_step": e_ms,
                  "h2d_bytes_per_step": (nq + nt) * 32, "d2h_bytes_per_step": int(nout.value) * 16 + 4,
                  "api": "pm_match_cross_hamming (host buffers, pinned; this rank's shard against the train set)"}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            import cv2
            cv2.setNumThreads(os.cpu_count() or 1)
            rows = 1500
            t0_ = time.perf_counter()
            cv2.batchDistance(q[:rows], t, cv2.CV_32S, None, None, cv2.NORM_HAMMING, 2)
            dt = time.perf_counter() - t0_
            res["cpu_baseline"] = {"value": rows * nt / dt, "unit": "pairs/s", "cores": cv2.getNumThreads(), "kind": "reference",
                                   "sample": f"cv2 {cv2.__version__} batchDistance(NORM_HAMMING, K=2) = BFMatcher(NORM_HAMMING).knnMatch, "
                                             f"{rows} of 12500 query rows x 100000 train rows"}
        except Exception as e:   # noqa: BLE001
            res["cpu_baseline"] = {"unavailable": str(e)}
    # keys of the default path at the top level (show_bench / earlier rounds read these)
    d = res["tensor"]
    res.update({"ms_per_step": d["ms_per_step"], "kernel_ms": d["kernel_ms"], "pairs_per_s_knn_kernel": d["pairs_per_s_knn_kernel"],
                "pairs_per_s_step": d["pairs_per_s_step"], "mutual_matches": d["mutual_matches"], "roofline": d["roofline"],
                "default_path": "tensor"})
    return res


def bench_split_mode(ctx, torch, dev, rank, stream, barrier, peaks):
    """The same cfg2-sized step on SURF-like descriptors (unit-norm signed floats -- what the reference's SURF extractor
    produces, main.cpp:37-40): bf16 is lossy there, so K2 runs three bf16 products per pair and K3 re-ranks in FP32 and
    certifies; uncertified rows get an exact scan (helper blocks of the filter kernel)."""
    from points_matching_b200 import synth
    q, t = synth.surf_pair(NQ, NT, seed=77 + rank)
    dq, dt_ = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
    knn = torch.zeros((NQ, 2, 4), dtype=torch.int32, device=dev)
    good = torch.zeros((NQ, 4), dtype=torch.int32, device=dev)
    ng = torch.zeros(4, dtype=torch.int32, device=dev)

    def step():
        ctx.knn2_ratio_l2_f32_dev(dq.data_ptr(), NQ, dt_.data_ptr(), NT, DIM, RATIO, knn.data_ptr(), good.data_ptr(), ng.data_ptr(), 0)

    for _ in range(20):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    ev0.record(stream)
    for _ in range(n):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1) / n
    ctx.profile_enable(True)
    for _ in range(20):
        step()
    k2_ms, k2_n = ctx.profile_read(0)
    ctx.profile_enable(False)
    k2_avg = k2_ms / max(k2_n, 1)
    st = ctx.l2_stats()
    flops = 2.0 * DIM * NQ * NT
    tf = flops / (k2_avg * 1e-3) / 1e12 if k2_n else None
    return {"workload": "SURF-like unit-norm float descriptors 10000 x 10000 x 128, L2 kNN-2 + ratio 0.75, one call per step (same inputs every step: L2-resident)",
            "ms_per_step": ms, "pairs_per_s": NQ * NT / (ms * 1e-3), "exact_integer_mode": st["exact_mode"],
            "mma_k_blocks_per_tile": st["k_blocks"], "exact_fallback_rows": st["fallback_rows"],
            "roofline": {"bound": "tensor", "kernel": "l2_tc_kernel, split mode (three bf16 products per pair: hi.hi + hi.lo + lo.hi)",
                         "kernel_ms": k2_avg, "achieved": tf, "unit": "TFLOP/s (algorithmic: 256 FLOP per pair; the extra passes count as overhead)",
                         "peak": peaks["bf16_burst"], "frac": tf / peaks["bf16_burst"] if tf else None,
                         "peak_source": peaks["source"] + ", burst bf16"}}


def bench_lmeds(ctx, rank, world, args):
    """The reference's LITERAL estimator call (main.cpp:95-98: findFundamentalMat(pts1, pts2, CV_FM_7POINT) with N > 7 =
    LMedS over 7-point samples) through the look-alike entry pm_find_fundamental_mat, host buffers."""
    import points_matching_b200 as pm
    from points_matching_b200 import synth
    n = 2000
    p1, p2, gt = synth.correspondences(n, seed=3, outlier_frac=0.3)
    for _ in range(2):
        F, mask = pm.findFundamentalMat(p1, p2, pm.FM_7POINT, ctx=ctx)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        F, mask = pm.findFundamentalMat(p1, p2, pm.FM_7POINT, ctx=ctx)
    ms = (time.perf_counter() - t0) / reps * 1e3
    niters = 300                                     # RANSACUpdateNumIters(0.99, 0.45, 7, 1000)
    out = {"workload": f"findFundamentalMat(FM_7POINT) on {n} correspondences (30% outliers) = LMedS, {niters} 7-point samples (<= {3 * niters} models), "
                       "median of the symmetric-epipolar error per model (FP64, radix select), host buffers; wall clock per call",
           "ms_per_call": ms, "models_per_s": 3 * niters / (ms * 1e-3), "inliers": int(mask.sum()), "inlier_recall_on_ground_truth": float(mask[gt].mean())}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            import cv2
            t0 = time.perf_counter()
            for _ in range(3):
                cv2.findFundamentalMat(p1, p2, cv2.FM_7POINT)
            dt = (time.perf_counter() - t0) / 3
            out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "calls/s", "ms_per_call": dt * 1e3, "cores": 1, "kind": "reference",
                                   "sample": f"cv2 {cv2.__version__} findFundamentalMat(FM_7POINT) on the same {n} correspondences (single-threaded)"}
        except Exception as e:   # noqa: BLE001
            out["cpu_baseline"] = {"unavailable": str(e)}
    return out


def _say(msg):
    print(f"bench.py:   {msg}", file=sys.stderr, flush=True)


def bench_cfg5(ctx, torch, dev, world, rank, barrier):
    """BASELINE config 5 (1024 pairs x 8k SIFT, match + RANSAC-F per pair, pairs partitioned across ranks): every
    rank runs its 1024 / world pairs, cycling through 4 distinct synthetic pairs (8192 x 8192 x 128), 4096 8-point
    hypotheses per pair, Sampson 1 px, refit.  Reports image pairs per second: device-resident (f32 descriptors in HBM)
    and end to end (u8 descriptors + keypoints in pinned host memory, uploaded inside the timed region)."""
    from points_matching_b200 import synth
    n, pool, hpool = 8192, [], []
    for k in range(4):
        d1, d2, k1, k2, _ = synth.image_pair(n, n, seed=100 + 10 * rank + k)
        pool.append(tuple(torch.from_numpy(a).to(dev) for a in (d1, d2, k1, k2)))
        hpool.append((torch.from_numpy(d1.astype(np.uint8)).pin_memory(), torch.from_numpy(d2.astype(np.uint8)).pin_memory(),
                      torch.from_numpy(k1).pin_memory(), torch.from_numpy(k2).pin_memory()))
    import points_matching_b200 as pm
    from points_matching_b200.pipeline import match_and_estimate_batch_native
    plist = [pool[p % 4] for p in range(1024)]      # the whole batch, identical on every rank: the calls below take
    pairs = len(plist) // world                     # this rank's contiguous share (shard_bounds) and nothing else

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        last = fn()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        tmax = torch.tensor([wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()), last

    # (a) the C ABI's batched entry: one call enqueues every pair, no host round trip inside a pair
    _say("cfg5: inputs ready")
    nctx = pm.Context(dev.index)
    lanes = int(os.environ.get("PM_BENCH_LANES", "0")) or 4       # groups of pairs in flight (internal streams of the batched call)
    nctx.set_batch_lanes(lanes)
    nctx.batch_warmup(n, n, 128, False, 4096)       # lanes and their workspaces exist before anything is timed
    match_and_estimate_batch_native(nctx, plist[:8 * world], n_hyp=4096)      # warm-up: 8 pairs on every rank
    _say("cfg5: warm")
    ms_n, res_n = timed(lambda: match_and_estimate_batch_native(nctx, plist, n_hyp=4096))
    last_n = res_n[-1][1]
    _say("cfg5: device-resident batch done")
    # (b) end to end from host memory: u8 descriptors (SIFT's native range) + keypoints in pinned buffers, every pair uploaded
    #     inside the timed region by the lane that processes it; 96 bytes per pair come back
    lo = rank * pairs
    hsel = [hpool[(lo + p) % 4] for p in range(pairs)]
    hargs = ([h[0].data_ptr() for h in hsel], [n] * pairs, [h[1].data_ptr() for h in hsel], [n] * pairs, 128, True,
             [h[2].data_ptr() for h in hsel], [h[3].data_ptr() for h in hsel], 0.75, 4096)
    nctx.batch_warmup(n, n, 128, True, 4096)
    nctx.match_estimate_batched(*[a[:16] if isinstance(a, list) else a for a in hargs], seed=lo)
    _say("cfg5: host batch warm")
    ms_e, rec = timed(lambda: nctx.match_estimate_batched(*hargs, seed=lo))
    _say("cfg5: host batch done")
    # parity: the host-buffer u8 run and the device-resident f32 run are the same pairs with the same seeds -> identical records
    same = all(int(r["n_matches"]) == o["n_matches"] and int(r["n_inliers"]) == o["n_inliers"] and
               (o["F"] is None) == (int(r["has_model"]) == 0) and (o["F"] is None or np.array_equal(o["F"], r["F"].reshape(3, 3)))
               for r, (_, o) in zip(rec, res_n))
    # ... and one pair against the oracle: same good matches; F at least as good as the oracle's RANSAC on the planted matches
    from oracle import oracle as orc
    d1, d2, k1, k2, (qi, _) = synth.image_pair(n, n, seed=100 + 10 * rank)
    o0 = res_n[0][1]
    rows = np.sort(np.random.default_rng(6).choice(n, 64, replace=False))
    gref = orc.ratio_filter(orc.knn2_l2(d1[rows], d2), 0.75)
    # (the sampled rows' verdicts must be a sub-list of the pair's good matches: count them through a second, sampled GPU call)
    gsub = ctx.ratio_filter(ctx.knn2(d1[rows], d2, pm.NORM_L2), 0.75)
    oracle_ok = len(gsub) == len(gref) and bool(np.array_equal(gsub["trainIdx"], gref["trainIdx"])) and o0["F"] is not None
    if oracle_ok:
        gall = ctx.ratio_filter(ctx.knn2(d1, d2, pm.NORM_L2), 0.75)
        planted = np.isin(gall["queryIdx"], qi)
        oracle_ok = o0["n_matches"] == len(gall) and float(orc.sampson_f64(o0["F"], k1[gall["queryIdx"]][planted], k2[gall["trainIdx"]][planted]).mean()) < 0.5
    ok = bool(same and oracle_ok)
    _say("cfg5: parity done")
    if world > 1:
        import torch.distributed as dist
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    nctx.close()
    return {"workload": f"cfg5: 1024 image pairs, {pairs} per rank (4 distinct synthetic pairs cycled), 8192 x 8192 SIFT-like descriptors, "
                        "kNN-2 + ratio 0.75 + gather + RANSAC-F (4096 8-point hypotheses, Sampson 1 px, refit) per pair; host wall clock "
                        "around the whole batch, max over ranks",
            "image_pairs_per_s": world * pairs / (ms_n * 1e-3), "ms_per_pair": ms_n / pairs,
            "native_batched": {"api": "pm_match_estimate_batched_dev (f32 descriptors resident in HBM, one call, match count stays on the device)",
                               "lanes": lanes, "ms_per_pair": ms_n / pairs, "image_pairs_per_s": world * pairs / (ms_n * 1e-3)},
            "e2e": {"value": world * pairs / (ms_e * 1e-3), "unit": "image pairs/s", "ms_per_pair": ms_e / pairs,
                    "h2d_bytes_per_step": 2 * n * 128 + 2 * n * 8, "d2h_bytes_per_step": 96, "step": "one image pair",
                    "api": "pm_match_estimate_batched (u8 descriptors + keypoints in pinned host memory; every lane uploads its pairs under "
                           "the other lanes' kernels)"},
            "parity": {"ok": ok, "host_u8_equals_device_f32": bool(same), "vs_oracle": bool(oracle_ok),
                       "against": "records of the two paths pair by pair; one pair's good matches (sampled rows) against oracle/pm_oracle.c and its F on the planted matches"},
            "last_pair": {"n_matches": last_n["n_matches"], "n_inliers": last_n["n_inliers"]},
            "full_config_s": ms_n * 1e-3, "full_config_e2e_s": ms_e * 1e-3}


def bench_ransac(ctx, torch, dist, dev, world, rank, stream, barrier, peaks, args, measured):
    import points_matching_b200 as pm
    from points_matching_b200 import synth
    p1, p2, _ = synth.correspondences(R_N, seed=0)
    nh = R_HYP // world                                # hypotheses shard by batch
    idx_all = synth.sample_index_sets(R_N, R_HYP, 8, seed=99)     # identical on every rank
    d1, d2 = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
    ds = torch.from_numpy(idx_all).to(dev)                       # the FULL index array on every rank: the winner is re-solved locally
    dF = torch.zeros(16, dtype=torch.float64, device=dev)
    dmask = torch.zeros(R_N, dtype=torch.uint8, device=dev)
    dn = torch.zeros(4, dtype=torch.int32, device=dev)
    dkey = torch.zeros(2, dtype=torch.int64, device=dev)

    def step():
        # solve + score this rank's slice, ncclAllReduce(max, u64) of the 8-byte winner key on the ctx stream (N > 1), local
        # re-solve of the winner, mask, refit -- one C-ABI call, nothing synchronised
        ctx.find_fundamental_sharded_dev(d1.data_ptr(), d2.data_ptr(), R_N, ds.data_ptr(), nh * world, rank * nh, nh, 8,
                                         pm.METRIC_SAMPSON, R_THR, True, dF.data_ptr(), dmask.data_ptr(), dn.data_ptr(), dkey.data_ptr())

    r_steps = max(2, min(args.steps, args.ransac_steps))
    for _ in range(2):
        step()
    barrier()
    ctx.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(r_steps):
        step()
    ev1.record(stream)
    barrier()
    k7_ms, k7_n = ctx.profile_read(2)
    ctx.profile_enable(False)
    tmax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item()) / r_steps
    n_inl = int(dn[0].item())
    key_sharded = int(dkey[0].item())
    F_sharded = dF[:9].cpu().numpy().copy()
    # k7_n counts the tiny winner re-score launches too? no: only ransac_score_kernel is bracketed; one (large) launch per step
    # parity: N > 1 -> rank 0 runs all 2^20 hypotheses alone; N == 1 -> inlier counts of sampled hypotheses against the C oracle
    ok, against = True, None
    if world > 1:
        if rank == 0:
            one = pm.Context(dev.index)
            F1 = torch.zeros(16, dtype=torch.float64, device=dev); m1 = torch.zeros(R_N, dtype=torch.uint8, device=dev)
            n1 = torch.zeros(4, dtype=torch.int32, device=dev); k1 = torch.zeros(2, dtype=torch.int64, device=dev)
            torch.cuda.synchronize()
            one.find_fundamental_dev(d1.data_ptr(), d2.data_ptr(), R_N, ds.data_ptr(), nh * world, 8, pm.METRIC_SAMPSON, R_THR, True,
                                     F1.data_ptr(), m1.data_ptr(), n1.data_ptr(), k1.data_ptr(), 0)
            one.sync()
            ok = int(k1[0].item()) == key_sharded and int(n1[0].item()) == n_inl and bool(np.array_equal(F1[:9].cpu().numpy(), F_sharded)) \
                and bool(torch.equal(m1, dmask))
            one.close()
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        against = "all 2^20 hypotheses on rank 0 alone (pm_find_fundamental_dev): winner key, inlier count, F and mask bit for bit"
    else:
        from oracle import oracle as orc
        hs = np.sort(np.random.default_rng(8).choice(R_HYP, 48, replace=False))
        sub = torch.from_numpy(np.ascontiguousarray(idx_all[hs])).to(dev)
        Fm = torch.zeros((len(hs), 12), dtype=torch.float32, device=dev)
        cn = torch.zeros(len(hs), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ctx.ransac_solve_dev(d1.data_ptr(), d2.data_ptr(), R_N, sub.data_ptr(), len(hs), 8, Fm.data_ptr())
        ctx.ransac_score_dev(d1.data_ptr(), d2.data_ptr(), R_N, Fm.data_ptr(), len(hs), R_THR, pm.METRIC_SAMPSON, cn.data_ptr())
        ctx.sync()
        Fh, ch = Fm.cpu().numpy()[:, :9], cn.cpu().numpy()
        ref = np.array([orc.count_inliers_f32(Fh[i], p1, p2, R_THR, 0) for i in range(len(hs))])
        ok = bool(np.array_equal(ch, ref))
        against = "oracle/pm_oracle.c: inlier counts of 48 sampled hypotheses over all 100k correspondences, bit-exact on identical F bits"
    # e2e: host call with pinned correspondences and host index sets (H2D 33.6 MB for 1M x 8 indices)
    idx = np.ascontiguousarray(idx_all[rank * nh:(rank + 1) * nh])
    h1, h2 = torch.from_numpy(p1).pin_memory(), torch.from_numpy(p2).pin_memory()
    hs_ = torch.from_numpy(idx).pin_memory()
    hF = torch.zeros(9, dtype=torch.float64).pin_memory()
    hmask = torch.zeros(R_N, dtype=torch.uint8).pin_memory()

    def time_host(sample_ptr):
        ctx.find_fundamental_ptr(h1.data_ptr(), h2.data_ptr(), R_N, sample_ptr, nh, 8, pm.METRIC_SAMPSON, R_THR, True, hF.data_ptr(), hmask.data_ptr())
        barrier()
        ev0.record(stream)
        for _ in range(2):
            ctx.find_fundamental_ptr(h1.data_ptr(), h2.data_ptr(), R_N, sample_ptr, nh, 8, pm.METRIC_SAMPSON, R_THR, True, hF.data_ptr(), hmask.data_ptr())
        ev1.record(stream)
        barrier()
        emax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(emax, op=dist.ReduceOp.MAX)
        return float(emax.item()) / 2

    e_ms = time_host(hs_.data_ptr())
    s_ms = time_host(None)                              # index sets generated on the device from the seed: only the points cross PCIe
    evals = float(nh) * R_N
    k7_avg = k7_ms / max(k7_n, 1)
    fp32_meas = (measured or {}).get("fp32_ffma_tflops") if isinstance(measured, dict) else None
    fp32_peak = fp32_meas if fp32_meas else 148 * 128 * 2 * 1.965e9 / 1e12
    out = {"metric": "ransac_f_hypotheses_per_sec", "value": world * nh / (ms * 1e-3), "unit": "hypotheses/s",
           "ms_per_step": ms, "steps": r_steps,
           "config": {"workload": "cfg4: 100k correspondences, 50% outliers, 8-point samples, Sampson thr 1 px, "
                                  "2^20 hypotheses (sharded by batch across ranks), refit on inliers",
                      "n_points": R_N, "n_hyp_total": nh * world, "winner_inliers": n_inl,
                      "api": "pm_find_fundamental_sharded_dev (N > 1: ncclAllReduce(max, u64) of the winner key inside the C ABI)"},
           "e2e": {"value": world * nh / (e_ms * 1e-3), "unit": "hypotheses/s", "ms_per_step": e_ms,
                   "h2d_bytes_per_step": 2 * R_N * 8 + nh * 8 * 4, "d2h_bytes_per_step": 72 + R_N + 32,
                   "api": "pm_find_fundamental (host buffers, pinned; index sets precomputed on the host)"},
           "e2e_seeded": {"value": world * nh / (s_ms * 1e-3), "unit": "hypotheses/s", "ms_per_step": s_ms,
                          "h2d_bytes_per_step": 2 * R_N * 8, "d2h_bytes_per_step": 72 + R_N + 32,
                          "api": "pm_find_fundamental, sample_idx = NULL (index sets generated on the device from the seed)"},
           "parity": {"ok": ok, "against": against},
           "roofline": {"bound": "fp32_issue", "kernel": "ransac_score_kernel", "kernel_ms": k7_avg,
                        "achieved": evals * 33 / (k7_avg * 1e-3) / 1e12 if k7_n else None, "unit": "TFLOP/s",
                        "peak": fp32_peak,
                        "peak_source": ("measured here (pm_measure_peak: FP32 FFMA issue rate x 2 FLOP)" if fp32_meas else
                                        "nominal 148 SM x 128 FMA/clk x 1965 MHz"),
                        "frac": (evals * 33 / (k7_avg * 1e-3) / 1e12) / fp32_peak if k7_n else None,
                        "flop_per_eval": 33, "fp32_instructions_per_eval": 18, "evals_per_launch": evals,
                        "note": "33 FLOP in 18 FP32 instructions (15 FMA + 3 MUL) + 1 IMAD: the instruction-bound ceiling of `frac` is 33 / (2 x 19) = 0.87",
                        "hbm_equivalent": {"bytes_per_eval": 16, "achieved_gbs": evals * 16 / (k7_avg * 1e-3) / 1e9 if k7_n else None,
                                           "peak_gbs": peaks["hbm"]}}}
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            import cv2
            cv2.setNumThreads(os.cpu_count() or 1)
            iters = 400
            t0 = time.perf_counter()
            cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, R_THR, 1 - 1e-15, iters)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": iters / dt, "unit": "iterations/s", "cores": 1, "kind": "reference",
                                   "sample": f"cv2 {cv2.__version__} findFundamentalMat(FM_RANSAC, 1 px, conf 1-1e-15, maxIters {iters}) "
                                             f"on the same 100k correspondences: single-threaded, 7-point samples with <=3 models each"}
        except Exception as e:   # noqa: BLE001
            out["cpu_baseline"] = {"unavailable": str(e)}
    return out


def start_watchdog(limit_s):
    """A GPU-side hang would block the process inside a CUDA synchronisation forever; the driver's clock is better spent
    elsewhere: after limit_s seconds the process reports and exits."""
    import threading

    def bark():
        print(f"bench.py: watchdog: still running after {limit_s} s -- giving up", file=sys.stderr, flush=True)
        try:
            import faulthandler
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        except Exception:
            pass
        os._exit(2)

    t = threading.Timer(limit_s, bark)
    t.daemon = True
    t.start()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-ransac", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--ransac-steps", type=int, default=5)
    ap.add_argument("--no-pipelining", action="store_true", help="consecutive steps strictly serial (no cross-step overlap)")
    ap.add_argument("--no-affinity", action="store_true", help="N > 1: do not bind the rank to its GPU's local cores")
    ap.add_argument("--lanes", type=int, default=4, help="contexts (stream + workspaces each) the timed steps and the e2e calls go round-robin over")
    ap.add_argument("--no-ramp", action="store_true", help="skip the 1 s clock ramp (profiling runs under ncu)")
    args = ap.parse_args()
    start_watchdog(float(os.environ.get("PM_BENCH_WATCHDOG_S", "1200")))
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
