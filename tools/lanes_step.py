#!/usr/bin/env python
"""cfg2 step throughput with L contexts (own stream + workspaces each) taking the steps in turn.
Usage: python tools/lanes_step.py [steps]"""
import ctypes as C
import sys
import time

import numpy as np
import torch

import points_matching_b200 as pm
from points_matching_b200 import _lib, synth

NQ, NT, DIM, RATIO = 10000, 10000, 128, 0.75
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda", 0)
q0, t0 = synth.sift_pair(NQ, NT, seed=1234)
gq, gt = torch.from_numpy(q0).to(dev), torch.from_numpy(t0).to(dev)
fn = _lib.lib().pm_knn2_ratio_l2_f32_dev
for L in (1, 2, 3, 4):
    for pipe in (False, True):
        lanes = []
        for l in range(L):
            ctx = pm.Context(0)
            st = torch.cuda.Stream(device=dev)
            ctx.set_stream(st.cuda_stream)
            ctx.set_pipelining(pipe)
            knn = torch.zeros((NQ, 2, 4), dtype=torch.int32, device=dev)
            good = torch.zeros((NQ, 4), dtype=torch.int32, device=dev)
            ngood = torch.zeros(4, dtype=torch.int32, device=dev)
            args = (ctx._h, C.c_void_p(gq.data_ptr()), C.c_int(NQ), C.c_void_p(gt.data_ptr()), C.c_int(NT), C.c_int(DIM),
                    C.c_float(RATIO), C.c_int(0), C.c_void_p(knn.data_ptr()), C.c_void_p(good.data_ptr()), C.c_void_p(ngood.data_ptr()))
            lanes.append((ctx, st, args, knn, good, ngood))
        for i in range(400):
            assert fn(*lanes[i % L][2]) == 0
        torch.cuda.synchronize()
        t = time.perf_counter()
        for i in range(steps):
            fn(*lanes[i % L][2])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        ng = [int(x[5][0].item()) for x in lanes]
        print(f"lanes {L} pipelining {int(pipe)}: {dt / steps * 1e6:.2f} us per step, good {ng}", flush=True)
        for x in lanes:
            x[0].set_pipelining(False)
        del lanes
