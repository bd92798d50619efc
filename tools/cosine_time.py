"""Times the 20k x 20k x 256 cosine force match (BASELINE configs[5]): whole call with CUDA events, the tcgen05 kernel alone through
ftk_set_profiling.  FTK_LIB_PATH selects an alternative build of the same library for A/B runs.
    python tools/cosine_time.py [reps] [n_ref n_cur dim]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import _capi, synthetic as S  # noqa: E402
from feature_tracker_b200.api import lib  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n_ref, n_cur, dim = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (20000, 20000, 256)
ctx = ft.Context(0)
L = lib()
dev = torch.device("cuda", 0)
rf, cf = S.make_float_sets(n_ref, n_cur, dim=dim, seed=5)
d_rf, d_cf = torch.from_numpy(rf).to(dev), torch.from_numpy(cf).to(dev)
d_idx = torch.full((n_ref,), -1, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
vp = C.c_void_p
fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT


def call():
    ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf.data_ptr()), n_ref, vp(d_cf.data_ptr()), n_cur, dim, 0.1, vp(d_idx.data_ptr()), fl))


for _ in range(5):
    call()
ctx.synchronize()
whole, kern = [], []
for profiling in (0, 1):
    L.ftk_set_profiling(ctx._h, profiling)
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        if profiling:
            call()
            kern.append(L.ftk_last_kernel_ms(ctx._h) * 1e3)
        else:
            ctx.synchronize()
            t0 = time.perf_counter()
            call()
            ctx.synchronize()
            whole.append((time.perf_counter() - t0) * 1e6)
print(json.dumps({"lib": os.path.basename(_capi.LIB_PATH), "matched": int((d_idx >= 0).sum()), "whole_call_us_median_host_clock": round(float(np.median(whole)), 1),
                  "tc_kernel_us_median": round(float(np.median(kern)), 1), "tc_kernel_us_min": round(float(np.min(kern)), 1)}))
