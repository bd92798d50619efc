"""Times BRIEF-256 ForceMatch and NearbyMatch (+-50 px) on 10k x 10k descriptors (BASELINE configs[4]) through the C ABI with device pointers:
back-to-back calls, CUDA-free host clock around n calls + one stream sync.  FTK_LIB_PATH selects an alternative build for A/B runs.
    python tools/c4_time.py [calls]"""
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

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = ft.Context(0)
L = lib()
dev = torch.device("cuda", 0)
vp = C.c_void_p
fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT
rb, cb, pred, pos, _ = S.make_brief_sets(10000, 10000, seed=99)
d_r = torch.from_numpy(ft.pack_brief(rb).view(np.int32)).to(dev)
d_c = torch.from_numpy(ft.pack_brief(cb).view(np.int32)).to(dev)
d_idx = torch.full((10000,), -1, dtype=torch.int32, device=dev)
d_pred, d_pos = torch.from_numpy(pred).to(dev), torch.from_numpy(pos).to(dev)


def force():
    ctx.check(L.ftk_match_hamming_force(ctx._h, vp(d_r.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, 60.0, vp(d_idx.data_ptr()), fl))


def nearby():
    ctx.check(L.ftk_match_hamming_nearby(ctx._h, vp(d_r.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, vp(d_pred.data_ptr()), vp(d_pos.data_ptr()), 50, 50, 60.0,
                                         vp(d_idx.data_ptr()), fl))


out = {"lib": os.path.basename(_capi.LIB_PATH)}
for name, fn in (("force", force), ("nearby", nearby)):
    for _ in range(10):
        fn()
    ctx.synchronize()
    best = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(calls):
            fn()
        ctx.synchronize()
        best.append((time.perf_counter() - t0) / calls * 1e6)
    out[name + "_us_per_call"] = round(min(best), 2)
    out[name + "_matched"] = int((d_idx >= 0).sum())
print(json.dumps(out))
