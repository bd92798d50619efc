"""One 20k x 20k x 256 cosine force match (for ncu captures)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import _capi, synthetic as S  # noqa: E402
from feature_tracker_b200.api import lib  # noqa: E402

ctx = ft.Context(0)
L = lib()
dev = torch.device("cuda", 0)
rf, cf = S.make_float_sets(20000, 20000, seed=5)
d_rf, d_cf = torch.from_numpy(rf).to(dev), torch.from_numpy(cf).to(dev)
d_idx = torch.full((20000,), -1, dtype=torch.int32, device=dev)
vp = C.c_void_p
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf.data_ptr()), 20000, vp(d_cf.data_ptr()), 20000, 256, 0.1, vp(d_idx.data_ptr()),
                                       _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT))
ctx.synchronize()
print("matched", int((d_idx >= 0).sum()))
