"""Back-to-back 20k x 20k x 256 cosine force matches (no L2 flush, like bench.py's timeit): us per call by the host clock."""
import ctypes as C, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import _capi, synthetic as S  # noqa: E402
from feature_tracker_b200.api import lib  # noqa: E402
ctx = ft.Context(0); L = lib(); dev = torch.device("cuda", 0); vp = C.c_void_p
rf, cf = S.make_float_sets(20000, 20000, seed=5)
d_rf, d_cf = torch.from_numpy(rf).to(dev), torch.from_numpy(cf).to(dev)
d_idx = torch.full((20000,), -1, dtype=torch.int32, device=dev)
fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT
def call():
    ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf.data_ptr()), 20000, vp(d_cf.data_ptr()), 20000, 256, 0.1, vp(d_idx.data_ptr()), fl))
for _ in range(10): call()
ctx.synchronize()
best = []
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(100): call()
    ctx.synchronize()
    best.append((time.perf_counter() - t0) / 100 * 1e6)
print(json.dumps({"lib": os.path.basename(_capi.LIB_PATH), "us_per_call_back_to_back": round(min(best), 2), "matched": int((d_idx >= 0).sum())}))
