import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import feature_tracker_b200 as ft
from feature_tracker_b200 import _capi
from feature_tracker_b200.api import lib
ctx = ft.Context(0); L = lib(); dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
vp = C.c_void_p
for n in (1024, 2048, 4096, 8192, 12288):
    d_s = torch.randn((n, n), dtype=torch.float32, device=dev) * 2 - 6
    d_i = torch.empty((n,), dtype=torch.int32, device=dev)
    fn = lambda: ctx.check(L.ftk_match_mutual_scores(ctx._h, vp(d_s.data_ptr()), n, n, -3.0, vp(d_i.data_ptr()), _capi.FLAG_DEVICE_POINTERS))
    for _ in range(3): fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20): fn()
    e1.record(stream); ctx.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(n, f"{ms*1e3:.1f} us", f"{4.0*n*n/(ms*1e-3)/1e9:.0f} GB/s")
    del d_s
