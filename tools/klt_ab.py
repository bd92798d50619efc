"""In-process A/B timing of one tracker configuration: the default library paths against FTK_DISABLE_FASTPATH=1 (general kernels
only), alternating on the same GPU so that box-to-box and clock noise cancels.  Also asserts that both give identical bits.
  python tools/klt_ab.py lssd inverse 10 720 1280 20 10000"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import _capi, synthetic as S  # noqa: E402
from feature_tracker_b200.api import lib  # noqa: E402

variant, method, half, rows, cols, n_pairs, n_feat = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
unique = 4
L = lib()
dev = torch.device("cuda", 0)
vp = C.c_void_p
os.environ.pop("FTK_DISABLE_FASTPATH", None)
ctx_a = ft.Context(0)
os.environ["FTK_DISABLE_FASTPATH"] = "1"
ctx_b = ft.Context(0)
os.environ.pop("FTK_DISABLE_FASTPATH", None)
pairs = [S.make_pair(rows, cols, n_feat, pair_id=100 + p) for p in range(unique)]
imgs = np.stack([pairs[p % unique][0] for p in range(n_pairs)] + [pairs[p % unique][1] for p in range(n_pairs)])
uv = np.concatenate([pairs[p % unique][2] for p in range(n_pairs)])
res = {}
state = {}
for name, ctx in (("default", ctx_a), ("general_only", ctx_b)):
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, 4, 2 * n_pairs)
    pyr.SetRawImages(imgs)
    pyr.CreateImagePyramid()
    d_ref = torch.from_numpy(uv).to(dev)
    d_cur = torch.empty_like(d_ref)
    d_st = torch.empty((uv.shape[0],), dtype=torch.uint8, device=dev)
    d_off = torch.from_numpy(np.arange(n_pairs + 1, dtype=np.int32) * n_feat).to(dev)
    d_ri = torch.arange(n_pairs, dtype=torch.int32, device=dev)
    d_ci = d_ri + n_pairs
    klt = {"basic": ft.OpticalFlowBasicKlt, "affine": ft.OpticalFlowAffineKlt, "lssd": ft.OpticalFlowLssdKlt}[variant](ctx)
    o = klt.options()
    o.kPatchRowHalfSize = o.kPatchColHalfSize = half
    o.kMethod = {"inverse": ft.OpticalFlowMethod.kInverse, "direct": ft.OpticalFlowMethod.kDirect, "fast": ft.OpticalFlowMethod.kFast}[method]
    o.kMaxTrackPointsNumber = max(500, n_feat)
    prm = klt._params()
    flags = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_PREDICTION | _capi.FLAG_NO_STATUS
    state[name] = (ctx, pyr, prm, d_ref, d_cur, d_st, d_off, d_ri, d_ci, flags)
    res[name] = []


def run(name):
    ctx, pyr, prm, d_ref, d_cur, d_st, d_off, d_ri, d_ci, flags = state[name]
    ctx.check(L.ftk_klt_track(ctx._h, C.byref(prm), pyr._h, pyr._h, n_pairs, vp(d_ri.data_ptr()), vp(d_ci.data_ptr()), vp(d_off.data_ptr()), vp(d_ref.data_ptr()),
                              vp(d_cur.data_ptr()), vp(d_st.data_ptr()), flags))


for name in state:
    run(name)
    state[name][0].synchronize()
for rep in range(5):
    for name in state:
        ctx = state[name][0]
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.synchronize()
        e0.record(stream)
        for _ in range(3):
            run(name)
        e1.record(stream)
        ctx.synchronize()
        res[name].append(e0.elapsed_time(e1) / 3)
a, b = state["default"], state["general_only"]
same = bool((a[4].cpu().numpy().view(np.uint32) == b[4].cpu().numpy().view(np.uint32)).all() and (a[5] == b[5]).all().item())
print({"config": sys.argv[1:], "features": int(uv.shape[0]), "ms_default": sorted(res["default"])[2], "ms_general_only": sorted(res["general_only"])[2],
       "speedup": sorted(res["general_only"])[2] / sorted(res["default"])[2], "identical_bits": same, "tracked": float((a[5] == 1).float().mean().item())})
