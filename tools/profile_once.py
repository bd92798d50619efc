"""One small run of a secondary workload for ncu captures (kernel names: KltKernel, HammingForceKernel, NearbyKernel, CosineTcKernel,
NormPrepKernel, RerankKernel).    python tools/profile_once.py c3|c4|c5 [repeats]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import _capi, synthetic as S  # noqa: E402
from feature_tracker_b200.api import lib  # noqa: E402

what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = ft.Context(0)
L = lib()
dev = torch.device("cuda", 0)
vp = C.c_void_p
fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT
if what == "c3":
    rows, cols, n_pairs, n_feat, unique = 720, 1280, 20, 10000, 2
    pairs = [S.make_pair(rows, cols, n_feat, pair_id=100 + p) for p in range(unique)]
    imgs = np.stack([pairs[p % unique][0] for p in range(n_pairs)] + [pairs[p % unique][1] for p in range(n_pairs)])
    uv = np.concatenate([pairs[p % unique][2] for p in range(n_pairs)])
    pyr = ft.ImagePyramidBatch(ctx, rows, cols, 4, 2 * n_pairs)
    pyr.SetRawImages(imgs)
    pyr.CreateImagePyramid()
    d_ref = torch.from_numpy(uv).to(dev)
    d_cur = torch.empty_like(d_ref)
    d_st = torch.empty((uv.shape[0],), dtype=torch.uint8, device=dev)
    d_off = torch.from_numpy(np.arange(n_pairs + 1, dtype=np.int32) * n_feat).to(dev)
    d_ri = torch.arange(n_pairs, dtype=torch.int32, device=dev)
    d_ci = d_ri + n_pairs
    klt = ft.OpticalFlowLssdKlt(ctx)
    o = klt.options()
    o.kPatchRowHalfSize = o.kPatchColHalfSize = 10
    o.kMethod = ft.OpticalFlowMethod.kInverse
    o.kMaxTrackPointsNumber = n_feat
    prm = klt._params()
    for _ in range(reps):
        ctx.check(L.ftk_klt_track(ctx._h, C.byref(prm), pyr._h, pyr._h, n_pairs, vp(d_ri.data_ptr()), vp(d_ci.data_ptr()), vp(d_off.data_ptr()), vp(d_ref.data_ptr()),
                                  vp(d_cur.data_ptr()), vp(d_st.data_ptr()), _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_PREDICTION | _capi.FLAG_NO_STATUS))
    ctx.synchronize()
    print("tracked", float((d_st == 1).float().mean().item()))
elif what == "c4":
    rb, cb, pred, pos, _ = S.make_brief_sets(10000, 10000, seed=99)
    d_r = torch.from_numpy(ft.pack_brief(rb).view(np.int32)).to(dev)
    d_c = torch.from_numpy(ft.pack_brief(cb).view(np.int32)).to(dev)
    d_idx = torch.full((10000,), -1, dtype=torch.int32, device=dev)
    d_pred, d_pos = torch.from_numpy(pred).to(dev), torch.from_numpy(pos).to(dev)
    for _ in range(reps):
        ctx.check(L.ftk_match_hamming_force(ctx._h, vp(d_r.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, 60.0, vp(d_idx.data_ptr()), fl))
        ctx.check(L.ftk_match_hamming_nearby(ctx._h, vp(d_r.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, vp(d_pred.data_ptr()), vp(d_pos.data_ptr()), 50, 50, 60.0,
                                             vp(d_idx.data_ptr()), fl))
    ctx.synchronize()
    print("matched", int((d_idx >= 0).sum()))
else:
    rf, cf = S.make_float_sets(20000, 20000, seed=5)
    d_rf, d_cf = torch.from_numpy(rf).to(dev), torch.from_numpy(cf).to(dev)
    d_idx = torch.full((20000,), -1, dtype=torch.int32, device=dev)
    for _ in range(reps):
        ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf.data_ptr()), 20000, vp(d_cf.data_ptr()), 20000, 256, 0.1, vp(d_idx.data_ptr()), fl))
    ctx.synchronize()
    print("matched", int((d_idx >= 0).sum()))
