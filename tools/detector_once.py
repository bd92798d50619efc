#!/usr/bin/env python
"""One detector + BRIEF pass on a 752x480 synthetic frame, timed with CUDA events (also the target of ncu captures).
    python tools/detector_once.py [min_response] [min_distance] [repeats]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_tracker_b200 as ft  # noqa: E402
from feature_tracker_b200 import _capi, synthetic as S  # noqa: E402
from feature_tracker_b200.api import lib  # noqa: E402


def main():
    thr = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
    dist = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    ctx = ft.Context(0)
    L = lib()
    img = S.make_pair(480, 752, 10, pair_id=301)[0]
    pyr = ft.ImagePyramidBatch(ctx, 480, 752, 4, 1)
    pyr.SetRawImages(img[None])
    pyr.CreateImagePyramid()
    det = ft.FeaturePointHarrisDetector(ctx)
    det.options().kMinValidResponse, det.options().kMinFeatureDistance = thr, dist
    prm = det._params()
    dev = torch.device("cuda:0")
    d_uv = torch.zeros((300, 2), dtype=torch.float32, device=dev)
    n = C.c_int32(0)
    stream = torch.cuda.ExternalStream(L.ftk_stream(ctx._h))
    vp = C.c_void_p

    def run():
        ctx.check(L.ftk_detect_features(ctx._h, C.byref(prm), pyr._h, 0, None, 0, 300, vp(d_uv.data_ptr()), None, C.byref(n), _capi.FLAG_DEVICE_POINTERS))

    run()
    l0 = L.ftk_kernel_launches(ctx._h)
    run()
    launches = L.ftk_kernel_launches(ctx._h) - l0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    ctx.synchronize()
    print(f"detect thr={thr} dist={dist}: {e0.elapsed_time(e1) / reps:.3f} ms per call, {n.value} features, {launches} launches per call")
    n_img = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    if n_img:
        imgs = np.stack([S.make_pair(480, 752, 10, pair_id=301 + (i % 8))[0] for i in range(8)])
        big = ft.ImagePyramidBatch(ctx, 480, 752, 4, n_img)
        big.SetRawImages(imgs[np.arange(n_img) % 8])
        big.CreateImagePyramid()
        d_buv = torch.zeros((n_img, 300, 2), dtype=torch.float32, device=dev)
        counts = np.zeros(n_img, np.int32)

        def run_batch():
            ctx.check(L.ftk_detect_features_batch(ctx._h, C.byref(prm), big._h, 0, n_img, 300, vp(d_buv.data_ptr()), None, vp(counts.ctypes.data), _capi.FLAG_DEVICE_POINTERS))

        run_batch()
        l0 = L.ftk_kernel_launches(ctx._h)
        e0.record(stream)
        for _ in range(3):
            run_batch()
        e1.record(stream)
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"batch of {n_img} images: {ms:.3f} ms per call = {ms / n_img * 1e3:.1f} us per image, {(L.ftk_kernel_launches(ctx._h) - l0) // 3} launches per call, "
              f"{int(counts.sum())} features")


if __name__ == "__main__":
    main()
