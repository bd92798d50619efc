#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see the task statement / DESIGN.md "Measurement").

Metric (BASELINE.json): tracked features/sec (4-level KLT, 752x480), whole job over N GPUs.
Workload (default, --workload configs1): BASELINE configs[1] -- affine KLT, kDirect AND kFast, 4-level pyramid, 2000
features per frame, a batch of 1000 frame pairs per GPU, 13x13 patches (the reference's default half size 6).  One "step" =
pyramid construction of all 2000 images of the batch + TrackFeatures with kDirect + TrackFeatures with kFast on every pair,
i.e. 2 x 2 000 000 tracked features per step and GPU.  Frame pairs shard across GPUs with no collective on the data path
(weak scaling: every rank owns its own 1000 pairs).

  value      : device-resident throughput (images + features already in HBM), CUDA-event timed on the library's stream
  e2e        : same metric through the C ABI with HOST buffers (pinned): H2D of the images and features and D2H of the
               results are inside the timed region (one ftk_track_image_pairs call per method)
  north_star : the same batch tracked with the tracker BASELINE's north_star sets its target on (basic KLT, kInverse, 15x15):
               value / e2e / roofline of that configuration (--workload north_star makes it the headline instead)
  --impl reference : the reference's own CPU implementation (oracle/_ref, built from the reference's sources) on all
               host cores, bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, LEVELS = 480, 752, 4
PAIRS_PER_GPU = 1000
FEATURES_PER_PAIR = 2000
UNIQUE_PAIRS = 32  # distinct synthetic pairs generated (seeded); the batch of 1000 cycles over them (distinct HBM addresses)
TRAFFIC_JSON = "r2_ncu_traffic.json"  # dram bytes per launch from `ncu --set full` captures of this round's kernels (tools/ncu_summary.py)
METRIC = "tracked features/sec (4-lvl KLT, 752x480)"
UNIT = "features/s"
WORKLOADS = {
    # BASELINE.json configs[1]
    "configs1": [("affine", "direct", 6), ("affine", "fast", 6)],
    # BASELINE.json north_star target: >= 1e8 features/s/GPU for 4-level inverse basic KLT with 15x15 patches
    "north_star": [("basic", "inverse", 7)],
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="configs1", choices=sorted(WORKLOADS) + ["match_sharded"],
                    help="match_sharded: BASELINE configs[3] / configs[4] with the reference rows split over the ranks (strong scaling, SURVEY 8(e))")
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--features", type=int, default=FEATURES_PER_PAIR)
    ap.add_argument("--variant", default=None, help="with --method / --half: time one tracker instead of a named workload (profiling)")
    ap.add_argument("--method", default=None)
    ap.add_argument("--half", type=int, default=None)
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (matchers, other trackers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-north-star", action="store_true", help="skip the north_star object of the configs1 workload")
    return ap.parse_args()


def trackers_of(args):
    if args.workload == "match_sharded":
        return WORKLOADS["configs1"]
    if args.variant or args.method or args.half is not None:
        return [(args.variant or "basic", args.method or "inverse", 7 if args.half is None else args.half)]
    return WORKLOADS[args.workload]


def tracker_name(t):
    return f"{t[0]} KLT {t[1]} {2 * t[2] + 1}x{2 * t[2] + 1}"


def workload_config(args):
    trackers = trackers_of(args)
    what = "BASELINE configs[1]" if trackers == WORKLOADS["configs1"] else ("BASELINE north_star target configuration on the configs[1] batch shape"
                                                                            if trackers == WORKLOADS["north_star"] else "single tracker on the configs[1] batch shape")
    return {
        "workload": f"{what}: {args.pairs} frame pairs/GPU x {args.features} features/frame, {COLS}x{ROWS}, {LEVELS}-level pyramid; trackers = "
                    + " + ".join(tracker_name(t) for t in trackers)
                    + "; step = pyramid build of both frames of every pair + one TrackFeatures per tracker and pair",
        "pairs_per_gpu": args.pairs, "features_per_pair": args.features, "image": [ROWS, COLS], "levels": LEVELS,
        "trackers": [{"variant": t[0], "method": t[1], "patch": 2 * t[2] + 1} for t in trackers],
        "tracked_features_per_step_per_gpu": len(trackers) * args.pairs * args.features,
        "l2_policy": "inputs larger than L2 (pyramid batch ~1 GB per GPU vs 126 MB L2)",
        "sharding": "frame pairs split across ranks, no collective on the data path",
    }


def make_unique_pairs(n_unique, n_features):
    from feature_tracker_b200 import synthetic as S
    refs, curs, uvs = [], [], []
    for p in range(n_unique):
        ref, cur, uv, _ = S.make_pair(ROWS, COLS, n_features, pair_id=p)
        refs.append(ref), curs.append(cur), uvs.append(uv)
    return refs, curs, uvs


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline (the only place bench.py touches oracle/): the reference's own code on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_checker():
    from oracle import pyoracle as po
    if os.path.exists(po.REF_SO) or os.path.isdir(po.REFERENCE_ROOT):
        try:
            return po.RefLib(), "reference"
        except Exception:
            pass
    return po.OracleLib(), "port"


def cpu_track_sample(lib, params_list, refs, curs, uvs, n_pairs, threads):
    """Runs the workload's trackers on n_pairs pairs (cycling over the unique ones) on `threads` host threads; a pair's pyramids
    are built once per tracker call, as the reference's demo does (test/test_optical_flow.cpp:69-73).  Returns (seconds, features)."""
    jobs = list(range(n_pairs))
    lock = threading.Lock()
    done = [0]

    def worker():
        while True:
            with lock:
                if not jobs:
                    return
                p = jobs.pop()
            u = p % len(refs)
            for params in params_list:
                ok, _, _ = lib.pyramid_and_track(params, LEVELS, refs[u], curs[u], uvs[u])
                assert ok
            with lock:
                done[0] += len(uvs[u]) * len(params_list)

    ts = [threading.Thread(target=worker) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, done[0]


def cpu_params(trackers, n_feat):
    from oracle import pyoracle as po
    return [po.make_params(v, m, half=h, max_points=max(500, n_feat)) for v, m, h in trackers]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, kind = cpu_checker()
    cores = os.cpu_count() or 1
    n_feat = args.features
    trackers = trackers_of(args)
    refs, curs, uvs = make_unique_pairs(min(UNIQUE_PAIRS, 4), n_feat)
    params_list = cpu_params(trackers, n_feat)
    # bounded sample: one frame pair per host core and step (a few seconds of CPU work per step)
    sample_pairs = max(1, min(cores, 64))
    for _ in range(max(args.warmup, 0) and 1):
        cpu_track_sample(lib, params_list, refs, curs, uvs, min(sample_pairs, cores), cores)
    times, feats = [], 0
    for _ in range(args.steps):
        dt, nf = cpu_track_sample(lib, params_list, refs, curs, uvs, sample_pairs, cores)
        times.append(dt)
        feats = nf
    total = sum(times)
    value = feats * len(times) / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample_pairs} frame pairs x {n_feat} features per step, per pair and tracker: CreateImagePyramid x2 + TrackFeatures; "
                                   "one tracker object per host thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # under load = the upper half of the samples (idle samples before/after the timed region drop out)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}



def run_extras(ctx, L, torch, local_rank, steps, peaks, keep):
    """Secondary workloads of BASELINE.json configs[1..4] (device-resident, CUDA-event timed).  Not the headline.
    `keep` receives the full-size C4 / C5 results (and their inputs) for the parity check against the reference in cpu_extras."""
    import feature_tracker_b200 as ft
    from feature_tracker_b200 import _capi, synthetic as S
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    vp = C.c_void_p
    out = {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    def timeit(fn, n):
        fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        ctx.synchronize()
        return e0.elapsed_time(e1) / n

    # ---- KLT variants on a reduced batch (100 pairs x 2000 features; 8 pairs for the heavy LSSD shape) ----
    def klt_case(name, variant, method, half, rows, cols, n_pairs, n_feat, unique=4, keep_key=None):
        pairs = [S.make_pair(rows, cols, n_feat, pair_id=100 + p) for p in range(unique)]
        imgs = np.stack([pairs[p % unique][0] for p in range(n_pairs)] + [pairs[p % unique][1] for p in range(n_pairs)])
        pyr = ft.ImagePyramidBatch(ctx, rows, cols, LEVELS, 2 * n_pairs)
        pyr.SetRawImages(imgs)
        pyr.CreateImagePyramid()
        uv = np.concatenate([pairs[p % unique][2] for p in range(n_pairs)])
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

        def fn():
            ctx.check(L.ftk_klt_track(ctx._h, C.byref(prm), pyr._h, pyr._h, n_pairs, vp(d_ri.data_ptr()), vp(d_ci.data_ptr()), vp(d_off.data_ptr()),
                                      vp(d_ref.data_ptr()), vp(d_cur.data_ptr()), vp(d_st.data_ptr()), flags))
        ms = timeit(fn, steps)
        out[name] = {"features_per_s": uv.shape[0] / (ms * 1e-3), "ms": ms, "pairs": n_pairs, "features_per_pair": n_feat,
                     "tracked_fraction": float((d_st.cpu().numpy() == 1).mean())}
        if keep_key:
            keep[keep_key] = {"tracker": (variant, method, half), "pairs": pairs, "n_pairs": n_pairs, "n_feat": n_feat, "unique": unique, "ms": ms,
                              "uv": d_cur.cpu().numpy(), "st": d_st.cpu().numpy()}
        pyr.close()

    klt_case("C2_affine_direct_13x13", "affine", "direct", 6, ROWS, COLS, 100, 2000)
    klt_case("C2_affine_fast_13x13", "affine", "fast", 6, ROWS, COLS, 100, 2000)
    klt_case("basic_fast_13x13_reference_default", "basic", "fast", 6, ROWS, COLS, 100, 2000)
    klt_case("basic_direct_15x15", "basic", "direct", 7, ROWS, COLS, 100, 2000)
    klt_case("basic_inverse_15x15_100pairs", "basic", "inverse", 7, ROWS, COLS, 100, 2000)
    # BASELINE configs[2]: 10k features per 1280x720 frame; 20 frame pairs = 200 000 features per launch (~50 waves of feature groups)
    klt_case("C3_lssd_inverse_21x21_1280x720", "lssd", "inverse", 10, 720, 1280, 20, 10000, unique=2, keep_key="c3")

    # ---- C4: BRIEF-256 force 10k x 10k + nearby ----
    rb, cb, pred, pos, truth = S.make_brief_sets(10000, 10000, seed=99)
    d_r = torch.from_numpy(ft.pack_brief(rb).view(np.int32)).to(dev)
    d_c = torch.from_numpy(ft.pack_brief(cb).view(np.int32)).to(dev)
    d_idx = torch.full((10000,), -1, dtype=torch.int32, device=dev)
    d_pred = torch.from_numpy(pred).to(dev)
    d_pos = torch.from_numpy(pos).to(dev)
    fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT
    ms = timeit(lambda: ctx.check(L.ftk_match_hamming_force(ctx._h, vp(d_r.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, 60.0, vp(d_idx.data_ptr()), fl)), steps * 4)
    idx = d_idx.cpu().numpy()
    has = truth >= 0
    popc_peak_pairs = 148 * 16 * 1.965e9 / 8  # measured POPC rate: 16 lanes/clk/SM (profiles/r1_microbench_pipe_rates.txt), 8 POPC per 256-bit pair
    out["C4_brief256_force_10k_x_10k"] = {"pairs_per_s": 1e8 / (ms * 1e-3), "ms": ms, "recovered_planted_matches": float((idx[has] == truth[has]).mean()),
                                          "popc_peak_pairs_per_s": popc_peak_pairs, "frac_of_popc_peak": 1e8 / (ms * 1e-3) / popc_peak_pairs,
                                          "roofline": {"bound": "integer pipe (POPC)", "achieved": 8e8 / (ms * 1e-3) / 1e12, "peak": popc_peak_pairs * 8 / 1e12,
                                                       "unit": "T POPC32/s", "frac": 1e8 / (ms * 1e-3) / popc_peak_pairs, "traffic": None,
                                                       "peak_source": "measured POPC issue rate 16 lanes/clk/SM (tools/microbench.cu, profiles/r1_microbench_pipe_rates.txt) x 148 SM x 1.965 GHz",
                                                       "algorithmic_work": "8 POPC32 per 256-bit pair x 1e8 pairs (SURVEY 8(d)); the kernel ISSUES 5 per pair (carry-save adders on the ALU "
                                                                           "pipe fold three words into two before counting), so this fraction can pass the naive XU-pipe bound",
                                                       "scope": "whole call (all launches)"}}
    keep["c4_force_idx"] = idx.copy()
    keep["c4_inputs"] = (rb, cb, pred, pos)
    ms = timeit(lambda: ctx.check(L.ftk_match_hamming_nearby(ctx._h, vp(d_r.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, vp(d_pred.data_ptr()),
                                                             vp(d_pos.data_ptr()), 50, 50, 60.0, vp(d_idx.data_ptr()), fl)), steps * 4)
    keep["c4_nearby_idx"] = d_idx.cpu().numpy().copy()
    # algorithmic bytes of the candidate search: every cur descriptor inside a ref row's window is read once (32 B descriptor + 8 B position)
    in_window = int(((np.abs(pos[None, :, 0] - pred[:, None, 0]) <= 50) & (np.abs(pos[None, :, 1] - pred[:, None, 1]) <= 50)).sum())
    nb_bytes = in_window * 40.0 + 10000 * (32 + 8 + 4)
    out["C4_brief256_nearby_10k_window50"] = {"ref_rows_per_s": 1e4 / (ms * 1e-3), "ms": ms, "candidate_pairs": in_window,
                                              "roofline": {"bound": "hbm", "achieved": nb_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                                           "frac": nb_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                                                           "algorithmic_bytes_per_launch": nb_bytes,
                                                           "note": "gather latency bound: the whole working set (400 KB) is L2 resident, so HBM is the nominal "
                                                                   "roofline only; candidates in window x 40 B + 44 B per ref row; whole call (grid build + search)"}}

    # ---- 1000 frame pairs x (300 x 300) BRIEF-256 NearbyMatch in one call (the KLT batch's counterpart for descriptor tracking) ----
    n_mp, per = 1000, 300
    rb1, cb1, pred1, pos1, _ = S.make_brief_sets(per, per, seed=9)
    d_rp = torch.from_numpy(np.tile(ft.pack_brief(rb1), (n_mp, 1)).astype(np.int32)).to(dev)
    d_cp = torch.from_numpy(np.tile(ft.pack_brief(cb1), (n_mp, 1)).astype(np.int32)).to(dev)
    d_predp = torch.from_numpy(np.tile(pred1, (n_mp, 1))).to(dev)
    d_posp = torch.from_numpy(np.tile(pos1, (n_mp, 1))).to(dev)
    offs = (np.arange(n_mp + 1) * per).astype(np.int32)
    d_idxp = torch.full((n_mp * per,), -1, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ctx.check(L.ftk_match_hamming_pairs(ctx._h, vp(d_rp.data_ptr()), vp(d_cp.data_ptr()), 8, n_mp, vp(offs.ctypes.data), vp(offs.ctypes.data),
                                                            vp(d_predp.data_ptr()), vp(d_posp.data_ptr()), 50, 50, 60.0, vp(d_idxp.data_ptr()), fl)), steps * 2)
    out["brief256_nearby_1000_pairs_x_300x300"] = {"ms": ms, "pairs_of_frames_per_s": n_mp / (ms * 1e-3), "ref_rows_per_s": n_mp * per / (ms * 1e-3),
                                                   "matched": int((d_idxp.cpu().numpy() >= 0).sum())}

    # ---- C5: float-256 force 20k x 20k ----
    rf, cf = S.make_float_sets(20000, 20000, seed=5)
    d_rf = torch.from_numpy(rf).to(dev)
    d_cf = torch.from_numpy(cf).to(dev)
    d_idx2 = torch.full((20000,), -1, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf.data_ptr()), 20000, vp(d_cf.data_ptr()), 20000, 256, 0.1, vp(d_idx2.data_ptr()), fl)), max(2, steps // 2))
    flop = 2.0 * 20000 * 20000 * 256
    keep["c5_idx"] = d_idx2.cpu().numpy().copy()
    keep["c5_inputs"] = (rf, cf)
    bf16_peak = peaks.get("bf16_tflops", 1650.0)
    c5 = {"pairs_per_s": 4e8 / (ms * 1e-3), "ms": ms, "tflops": flop / (ms * 1e-3) / 1e12, "matched": int((keep["c5_idx"] >= 0).sum()),
          "exact_scan_rows": int(L.ftk_last_cosine_exact_scan_items(ctx._h)),
          "roofline_call": {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s", "frac": flop / (ms * 1e-3) / 1e12 / bf16_peak,
                            "traffic": None, "scope": "whole ftk_match_cosine_force call: normalise + BF16 copies, tcgen05 GEMM with top-2 epilogue, exact re-rank",
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if "bf16_tflops" in peaks else "fallback"}}
    if hasattr(L, "ftk_set_profiling"):
        L.ftk_set_profiling(ctx._h, 1)
        kms = []
        for _ in range(max(3, steps // 2)):
            ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf.data_ptr()), 20000, vp(d_cf.data_ptr()), 20000, 256, 0.1, vp(d_idx2.data_ptr()), fl))
            ctx.synchronize()
            kms.append(float(L.ftk_last_kernel_ms(ctx._h)))
        L.ftk_set_profiling(ctx._h, 0)
        k_ms = statistics.median(kms)
        c5["roofline"] = {"bound": "tensor", "achieved": flop / (k_ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s", "frac": flop / (k_ms * 1e-3) / 1e12 / bf16_peak,
                          "traffic": None, "kernel": "CosineTcKernel (tcgen05.mma BF16, TMEM accumulators, TMA operands)", "launch_ms": k_ms,
                          "algorithmic_flops_per_launch": flop, "scope": "dominant kernel alone, CUDA events around its launch on the library's stream"}
    out["C5_float256_force_20k_x_20k"] = c5

    # ---- direct-method pose tracker: one 6-DoF pose per frame pair from 300 features, 13x13 patches, 4 levels ----
    n_dm, f_dm = 600, 300
    scenes = [S.make_direct_method_scene(ROWS, COLS, f_dm, pair_id=200 + p) for p in range(4)]
    pyr_dm = ft.ImagePyramidBatch(ctx, ROWS, COLS, LEVELS, 2 * n_dm)
    pyr_dm.SetRawImages(np.stack([scenes[p % 4][0] for p in range(n_dm)] + [scenes[p % 4][1] for p in range(n_dm)]))
    pyr_dm.CreateImagePyramid()
    counts = [scenes[p % 4][2].shape[0] for p in range(n_dm)]
    d_off = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)).to(dev)
    d_uv = torch.from_numpy(np.concatenate([scenes[p % 4][2] for p in range(n_dm)])).to(dev)
    d_pts = torch.from_numpy(np.concatenate([scenes[p % 4][4] for p in range(n_dm)])).to(dev)
    d_K = torch.from_numpy(np.stack([scenes[p % 4][3] for p in range(n_dm)])).to(dev)
    d_cur = torch.empty_like(d_uv)
    d_st = torch.empty((d_uv.shape[0],), dtype=torch.uint8, device=dev)
    d_ri = torch.arange(n_dm, dtype=torch.int32, device=dev)
    d_ci = d_ri + n_dm
    q_init = torch.tensor([1.0, 0.0, 0.0, 0.0], device=dev).repeat(n_dm, 1).contiguous()
    d_q, d_p = q_init.clone(), torch.zeros((n_dm, 3), device=dev)
    dprm = ft.DirectMethod(ctx)._params()
    fl_dm = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_PREDICTION | _capi.FLAG_NO_STATUS

    def dm_fn():
        d_q.copy_(q_init)  # every run starts from the identity pose (tiny device copies on torch's stream, ordered by the syncs of timeit)
        d_p.zero_()
        torch.cuda.current_stream().synchronize()
        ctx.check(L.ftk_direct_method_track(ctx._h, C.byref(dprm), pyr_dm._h, pyr_dm._h, n_dm, vp(d_ri.data_ptr()), vp(d_ci.data_ptr()), vp(d_off.data_ptr()),
                                            vp(d_K.data_ptr()), vp(d_pts.data_ptr()), vp(d_uv.data_ptr()), vp(d_cur.data_ptr()), vp(d_q.data_ptr()),
                                            vp(d_p.data_ptr()), vp(d_st.data_ptr()), fl_dm))
    ms = timeit(dm_fn, max(2, steps // 2))
    out["direct_method_600_pairs_x_300_features"] = {"ms": ms, "pairs_per_s": n_dm / (ms * 1e-3), "features_per_s": int(d_uv.shape[0]) / (ms * 1e-3),
                                                     "tracked_fraction": float((d_st.cpu().numpy() == 1).mean()),
                                                     "pose_translation_of_pair_0": [float(x) for x in d_p[0].cpu().numpy()]}
    pyr_dm.close()

    # ---- dense optical flow: Farneback, 752x480, 4 levels, reference defaults ----
    pair0 = S.make_pair(ROWS, COLS, 10, pair_id=300)
    pyr_df = ft.ImagePyramidBatch(ctx, ROWS, COLS, LEVELS, 2)
    pyr_df.SetRawImages(np.stack([pair0[0], pair0[1]]))
    pyr_df.CreateImagePyramid()
    d_fr = torch.empty((ROWS, COLS), dtype=torch.float32, device=dev)
    d_fc = torch.empty_like(d_fr)
    fprm = ft.DenseOpticalFlow(ctx)._params()
    ms = timeit(lambda: ctx.check(L.ftk_dense_flow_track(ctx._h, C.byref(fprm), pyr_df._h, pyr_df._h, 0, 1, vp(d_fr.data_ptr()), vp(d_fc.data_ptr()),
                                                         _capi.FLAG_DEVICE_POINTERS)), steps * 2)
    out["dense_flow_752x480_4_levels"] = {"ms": ms, "pixels_per_s": ROWS * COLS / (ms * 1e-3), "mean_abs_flow_px": float(d_fr.abs().mean().item())}
    pyr_df.close()

    # ---- front end (parity unpinned): detect 300 Harris corners (demo options) on the device pyramid + BRIEF-256 ----
    img0 = S.make_pair(ROWS, COLS, 10, pair_id=301)[0]
    pyr_det = ft.ImagePyramidBatch(ctx, ROWS, COLS, LEVELS, 1)
    pyr_det.SetRawImages(img0[None])
    pyr_det.CreateImagePyramid()
    det = ft.FeaturePointHarrisDetector(ctx)
    d_duv = torch.zeros((300, 2), dtype=torch.float32, device=dev)
    d_dresp = torch.zeros((300,), dtype=torch.float32, device=dev)
    n_found = C.c_int32(0)
    for thr, tag in ((40.0, "demo_threshold_40"), (1e5, "threshold_1e5")):
        det.options().kMinValidResponse = thr
        dprm2 = det._params()
        ms = timeit(lambda: ctx.check(L.ftk_detect_features(ctx._h, C.byref(dprm2), pyr_det._h, 0, None, 0, 300, vp(d_duv.data_ptr()), vp(d_dresp.data_ptr()),
                                                            C.byref(n_found), _capi.FLAG_DEVICE_POINTERS)), max(2, steps // 2))
        out[f"detect_300_harris_752x480_{tag}"] = {"ms": ms, "found": int(n_found.value), "pixels_per_s": ROWS * COLS / (ms * 1e-3)}
    n_batch = 256  # the same detector over a batch of frames in one call: every kernel covers all images
    imgs8 = np.stack([S.make_pair(ROWS, COLS, 10, pair_id=301 + i)[0] for i in range(8)])
    pyr_batch = ft.ImagePyramidBatch(ctx, ROWS, COLS, LEVELS, n_batch)
    pyr_batch.SetRawImages(imgs8[np.arange(n_batch) % 8])
    pyr_batch.CreateImagePyramid()
    d_buv = torch.zeros((n_batch, 300, 2), dtype=torch.float32, device=dev)
    counts = np.zeros(n_batch, np.int32)
    ms = timeit(lambda: ctx.check(L.ftk_detect_features_batch(ctx._h, C.byref(dprm2), pyr_batch._h, 0, n_batch, 300, vp(d_buv.data_ptr()), None,
                                                              vp(counts.ctypes.data), _capi.FLAG_DEVICE_POINTERS)), 3)
    out["detect_300_harris_batch_of_256_frames"] = {"ms": ms, "us_per_frame": ms / n_batch * 1e3, "frames_per_s": n_batch / (ms * 1e-3), "found": int(counts.sum())}
    pyr_batch.close()
    d_resp_map = torch.empty((ROWS, COLS), dtype=torch.float32, device=dev)
    ms = timeit(lambda: ctx.check(L.ftk_detect_response(ctx._h, C.byref(dprm2), pyr_det._h, 0, vp(d_resp_map.data_ptr()), _capi.FLAG_DEVICE_POINTERS)), steps * 2)
    out["harris_response_752x480"] = {"ms": ms, "gb_per_s": 5.0 * ROWS * COLS / (ms * 1e-3) / 1e9, "algorithmic_bytes": 5 * ROWS * COLS}
    pattern = ft.brief_pattern(256, 8, 0)
    uv_many = np.tile(d_duv.cpu().numpy(), (200, 1))  # 60 000 features
    d_many = torch.from_numpy(uv_many).to(dev)
    d_desc = torch.empty((len(uv_many), 8), dtype=torch.int32, device=dev)
    ms = timeit(lambda: ctx.check(L.ftk_describe_brief(ctx._h, pyr_det._h, 0, vp(d_many.data_ptr()), len(uv_many), vp(pattern.ctypes.data), 256, 8,
                                                       vp(d_desc.data_ptr()), None, _capi.FLAG_DEVICE_POINTERS)), steps * 2)
    out["brief256_60k_features"] = {"ms": ms, "descriptors_per_s": len(uv_many) / (ms * 1e-3)}
    pyr_det.close()

    # ---- score-matrix mutual arg-max (NNFeatureMatcher post-processing): HBM bound, 4 B per matrix element ----
    for n in (2048, 12288):  # LightGlue's usual size (16 MB, L2 resident) and a matrix far larger than L2 (604 MB)
        d_s = torch.randn((n, n), dtype=torch.float32, device=dev) * 2.0 - 6.0
        d_i = torch.empty((n,), dtype=torch.int32, device=dev)
        ms = timeit(lambda: ctx.check(L.ftk_match_mutual_scores(ctx._h, vp(d_s.data_ptr()), n, n, -3.0, vp(d_i.data_ptr()), _capi.FLAG_DEVICE_POINTERS)), steps * 2)
        out[f"mutual_scores_{n}x{n}"] = {"ms": ms, "gb_per_s": 4.0 * n * n / (ms * 1e-3) / 1e9, "algorithmic_bytes": 4 * n * n}
        del d_s
    return out


def cpu_extras(out, keep, fp32_peak):
    """Single-thread CPU figures of the reference (oracle/_ref; the C restatement for the mutual scores) on bounded samples of the
    secondary workloads (2k x 2k sub-problems for the matchers), written next to the GPU figures; FULL-SIZE parity of
    the C4 / C5 results of the timed runs against the reference (ref rows split over the host threads); C3's roofline."""
    from feature_tracker_b200 import synthetic as S
    from oracle import pyoracle as po
    lib_cpu, kind = cpu_checker()
    cores = os.cpu_count() or 1

    if "c4_inputs" in keep:
        rb, cb, pred, pos = keep["c4_inputs"]
        t0 = time.perf_counter()
        ok, exp = lib_cpu.match_rows_threaded("brief_force", rb, cb, 60.0)
        dt = time.perf_counter() - t0
        out["C4_brief256_force_10k_x_10k"]["parity"] = {"checked_rows": 10000, "index_mismatch": int((exp != keep["c4_force_idx"]).sum()), "matched": int((exp >= 0).sum()),
                                                        "checker": f"oracle/_ref ForceMatch, full 10k x 10k, {cores} host threads, {dt:.1f} s", "kind": kind}
        ok, exp = lib_cpu.match_rows_threaded("brief_nearby", rb, cb, 60.0, pred_uv=pred, cur_uv=pos, max_drow=50, max_dcol=50)
        out["C4_brief256_nearby_10k_window50"]["parity"] = {"checked_rows": 10000, "index_mismatch": int((exp != keep["c4_nearby_idx"]).sum()),
                                                            "matched": int((exp >= 0).sum()), "checker": "oracle/_ref NearbyMatch, full size", "kind": kind}
    if "c5_inputs" in keep:
        rf, cf = keep["c5_inputs"]
        t0 = time.perf_counter()
        ok, exp = lib_cpu.match_rows_threaded("cosine_force", rf, cf, 0.1)
        dt = time.perf_counter() - t0
        out["C5_float256_force_20k_x_20k"]["parity"] = {"checked_rows": 20000, "index_mismatch": int((exp != keep["c5_idx"]).sum()), "matched": int((exp >= 0).sum()),
                                                        "checker": f"oracle/_ref ForceMatch (sequential fp32 cosine), full 20k x 20k, {cores} host threads, {dt:.1f} s",
                                                        "kind": kind}
        out["C5_float256_force_20k_x_20k"]["cpu_reference_all_cores"] = {"pairs_per_s": 4e8 / dt, "cores": cores, "sample": "full 20k x 20k", "kind": kind}
    if "c3" in keep:
        c3 = keep["c3"]
        v, m, h = c3["tracker"]
        oc = po.OracleLib()
        params = po.make_params(v, m, half=h, max_points=max(500, c3["n_feat"]))
        sub = 1500  # oracle iteration trace + parity on the first `sub` features of each unique pair (LSSD 21x21 costs ~1 ms per feature on one core)
        iters, mism = 0, 0
        for u in range(c3["unique"]):
            ref, cur, uv, _ = c3["pairs"][u]
            cu, st, it = oracle_trace(oc, params, [ref], [cur], [uv[:sub]], 0, sub)
            iters += int(it.sum())
            got_uv, got_st = c3["uv"][u * c3["n_feat"]:u * c3["n_feat"] + sub], c3["st"][u * c3["n_feat"]:u * c3["n_feat"] + sub]
            same = (got_uv.view(np.uint32) == cu.view(np.uint32)) | (np.isnan(got_uv) & np.isnan(cu))
            mism += int((~same).any(1).sum()) + int((got_st != st).sum())
        n_total = c3["n_pairs"] * c3["n_feat"]
        per_feature = iters / (sub * c3["unique"])
        flops, formula = algorithmic_flops(c3["tracker"], per_feature * n_total, n_total)
        tfs = flops / (c3["ms"] * 1e-3) / 1e12
        o = out["C3_lssd_inverse_21x21_1280x720"]
        o["roofline"] = {"bound": "fp32", "bound_note": "FP32 CUDA-core issue (no FMA: bit-exact arithmetic) + L1/shared bandwidth; neither HBM nor tensor",
                         "achieved": tfs, "peak": fp32_peak["tflops"], "unit": "TFLOP/s", "frac": tfs / fp32_peak["tflops"], "traffic": None,
                         "peak_source": fp32_peak["source"], "algorithmic_flops_per_launch": flops, "algorithmic_flops_formula": formula,
                         "patch_iterations_per_feature": per_feature, "launch_ms": c3["ms"], "features_per_launch": n_total}
        o["parity"] = {"checked_features": sub * c3["unique"], "mismatch": mism, "checker": "oracle C restatement (bit-identical to oracle/_ref)"}
        dt1 = clock_once(lambda: lib_cpu.pyramid_and_track(params, LEVELS, c3["pairs"][0][0], c3["pairs"][0][1], c3["pairs"][0][2][:sub]))
        o["cpu_reference"] = {"features_per_s": sub / dt1, "sample": f"1 frame pair x {sub} features incl. both pyramids, 1 thread", "kind": kind}

    def clock(fn):
        t0 = time.perf_counter()
        fn()
        return time.perf_counter() - t0

    rb, cb, _, _, _ = S.make_brief_sets(2000, 2000, seed=99)
    dt = clock(lambda: lib_cpu.match_brief_force(rb, cb, 60.0))
    out["C4_brief256_force_10k_x_10k"]["cpu_reference"] = {"pairs_per_s": 4e6 / dt, "sample": "2000 x 2000, 1 thread", "kind": kind}
    rf, cf = S.make_float_sets(2000, 2000, seed=5)
    dt = clock(lambda: lib_cpu.match_cosine_force(rf, cf, 0.1))
    out["C5_float256_force_20k_x_20k"]["cpu_reference"] = {"pairs_per_s": 4e6 / dt, "sample": "2000 x 2000, 1 thread", "kind": kind}
    ref, cur, uv, K, pts = S.make_direct_method_scene(ROWS, COLS, 300, pair_id=200)
    rl, cl = lib_cpu.pyramid_build(ref, LEVELS), lib_cpu.pyramid_build(cur, LEVELS)
    dt = clock(lambda: lib_cpu.direct_method_track(po.make_direct_params(), rl, cl, K, pts, uv, [1, 0, 0, 0], [0, 0, 0]))
    out["direct_method_600_pairs_x_300_features"]["cpu_reference"] = {"pairs_per_s": 1.0 / dt, "sample": "1 frame pair x 300 features, 1 thread", "kind": kind}
    dt = clock(lambda: lib_cpu.dense_flow_track(po.make_dense_flow_params(), rl, cl))
    out["dense_flow_752x480_4_levels"]["cpu_reference"] = {"pixels_per_s": ROWS * COLS / dt, "sample": "1 frame pair, 1 thread", "kind": kind}
    oc = po.OracleLib()
    img0 = S.make_pair(ROWS, COLS, 10, pair_id=301)[0]
    dt = clock(lambda: oc.detect_features(po.make_detector_params("harris", 1, 0.04, 40.0, 20), img0, 300))
    out["detect_300_harris_752x480_demo_threshold_40"]["cpu_reference"] = {"ms": dt * 1e3, "sample": "1 image, 1 thread", "kind": "port (parity unpinned)"}
    scores = (np.random.default_rng(1).normal(-6, 2, (2048, 2048))).astype(np.float32)
    dt = clock(lambda: oc.mutual_scores(scores, -3.0))
    out["mutual_scores_2048x2048"]["cpu_reference"] = {"ms": dt * 1e3, "sample": "2048 x 2048, 1 thread", "kind": "port"}


def pyramid_traffic(n_images):
    """dram bytes per launch of the pyramid kernel from this round's `ncu --set full` capture at bench scale (2 000 images, far larger
    than L2), scaled by the image count when the launch differs; None when no capture is committed."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_JSON)))["PyramidKernel"]
        return (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["images_per_launch"] * n_images
    except Exception:
        return None


def c1_latency(cpu_lib, kind):
    """BASELINE configs[0]: one 752x480 frame pair, 200 features, 4 levels, basic KLT kInverse 15x15 -- latency of the reference-facing
    routes (tools/c1_latency: the C++ facade on ordinary host memory, and the fused C-ABI call on pinned memory; host clock, median of
    300 calls), the reference's own time for the same pair beside it (test/test_optical_flow.cpp:69-73: pyramids x2 + TrackFeatures)."""
    import struct
    import tempfile
    from feature_tracker_b200 import synthetic as S
    from oracle import pyoracle as po
    ref, cur, uv, _ = S.make_pair(ROWS, COLS, 200, pair_id=7)
    exe = os.path.join(ROOT, "tools", "c1_latency")
    os.chmod(exe, os.stat(exe).st_mode | 0o111)
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(struct.pack("5i", ROWS, COLS, LEVELS, len(uv), 7))
        for a in (ref, cur, uv):
            f.write(np.ascontiguousarray(a).tobytes())
        path = f.name
    try:
        res = json.loads(subprocess.run([exe, path, "300"], capture_output=True, text=True, timeout=300).stdout.strip().splitlines()[-1])
    finally:
        os.unlink(path)
    out = {"config": "BASELINE configs[0]: 1 frame pair 752x480, 200 features, 4 levels, basic KLT kInverse 15x15; per call: pyramid x2 + TrackFeatures, host buffers in "
                     "and out; host steady_clock, median of 300 calls after 20 warm-up calls", "gpu": res}
    if cpu_lib is not None:
        params = po.make_params("basic", "inverse", half=7, max_points=500)
        cpu_lib.pyramid_and_track(params, LEVELS, ref, cur, uv)
        ts = sorted(clock_once(lambda: cpu_lib.pyramid_and_track(params, LEVELS, ref, cur, uv)) for _ in range(15))
        out["cpu_reference"] = {"ms": ts[len(ts) // 2] * 1e3, "kind": kind, "cores": 1, "sample": "the same pair, median of 15 calls, 1 thread (the reference is single-threaded)"}
        out["speedup_facade"] = out["cpu_reference"]["ms"] * 1e3 / res["facade_us"]["median"]
        out["speedup_one_call"] = out["cpu_reference"]["ms"] * 1e3 / res["one_call_us"]["median"]
    return out


def clock_once(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def oracle_trace(oracle, cparams, refs, curs, uvs, u, n_feat):
    """The oracle's results and per-feature patch-iteration counts (SURVEY 8(d) unit) for unique pair u."""
    rl, cl = oracle.pyramid_build(refs[u], LEVELS), oracle.pyramid_build(curs[u], LEVELS)
    it = np.zeros(n_feat, np.int32)
    levels = len(rl)
    rows = np.array([a.shape[0] for a in rl], np.int32)
    cols = np.array([a.shape[1] for a in rl], np.int32)
    PtrArr = C.POINTER(C.c_uint8) * levels
    rp = PtrArr(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in rl])
    cp = PtrArr(*[a.ctypes.data_as(C.POINTER(C.c_uint8)) for a in cl])
    cu = np.zeros((n_feat, 2), np.float32)
    st = np.zeros(n_feat, np.uint8)
    f = oracle.lib.ftko_klt_track_traced
    f.restype = C.c_int
    f(C.byref(cparams), C.c_int32(levels), rp, cp, rows.ctypes.data_as(C.c_void_p), cols.ctypes.data_as(C.c_void_p), C.c_int32(n_feat),
      uvs[u].ctypes.data_as(C.c_void_p), cu.ctypes.data_as(C.c_void_p), C.c_int32(0), st.ctypes.data_as(C.c_void_p), C.c_int32(0), C.c_int32(0),
      it.ctypes.data_as(C.c_void_p))
    return cu, st, it


def measured_fp32_peak(device):
    """The FP32 roofline denominator of the KLT kernels.  They are compiled without FMA contraction (bit-exact arithmetic), so the
    reachable ceiling is the FADD / FMUL issue rate, measured live by tools/microbench (built by `make`; dependent-free FADD and FMUL
    streams, 8 independent chains per thread): warp-instructions/clk/SM x 32 lanes x SMs x clock = flop/s (1 flop per instruction)."""
    exe = os.path.join(ROOT, "tools", "microbench")
    derived = 148 * 128 * 1.965e9 / 1e12
    try:
        os.chmod(exe, os.stat(exe).st_mode | 0o111)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
        res = json.loads(subprocess.run([exe, "--json", str(device)], capture_output=True, text=True, timeout=120, env=env).stdout.strip().splitlines()[-1])
        rate = 0.5 * (res["fadd"]["flops_per_s"] + res["fmul"]["flops_per_s"]) / 1e12
        return {"tflops": rate, "source": f"measured live by tools/microbench: FADD {res['fadd']['warp_instr_per_clk_sm']:.2f} / FMUL "
                                          f"{res['fmul']['warp_instr_per_clk_sm']:.2f} warp-instr/clk/SM x 32 lanes x {res['sm_count']} SMs (timed, no clock assumed); 1 flop per "
                                          "instruction because the kernels run without FMA (bit-exact fp32); the nominal FMA peak would be "
                                          f"{2 * derived:.1f} TFLOP/s", "popc_per_s": res["popc"]["ops_per_s"], "raw": res}
    except Exception as e:  # noqa: BLE001
        return {"tflops": derived * 3.75 / 4.0, "source": f"fallback: 3.75 of 4 warp-instr/clk/SM (profiles/r1_microbench_pipe_rates.txt) x 148 SMs x 1.965 GHz; "
                                                          f"live microbench failed ({e!r})", "popc_per_s": None, "raw": None}


def algorithmic_flops(tracker, iters, n_total):
    """SURVEY 8(d) per-unit figures x executed units (patch iterations from the oracle's trace, feature-levels)."""
    v, m, h = tracker
    P = (2 * h + 1) ** 2
    if v == "basic":  # hoisted form: 20*P flop per iteration + ((2h+3)^2-4)*15 + 8*P flop per feature-level
        return iters * 20.0 * P + n_total * LEVELS * ((((2 * h + 3) ** 2) - 4) * 15.0 + 8.0 * P), "20*P flop/iteration + ((2h+3)^2-4)*15 + 8*P flop/feature-level"
    if v == "affine":  # ~90*P flop per iteration + 6x6 LDLT ~250 flop
        return iters * (90.0 * P + 250.0), "90*P + 250 flop/iteration"
    return iters * 60.0 * P, "60*P flop/iteration (hoisted form)"


def run_b200(args):
    import torch
    import torch.distributed as dist

    import feature_tracker_b200 as ft
    from feature_tracker_b200 import _capi
    from feature_tracker_b200.api import lib as ftk_lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created: point fd 1 at stderr until the first
        # collective has run, so that stdout carries the JSON line only.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    ctx = ft.Context(local_rank)
    L = ftk_lib()
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    vp = C.c_void_p

    trackers = trackers_of(args)
    n_pairs, n_feat = args.pairs, args.features
    refs, curs, uvs = make_unique_pairs(UNIQUE_PAIRS, n_feat)
    # this rank's shard: pairs [rank*n_pairs, (rank+1)*n_pairs) of the global batch, cycling over the unique pairs
    uniq = [(rank * n_pairs + p) % UNIQUE_PAIRS for p in range(n_pairs)]
    plane = ROWS * COLS
    host_images = torch.empty((2 * n_pairs, ROWS, COLS), dtype=torch.uint8).pin_memory()
    hi = host_images.numpy()
    for p, u in enumerate(uniq):
        hi[p] = refs[u]
        hi[n_pairs + p] = curs[u]
    n_total = n_pairs * n_feat
    host_ref_uv = torch.empty((n_total, 2), dtype=torch.float32).pin_memory()
    host_cur_uv = torch.empty((n_total, 2), dtype=torch.float32).pin_memory()
    host_status = torch.empty((n_total,), dtype=torch.uint8).pin_memory()
    hr = host_ref_uv.numpy()
    for p, u in enumerate(uniq):
        hr[p * n_feat:(p + 1) * n_feat] = uvs[u]
    offsets = (np.arange(n_pairs + 1, dtype=np.int32) * n_feat)

    pyr = ft.ImagePyramidBatch(ctx, ROWS, COLS, LEVELS, 2 * n_pairs)
    d_ref_uv = torch.from_numpy(hr).to(dev)
    d_offsets = torch.from_numpy(offsets).to(dev)
    d_ref_idx = torch.arange(n_pairs, dtype=torch.int32, device=dev)
    d_cur_idx = d_ref_idx + n_pairs
    pyr.set_images_ptr(host_images.data_ptr(), 2 * n_pairs)
    ctx.synchronize()

    def klt_params(tracker):
        v, m, h = tracker
        klt = {"basic": ft.OpticalFlowBasicKlt, "affine": ft.OpticalFlowAffineKlt, "lssd": ft.OpticalFlowLssdKlt}[v](ctx)
        o = klt.options()
        o.kPatchRowHalfSize = o.kPatchColHalfSize = h
        o.kMethod = {"inverse": ft.OpticalFlowMethod.kInverse, "direct": ft.OpticalFlowMethod.kDirect, "fast": ft.OpticalFlowMethod.kFast}[m]
        o.kMaxTrackPointsNumber = max(500, n_feat)
        return klt._params()

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """Times `steps` calls with CUDA events on the library's stream, bracketed by barrier + synchronize; max over ranks."""
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ctx.kernel_launches
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if os.environ.get("FTK_BENCH_VERBOSE"):
            print(f"[rank {rank}] {getattr(fn, '__name__', 'step')}: {ms / steps:.3f} ms/step", file=sys.stderr)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.kernel_launches - launches0

    def pyramid_only():
        ctx.check(L.ftk_pyramid_build(ctx._h, pyr._h, 0, 2 * n_pairs))

    class Run:
        """One tracker list on the resident batch: device outputs per tracker, the resident step, the end-to-end step."""

        def __init__(self, tracker_list):
            self.trackers = tracker_list
            self.params = [klt_params(t) for t in tracker_list]
            self.d_cur = [torch.empty_like(d_ref_uv) for _ in tracker_list]
            self.d_st = [torch.empty((n_total,), dtype=torch.uint8, device=dev) for _ in tracker_list]
            self.h_cur = torch.empty((len(tracker_list), n_total, 2), dtype=torch.float32).pin_memory()
            self.h_st = torch.empty((len(tracker_list), n_total), dtype=torch.uint8).pin_memory()

        def klt_only(self, i):
            flags = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_PREDICTION | _capi.FLAG_NO_STATUS
            ctx.check(L.ftk_klt_track(ctx._h, C.byref(self.params[i]), pyr._h, pyr._h, n_pairs, vp(d_ref_idx.data_ptr()), vp(d_cur_idx.data_ptr()),
                                      vp(d_offsets.data_ptr()), vp(d_ref_uv.data_ptr()), vp(self.d_cur[i].data_ptr()), vp(self.d_st[i].data_ptr()), flags))

        def step_resident(self):
            pyramid_only()
            for i in range(len(self.trackers)):
                self.klt_only(i)

        def step_e2e(self):
            # the user-facing call: host images + host features in, host results of every tracker out (H2D of chunk k+1 overlaps
            # compute of chunk k; the images cross PCIe once for all trackers of the workload)
            flags = _capi.FLAG_NO_PREDICTION | _capi.FLAG_NO_STATUS
            arr = (_capi.KltParams * len(self.params))(*self.params)
            ctx.check(L.ftk_track_image_pairs_multi(ctx._h, arr, len(self.params), ROWS, COLS, LEVELS, n_pairs, vp(host_images.data_ptr()),
                                                    vp(host_images.data_ptr() + n_pairs * plane), vp(offsets.ctypes.data), vp(host_ref_uv.data_ptr()),
                                                    vp(self.h_cur.data_ptr()), vp(self.h_st.data_ptr()), flags))

        def measure(self, steps, warmup, sample_clocks):
            n_track = len(self.trackers)
            for _ in range(max(warmup, 3)):
                self.step_resident()
            sampler = ClockSampler(local_rank) if sample_clocks else None
            if sampler:
                sampler.start()
            ms, launches = timed(self.step_resident, steps)
            clocks = sampler.stop() if sampler else None
            per_step = n_track * n_total * world
            res = {"value": per_step * steps / (ms * 1e-3), "ms_per_step": ms / steps, "launches": launches, "clocks": clocks}
            ms_pyr, _ = timed(pyramid_only, steps)
            res["kernel_ms"] = {"pyramid": ms_pyr / steps}
            for i, t in enumerate(self.trackers):
                ms_k, _ = timed(lambda i=i: self.klt_only(i), steps)
                res["kernel_ms"][f"klt_{t[0]}_{t[1]}"] = ms_k / steps
            res["tracked_fraction"] = {f"{t[0]}_{t[1]}": float((self.d_st[i] == 1).float().mean().item()) for i, t in enumerate(self.trackers)}
            for _ in range(2):
                self.step_e2e()
            ms_e2e, _ = timed(self.step_e2e, steps)
            h2d = 2 * n_pairs * plane + n_total * 8 + (n_pairs + 1) * 4 + 2 * n_pairs * 4
            d2h = n_track * n_total * 9
            res["e2e"] = {"value": per_step * steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                          "ms_per_step": ms_e2e / steps, "h2d_gbs_per_rank": h2d / (ms_e2e / steps * 1e-3) / 1e9,
                          "api": "ftk_track_image_pairs_multi: one call for all trackers of the workload (pinned host images + features in, host results "
                                 "of every tracker out; H2D of chunk k+1 overlaps compute of chunk k)"}
            return res

        def rooflines(self, res, oracle, fp32_peak):
            """Per tracker: algorithmic flops (oracle iteration trace) / kernel time, against the derived fp32 issue peak; plus the
            at-scale parity spot check of the timed run's outputs against the oracle."""
            from oracle import pyoracle as po
            out, parity = {}, {"pairs": 0, "status_mismatch": 0, "position_bits_mismatch": 0}
            for i, t in enumerate(self.trackers):
                cparams = po.make_params(t[0], t[1], half=t[2], max_points=max(500, n_feat))
                st_host = self.d_st[i].cpu().numpy()
                uv_host = self.d_cur[i].cpu().numpy()
                iters = 0
                for u in range(UNIQUE_PAIRS):
                    if u not in uniq:
                        continue
                    cu, st, it = oracle_trace(oracle, cparams, refs, curs, uvs, u, n_feat)
                    iters += int(it.sum()) * sum(1 for x in uniq if x == u)
                    p0 = uniq.index(u)
                    sl = slice(p0 * n_feat, (p0 + 1) * n_feat)
                    parity["pairs"] += 1
                    parity["status_mismatch"] += int((st_host[sl] != st).sum())
                    same = (uv_host[sl].view(np.uint32) == cu.view(np.uint32)) | (np.isnan(uv_host[sl]) & np.isnan(cu))
                    parity["position_bits_mismatch"] += int((~same).any(1).sum())
                flops, formula = algorithmic_flops(t, iters, n_total)
                traffic = None
                try:  # dram bytes of one ncu --set full capture of this kernel, scaled from its 200 000-feature launch to this launch
                    key = {("basic", "inverse"): "BasicInverseFastKernel", ("affine", "direct"): "KltKernel_affine_direct",
                           ("affine", "fast"): "KltKernel_affine_fast"}[(t[0], t[1])]
                    tr = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_JSON)))[key]
                    traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["features_per_launch"] * n_total
                except Exception:
                    pass
                ms_k = res["kernel_ms"][f"klt_{t[0]}_{t[1]}"]
                tfs = flops / (ms_k * 1e-3) / 1e12
                out[f"{t[0]}_{t[1]}"] = {"bound": "fp32", "bound_note": "FP32 CUDA-core issue + L1/shared bandwidth; neither HBM nor tensor (SURVEY 8(d))", "achieved": tfs,
                                         "peak": fp32_peak["tflops"], "unit": "TFLOP/s", "frac": tfs / fp32_peak["tflops"], "traffic": traffic,
                                         "peak_source": fp32_peak["source"],
                                         "algorithmic_flops_per_launch": flops, "algorithmic_flops_formula": formula,
                                         "patch_iterations_per_feature": iters / n_total, "launch_ms": ms_k,
                                         "kernel": "BasicInverseFastKernel<15,15>" if t == ("basic", "inverse", 7) else f"KltKernel ({tracker_name(t)})"}
            return out, parity

    head = Run(trackers)
    res = head.measure(args.steps, args.warmup, sample_clocks=True)

    north = north_res = None
    if trackers == WORKLOADS["configs1"] and not args.no_north_star:
        north = Run(WORKLOADS["north_star"])
        north_res = north.measure(args.steps, args.warmup, sample_clocks=False)
        # the temporal form: n_pairs + 1 host frames, every frame uploaded once.  The sequence runs through ALL unique pairs in
        # blocks (ref_u, cur_u, ref_u, cur_u, ...: consecutive frames form real pairs in alternating direction); the one pair that joins two
        # blocks shows unrelated images and carries no features.  Every rank sees the same mix, rotated by its rank.
        block = 31
        hs = host_images.numpy()[:n_pairs + 1]  # reuse the pinned buffer: the resident pyramids are not needed any more
        seq_counts = np.zeros(n_pairs, np.int64)
        hru = host_ref_uv.numpy()
        f = 0
        for k in range(n_pairs + 1):
            b, j = divmod(k, block)
            u = (b + rank) % UNIQUE_PAIRS
            hs[k] = refs[u] if j % 2 == 0 else curs[u]
            if k < n_pairs and j != block - 1:  # pair k = frame k -> k + 1 inside one block
                seq_counts[k] = n_feat
                hru[f:f + n_feat] = uvs[u]
                f += n_feat
        seq_offsets = np.concatenate([[0], np.cumsum(seq_counts)]).astype(np.int32)
        n_seq = int(seq_offsets[-1])

        def step_e2e_sequence():
            flags = _capi.FLAG_NO_PREDICTION | _capi.FLAG_NO_STATUS
            ctx.check(L.ftk_track_image_sequence(ctx._h, C.byref(north.params[0]), ROWS, COLS, LEVELS, n_pairs + 1, vp(host_images.data_ptr()),
                                                 vp(seq_offsets.ctypes.data), vp(host_ref_uv.data_ptr()), vp(host_cur_uv.data_ptr()), vp(host_status.data_ptr()), flags))

        for _ in range(2):
            step_e2e_sequence()
        ms_seq, _ = timed(step_e2e_sequence, args.steps)
        north_res["e2e_sequence"] = {"value": n_seq * world * args.steps / (ms_seq * 1e-3), "unit": UNIT, "ms_per_step": ms_seq / args.steps,
                                     "h2d_bytes_per_step": ((n_pairs + 1) * plane + n_seq * 8) * world, "d2h_bytes_per_step": n_seq * 9 * world,
                                     "h2d_gbs_per_rank": ((n_pairs + 1) * plane + n_seq * 8) / (ms_seq / args.steps * 1e-3) / 1e9,
                                     "tracked_features_per_step_per_gpu": n_seq, "tracked_fraction": float((host_status.numpy()[:n_seq] == 1).mean()),
                                     "api": "ftk_track_image_sequence (n_pairs + 1 host frames, pair k = frame k -> k+1; every frame uploaded and its pyramid "
                                            "built once; the sequence walks through all unique pairs in blocks of 31 frames, the pair joining two blocks has no features)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    # pyramid: algorithmic bytes = read H*W + write H*W*(1/4+1/16+1/64) per image (479 400 B per 752x480 image)
    pyr_bytes = sum((ROWS >> l) * (COLS >> l) for l in range(LEVELS)) * 2 * n_pairs
    pyr_gbs = pyr_bytes / (res["kernel_ms"]["pyramid"] * 1e-3) / 1e9

    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "clocks": res["clocks"], "gpu_launches": res["launches"],
        "e2e": res["e2e"],
        "tracked_fraction": res["tracked_fraction"],
        "kernel_ms": res["kernel_ms"],
        "roofline_pyramid": {"bound": "hbm", "achieved": pyr_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": pyr_gbs / hbm_peak,
                             "traffic": pyramid_traffic(2 * n_pairs), "peak_source": hbm_src, "algorithmic_bytes_per_launch": pyr_bytes},
    }
    fp32_peak = measured_fp32_peak(local_rank)
    line["fp32_issue_peak"] = {"tflops": fp32_peak["tflops"], "source": fp32_peak["source"]}

    if not args.no_cpu_baseline:
        from oracle import pyoracle as po
        lib_cpu, kind = cpu_checker()
        cores = os.cpu_count() or 1
        plist = cpu_params(trackers, n_feat)
        sample_pairs = max(1, min(cores, 64))
        cpu_track_sample(lib_cpu, plist, refs, curs, uvs, min(sample_pairs, cores), cores)
        dt, nf = cpu_track_sample(lib_cpu, plist, refs, curs, uvs, sample_pairs, cores)
        dt1, nf1 = cpu_track_sample(lib_cpu, plist, refs, curs, uvs, 1, 1)
        line["cpu_baseline"] = {"value": nf / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{sample_pairs} frame pairs x {n_feat} features, per pair and tracker: CreateImagePyramid x2 + TrackFeatures; one "
                                          "tracker object per host thread",
                                "single_thread_value": nf1 / dt1}
        oracle = po.OracleLib()
        roofs, parity = head.rooflines(res, oracle, fp32_peak)
        # the dominant kernel of the step = the tracker kernel with the longest launch
        dominant = max(roofs, key=lambda k: roofs[k]["launch_ms"])
        line["roofline"] = roofs[dominant]
        line["roofline_by_tracker"] = roofs
        line["parity_check"] = parity
        if north is not None:
            nroofs, nparity = north.rooflines(north_res, oracle, fp32_peak)
            dt, nf = cpu_track_sample(lib_cpu, cpu_params(WORKLOADS["north_star"], n_feat), refs, curs, uvs, sample_pairs, cores)
            north_res["roofline"] = nroofs["basic_inverse"]
            north_res["parity_check"] = nparity
            north_res["cpu_baseline"] = {"value": nf / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{sample_pairs} frame pairs x {n_feat} features"}
    if north_res is not None:
        north_res.pop("clocks", None)
        north_res["config"] = "BASELINE north_star target: basic KLT kInverse 15x15, 4 levels, the same 1000 x 2000 batch; target >= 1e8 features/s/GPU"
        north_res["unit"] = UNIT
        line["north_star"] = north_res
        # a compact copy inside `e2e` (the scaling harness keeps that object whole): whole-job north-star throughput at this N
        line["e2e"]["north_star"] = {
            "tracker": "basic KLT kInverse 15x15", "device_resident_value": north_res["value"],
            "e2e_sequence_value": north_res["e2e_sequence"]["value"], "e2e_sequence_h2d_gbs_per_rank": north_res["e2e_sequence"]["h2d_gbs_per_rank"],
            "e2e_pairs_value": north_res["e2e"]["value"], "e2e_pairs_h2d_gbs_per_rank": north_res["e2e"]["h2d_gbs_per_rank"], "unit": UNIT,
            "note": "sequence = ftk_track_image_sequence (every frame uploaded once: the VIO use); pairs = ftk_track_image_pairs_multi (both frames of every pair uploaded)"}
    if not args.no_extras and world == 1:
        try:
            ctx.synchronize()
            line["c1_latency"] = c1_latency(*(cpu_checker() if not args.no_cpu_baseline else (None, None)))
        except Exception as e:  # noqa: BLE001
            line["c1_latency"] = {"error": repr(e)}
        try:
            keep = {}
            line["other_workloads"] = run_extras(ctx, L, torch, local_rank, args.steps, peaks, keep)
            if not args.no_cpu_baseline:
                cpu_extras(line["other_workloads"], keep, fp32_peak)
        except Exception as e:  # the headline must survive a failure in the extras
            import traceback
            line["other_workloads"] = {"error": repr(e), "traceback": traceback.format_exc()[-1500:]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_match_sharded(args):
    """SURVEY 8(e), matching: the reference rows of BASELINE configs[3] (BRIEF-256 10k x 10k force + nearby) and configs[4] (float-256
    20k x 20k force) are split into contiguous blocks over the ranks, the current set is replicated on every GPU, no collective on
    the data path; the index vectors are gathered on rank 0 (NCCL) and compared with the whole problem solved on rank 0's GPU.  Strong
    scaling: the total work is fixed.  Timing: CUDA events per rank, barrier on both sides, max over ranks."""
    import torch
    import torch.distributed as dist

    import feature_tracker_b200 as ft
    from feature_tracker_b200 import _capi, sharding, synthetic as S
    from feature_tracker_b200.api import lib as ftk_lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    ctx = ft.Context(local_rank)
    L = ftk_lib()
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    vp = C.c_void_p
    fl = _capi.FLAG_DEVICE_POINTERS | _capi.FLAG_NO_INDEX_INPUT
    steps = max(args.steps, 5)

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    out = {}
    # ---- C4 ----
    rb, cb, pred, pos, _ = S.make_brief_sets(10000, 10000, seed=99)
    lo, hi = sharding.shard_bounds(10000, world, rank)
    d_c = torch.from_numpy(ft.pack_brief(cb).view(np.int32)).to(dev)
    d_pos = torch.from_numpy(pos).to(dev)
    d_r = torch.from_numpy(ft.pack_brief(rb[lo:hi]).view(np.int32)).to(dev)
    d_pred = torch.from_numpy(pred[lo:hi]).to(dev)
    d_idx = torch.full((hi - lo,), -1, dtype=torch.int32, device=dev)
    full = {}
    if rank == 0:
        d_rf_all = torch.from_numpy(ft.pack_brief(rb).view(np.int32)).to(dev)
        d_pred_all = torch.from_numpy(pred).to(dev)
        d_all = torch.full((10000,), -1, dtype=torch.int32, device=dev)
        ctx.check(L.ftk_match_hamming_force(ctx._h, vp(d_rf_all.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, 60.0, vp(d_all.data_ptr()), fl))
        ctx.synchronize()
        full["c4_force"] = d_all.cpu().numpy().copy()
        ctx.check(L.ftk_match_hamming_nearby(ctx._h, vp(d_rf_all.data_ptr()), 10000, vp(d_c.data_ptr()), 10000, 8, vp(d_pred_all.data_ptr()), vp(d_pos.data_ptr()), 50, 50,
                                             60.0, vp(d_all.data_ptr()), fl))
        ctx.synchronize()
        full["c4_nearby"] = d_all.cpu().numpy().copy()
    ms = timed(lambda: ctx.check(L.ftk_match_hamming_force(ctx._h, vp(d_r.data_ptr()), hi - lo, vp(d_c.data_ptr()), 10000, 8, 60.0, vp(d_idx.data_ptr()), fl)))
    ctx.synchronize()
    got = sharding.match_sharded(lambda a, b: d_idx.cpu().numpy(), 10000, world, rank)
    out["C4_brief256_force_10k_x_10k"] = {"ms": ms, "pairs_per_s": 1e8 / (ms * 1e-3), "rows_per_rank": hi - lo}
    if rank == 0:
        out["C4_brief256_force_10k_x_10k"]["index_mismatch_vs_single_gpu"] = int((got != full["c4_force"]).sum())
    ms = timed(lambda: ctx.check(L.ftk_match_hamming_nearby(ctx._h, vp(d_r.data_ptr()), hi - lo, vp(d_c.data_ptr()), 10000, 8, vp(d_pred.data_ptr()), vp(d_pos.data_ptr()),
                                                            50, 50, 60.0, vp(d_idx.data_ptr()), fl)))
    ctx.synchronize()
    got = sharding.match_sharded(lambda a, b: d_idx.cpu().numpy(), 10000, world, rank)
    out["C4_brief256_nearby_10k_window50"] = {"ms": ms, "ref_rows_per_s": 1e4 / (ms * 1e-3)}
    if rank == 0:
        out["C4_brief256_nearby_10k_window50"]["index_mismatch_vs_single_gpu"] = int((got != full["c4_nearby"]).sum())
    # ---- C5 ----
    rf, cf = S.make_float_sets(20000, 20000, seed=5)
    lo, hi = sharding.shard_bounds(20000, world, rank)
    d_cf = torch.from_numpy(cf).to(dev)
    d_rf = torch.from_numpy(np.ascontiguousarray(rf[lo:hi])).to(dev)
    d_idx2 = torch.full((hi - lo,), -1, dtype=torch.int32, device=dev)
    if rank == 0:
        d_rf_all = torch.from_numpy(rf).to(dev)
        d_all = torch.full((20000,), -1, dtype=torch.int32, device=dev)
        ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf_all.data_ptr()), 20000, vp(d_cf.data_ptr()), 20000, 256, 0.1, vp(d_all.data_ptr()), fl))
        ctx.synchronize()
        full["c5"] = d_all.cpu().numpy().copy()
    ms5 = timed(lambda: ctx.check(L.ftk_match_cosine_force(ctx._h, vp(d_rf.data_ptr()), hi - lo, vp(d_cf.data_ptr()), 20000, 256, 0.1, vp(d_idx2.data_ptr()), fl)))
    ctx.synchronize()
    got = sharding.match_sharded(lambda a, b: d_idx2.cpu().numpy(), 20000, world, rank)
    flop = 2.0 * 20000 * 20000 * 256
    out["C5_float256_force_20k_x_20k"] = {"ms": ms5, "pairs_per_s": 4e8 / (ms5 * 1e-3), "tflops_whole_job": flop / (ms5 * 1e-3) / 1e12, "rows_per_rank": hi - lo}
    if rank == 0:
        out["C5_float256_force_20k_x_20k"]["index_mismatch_vs_single_gpu"] = int((got != full["c5"]).sum())
        line = {"metric": "descriptor pairs/sec (float-256 force match 20k x 20k, ref rows sharded)", "value": 4e8 / (ms5 * 1e-3), "unit": "pairs/s", "n_gpus": world,
                "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms5, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16 screen + f32 exact",
                "data": "synthetic", "config": {"workload": "BASELINE configs[3] + configs[4], reference rows split over the ranks, current set replicated, no collective on "
                                                             "the data path; results gathered on rank 0 and compared with the single-GPU result",
                                                "sharding": "ref rows, contiguous blocks"}, "workloads": out}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "match_sharded":
        run_match_sharded(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
