"""ORACLE bindings (test infrastructure -- never imported by the product package).

ctypes wrappers for the two CPU checkers:
  * ``RefLib``    -> oracle/_ref/libftk_ref.so, the reference's own .cpp compiled in place (kind "reference")
  * ``OracleLib`` -> oracle/_build/libftk_oracle.so, the plain-C restatement (kind "port")
Both expose the same Python methods so tests can compare them with each other and with the CUDA path.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libftk_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "libftk_oracle.so")
REFERENCE_ROOT = "/root/reference"

VARIANTS = {"basic": 0, "affine": 1, "lssd": 2}
METHODS = {"inverse": 0, "direct": 1, "fast": 2}


class KltParams(C.Structure):
    """Mirror of ftko_klt_params / ftk_klt_params (same layout in all three libraries)."""

    _fields_ = [
        ("variant", C.c_int32),
        ("method", C.c_int32),
        ("max_track_points", C.c_uint32),
        ("max_iteration", C.c_uint32),
        ("max_tolerance_large_step", C.c_uint32),
        ("patch_row_half", C.c_int32),
        ("patch_col_half", C.c_int32),
        ("max_converge_step", C.c_float),
        ("predict", C.c_float * 4),
        ("consider_patch_luminance", C.c_int32),
    ]


class DirectParams(C.Structure):
    """ftko_direct_params == DirectMethodOptions (src/direct_method_tracker/direct_method_tracker.h:20-28)."""

    _fields_ = [
        ("max_track_points", C.c_uint32),
        ("max_iteration", C.c_uint32),
        ("patch_row_half", C.c_int32),
        ("patch_col_half", C.c_int32),
        ("max_converge_step", C.c_float),
        ("max_converge_residual", C.c_float),
        ("method", C.c_int32),
    ]


def make_direct_params(half=6, half_col=None, max_points=500, max_iter=15, converge=1e-6, method="direct"):
    p = DirectParams()
    p.max_track_points = max_points
    p.max_iteration = max_iter
    p.patch_row_half = half
    p.patch_col_half = half if half_col is None else half_col
    p.max_converge_step = converge
    p.max_converge_residual = 2.0
    p.method = {"inverse": 0, "direct": 1, "fast": 2}[method]
    return p


class DenseFlowParams(C.Structure):
    """ftko_dense_flow_params == DenseOpticalFlow::Options (src/dense_optical_flow_tracker/dense_optical_flow.h:15-20)."""

    _fields_ = [("max_iteration", C.c_int32), ("half_patch_size", C.c_int32), ("max_converge_step", C.c_float), ("max_delta_flow_step", C.c_float)]


class DetectorParams(C.Structure):
    """ftko_detector_params (parity unpinned: Feature_Detector is absent; see oracle/ftk_oracle.c)."""

    _fields_ = [("kind", C.c_int32), ("half_patch", C.c_int32), ("harris_k", C.c_float), ("min_response", C.c_float), ("min_distance", C.c_int32)]


def make_detector_params(kind="harris", half=1, k=0.04, min_response=40.0, min_distance=20):
    p = DetectorParams()
    p.kind = {"harris": 0, "shi_tomasi": 1}[kind] if isinstance(kind, str) else int(kind)
    p.half_patch, p.harris_k, p.min_response, p.min_distance = half, k, min_response, min_distance
    return p


def make_dense_flow_params(max_iter=10, half=2, converge=1e-6, max_step=1.0):
    p = DenseFlowParams()
    p.max_iteration, p.half_patch_size, p.max_converge_step, p.max_delta_flow_step = max_iter, half, converge, max_step
    return p


def make_params(variant="basic", method="fast", half=6, half_col=None, max_points=500, max_iter=15, max_large=3,
                converge=4e-2, predict=(1.0, 0.0, 0.0, 1.0), luminance=False):
    """Defaults are the reference's OpticalFlowOptions defaults (optical_flow.h:20-28)."""
    p = KltParams()
    p.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    p.method = METHODS[method] if isinstance(method, str) else int(method)
    p.max_track_points = max_points
    p.max_iteration = max_iter
    p.max_tolerance_large_step = max_large
    p.patch_row_half = half
    p.patch_col_half = half if half_col is None else half_col
    p.max_converge_step = converge
    for i in range(4):
        p.predict[i] = predict[i]
    p.consider_patch_luminance = 1 if luminance else 0
    return p


def build_ref(force=False):
    """Compile the reference sources in place (only possible where /root/reference exists)."""
    if os.path.exists(REF_SO) and not force:
        return True
    if not os.path.isdir(REFERENCE_ROOT):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return os.path.exists(REF_SO)


def build_oracle(force=False):
    if force and os.path.exists(ORACLE_SO):
        os.remove(ORACLE_SO)
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return os.path.exists(ORACLE_SO)


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def pyramid_level_shapes(rows, cols, levels):
    return [(rows >> i, cols >> i) for i in range(levels)]


class _CpuChecker:
    """Common Python surface over a CPU checker library with symbol prefix ``self.prefix``."""

    prefix = None

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.path = path

    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    # -- pyramid -----------------------------------------------------------------------------------------------
    def pyramid_build(self, image, levels):
        """Returns the list of level images [level0(view), level1, ...] (uint8 2-D arrays)."""
        image = np.ascontiguousarray(image, dtype=np.uint8)
        rows, cols = image.shape
        shapes = pyramid_level_shapes(rows, cols, levels)
        total = sum(r * c for r, c in shapes[1:])
        out = np.zeros(max(total, 1), dtype=np.uint8)
        ok = self._fn("pyramid_build")(_u8p(image), C.c_int32(rows), C.c_int32(cols), C.c_int32(levels), _u8p(out))
        assert ok == 1
        res = [image]
        off = 0
        for r, c in shapes[1:]:
            res.append(out[off:off + r * c].reshape(r, c).copy())
            off += r * c
        return res

    # -- KLT ---------------------------------------------------------------------------------------------------
    def klt_track(self, params, ref_levels, cur_levels, ref_uv, cur_uv=None, status=None, single_level=False):
        """ref_levels/cur_levels: lists of uint8 2-D arrays.  cur_uv/status None == empty vectors on entry.
        Returns (ok, cur_uv[n,2] float32, status[n] uint8)."""
        levels = len(ref_levels)
        ref_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in ref_levels]
        cur_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in cur_levels]
        rows = np.array([a.shape[0] for a in ref_levels], dtype=np.int32)
        cols = np.array([a.shape[1] for a in ref_levels], dtype=np.int32)
        PtrArr = C.POINTER(C.c_uint8) * levels
        rp = PtrArr(*[_u8p(a) for a in ref_levels])
        cp = PtrArr(*[_u8p(a) for a in cur_levels])
        ref_uv = np.ascontiguousarray(ref_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        if cur_uv is None:
            cur_buf, cur_count = np.zeros((max(n, 1), 2), np.float32), 0
        else:
            cur_in = np.ascontiguousarray(cur_uv, dtype=np.float32).reshape(-1, 2)
            cur_count = cur_in.shape[0]
            cur_buf = np.zeros((max(n, cur_count, 1), 2), np.float32)
            cur_buf[:cur_count] = cur_in
        if status is None:
            st_buf, st_count = np.zeros(max(n, 1), np.uint8), 0
        else:
            st_in = np.ascontiguousarray(status, dtype=np.uint8).reshape(-1)
            st_count = st_in.shape[0]
            st_buf = np.zeros(max(n, st_count, 1), np.uint8)
            st_buf[:st_count] = st_in
        ok = self._fn("klt_track")(C.byref(params), C.c_int32(levels), rp, cp, _i32p(rows), _i32p(cols), C.c_int32(n), _f32p(ref_uv),
                                    _f32p(cur_buf), C.c_int32(cur_count), _u8p(st_buf), C.c_int32(st_count),
                                    C.c_int32(1 if single_level else 0))
        return ok == 1, cur_buf[:n].copy(), st_buf[:n].copy()

    def direct_method_track(self, params, ref_levels, cur_levels, K, p_c_in_ref, ref_uv, q_rc, p_rc, cur_uv=None, status=None):
        """DirectMethod::TrackFeatures, camera-frame overload (direct_method_tracker.cpp:41-95).  q_rc = (w, x, y, z).
        Returns (ok, cur_uv, q_rc, p_rc, status)."""
        levels = len(ref_levels)
        ref_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in ref_levels]
        cur_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in cur_levels]
        rows = np.array([a.shape[0] for a in ref_levels], dtype=np.int32)
        cols = np.array([a.shape[1] for a in ref_levels], dtype=np.int32)
        PtrArr = C.POINTER(C.c_uint8) * levels
        rp = PtrArr(*[_u8p(a) for a in ref_levels])
        cp = PtrArr(*[_u8p(a) for a in cur_levels])
        ref_uv = np.ascontiguousarray(ref_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        pts = np.ascontiguousarray(p_c_in_ref, dtype=np.float32).reshape(-1, 3)
        assert pts.shape[0] == n
        Kc = np.ascontiguousarray(K, dtype=np.float32).reshape(4)
        q = np.ascontiguousarray(q_rc, dtype=np.float32).reshape(4).copy()
        p = np.ascontiguousarray(p_rc, dtype=np.float32).reshape(3).copy()
        if cur_uv is None:
            cur_buf, cur_count = np.zeros((max(n, 1), 2), np.float32), 0
        else:
            cur_in = np.ascontiguousarray(cur_uv, dtype=np.float32).reshape(-1, 2)
            cur_count = cur_in.shape[0]
            cur_buf = np.zeros((max(n, cur_count, 1), 2), np.float32)
            cur_buf[:cur_count] = cur_in
        if status is None:
            st_buf, st_count = np.zeros(max(n, 1), np.uint8), 0
        else:
            st_in = np.ascontiguousarray(status, dtype=np.uint8).reshape(-1)
            st_count = st_in.shape[0]
            st_buf = np.zeros(max(n, st_count, 1), np.uint8)
            st_buf[:st_count] = st_in
        ok = self._fn("direct_method_track")(C.byref(params), C.c_int32(levels), rp, cp, _i32p(rows), _i32p(cols), _f32p(Kc), C.c_int32(n), _f32p(pts),
                                              _f32p(ref_uv), _f32p(cur_buf), C.c_int32(cur_count), _f32p(q), _f32p(p), _u8p(st_buf), C.c_int32(st_count))
        return ok == 1, cur_buf[:n].copy(), q, p, st_buf[:n].copy()

    def dense_flow_track(self, params, ref_levels, cur_levels, single_level=False, flow=None):
        """DenseOpticalFlow::Track (dense_optical_flow.cpp:7-85).  flow = (flow_row, flow_col) initial content for the single-level
        overload (None = matrices of the wrong size).  Returns (ok, flow_row, flow_col), each rows x cols float32."""
        levels = len(ref_levels)
        ref_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in ref_levels]
        cur_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in cur_levels]
        rows = np.array([a.shape[0] for a in ref_levels], dtype=np.int32)
        cols = np.array([a.shape[1] for a in ref_levels], dtype=np.int32)
        PtrArr = C.POINTER(C.c_uint8) * levels
        rp = PtrArr(*[_u8p(a) for a in ref_levels])
        cp = PtrArr(*[_u8p(a) for a in cur_levels])
        fr = np.zeros((rows[0], cols[0]), np.float32)
        fc = np.zeros((rows[0], cols[0]), np.float32)
        valid = 0
        if flow is not None:
            fr[:] = flow[0]
            fc[:] = flow[1]
            valid = 1
        ok = self._fn("dense_flow_track")(C.byref(params), C.c_int32(levels), rp, cp, _i32p(rows), _i32p(cols), C.c_int32(1 if single_level else 0),
                                           C.c_int32(valid), _f32p(fr), _f32p(fc))
        return ok == 1, fr, fc

    def pyramid_and_track(self, params, levels, ref_image, cur_image, ref_uv):
        ref_image = np.ascontiguousarray(ref_image, dtype=np.uint8)
        cur_image = np.ascontiguousarray(cur_image, dtype=np.uint8)
        rows, cols = ref_image.shape
        ref_uv = np.ascontiguousarray(ref_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        cur = np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.uint8)
        ok = self._fn("pyramid_and_track")(C.byref(params), C.c_int32(levels), _u8p(ref_image), _u8p(cur_image), C.c_int32(rows), C.c_int32(cols),
                                            C.c_int32(n), _f32p(ref_uv), _f32p(cur), _u8p(st))
        return ok == 1, cur, st

    # -- matching ----------------------------------------------------------------------------------------------
    @staticmethod
    def _idx(idx, n_ref):
        if idx is None:
            return np.zeros(max(n_ref, 1), np.int32), 0
        idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1)
        buf = np.zeros(max(n_ref, idx.shape[0], 1), np.int32)
        buf[:idx.shape[0]] = idx
        return buf, idx.shape[0]

    def match_brief_force(self, ref_bits, cur_bits, max_dist, idx=None):
        """ref_bits/cur_bits: [n, len] uint8 arrays of 0/1.  idx None == empty index vector on entry."""
        ref_bits = np.ascontiguousarray(ref_bits, dtype=np.uint8)
        cur_bits = np.ascontiguousarray(cur_bits, dtype=np.uint8)
        n_ref, ln = ref_bits.shape if ref_bits.ndim == 2 else (0, cur_bits.shape[1])
        n_cur = cur_bits.shape[0]
        buf, cnt = self._idx(idx, n_ref)
        ok = self._fn("match_brief_force")(_u8p(ref_bits), C.c_int32(n_ref), _u8p(cur_bits), C.c_int32(n_cur), C.c_int32(ln), C.c_float(max_dist),
                                            _i32p(buf), C.c_int32(cnt))
        return ok == 1, buf[:n_ref].copy()

    def match_brief_nearby(self, ref_bits, cur_bits, pred_uv, cur_uv, max_drow, max_dcol, max_dist, idx=None):
        ref_bits = np.ascontiguousarray(ref_bits, dtype=np.uint8)
        cur_bits = np.ascontiguousarray(cur_bits, dtype=np.uint8)
        n_ref, ln = ref_bits.shape
        n_cur = cur_bits.shape[0]
        pred_uv = np.ascontiguousarray(pred_uv, dtype=np.float32).reshape(-1, 2)
        cur_uv = np.ascontiguousarray(cur_uv, dtype=np.float32).reshape(-1, 2)
        assert pred_uv.shape[0] == n_ref and cur_uv.shape[0] == n_cur
        buf, cnt = self._idx(idx, n_ref)
        ok = self._fn("match_brief_nearby")(_u8p(ref_bits), C.c_int32(n_ref), _u8p(cur_bits), C.c_int32(n_cur), C.c_int32(ln), _f32p(pred_uv),
                                             _f32p(cur_uv), C.c_int32(max_drow), C.c_int32(max_dcol), C.c_float(max_dist), _i32p(buf), C.c_int32(cnt))
        return ok == 1, buf[:n_ref].copy()

    def match_cosine_force(self, ref, cur, max_dist, idx=None):
        ref = np.ascontiguousarray(ref, dtype=np.float32)
        cur = np.ascontiguousarray(cur, dtype=np.float32)
        n_ref, dim = ref.shape
        n_cur = cur.shape[0]
        buf, cnt = self._idx(idx, n_ref)
        ok = self._fn("match_cosine_force")(_f32p(ref), C.c_int32(n_ref), _f32p(cur), C.c_int32(n_cur), C.c_int32(dim), C.c_float(max_dist), _i32p(buf),
                                             C.c_int32(cnt))
        return ok == 1, buf[:n_ref].copy()

    def match_cosine_nearby(self, ref, cur, pred_uv, cur_uv, max_drow, max_dcol, max_dist, idx=None):
        ref = np.ascontiguousarray(ref, dtype=np.float32)
        cur = np.ascontiguousarray(cur, dtype=np.float32)
        n_ref, dim = ref.shape
        n_cur = cur.shape[0]
        pred_uv = np.ascontiguousarray(pred_uv, dtype=np.float32).reshape(-1, 2)
        cur_uv = np.ascontiguousarray(cur_uv, dtype=np.float32).reshape(-1, 2)
        buf, cnt = self._idx(idx, n_ref)
        ok = self._fn("match_cosine_nearby")(_f32p(ref), C.c_int32(n_ref), _f32p(cur), C.c_int32(n_cur), C.c_int32(dim), _f32p(pred_uv), _f32p(cur_uv),
                                              C.c_int32(max_drow), C.c_int32(max_dcol), C.c_float(max_dist), _i32p(buf), C.c_int32(cnt))
        return ok == 1, buf[:n_ref].copy()

    def match_rows_threaded(self, kind, ref, cur, max_dist, threads=None, pred_uv=None, cur_uv=None, max_drow=0, max_dcol=0):
        """A full ForceMatch / NearbyMatch of `kind` in {"brief_force", "brief_nearby", "cosine_force", "cosine_nearby"} with the ref rows
        split over host threads (rows are independent: descriptor_matcher.h:67-76, :103-121, and ctypes drops the GIL), every slice
        computed by the checker's own single-threaded entry point with an empty index vector on entry.  Returns (ok, idx[n_ref])."""
        import threading
        n_ref = len(ref)
        threads = max(1, min(threads or (os.cpu_count() or 1), n_ref))
        bounds = [n_ref * t // threads for t in range(threads + 1)]
        out, oks = np.full(n_ref, -1, np.int32), [True] * threads

        def work(t):
            a, b = bounds[t], bounds[t + 1]
            if a == b:
                return
            if kind == "brief_force":
                ok, idx = self.match_brief_force(ref[a:b], cur, max_dist)
            elif kind == "brief_nearby":
                ok, idx = self.match_brief_nearby(ref[a:b], cur, pred_uv[a:b], cur_uv, max_drow, max_dcol, max_dist)
            elif kind == "cosine_force":
                ok, idx = self.match_cosine_force(ref[a:b], cur, max_dist)
            elif kind == "cosine_nearby":
                ok, idx = self.match_cosine_nearby(ref[a:b], cur, pred_uv[a:b], cur_uv, max_drow, max_dcol, max_dist)
            else:
                raise ValueError(kind)
            oks[t] = ok
            out[a:b] = idx

        ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return all(oks), out

    def match_brief_nearby_uv(self, ref_bits, cur_bits, pred_uv, cur_uv, max_drow, max_dcol, max_dist, status=None):
        ref_bits = np.ascontiguousarray(ref_bits, dtype=np.uint8)
        cur_bits = np.ascontiguousarray(cur_bits, dtype=np.uint8)
        n_ref, ln = ref_bits.shape
        n_cur = cur_bits.shape[0]
        pred_uv = np.ascontiguousarray(pred_uv, dtype=np.float32).reshape(-1, 2)
        cur_uv = np.ascontiguousarray(cur_uv, dtype=np.float32).reshape(-1, 2)
        matched = np.zeros((n_ref, 2), np.float32)
        if status is None:
            st, cnt = np.zeros(max(n_ref, 1), np.uint8), 0
        else:
            s_in = np.ascontiguousarray(status, dtype=np.uint8)
            cnt = s_in.shape[0]
            st = np.zeros(max(n_ref, cnt, 1), np.uint8)
            st[:cnt] = s_in
        ok = self._fn("match_brief_nearby_uv")(_u8p(ref_bits), C.c_int32(n_ref), _u8p(cur_bits), C.c_int32(n_cur), C.c_int32(ln), _f32p(pred_uv),
                                                _f32p(cur_uv), C.c_int32(max_drow), C.c_int32(max_dcol), C.c_float(max_dist), _f32p(matched), _u8p(st),
                                                C.c_int32(cnt))
        return ok == 1, matched, st[:n_ref].copy()

    def match_brief_force_uv(self, ref_bits, cur_bits, cur_uv, max_dist, status=None):
        ref_bits = np.ascontiguousarray(ref_bits, dtype=np.uint8)
        cur_bits = np.ascontiguousarray(cur_bits, dtype=np.uint8)
        n_ref, ln = ref_bits.shape
        n_cur = cur_bits.shape[0]
        cur_uv = np.ascontiguousarray(cur_uv, dtype=np.float32).reshape(-1, 2)
        matched = np.zeros((n_ref, 2), np.float32)
        if status is None:
            st, cnt = np.zeros(max(n_ref, 1), np.uint8), 0
        else:
            s_in = np.ascontiguousarray(status, dtype=np.uint8)
            cnt = s_in.shape[0]
            st = np.zeros(max(n_ref, cnt, 1), np.uint8)
            st[:cnt] = s_in
        ok = self._fn("match_brief_force_uv")(_u8p(ref_bits), C.c_int32(n_ref), _u8p(cur_bits), C.c_int32(n_cur), C.c_int32(ln), _f32p(cur_uv),
                                               C.c_float(max_dist), _f32p(matched), _u8p(st), C.c_int32(cnt))
        return ok == 1, matched, st[:n_ref].copy()


class RefLib(_CpuChecker):
    prefix = "ftkref_"

    def __init__(self):
        if not build_ref():
            raise FileNotFoundError("oracle/_ref/libftk_ref.so is absent and /root/reference is not available to build it")
        super().__init__(REF_SO)

    def direct_method_track_world(self, params, ref_levels, cur_levels, K, ref_q_wc, ref_p_wc, p_w, ref_uv, cur_q_wc, cur_p_wc):
        """The world-frame overload (direct_method_tracker.cpp:8-39); empty cur_pixel_uv / status vectors on entry."""
        levels = len(ref_levels)
        ref_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in ref_levels]
        cur_levels = [np.ascontiguousarray(a, dtype=np.uint8) for a in cur_levels]
        rows = np.array([a.shape[0] for a in ref_levels], dtype=np.int32)
        cols = np.array([a.shape[1] for a in ref_levels], dtype=np.int32)
        PtrArr = C.POINTER(C.c_uint8) * levels
        rp = PtrArr(*[_u8p(a) for a in ref_levels])
        cp = PtrArr(*[_u8p(a) for a in cur_levels])
        ref_uv = np.ascontiguousarray(ref_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        pts = np.ascontiguousarray(p_w, dtype=np.float32).reshape(n, 3)
        f = lambda a, k: np.ascontiguousarray(a, dtype=np.float32).reshape(k).copy()
        Kc, rq, rpw, q, p = f(K, 4), f(ref_q_wc, 4), f(ref_p_wc, 3), f(cur_q_wc, 4), f(cur_p_wc, 3)
        cur_buf, st_buf = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
        ok = self._fn("direct_method_track_world")(C.byref(params), C.c_int32(levels), rp, cp, _i32p(rows), _i32p(cols), _f32p(Kc), _f32p(rq), _f32p(rpw),
                                                    C.c_int32(n), _f32p(pts), _f32p(ref_uv), _f32p(cur_buf), C.c_int32(0), _f32p(q), _f32p(p), _u8p(st_buf),
                                                    C.c_int32(0))
        return ok == 1, cur_buf, q, p, st_buf


    def nn_match_scores(self, scores, min_score):
        """NNFeatureMatcher::Match, score-matrix branch (nn_feature_matcher.cpp:180-216), by the reference's own code compiled in place
        against the ONNX Runtime stub of oracle/shim/onnx_run_time.h.  Needs n_ref <= n_cur.  Returns (ok, idx[n_ref])."""
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        n_ref, n_cur = scores.shape
        idx = np.full(max(n_ref, 1), -1, np.int32)
        ok = self._fn("nn_match_scores")(_f32p(scores), C.c_int32(n_ref), C.c_int32(n_cur), C.c_float(min_score), _i32p(idx))
        return ok == 1, idx[:n_ref].copy()

    def nn_match_pairs(self, matches, n_ref, n_cur):
        """The "matches" branch (nn_feature_matcher.cpp:160-178): matches = [n, 2] int64 (idx_ref, idx_cur).  Returns (ok, idx[n_ref])."""
        matches = np.ascontiguousarray(matches, dtype=np.int64).reshape(-1, 2)
        idx = np.full(max(n_ref, 1), -1, np.int32)
        ok = self._fn("nn_match_pairs")(matches.ctypes.data_as(C.c_void_p), C.c_int32(len(matches)), C.c_int32(n_ref), C.c_int32(n_cur), _i32p(idx))
        return ok == 1, idx[:n_ref].copy()


class OracleLib(_CpuChecker):
    prefix = "ftko_"

    def __init__(self):
        build_oracle()
        super().__init__(ORACLE_SO)

    def mutual_scores(self, scores, min_score):
        """nn_feature_matcher.cpp:180-216 (C restatement; pinned against the reference's own Match() -- RefLib.nn_match_scores -- by
        tests/test_oracle.py)."""
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        n_ref, n_cur = scores.shape
        idx = np.full(max(n_ref, 1), -1, np.int32)
        ok = self._fn("mutual_scores")(_f32p(scores), C.c_int32(n_ref), C.c_int32(n_cur), C.c_float(min_score), _i32p(idx))
        return ok == 1, idx[:n_ref].copy()

    # -- feature detection + BRIEF (restatement of the published algorithm; Feature_Detector's sources are absent) --------------
    def detect_response(self, params, image):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        out = np.zeros(image.shape, np.float32)
        ok = self._fn("detect_response")(C.byref(params), _u8p(image), C.c_int32(image.shape[0]), C.c_int32(image.shape[1]), _f32p(out))
        return ok == 1, out

    def detect_features(self, params, image, needed, existing=None):
        """Returns (ok, uv [n][2], response [n])."""
        image = np.ascontiguousarray(image, dtype=np.uint8)
        existing = np.zeros((0, 2), np.float32) if existing is None else np.ascontiguousarray(existing, np.float32).reshape(-1, 2)
        uv = np.zeros((max(needed, 1), 2), np.float32)
        resp = np.zeros(max(needed, 1), np.float32)
        n = self._fn("detect_features")(C.byref(params), _u8p(image), C.c_int32(image.shape[0]), C.c_int32(image.shape[1]),
                                        _f32p(existing) if len(existing) else None, C.c_int32(len(existing)), C.c_int32(needed), _f32p(uv), _f32p(resp))
        return n >= 0, uv[:max(n, 0)].copy(), resp[:max(n, 0)].copy()

    def brief_pattern(self, n_bits, half_patch, seed=0):
        pattern = np.zeros((n_bits, 4), np.int8)
        f = getattr(self.lib, self.prefix + "brief_pattern")
        f.restype = None
        f(C.c_int32(n_bits), C.c_int32(half_patch), C.c_uint32(seed), pattern.ctypes.data_as(C.c_void_p))
        return pattern

    def describe_brief(self, image, uv, pattern, half_patch):
        """Returns (ok, desc uint32 [n][n_bits / 32], valid uint8 [n])."""
        image = np.ascontiguousarray(image, dtype=np.uint8)
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        pattern = np.ascontiguousarray(pattern, np.int8).reshape(-1, 4)
        n_bits = len(pattern)
        desc = np.zeros((max(len(uv), 1), max(n_bits // 32, 1)), np.uint32)
        valid = np.zeros(max(len(uv), 1), np.uint8)
        ok = self._fn("describe_brief")(_u8p(image), C.c_int32(image.shape[0]), C.c_int32(image.shape[1]), _f32p(uv), C.c_int32(len(uv)),
                                        pattern.ctypes.data_as(C.c_void_p), C.c_int32(n_bits), C.c_int32(half_patch), desc.ctypes.data_as(C.c_void_p), _u8p(valid))
        return ok == 1, desc[:len(uv)].copy(), valid[:len(uv)].copy()


def have_ref():
    return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_ROOT)
