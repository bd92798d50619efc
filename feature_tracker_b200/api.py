"""Host-side mirror of the reference's tracker / matcher interface on top of the C ABI.

Names, option fields, argument meaning and error behaviour follow the reference's C++ classes
(src/optical_flow_tracker/optical_flow.h, the three subclasses, src/descriptor_matcher/descriptor_matcher.h); the C++
facade with the same names lives in include/feature_tracker_b200/.  In/out std::vector arguments become "pass the
current content (or None for an empty vector), get the new content back".
"""
import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum

import numpy as np

from . import _capi


class TrackStatus(IntEnum):
    """src/feature_tracker.h:8-14"""
    kNotTracked = 0
    kTracked = 1
    kLargeResidual = 2
    kOutside = 3
    kNumericError = 4


class OpticalFlowMethod(IntEnum):
    """src/optical_flow_tracker/optical_flow.h:12-18"""
    kInverse = 0
    kDirect = 1
    kFast = 2
    kSse = 3
    kNeon = 4


@dataclass
class OpticalFlowOptions:
    """src/optical_flow_tracker/optical_flow.h:20-28 (same defaults)."""
    kMaxTrackPointsNumber: int = 500
    kMaxIteration: int = 15
    kMaxToleranceLargeStep: int = 3
    kPatchRowHalfSize: int = 6
    kPatchColHalfSize: int = 6
    kMaxConvergeStep: float = 4e-2
    kMethod: OpticalFlowMethod = OpticalFlowMethod.kFast


class FtkError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _capi.load_library()
    return _lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One GPU + one stream (ftk_context).  Not re-entrant, like the reference's tracker objects."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().ftk_create(int(device), C.byref(self._h))
        if rc != _capi.OK:
            self._h = None
            raise FtkError(f"ftk_create(device={device}) failed with code {rc}: a B200 (sm_100) GPU is required; there is no CPU fallback")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            lib().ftk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, soft=()):
        """Raises on hard errors; returns False for the codes the reference maps to `return false`."""
        if rc == _capi.OK:
            return True
        if rc in soft:
            return False
        raise FtkError(f"ftk error {rc}: {lib().ftk_last_error(self._h).decode()}")

    def synchronize(self):
        self.check(lib().ftk_synchronize(self._h))

    @property
    def stream(self):
        return lib().ftk_stream(self._h)

    @property
    def kernel_launches(self):
        return int(lib().ftk_kernel_launches(self._h))


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class ImagePyramidBatch:
    """Device-resident batch of same-sized image pyramids (ImagePyramid of the reference, batched)."""

    def __init__(self, ctx, rows, cols, levels, n_images=1):
        self.ctx = ctx
        self.rows, self.cols, self.n_levels, self.n_images = int(rows), int(cols), int(levels), int(n_images)
        self._h = C.c_void_p()
        ctx.check(lib().ftk_pyramid_create(ctx._h, self.rows, self.cols, self.n_levels, self.n_images, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib().ftk_pyramid_destroy(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level(self):
        return self.n_levels

    def SetRawImages(self, images, first=0):
        """images: uint8 array [n, rows, cols] (host)."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        if images.ndim == 2:
            images = images[None]
        assert images.shape[1:] == (self.rows, self.cols), images.shape
        self.ctx.check(lib().ftk_pyramid_set_images(self.ctx._h, self._h, int(first), images.shape[0], _ptr(images), 0))
        self.ctx.synchronize()  # the source array may be freed by the caller

    def set_images_ptr(self, ptr, count, first=0, device=False):
        """Raw pointer variant (pinned host memory or device memory) used by the benchmark; asynchronous."""
        flags = _capi.FLAG_DEVICE_POINTERS if device else 0
        self.ctx.check(lib().ftk_pyramid_set_images(self.ctx._h, self._h, int(first), int(count), C.c_void_p(ptr), flags))

    def CreateImagePyramid(self, first=0, count=None):
        count = self.n_images - first if count is None else count
        self.ctx.check(lib().ftk_pyramid_build(self.ctx._h, self._h, int(first), int(count)))

    def SetLevel(self, image, level, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.shape == (self.rows >> level, self.cols >> level), data.shape
        self.ctx.check(lib().ftk_pyramid_set_level(self.ctx._h, self._h, int(image), int(level), _ptr(data)))

    def GetLevel(self, image, level):
        out = np.zeros((self.rows >> level, self.cols >> level), np.uint8)
        if out.size:
            self.ctx.check(lib().ftk_pyramid_get_level(self.ctx._h, self._h, int(image), int(level), _ptr(out)))
        return out


class OpticalFlow:
    """Base of the three trackers (src/optical_flow_tracker/optical_flow.h:30-104)."""

    _variant = None
    _name = "None"

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._options = OpticalFlowOptions()
        self._predict = np.eye(2, dtype=np.float32)
        self._consider_patch_luminance = False
        # > 0: forward-backward consistency pass (not in the reference; see ftk_klt_params::forward_backward_max_error)
        self.forward_backward_max_error = 0.0

    def OpticalFlowMethodName(self):
        return self._name

    def options(self):
        return self._options

    def _params(self):
        o = self._options
        p = _capi.KltParams()
        p.variant = self._variant
        p.method = int(o.kMethod)
        p.max_track_points = int(o.kMaxTrackPointsNumber)
        p.max_iteration = int(o.kMaxIteration)
        p.max_tolerance_large_step = int(o.kMaxToleranceLargeStep)
        p.patch_row_half = int(o.kPatchRowHalfSize)
        p.patch_col_half = int(o.kPatchColHalfSize)
        p.max_converge_step = float(o.kMaxConvergeStep)
        pr = np.asarray(self._predict, dtype=np.float32).reshape(4)
        for i in range(4):
            p.predict[i] = float(pr[i])
        p.consider_patch_luminance = 1 if self._consider_patch_luminance else 0
        p.forward_backward_max_error = float(self.forward_backward_max_error)
        return p

    def TrackFeatures(self, ref_pyramid, cur_pyramid, ref_pixel_uv, cur_pixel_uv=None, status=None, single_level=False, ref_image=0,
                      cur_image=0):
        """optical_flow.cpp:6-47.  ref_pixel_uv [n,2] float32; cur_pixel_uv / status = the vectors' content on entry (None or a
        wrong length behaves like the reference: no prediction / all kNotTracked).  `single_level=True` is the GrayImage overload.
        Returns (ok, cur_pixel_uv, status)."""
        ref_uv = np.ascontiguousarray(ref_pixel_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        offsets = np.array([0, n], dtype=np.int32)
        ok, cur, st = self.TrackFeaturesBatch(ref_pyramid, cur_pyramid, offsets, ref_uv, cur_pixel_uv, status, single_level,
                                              np.array([ref_image], np.int32), np.array([cur_image], np.int32))
        return ok, cur, st

    def TrackFeaturesBatch(self, ref_pyramid, cur_pyramid, feat_offsets, ref_pixel_uv, cur_pixel_uv=None, status=None, single_level=False,
                           ref_image=None, cur_image=None):
        """Many independent frame pairs in one call (pair p owns features feat_offsets[p]:feat_offsets[p+1])."""
        ref_uv = np.ascontiguousarray(ref_pixel_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        feat_offsets = np.ascontiguousarray(feat_offsets, dtype=np.int32)
        n_pairs = feat_offsets.shape[0] - 1
        flags = _capi.FLAG_SINGLE_LEVEL if single_level else 0
        cur = np.zeros((max(n, 1), 2), np.float32)
        if cur_pixel_uv is not None and np.asarray(cur_pixel_uv).reshape(-1, 2).shape[0] == n and n > 0:
            cur[:n] = np.asarray(cur_pixel_uv, dtype=np.float32).reshape(-1, 2)
        else:
            flags |= _capi.FLAG_NO_PREDICTION
        st = np.zeros(max(n, 1), np.uint8)
        if status is not None and np.asarray(status).reshape(-1).shape[0] == n and n > 0:
            st[:n] = np.asarray(status, dtype=np.uint8).reshape(-1)
        else:
            flags |= _capi.FLAG_NO_STATUS
        ri = None if ref_image is None else np.ascontiguousarray(ref_image, dtype=np.int32)
        ci = None if cur_image is None else np.ascontiguousarray(cur_image, dtype=np.int32)
        p = self._params()
        rc = lib().ftk_klt_track(self.ctx._h, C.byref(p), ref_pyramid._h, cur_pyramid._h, n_pairs, _ptr(ri), _ptr(ci), _ptr(feat_offsets),
                                 _ptr(ref_uv), _ptr(cur), _ptr(st), flags)
        ok = self.ctx.check(rc, soft=(_capi.ERR_EMPTY_INPUT, _capi.ERR_LEVEL_MISMATCH))
        return ok, cur[:n], st[:n]


    def TrackImagePairs(self, levels, ref_images, cur_images, feat_offsets, ref_pixel_uv, cur_pixel_uv=None, status=None):
        """The reference demo's timed region (CreateImagePyramid x2 + TrackFeatures, test/test_optical_flow.cpp:69-73) for a batch of
        frame pairs given as host images [n_pairs, rows, cols]; H2D copies overlap compute (ftk_track_image_pairs)."""
        ref_images = np.ascontiguousarray(ref_images, dtype=np.uint8)
        cur_images = np.ascontiguousarray(cur_images, dtype=np.uint8)
        n_pairs, rows, cols = ref_images.shape
        ref_uv = np.ascontiguousarray(ref_pixel_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        feat_offsets = np.ascontiguousarray(feat_offsets, dtype=np.int32)
        flags = 0
        cur = np.zeros((max(n, 1), 2), np.float32)
        if cur_pixel_uv is not None and np.asarray(cur_pixel_uv).reshape(-1, 2).shape[0] == n and n > 0:
            cur[:n] = np.asarray(cur_pixel_uv, dtype=np.float32).reshape(-1, 2)
        else:
            flags |= _capi.FLAG_NO_PREDICTION
        st = np.zeros(max(n, 1), np.uint8)
        if status is not None and np.asarray(status).reshape(-1).shape[0] == n and n > 0:
            st[:n] = np.asarray(status, dtype=np.uint8).reshape(-1)
        else:
            flags |= _capi.FLAG_NO_STATUS
        p = self._params()
        rc = lib().ftk_track_image_pairs(self.ctx._h, C.byref(p), rows, cols, int(levels), n_pairs, _ptr(ref_images), _ptr(cur_images),
                                         _ptr(feat_offsets), _ptr(ref_uv), _ptr(cur), _ptr(st), flags)
        ok = self.ctx.check(rc, soft=(_capi.ERR_EMPTY_INPUT,))
        return ok, cur[:n], st[:n]

    @staticmethod
    def TrackImagePairsMulti(trackers, levels, ref_images, cur_images, feat_offsets, ref_pixel_uv):
        """Several trackers (any variants / methods / patch sizes, one context) over the same batch of host frame pairs in one pipelined
        call: the images are uploaded and their pyramids built once (ftk_track_image_pairs_multi).  No prediction / status input.
        Returns (ok, cur_pixel_uv [n_trackers, n, 2], status [n_trackers, n])."""
        ctx = trackers[0].ctx
        ref_images = np.ascontiguousarray(ref_images, dtype=np.uint8)
        cur_images = np.ascontiguousarray(cur_images, dtype=np.uint8)
        n_pairs, rows, cols = ref_images.shape
        ref_uv = np.ascontiguousarray(ref_pixel_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        feat_offsets = np.ascontiguousarray(feat_offsets, dtype=np.int32)
        params = (_capi.KltParams * len(trackers))(*[t._params() for t in trackers])
        cur = np.zeros((len(trackers), max(n, 1), 2), np.float32)
        st = np.zeros((len(trackers), max(n, 1)), np.uint8)
        rc = lib().ftk_track_image_pairs_multi(ctx._h, params, len(trackers), rows, cols, int(levels), n_pairs, _ptr(ref_images), _ptr(cur_images),
                                               _ptr(feat_offsets), _ptr(ref_uv), _ptr(cur), _ptr(st), _capi.FLAG_NO_PREDICTION | _capi.FLAG_NO_STATUS)
        ok = ctx.check(rc, soft=(_capi.ERR_EMPTY_INPUT,))
        return ok, cur[:, :n], st[:, :n]

    def TrackImageSequence(self, levels, frames, feat_offsets, ref_pixel_uv, cur_pixel_uv=None, status=None):
        """Temporal form of TrackImagePairs: host frames [n_frames, rows, cols]; pair k tracks features
        feat_offsets[k]:feat_offsets[k+1] from frame k to frame k+1; every frame is uploaded and its pyramid built once
        (ftk_track_image_sequence)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n_frames, rows, cols = frames.shape
        ref_uv = np.ascontiguousarray(ref_pixel_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        feat_offsets = np.ascontiguousarray(feat_offsets, dtype=np.int32)
        flags = 0
        cur = np.zeros((max(n, 1), 2), np.float32)
        if cur_pixel_uv is not None and np.asarray(cur_pixel_uv).reshape(-1, 2).shape[0] == n and n > 0:
            cur[:n] = np.asarray(cur_pixel_uv, dtype=np.float32).reshape(-1, 2)
        else:
            flags |= _capi.FLAG_NO_PREDICTION
        st = np.zeros(max(n, 1), np.uint8)
        if status is not None and np.asarray(status).reshape(-1).shape[0] == n and n > 0:
            st[:n] = np.asarray(status, dtype=np.uint8).reshape(-1)
        else:
            flags |= _capi.FLAG_NO_STATUS
        p = self._params()
        rc = lib().ftk_track_image_sequence(self.ctx._h, C.byref(p), rows, cols, int(levels), n_frames, _ptr(frames), _ptr(feat_offsets),
                                            _ptr(ref_uv), _ptr(cur), _ptr(st), flags)
        ok = self.ctx.check(rc, soft=(_capi.ERR_EMPTY_INPUT,))
        return ok, cur[:n], st[:n]


class OpticalFlowBasicKlt(OpticalFlow):
    """basic_klt/optical_flow_basic_klt.h:9-41"""
    _variant = 0
    _name = "Basic-Klt"


class OpticalFlowAffineKlt(OpticalFlow):
    """affine_klt/optical_flow_affine_klt.h:9-52"""
    _variant = 1
    _name = "Affine-Klt"

    def predict_affine(self):
        return self._predict


class OpticalFlowLssdKlt(OpticalFlow):
    """lssd_klt/optical_flow_lssd_klt.h:9-56"""
    _variant = 2
    _name = "Lssd-Klt"

    def predict_R_cr(self):
        return self._predict

    @property
    def consider_patch_luminance(self):
        return self._consider_patch_luminance

    @consider_patch_luminance.setter
    def consider_patch_luminance(self, v):
        self._consider_patch_luminance = bool(v)


@dataclass
class MatcherOptions:
    """descriptor_matcher.h:16-20 (same defaults: nothing matches until kMaxValidDescriptorDistance is raised)."""
    kMaxValidPredictRowDistance: int = 40
    kMaxValidPredictColDistance: int = 40
    kMaxValidDescriptorDistance: float = 0.0


def pack_brief(bits):
    """[n, len] array of 0/1 -> [n, ceil(len/32)] uint32, element k = bit k%32 of word k//32."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    n, ln = bits.shape
    words = (ln + 31) // 32
    padded = np.zeros((n, words * 32), np.uint8)
    padded[:, :ln] = bits != 0
    return np.packbits(padded.reshape(n, words, 32), axis=2, bitorder="little").view(np.uint32).reshape(n, words)


class DescriptorMatcher:
    """descriptor_matcher.h:12-52.  Subclasses fix the distance instead of overriding a per-pair virtual."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._options = MatcherOptions()

    def options(self):
        return self._options

    def _idx(self, index_pairs_in_cur, n_ref):
        buf = np.full(max(n_ref, 1), -1, np.int32)
        flags = 0
        if index_pairs_in_cur is not None and np.asarray(index_pairs_in_cur).reshape(-1).shape[0] == n_ref and n_ref > 0:
            buf[:n_ref] = np.asarray(index_pairs_in_cur, dtype=np.int32).reshape(-1)
        else:
            flags = _capi.FLAG_NO_INDEX_INPUT
        return buf, flags

    def _force(self, ref, cur, idx, flags):
        raise NotImplementedError

    def _nearby(self, ref, cur, pred, pos, idx, flags):
        raise NotImplementedError

    def ForceMatch(self, descriptors_ref, descriptors_cur, index_pairs_in_cur=None):
        """descriptor_matcher.h:55-79.  Returns (ok, index_pairs_in_cur)."""
        ref, cur = self._prep(descriptors_ref), self._prep(descriptors_cur)
        n_ref = ref.shape[0]
        idx, flags = self._idx(index_pairs_in_cur, n_ref)
        ok = self.ctx.check(self._force(ref, cur, idx, flags), soft=(_capi.ERR_EMPTY_INPUT,))
        return ok, idx[:n_ref]

    def NearbyMatch(self, descriptors_ref, descriptors_cur, pixel_uv_pred_in_cur, pixel_uv_cur, index_pairs_in_cur=None):
        """descriptor_matcher.h:90-124.  Returns (ok, index_pairs_in_cur)."""
        ref, cur = self._prep(descriptors_ref), self._prep(descriptors_cur)
        n_ref, n_cur = ref.shape[0], cur.shape[0]
        pred = np.ascontiguousarray(pixel_uv_pred_in_cur, dtype=np.float32).reshape(-1, 2)
        pos = np.ascontiguousarray(pixel_uv_cur, dtype=np.float32).reshape(-1, 2)
        idx, flags = self._idx(index_pairs_in_cur, n_ref)
        if n_cur == 0 or pred.shape[0] != n_ref or pos.shape[0] != n_cur:  # :94-96
            return False, idx[:n_ref]
        ok = self.ctx.check(self._nearby(ref, cur, pred, pos, idx, flags), soft=(_capi.ERR_EMPTY_INPUT, _capi.ERR_SIZE_MISMATCH))
        return ok, idx[:n_ref]

    def CrossCheckForceMatch(self, descriptors_ref, descriptors_cur):
        """Not in the reference: ForceMatch in both directions, keeping ref i -> cur j only when cur j -> ref i too
        (ftk_match_cross_check).  Returns (ok, index_pairs_in_cur)."""
        ok_f, fwd = self.ForceMatch(descriptors_ref, descriptors_cur)
        ok_b, bwd = self.ForceMatch(descriptors_cur, descriptors_ref)
        if not (ok_f and ok_b):
            return False, fwd
        fwd = np.ascontiguousarray(fwd, dtype=np.int32)
        bwd = np.ascontiguousarray(bwd, dtype=np.int32)
        self.ctx.check(lib().ftk_match_cross_check(self.ctx._h, _ptr(fwd), fwd.shape[0], _ptr(bwd), bwd.shape[0], 0))
        return True, fwd

    def FillMatchedPixelByPairIndices(self, index_pairs_in_cur, pixel_uv_cur, status=None):
        """descriptor_matcher.h:135-157.  Returns (matched_pixel_uv_cur, status)."""
        idx = np.ascontiguousarray(index_pairs_in_cur, dtype=np.int32)
        pos = np.ascontiguousarray(pixel_uv_cur, dtype=np.float32).reshape(-1, 2)
        n_ref = idx.shape[0]
        matched = np.zeros((max(n_ref, 1), 2), np.float32)
        st = np.zeros(max(n_ref, 1), np.uint8)
        valid = 0
        if status is not None and np.asarray(status).reshape(-1).shape[0] == n_ref:
            st[:n_ref] = np.asarray(status, dtype=np.uint8)
            valid = 1
        rc = lib().ftk_fill_matched(_ptr(idx), n_ref, _ptr(pos), pos.shape[0], _ptr(matched), _ptr(st), valid)
        self.ctx.check(rc)
        return matched[:n_ref], st[:n_ref]

    def ForceMatchUv(self, descriptors_ref, descriptors_cur, pixel_uv_cur, status=None):
        """descriptor_matcher.h:81-88: returns (ok, matched_pixel_uv_cur, status)."""
        ok, idx = self.ForceMatch(descriptors_ref, descriptors_cur, None)
        if not ok:
            return False, None, status
        matched, st = self.FillMatchedPixelByPairIndices(idx, pixel_uv_cur, status)
        return True, matched, st

    def NearbyMatchUv(self, descriptors_ref, descriptors_cur, pixel_uv_pred_in_cur, pixel_uv_cur, status=None):
        """descriptor_matcher.h:126-133: returns (ok, matched_pixel_uv_cur, status)."""
        ok, idx = self.NearbyMatch(descriptors_ref, descriptors_cur, pixel_uv_pred_in_cur, pixel_uv_cur, None)
        if not ok:
            return False, None, status
        matched, st = self.FillMatchedPixelByPairIndices(idx, pixel_uv_cur, status)
        return True, matched, st


class BriefMatcher(DescriptorMatcher):
    """DescriptorMatcher<BriefType> with the Hamming ComputeDistance of test/test_descriptor_matcher_brief.cpp:33-45.
    Descriptors are [n, words] uint32 (see pack_brief)."""

    def _prep(self, d):
        d = np.ascontiguousarray(d, dtype=np.uint32)
        return d if d.ndim == 2 else d.reshape(0, 8)

    def _force(self, ref, cur, idx, flags):
        o = self._options
        words = cur.shape[1]
        return lib().ftk_match_hamming_force(self.ctx._h, _ptr(ref), ref.shape[0], _ptr(cur), cur.shape[0], words,
                                             float(o.kMaxValidDescriptorDistance), _ptr(idx), flags)

    def _nearby(self, ref, cur, pred, pos, idx, flags):
        o = self._options
        words = cur.shape[1]
        return lib().ftk_match_hamming_nearby(self.ctx._h, _ptr(ref), ref.shape[0], _ptr(cur), cur.shape[0], words, _ptr(pred), _ptr(pos),
                                              int(o.kMaxValidPredictRowDistance), int(o.kMaxValidPredictColDistance),
                                              float(o.kMaxValidDescriptorDistance), _ptr(idx), flags)

    def MatchPairs(self, descriptors_ref, ref_offsets, descriptors_cur, cur_offsets, pixel_uv_pred_in_cur=None, pixel_uv_cur=None, index_pairs_in_cur=None):
        """Many frame pairs in one call (ftk_match_hamming_pairs): pair p matches ref rows ref_offsets[p]:ref_offsets[p+1] against cur rows
        cur_offsets[p]:cur_offsets[p+1]; ForceMatch per pair, or NearbyMatch per pair when the positions are given.  Returns (ok,
        index_pairs_in_cur) with indices relative to each pair's cur block."""
        ref, cur = self._prep(descriptors_ref), self._prep(descriptors_cur)
        ro = np.ascontiguousarray(ref_offsets, dtype=np.int32)
        co = np.ascontiguousarray(cur_offsets, dtype=np.int32)
        n_ref = ref.shape[0]
        assert ro[-1] == n_ref and co[-1] == cur.shape[0] and len(ro) == len(co)
        idx, flags = self._idx(index_pairs_in_cur, n_ref)
        pred = pos = None
        if pixel_uv_pred_in_cur is not None:
            pred = np.ascontiguousarray(pixel_uv_pred_in_cur, dtype=np.float32).reshape(-1, 2)
            pos = np.ascontiguousarray(pixel_uv_cur, dtype=np.float32).reshape(-1, 2)
            assert pred.shape[0] == n_ref and pos.shape[0] == cur.shape[0]
        o = self._options
        rc = lib().ftk_match_hamming_pairs(self.ctx._h, _ptr(ref), _ptr(cur), ref.shape[1], len(ro) - 1, _ptr(ro), _ptr(co), _ptr(pred), _ptr(pos),
                                           int(o.kMaxValidPredictRowDistance), int(o.kMaxValidPredictColDistance), float(o.kMaxValidDescriptorDistance),
                                           _ptr(idx), flags)
        return self.ctx.check(rc), idx[:n_ref]


class CosineMatcher(DescriptorMatcher):
    """DescriptorMatcher<SuperpointDescriptorType / DiskDescriptorType> with the 0.5 - 0.5*cos ComputeDistance of
    test/test_descriptor_matcher_superpoint.cpp:32-34.  Descriptors are [n, dim] float32."""

    def _prep(self, d):
        d = np.ascontiguousarray(d, dtype=np.float32)
        return d if d.ndim == 2 else d.reshape(0, 256)

    def _force(self, ref, cur, idx, flags):
        o = self._options
        return lib().ftk_match_cosine_force(self.ctx._h, _ptr(ref), ref.shape[0], _ptr(cur), cur.shape[0], cur.shape[1],
                                            float(o.kMaxValidDescriptorDistance), _ptr(idx), flags)

    def _nearby(self, ref, cur, pred, pos, idx, flags):
        o = self._options
        return lib().ftk_match_cosine_nearby(self.ctx._h, _ptr(ref), ref.shape[0], _ptr(cur), cur.shape[0], cur.shape[1], _ptr(pred), _ptr(pos),
                                             int(o.kMaxValidPredictRowDistance), int(o.kMaxValidPredictColDistance),
                                             float(o.kMaxValidDescriptorDistance), _ptr(idx), flags)

    def MatchPairs(self, descriptors_ref, ref_offsets, descriptors_cur, cur_offsets, pixel_uv_pred_in_cur=None, pixel_uv_cur=None, index_pairs_in_cur=None):
        """Many frame pairs in one call (ftk_match_cosine_pairs), the float-descriptor twin of BriefMatcher.MatchPairs."""
        ref, cur = self._prep(descriptors_ref), self._prep(descriptors_cur)
        ro = np.ascontiguousarray(ref_offsets, dtype=np.int32)
        co = np.ascontiguousarray(cur_offsets, dtype=np.int32)
        n_ref = ref.shape[0]
        assert ro[-1] == n_ref and co[-1] == cur.shape[0] and len(ro) == len(co) and ref.shape[1] == cur.shape[1]
        idx, flags = self._idx(index_pairs_in_cur, n_ref)
        pred = pos = None
        if pixel_uv_pred_in_cur is not None:
            pred = np.ascontiguousarray(pixel_uv_pred_in_cur, dtype=np.float32).reshape(-1, 2)
            pos = np.ascontiguousarray(pixel_uv_cur, dtype=np.float32).reshape(-1, 2)
            assert pred.shape[0] == n_ref and pos.shape[0] == cur.shape[0]
        o = self._options
        rc = lib().ftk_match_cosine_pairs(self.ctx._h, _ptr(ref), _ptr(cur), cur.shape[1], len(ro) - 1, _ptr(ro), _ptr(co), _ptr(pred), _ptr(pos),
                                          int(o.kMaxValidPredictRowDistance), int(o.kMaxValidPredictColDistance), float(o.kMaxValidDescriptorDistance),
                                          _ptr(idx), flags)
        return self.ctx.check(rc), idx[:n_ref]


    def last_exact_scan_items(self):
        """Diagnostics: rows x splits the last ForceMatch re-scanned exactly (0 = the tensor-core pass decided everything)."""
        return int(lib().ftk_last_cosine_exact_scan_items(self.ctx._h))


SuperpointMatcher = CosineMatcher
DiskMatcher = CosineMatcher


def nn_match_pairs_host(matches, n_ref, n_cur):
    """The fused-model branch of NNFeatureMatcher::Match (nn_feature_matcher.cpp:160-178): matches [n, 2] int64 (idx_ref, idx_cur)
    -> index_pairs_in_cur [n_ref]; pairs with an index outside either set are ignored, later pairs overwrite earlier ones.  Host
    code, as in the reference: the network already produced at most kMaxNumberOfMatches pairs."""
    idx = np.full(n_ref, -1, np.int32)
    for i_ref, i_cur in np.asarray(matches, np.int64).reshape(-1, 2):
        # :169 bounds idx_ref by matched_pixel_uv_cur.size(), which is pixel_uv_cur.size() (:158)
        if 0 <= i_ref < min(n_ref, n_cur) and 0 <= i_cur < n_cur:
            idx[i_ref] = i_cur
    return idx


class NNFeatureMatcherOptions:
    """nn_feature_matcher.h:23-27 (the fields the score-matrix post-processing uses)."""

    def __init__(self):
        self.kMinValidMatchScore = -3.0


class NNFeatureMatcher:
    """The score-matrix post-processing of NNFeatureMatcher::Match (src/nn_feature_matcher/nn_feature_matcher.cpp:154-216):
    mutual row / column arg-max with a minimum score.  The LightGlue network itself (ONNX Runtime in the reference) is out of
    scope; its output matrix is the input here."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._options = NNFeatureMatcherOptions()

    def options(self):
        return self._options

    def MatchScores(self, scores):
        """scores [n_ref, n_cur] float32 -> index_pairs_in_cur [n_ref] (-1 = no mutual match)."""
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        n_ref, n_cur = scores.shape
        idx = np.full(max(n_ref, 1), -1, np.int32)
        rc = lib().ftk_match_mutual_scores(self.ctx._h, _ptr(scores), n_ref, n_cur, float(self._options.kMinValidMatchScore), _ptr(idx), 0)
        ok = self.ctx.check(rc, soft=(_capi.ERR_EMPTY_INPUT,))
        return ok, idx[:n_ref]

    def Match(self, scores, pixel_uv_ref, pixel_uv_cur):
        """nn_feature_matcher.cpp:154-216 after InferenceSession: returns (ok, matched_pixel_uv_cur [n_ref, 2], status [n_ref]) with
        status kLargeResidual for unmatched rows (:157) and kTracked for matched ones (:213).  (The reference copies
        pixel_uv_cur into matched_pixel_uv_cur first, :158; unmatched rows here keep their own reference position.)"""
        ok, idx = self.MatchScores(scores)
        ref = np.ascontiguousarray(pixel_uv_ref, dtype=np.float32).reshape(-1, 2)
        cur = np.ascontiguousarray(pixel_uv_cur, dtype=np.float32).reshape(-1, 2)
        matched = ref.copy()
        status = np.full(ref.shape[0], 2, np.uint8)
        if ok:
            good = idx >= 0
            matched[good] = cur[idx[good]]
            status[good] = 1
        return ok, matched, status


class DirectMethodMethod:
    """direct_method_tracker.h:14-18"""
    kInverse = 0
    kDirect = 1
    kFast = 2


class DirectMethodOptions:
    """direct_method_tracker.h:20-28"""

    def __init__(self):
        self.kMaxTrackPointsNumber = 500
        self.kMaxIteration = 15
        self.kPatchRowHalfSize = 6
        self.kPatchColHalfSize = 6
        self.kMaxConvergeStep = 1e-6
        self.kMaxConvergeResidual = 2.0
        self.kMethod = DirectMethodMethod.kDirect


class DirectMethod:
    """DirectMethod (src/direct_method_tracker/direct_method_tracker.h:30-77): one 6-DoF pose of the current frame relative to the
    reference frame from all features of a frame pair.  Quaternions are (w, x, y, z)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._options = DirectMethodOptions()

    def options(self):
        return self._options

    def _params(self):
        o = self._options
        p = _capi.DirectParams()
        p.max_track_points = int(o.kMaxTrackPointsNumber)
        p.max_iteration = int(o.kMaxIteration)
        p.patch_row_half = int(o.kPatchRowHalfSize)
        p.patch_col_half = int(o.kPatchColHalfSize)
        p.max_converge_step = float(o.kMaxConvergeStep)
        p.max_converge_residual = float(o.kMaxConvergeResidual)
        p.method = int(o.kMethod)
        return p

    def TrackFeaturesBatch(self, ref_pyramid, cur_pyramid, feat_offsets, K, p_c_in_ref, ref_pixel_uv, q_rc, p_rc, cur_pixel_uv=None, status=None,
                           ref_image=None, cur_image=None):
        """n_pairs independent problems in one call (ftk_direct_method_track).  K [n_pairs, 4], q_rc [n_pairs, 4], p_rc [n_pairs, 3].
        Returns (ok, cur_pixel_uv, q_rc, p_rc, status)."""
        ref_uv = np.ascontiguousarray(ref_pixel_uv, dtype=np.float32).reshape(-1, 2)
        n = ref_uv.shape[0]
        feat_offsets = np.ascontiguousarray(feat_offsets, dtype=np.int32)
        n_pairs = feat_offsets.shape[0] - 1
        pts = np.ascontiguousarray(p_c_in_ref, dtype=np.float32).reshape(-1, 3)
        Kc = np.ascontiguousarray(K, dtype=np.float32).reshape(n_pairs, 4)
        q = np.ascontiguousarray(q_rc, dtype=np.float32).reshape(n_pairs, 4).copy()
        p = np.ascontiguousarray(p_rc, dtype=np.float32).reshape(n_pairs, 3).copy()
        flags = 0
        cur = np.zeros((max(n, 1), 2), np.float32)
        if cur_pixel_uv is not None and np.asarray(cur_pixel_uv).reshape(-1, 2).shape[0] == n and n > 0:
            cur[:n] = np.asarray(cur_pixel_uv, dtype=np.float32).reshape(-1, 2)
        else:
            flags |= _capi.FLAG_NO_PREDICTION
        st = np.zeros(max(n, 1), np.uint8)
        if status is not None and np.asarray(status).reshape(-1).shape[0] == n and n > 0:
            st[:n] = np.asarray(status, dtype=np.uint8).reshape(-1)
        else:
            flags |= _capi.FLAG_NO_STATUS
        ri = None if ref_image is None else np.ascontiguousarray(ref_image, dtype=np.int32)
        ci = None if cur_image is None else np.ascontiguousarray(cur_image, dtype=np.int32)
        prm = self._params()
        rc = lib().ftk_direct_method_track(self.ctx._h, C.byref(prm), ref_pyramid._h, cur_pyramid._h, n_pairs, _ptr(ri), _ptr(ci), _ptr(feat_offsets),
                                           _ptr(Kc), _ptr(pts), _ptr(ref_uv), _ptr(cur), _ptr(q), _ptr(p), _ptr(st), flags)
        ok = self.ctx.check(rc, soft=(_capi.ERR_EMPTY_INPUT, _capi.ERR_LEVEL_MISMATCH))
        return ok, cur[:n], q, p, st[:n]

    def TrackFeatures(self, ref_pyramid, cur_pyramid, K, p_c_in_ref, ref_pixel_uv, q_rc, p_rc, cur_pixel_uv=None, status=None, ref_image=0,
                      cur_image=0):
        """direct_method_tracker.cpp:41-95 (camera-frame overload) for one frame pair."""
        ref_uv = np.ascontiguousarray(ref_pixel_uv, dtype=np.float32).reshape(-1, 2)
        ok, cur, q, p, st = self.TrackFeaturesBatch(ref_pyramid, cur_pyramid, np.array([0, ref_uv.shape[0]], np.int32), np.asarray(K).reshape(1, 4),
                                                    p_c_in_ref, ref_uv, np.asarray(q_rc).reshape(1, 4), np.asarray(p_rc).reshape(1, 3), cur_pixel_uv,
                                                    status, np.array([ref_image], np.int32), np.array([cur_image], np.int32))
        return ok, cur, q[0], p[0], st

    def TrackFeaturesWorld(self, ref_pyramid, cur_pyramid, K, ref_q_wc, ref_p_wc, p_w, ref_pixel_uv, cur_q_wc, cur_p_wc, cur_pixel_uv=None,
                           status=None, ref_image=0, cur_image=0):
        """direct_method_tracker.cpp:8-39 (world-frame overload): lifts the points and the predicted pose into the reference camera
        frame, runs the camera-frame overload and maps the pose back.  Returns (ok, cur_pixel_uv, cur_q_wc, cur_p_wc, status)."""
        from . import quat
        ref_q_wc = np.asarray(ref_q_wc, np.float32).reshape(4)
        ref_p_wc = np.asarray(ref_p_wc, np.float32).reshape(3)
        ref_q_cw = quat.inverse(ref_q_wc)
        p_c_in_ref = quat.rotate(ref_q_cw, np.asarray(p_w, np.float32).reshape(-1, 3) - ref_p_wc)
        q_rc = quat.multiply(ref_q_cw, cur_q_wc)
        p_rc = quat.rotate(ref_q_cw, np.asarray(cur_p_wc, np.float32).reshape(3) - ref_p_wc)
        ok, cur, q_rc, p_rc, st = self.TrackFeatures(ref_pyramid, cur_pyramid, K, p_c_in_ref, ref_pixel_uv, q_rc, p_rc, cur_pixel_uv, status, ref_image,
                                                     cur_image)
        if not ok:  # RETURN_FALSE_IF_FALSE: the caller's pose stays untouched
            return False, cur, np.asarray(cur_q_wc, np.float32).reshape(4), np.asarray(cur_p_wc, np.float32).reshape(3), st
        return True, cur, quat.multiply(ref_q_wc, q_rc), quat.rotate(ref_q_wc, p_rc) + ref_p_wc, st


class DenseOpticalFlowOptions:
    """dense_optical_flow.h:15-20"""

    def __init__(self):
        self.kMaxIteration = 10
        self.kHalfPatchSize = 2
        self.kMaxConvergeStep = 1e-6
        self.kMaxDeltaFlowStep = 1.0


class DenseOpticalFlow:
    """DenseOpticalFlow (src/dense_optical_flow_tracker/dense_optical_flow.h:12-65), Gunnar Farneback's method."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._options = DenseOpticalFlowOptions()

    def OpticalFlowMethodName(self):
        return "Gunnar Farneback"

    def options(self):
        return self._options

    def _params(self):
        o = self._options
        p = _capi.DenseFlowParams()
        p.max_iteration, p.half_patch_size = int(o.kMaxIteration), int(o.kHalfPatchSize)
        p.max_converge_step, p.max_delta_flow_step = float(o.kMaxConvergeStep), float(o.kMaxDeltaFlowStep)
        return p

    def Track(self, ref_pyramid, cur_pyramid, flow_rc=None, single_level=False, ref_image=0, cur_image=0):
        """dense_optical_flow.cpp:35-85 (pyramids) or, with single_level=True, :7-33 (the GrayImage overload on level 0, where flow_rc
        = (flow_row, flow_col) of the image's size is the initial flow).  Returns (ok, flow_row, flow_col)."""
        rows, cols = ref_pyramid.rows, ref_pyramid.cols
        fr = np.zeros((rows, cols), np.float32)
        fc = np.zeros((rows, cols), np.float32)
        flags = _capi.FLAG_SINGLE_LEVEL if single_level else 0
        if single_level and flow_rc is not None and np.asarray(flow_rc[0]).shape == (rows, cols) and np.asarray(flow_rc[1]).shape == (rows, cols):
            fr[:] = flow_rc[0]
            fc[:] = flow_rc[1]
        else:
            flags |= _capi.FLAG_NO_PREDICTION
        prm = self._params()
        rc = lib().ftk_dense_flow_track(self.ctx._h, C.byref(prm), ref_pyramid._h, cur_pyramid._h, int(ref_image), int(cur_image), _ptr(fr), _ptr(fc), flags)
        ok = self.ctx.check(rc, soft=(_capi.ERR_LEVEL_MISMATCH,))
        return ok, fr, fc



# ---------------------------------------------------------------------------------------------------------------------------
# Feature detection + BRIEF description.  Class / option names follow the reference's call sites
# (test/test_descriptor_matcher_brief.cpp:59-76, test/test_optical_flow.cpp:60-66); the classes themselves belong to the absent
# sibling repository Feature_Detector, so parity is unpinned: the checker is oracle/ftk_oracle.c's restatement of the published
# algorithm (see include/ftk_c.h).
# ---------------------------------------------------------------------------------------------------------------------------
class FeaturePointDetectorOptions:
    def __init__(self):
        self.kMinValidResponse = 40.0
        self.kMinFeatureDistance = 20
        self.kHalfPatchSize = 1
        self.kHarrisK = 0.04


class FeaturePointHarrisDetector:
    """feature_detector::FeaturePointHarrisDetector as the reference uses it: options().kMinFeatureDistance / kMinValidResponse and
    DetectGoodFeatures(image, needed, features)."""

    KIND = 0

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._options = FeaturePointDetectorOptions()

    def options(self):
        return self._options

    def _params(self):
        o = self._options
        p = _capi.DetectorParams()
        p.kind, p.half_patch, p.harris_k = self.KIND, int(o.kHalfPatchSize), float(o.kHarrisK)
        p.min_response, p.min_distance = float(o.kMinValidResponse), int(o.kMinFeatureDistance)
        return p

    def DetectGoodFeatures(self, image, needed_feature_num, features=None, image_index=0, return_response=False):
        """`image` is an ImagePyramid (level 0 is searched; the trackers then reuse it).  `features` already held by the caller block
        their neighbourhood and are kept: returns (ok, features) with the new ones appended, so that at most `needed_feature_num`
        are held afterwards (the reference's in/out vector)."""
        existing = np.zeros((0, 2), np.float32) if features is None else np.ascontiguousarray(features, np.float32).reshape(-1, 2)
        want = max(0, int(needed_feature_num) - len(existing))
        out = np.zeros((max(want, 1), 2), np.float32)
        resp = np.zeros(max(want, 1), np.float32)
        n_out = C.c_int32(0)
        prm = self._params()
        rc = lib().ftk_detect_features(self.ctx._h, C.byref(prm), image._h, int(image_index), _ptr(existing) if len(existing) else None, len(existing),
                                       want, _ptr(out), _ptr(resp), C.byref(n_out), 0)
        ok = self.ctx.check(rc)
        merged = np.concatenate([existing, out[:n_out.value]], axis=0)
        if return_response:
            return ok, merged, resp[:n_out.value]
        return ok, merged

    def DetectGoodFeaturesBatch(self, images, needed_feature_num, first=0, count=None):
        """Every image of a pyramid batch in one call (ftk_detect_features_batch).  Returns (ok, [features of image first, first + 1, ...],
        [responses ...])."""
        count = images.n_images - first if count is None else int(count)
        needed = max(0, int(needed_feature_num))
        out = np.zeros((count, max(needed, 1), 2), np.float32)
        resp = np.zeros((count, max(needed, 1)), np.float32)
        n_out = np.zeros(count, np.int32)
        prm = self._params()
        rc = lib().ftk_detect_features_batch(self.ctx._h, C.byref(prm), images._h, int(first), count, needed, _ptr(out), _ptr(resp), _ptr(n_out), 0)
        ok = self.ctx.check(rc)
        return ok, [out[i, :n_out[i]].copy() for i in range(count)], [resp[i, :n_out[i]].copy() for i in range(count)]

    def ComputeResponse(self, image, image_index=0):
        """The response map of level 0 (rows x cols float32, -inf where the window leaves the image)."""
        out = np.zeros((image.rows, image.cols), np.float32)
        prm = self._params()
        self.ctx.check(lib().ftk_detect_response(self.ctx._h, C.byref(prm), image._h, int(image_index), _ptr(out), 0))
        return out


class FeaturePointShiTomasDetector(FeaturePointHarrisDetector):
    """Same selection, response = smaller eigenvalue of the structure tensor."""

    KIND = 1


class BriefDescriptorOptions:
    def __init__(self):
        self.kLength = 256
        self.kHalfPatchSize = 8
        self.kPatternSeed = 0


def brief_pattern(length, half_patch, seed=0):
    """The library's default pair list, int8 [length][4] = (drow_a, dcol_a, drow_b, dcol_b)."""
    pattern = np.zeros((int(length), 4), np.int8)
    lib().ftk_brief_pattern_default(int(length), int(half_patch), int(seed) & 0xFFFFFFFF, _ptr(pattern))
    return pattern


class BriefDescriptor:
    """feature_detector::BriefDescriptor as the reference uses it: options().kLength / kHalfPatchSize and Compute(image, features,
    descriptors).  Descriptors come back packed (uint32 [n][kLength / 32]), the layout the Hamming matchers take."""

    def __init__(self, ctx=None, pattern=None):
        self.ctx = ctx or default_context()
        self._options = BriefDescriptorOptions()
        self._pattern = None if pattern is None else np.ascontiguousarray(pattern, np.int8).reshape(-1, 4)

    def options(self):
        return self._options

    def pattern(self):
        o = self._options
        if self._pattern is not None:
            return self._pattern
        return brief_pattern(o.kLength, o.kHalfPatchSize, o.kPatternSeed)

    def Compute(self, image, features, image_index=0):
        """Returns (ok, descriptors, valid)."""
        o = self._options
        uv = np.ascontiguousarray(features, np.float32).reshape(-1, 2)
        pattern = self.pattern()
        n_bits = len(pattern)
        desc = np.zeros((len(uv), max(n_bits // 32, 1)), np.uint32)
        valid = np.zeros(len(uv), np.uint8)
        rc = lib().ftk_describe_brief(self.ctx._h, image._h, int(image_index), _ptr(uv) if len(uv) else None, len(uv), _ptr(pattern), n_bits,
                                      int(o.kHalfPatchSize), _ptr(desc) if len(uv) else None, _ptr(valid) if len(uv) else None, 0)
        ok = self.ctx.check(rc)
        return ok, desc, valid

    def ComputeBatch(self, images, feature_lists, first=0):
        """Descriptors of several images of a pyramid batch in one launch (ftk_describe_brief_batch): feature_lists[i] belongs to image
        first + i.  Returns (ok, [descriptors ...], [valid ...])."""
        o = self._options
        lists = [np.ascontiguousarray(f, np.float32).reshape(-1, 2) for f in feature_lists]
        offsets = np.concatenate([[0], np.cumsum([len(f) for f in lists])]).astype(np.int32)
        uv = np.concatenate(lists) if lists else np.zeros((0, 2), np.float32)
        pattern = self.pattern()
        n_bits = len(pattern)
        desc = np.zeros((len(uv), max(n_bits // 32, 1)), np.uint32)
        valid = np.zeros(len(uv), np.uint8)
        rc = lib().ftk_describe_brief_batch(self.ctx._h, images._h, int(first), len(lists), _ptr(offsets), _ptr(uv) if len(uv) else None, _ptr(pattern), n_bits,
                                            int(o.kHalfPatchSize), _ptr(desc) if len(uv) else None, _ptr(valid) if len(uv) else None, 0)
        ok = self.ctx.check(rc)
        return ok, [desc[offsets[i]:offsets[i + 1]] for i in range(len(lists))], [valid[offsets[i]:offsets[i + 1]] for i in range(len(lists))]
