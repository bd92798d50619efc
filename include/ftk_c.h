/* ftk_c.h -- C ABI of the B200-native sparse feature tracking / descriptor matching library (libftk_b200.so).
 *
 * This is the drop-in boundary for the reference's data-parallel hot path (Horizon1026/Feature_Tracker):
 * plain pointers and sizes, no C++ / torch types.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference root).  The reference has no FFI today (single-threaded C++ loops); the
 * seam a maintainer would bind is shown in INTEGRATION.md, and include/feature_tracker_b200/ holds C++ facade
 * classes with the reference's class names on top of this ABI.
 *
 * Conventions
 *   - All functions return FTK_OK (0) or a negative error code; ftk_last_error() gives the text.  The facade maps
 *     FTK_ERR_EMPTY_INPUT / FTK_ERR_LEVEL_MISMATCH / FTK_ERR_SIZE_MISMATCH to the reference's `return false`.
 *   - uv arrays are interleaved float32 (x = column, y = row), like std::vector<Vec2>.
 *   - status bytes are feature_tracker::TrackStatus values (src/feature_tracker.h:8-14).
 *   - Pointers are HOST pointers unless FTK_FLAG_DEVICE_POINTERS is given; host calls are synchronous (results are
 *     valid on return), device-pointer calls are asynchronous on the context's stream.  Inside a call the matcher kernels may
 *     overlap each other's launch (programmatic dependent launch); the first kernel of every call is an ordinary launch, so work
 *     the caller queued on ftk_stream() before the call is complete before the call's kernels read anything.
 *   - A context is bound to one GPU and one stream; it is not re-entrant (the reference objects are not either:
 *     src/optical_flow_tracker/optical_flow.h:94-103).  Use one context per host thread / per GPU.
 *   - There is no CPU fallback: every call fails with FTK_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef FTK_C_H_
#define FTK_C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTK_ABI_VERSION 4

/* ---- error codes ------------------------------------------------------------------------------------------ */
#define FTK_OK 0
#define FTK_ERR_INVALID_ARGUMENT (-1)
#define FTK_ERR_EMPTY_INPUT (-2)    /* reference: RETURN_FALSE_IF(ref_pixel_uv.empty()) / descriptors_cur.empty() */
#define FTK_ERR_LEVEL_MISMATCH (-3) /* reference: cur_pyramid.level() != ref_pyramid.level() */
#define FTK_ERR_SIZE_MISMATCH (-4)  /* reference: descriptor/uv vector sizes differ (descriptor_matcher.h:95-96) */
#define FTK_ERR_CUDA (-5)
#define FTK_ERR_UNSUPPORTED (-6)

/* ---- TrackStatus (src/feature_tracker.h:8-14) ---------------------------------------------------------------- */
#define FTK_STATUS_NOT_TRACKED 0
#define FTK_STATUS_TRACKED 1
#define FTK_STATUS_LARGE_RESIDUAL 2
#define FTK_STATUS_OUTSIDE 3
#define FTK_STATUS_NUMERIC_ERROR 4

/* ---- flags -------------------------------------------------------------------------------------------------- */
#define FTK_FLAG_DEVICE_POINTERS 1u /* feature / descriptor / index arrays already live on the context's GPU */
#define FTK_FLAG_NO_PREDICTION 2u   /* cur_uv holds no prediction: behave as cur_pixel_uv.size() != ref size (optical_flow.cpp:12-14) */
#define FTK_FLAG_NO_STATUS 4u       /* status holds nothing: behave as status.size() != ref size (optical_flow.cpp:17-19) */
#define FTK_FLAG_SINGLE_LEVEL 8u    /* the GrayImage overload: TrackSingleLevel on level 0 (optical_flow.cpp:28-47) */
#define FTK_FLAG_NO_INDEX_INPUT 16u /* idx holds nothing: behave as index_pairs_in_cur.size() != ref size (descriptor_matcher.h:60-62) */

/* ---- KLT parameters ------------------------------------------------------------------------------------------
 * OpticalFlowOptions (src/optical_flow_tracker/optical_flow.h:20-28), OpticalFlowMethod (:12-18) and the subclass
 * extras predict_affine() (affine_klt/optical_flow_affine_klt.h:18), predict_R_cr() / consider_patch_luminance()
 * (lssd_klt/optical_flow_lssd_klt.h:18-19). */
#define FTK_VARIANT_BASIC 0  /* OpticalFlowBasicKlt  */
#define FTK_VARIANT_AFFINE 1 /* OpticalFlowAffineKlt */
#define FTK_VARIANT_LSSD 2   /* OpticalFlowLssdKlt   */
#define FTK_METHOD_INVERSE 0
#define FTK_METHOD_DIRECT 1
#define FTK_METHOD_FAST 2 /* kSse(3) / kNeon(4) take the reference's `default:` branch, i.e. kFast */

typedef struct ftk_klt_params {
    int32_t variant;
    int32_t method;
    uint32_t max_track_points;         /* kMaxTrackPointsNumber, applied per frame pair */
    uint32_t max_iteration;            /* kMaxIteration */
    uint32_t max_tolerance_large_step; /* kMaxToleranceLargeStep */
    int32_t patch_row_half;            /* kPatchRowHalfSize */
    int32_t patch_col_half;            /* kPatchColHalfSize */
    float max_converge_step;           /* kMaxConvergeStep, compared with the SQUARED step */
    float predict[4];                  /* row-major 2x2: predict_affine_ (affine, single level) / predict_R_cr_ (lssd) */
    int32_t consider_patch_luminance;  /* lssd fast only */
    /* Not in the reference (SURVEY 8(f) "next" row): > 0 adds a forward-backward consistency pass.  Every feature the forward
     * pass left kTracked is tracked back from its result in the current frame into the reference frame -- exactly
     * TrackFeatures(cur_pyramid, ref_pyramid, result, prediction = its reference position) with identity `predict` -- and is
     * downgraded to kLargeResidual unless the backward track is kTracked and ends within this many pixels of the
     * reference position (dx*dx + dy*dy <= max*max in fp32).  0 (default) = the reference's behaviour. */
    float forward_backward_max_error;
} ftk_klt_params;

/* Fills the reference defaults (optical_flow.h:20-28; identity predictions; luminance off). */
void ftk_klt_params_default(ftk_klt_params *params);

/* ---- context ------------------------------------------------------------------------------------------------ */
typedef struct ftk_context ftk_context;

int ftk_abi_version(void);
int ftk_create(int device, ftk_context **out);
void ftk_destroy(ftk_context *ctx);
const char *ftk_last_error(const ftk_context *ctx);
int ftk_synchronize(ftk_context *ctx);
/* The context's cudaStream_t, for callers that time with CUDA events or enqueue their own work. */
void *ftk_stream(ftk_context *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t ftk_kernel_launches(const ftk_context *ctx);

/* Page-locked host memory for the host-pointer entry points: copies from / to it are asynchronous DMA transfers (what "pinned memory
 * recommended" means below); ordinary memory works everywhere but is staged by the driver.  Not tied to a context. */
int ftk_alloc_pinned(size_t bytes, void **out);
void ftk_free_pinned(void *ptr);
/* Measurement hook (bench.py's per-kernel roofline figures): while enabled, the entry points that have one dominant kernel
 * (ftk_match_cosine_force: the tcgen05 GEMM; ftk_match_hamming_force: the POPC kernel; ftk_klt_track: the tracker kernel) record
 * CUDA events around that kernel's launch on the context's stream.  ftk_last_kernel_ms synchronises and returns the duration of the
 * last such launch in milliseconds (negative when none was recorded).  Off by default: no events are recorded. */
int ftk_set_profiling(ftk_context *ctx, int enabled);
float ftk_last_kernel_ms(ftk_context *ctx);

/* ---- image pyramids (replaces ImagePyramid::SetRawImage / CreateImagePyramid, call sites
 *      test/test_optical_flow.cpp:49-53,70-71; semantics: oracle/shim/datatype_image_pyramid.h) -------------------
 * A pyramid batch holds n_images same-sized grayscale images, device resident, level-major. */
typedef struct ftk_pyramid ftk_pyramid;

int ftk_pyramid_create(ftk_context *ctx, int32_t rows, int32_t cols, int32_t levels, int32_t n_images, ftk_pyramid **out);
void ftk_pyramid_destroy(ftk_context *ctx, ftk_pyramid *pyr);
/* Copies `count` tightly packed rows*cols images into level 0 of images [first, first+count).
 * `images` is a host pointer (pinned memory makes the copy asynchronous) or a device pointer with
 * FTK_FLAG_DEVICE_POINTERS. */
int ftk_pyramid_set_images(ftk_context *ctx, ftk_pyramid *pyr, int32_t first, int32_t count, const uint8_t *images, uint32_t flags);
/* Builds levels 1..levels-1 of images [first, first+count) from their level 0 (asynchronous on the stream). */
int ftk_pyramid_build(ftk_context *ctx, ftk_pyramid *pyr, int32_t first, int32_t count);
/* Verbatim access to one level of one image (tightly packed (rows>>level)*(cols>>level) bytes, host memory). */
int ftk_pyramid_set_level(ftk_context *ctx, ftk_pyramid *pyr, int32_t image, int32_t level, const uint8_t *data);
int ftk_pyramid_get_level(ftk_context *ctx, const ftk_pyramid *pyr, int32_t image, int32_t level, uint8_t *data);
int32_t ftk_pyramid_levels(const ftk_pyramid *pyr);
int32_t ftk_pyramid_images(const ftk_pyramid *pyr);

/* ---- sparse KLT tracking (replaces OpticalFlow::TrackFeatures, src/optical_flow_tracker/optical_flow.cpp:6-47,
 *      and the TrackMultipleLevel / TrackSingleLevel overrides of the three subclasses) --------------------------
 * Tracks n_pairs independent frame pairs in one call.  Pair p uses image ref_image[p] of `ref` and cur_image[p]
 * of `cur` (NULL = image p) and owns features [feat_offsets[p], feat_offsets[p+1]).  `ref` and `cur` may be the
 * same batch.  cur_uv is in/out (prediction in, result out), status is in/out (entries > kTracked are skipped
 * untouched; only the first max_track_points features of each pair are tracked). */
int ftk_klt_track(ftk_context *ctx, const ftk_klt_params *params, const ftk_pyramid *ref, const ftk_pyramid *cur, int32_t n_pairs,
                  const int32_t *ref_image, const int32_t *cur_image, const int32_t *feat_offsets, const float *ref_uv, float *cur_uv,
                  uint8_t *status, uint32_t flags);

/* One call for the reference demo's whole timed region -- CreateImagePyramid x2 + TrackFeatures
 * (test/test_optical_flow.cpp:69-73) -- over a batch of frame pairs given as HOST images (n_pairs tightly packed rows*cols
 * images each; pinned memory recommended).  The batch is processed in chunks on two streams so the host-to-device copy of
 * chunk k+1 overlaps pyramid construction and tracking of chunk k.  ref_uv / cur_uv / status / feat_offsets are host arrays
 * with the semantics of ftk_klt_track (FTK_FLAG_NO_PREDICTION / FTK_FLAG_NO_STATUS apply; multi-level only). */
int ftk_track_image_pairs(ftk_context *ctx, const ftk_klt_params *params, int32_t rows, int32_t cols, int32_t levels, int32_t n_pairs,
                          const uint8_t *ref_images, const uint8_t *cur_images, const int32_t *feat_offsets, const float *ref_uv, float *cur_uv,
                          uint8_t *status, uint32_t flags);

/* The same for several parameter sets at once (e.g. kDirect and kFast of one tracker, as BASELINE configs[1] runs them): every
 * tracker k runs on each chunk while its images are resident, reading / writing cur_uv + k * 2 * n_features and
 * status + k * n_features (n_features = feat_offsets[n_pairs]), so the result equals n_trackers separate calls while the images
 * cross the host link once. */
int ftk_track_image_pairs_multi(ftk_context *ctx, const ftk_klt_params *params, int32_t n_trackers, int32_t rows, int32_t cols, int32_t levels,
                                int32_t n_pairs, const uint8_t *ref_images, const uint8_t *cur_images, const int32_t *feat_offsets, const float *ref_uv,
                                float *cur_uv, uint8_t *status, uint32_t flags);

/* The temporal form of the same pipeline (SURVEY 8(f) "next" row: the current frame of pair k is the reference frame of pair
 * k+1, but the reference's callers rebuild both pyramids for every pair, test/test_optical_flow.cpp:70-71).  `frames` holds
 * n_frames >= 2 tightly packed HOST images; pair k = (frame k -> frame k+1), k = 0 .. n_frames-2, owns features
 * [feat_offsets[k], feat_offsets[k+1]).  Every frame is uploaded once and its pyramid is built once, which halves the
 * host-to-device traffic and the pyramid work of ftk_track_image_pairs; results are identical to it. */
int ftk_track_image_sequence(ftk_context *ctx, const ftk_klt_params *params, int32_t rows, int32_t cols, int32_t levels, int32_t n_frames,
                             const uint8_t *frames, const int32_t *feat_offsets, const float *ref_uv, float *cur_uv, uint8_t *status,
                             uint32_t flags);

/* ---- direct-method pose tracker (SURVEY 8(f) "next" row; replaces DirectMethod::TrackFeatures, camera-frame overload,
 *      src/direct_method_tracker/direct_method_tracker.cpp:41-95, and TrackAllFeaturesDirect, :115-192) --------------------
 * DirectMethodOptions (direct_method_tracker.h:20-28).  Only kDirect does work upstream; kInverse / kFast are empty there
 * (:107-113, :194-199) and leave pose and positions untouched here too. */
#define FTK_DIRECT_METHOD_INVERSE 0
#define FTK_DIRECT_METHOD_DIRECT 1
#define FTK_DIRECT_METHOD_FAST 2
typedef struct ftk_direct_params {
    uint32_t max_track_points;   /* kMaxTrackPointsNumber (per frame pair) */
    uint32_t max_iteration;      /* kMaxIteration (per pyramid level) */
    int32_t patch_row_half;      /* kPatchRowHalfSize */
    int32_t patch_col_half;      /* kPatchColHalfSize */
    float max_converge_step;     /* kMaxConvergeStep, compared with the squared 6-vector step */
    float max_converge_residual; /* kMaxConvergeResidual (unused by the reference) */
    int32_t method;              /* DirectMethodMethod */
} ftk_direct_params;
void ftk_direct_params_default(ftk_direct_params *params);
/* One 6-DoF pose per frame pair, estimated from ALL features of the pair (n_pairs independent problems per call).  Pair p
 * uses image ref_image[p] / cur_image[p] (NULL = image p), features [feat_offsets[p], feat_offsets[p+1]), intrinsics
 * K[4p..4p+3] = (fx, fy, cx, cy), and updates q_rc[4p..] = (w, x, y, z) and p_rc[3p..] (current frame in the reference frame)
 * in place.  p_c_in_ref: feature positions in the reference camera frame (n x 3).  cur_uv (in/out; out = the projection with
 * the pose of the last linearisation) and status follow the reference: FTK_FLAG_NO_PREDICTION = cur_pixel_uv.size() differs
 * (:48-50), FTK_FLAG_NO_STATUS = status.size() differs => all kTracked (:82-84); features projecting outside become
 * kOutside (:86-91).  q_rc and K must be 16-byte aligned when passed as device pointers. */
int ftk_direct_method_track(ftk_context *ctx, const ftk_direct_params *params, const ftk_pyramid *ref, const ftk_pyramid *cur, int32_t n_pairs,
                            const int32_t *ref_image, const int32_t *cur_image, const int32_t *feat_offsets, const float *K, const float *p_c_in_ref,
                            const float *ref_uv, float *cur_uv, float *q_rc, float *p_rc, uint8_t *status, uint32_t flags);

/* ---- dense optical flow (SURVEY 8(f) "next" row; replaces DenseOpticalFlow::Track, both overloads,
 *      src/dense_optical_flow_tracker/dense_optical_flow.cpp:7-85) ---------------------------------------------------------
 * DenseOpticalFlow::Options (dense_optical_flow.h:15-20). */
typedef struct ftk_dense_flow_params {
    int32_t max_iteration;     /* kMaxIteration */
    int32_t half_patch_size;   /* kHalfPatchSize (<= 7 in this build) */
    float max_converge_step;   /* kMaxConvergeStep */
    float max_delta_flow_step; /* kMaxDeltaFlowStep */
} ftk_dense_flow_params;
void ftk_dense_flow_params_default(ftk_dense_flow_params *params);
/* Flow from image ref_image of `ref` to image cur_image of `cur`: flow_row / flow_col = rows x cols tightly packed floats (the
 * reference's flow_rc[0] / flow_rc[1]).  Default: the pyramid overload (coarse to fine over all levels; the flow is an output
 * only).  FTK_FLAG_SINGLE_LEVEL: the GrayImage overload on level 0; flow_* then are in/out unless FTK_FLAG_NO_PREDICTION says
 * the caller's matrices have the wrong size (:18-23: start from zero). */
int ftk_dense_flow_track(ftk_context *ctx, const ftk_dense_flow_params *params, const ftk_pyramid *ref, const ftk_pyramid *cur, int32_t ref_image,
                         int32_t cur_image, float *flow_row, float *flow_col, uint32_t flags);

/* ---- feature detection + BRIEF description (the callers' upstream step
 *      feature_detector::FeaturePointHarrisDetector::DetectGoodFeatures / feature_detector::BriefDescriptor::Compute,
 *      test/test_descriptor_matcher_brief.cpp:59-76, test/test_optical_flow.cpp:60-66) -------------------------------------
 * PARITY UNPINNED: those classes live in the sibling repository Feature_Detector, which is absent from the reference tree
 * (CMakeLists.txt:24-29) and unpinned.  The option names follow the call sites; the arithmetic is the published algorithm as
 * frozen in the detector section of oracle/ftk_oracle.c, to which these entry points are bit-exact. */
#define FTK_DETECTOR_HARRIS 0
#define FTK_DETECTOR_SHI_TOMASI 1
typedef struct ftk_detector_params {
    int32_t kind;         /* FTK_DETECTOR_* */
    int32_t half_patch;   /* structure-tensor window half size, 1..3 */
    float harris_k;       /* 0.04; 0 <= k <= 1 */
    float min_response;   /* kMinValidResponse (finite) */
    int32_t min_distance; /* kMinFeatureDistance: a taken feature blocks |drow| < d and |dcol| < d */
} ftk_detector_params;
void ftk_detector_params_default(ftk_detector_params *params);
/* Detects up to `needed` features on level 0 of image `image` of `pyr` (the pyramid the trackers use: no second upload).
 * existing_uv [n_existing][2] (x, y; may be NULL when n_existing == 0) are features the caller already tracks: they block
 * their neighbourhood and are not returned.  out_uv [needed][2] receives (x = col, y = row) in falling response order,
 * out_response [needed] the responses (may be NULL), *n_out (HOST pointer, always) the count.  FTK_FLAG_DEVICE_POINTERS
 * applies to existing_uv / out_uv / out_response.  The call synchronises the context. */
int ftk_detect_features(ftk_context *ctx, const ftk_detector_params *params, const ftk_pyramid *pyr, int32_t image, const float *existing_uv,
                        int32_t n_existing, int32_t needed, float *out_uv, float *out_response, int32_t *n_out, uint32_t flags);
/* The same for images first_image .. first_image + n_images - 1 of the batch in one call (every kernel covers all images, so the
 * launch count does not grow with the batch): out_uv [n_images][needed][2], out_response [n_images][needed] (may be NULL), n_out
 * [n_images] (HOST).  Slots past n_out[i] are zero for host callers and untouched for device callers. */
int ftk_detect_features_batch(ftk_context *ctx, const ftk_detector_params *params, const ftk_pyramid *pyr, int32_t first_image, int32_t n_images,
                              int32_t needed, float *out_uv, float *out_response, int32_t *n_out, uint32_t flags);
/* The response map alone: rows x cols tightly packed floats, -inf where the window leaves the image. */
int ftk_detect_response(ftk_context *ctx, const ftk_detector_params *params, const ftk_pyramid *pyr, int32_t image, float *response, uint32_t flags);
/* Default BRIEF pair list (HOST memory): pattern [n_bits][4] = (drow_a, dcol_a, drow_b, dcol_b) inside +-half_patch, from
 * xorshift32 seeded with `seed`. */
void ftk_brief_pattern_default(int32_t n_bits, int32_t half_patch, uint32_t seed, int8_t *pattern);
/* BRIEF descriptors of n features on level 0 of image `image`: bit k = I(p + a_k) < I(p + b_k) at the truncated position p;
 * desc [n][n_bits / 32] in the packed layout ftk_match_hamming_* takes (n_bits a multiple of 32, <= 1024); a feature whose
 * +-half_patch patch leaves the image gets zeros and valid[i] = 0 (valid may be NULL).  `pattern` is always a HOST pointer;
 * FTK_FLAG_DEVICE_POINTERS applies to uv / desc / valid. */
int ftk_describe_brief(ftk_context *ctx, const ftk_pyramid *pyr, int32_t image, const float *uv, int32_t n, const int8_t *pattern, int32_t n_bits,
                       int32_t half_patch, uint32_t *desc, uint8_t *valid, uint32_t flags);

/* The same for images first_image .. first_image + n_images - 1 in one launch: the features of image first_image + i are
 * uv[feat_offsets[i] .. feat_offsets[i+1]-1] (feat_offsets: HOST, n_images + 1 entries starting at 0); desc / valid are indexed like uv. */
int ftk_describe_brief_batch(ftk_context *ctx, const ftk_pyramid *pyr, int32_t first_image, int32_t n_images, const int32_t *feat_offsets, const float *uv,
                             const int8_t *pattern, int32_t n_bits, int32_t half_patch, uint32_t *desc, uint8_t *valid, uint32_t flags);

/* ---- descriptor matching (replaces DescriptorMatcher<T>::ForceMatch / NearbyMatch,
 *      src/descriptor_matcher/descriptor_matcher.h:55-79,90-124, with the ComputeDistance bodies of
 *      test/test_descriptor_matcher_brief.cpp:33-45 and test/test_descriptor_matcher_superpoint.cpp:32-34) --------
 * Binary descriptors are packed little-endian: element k of a BriefType is bit (k % 32) of word (k / 32).
 * idx is in/out: idx[i] is overwritten only when ref descriptor i finds a match with distance < max_dist
 * (strict <, lowest cur index wins ties). */
int ftk_match_hamming_force(ftk_context *ctx, const uint32_t *ref, int32_t n_ref, const uint32_t *cur, int32_t n_cur, int32_t words,
                            float max_dist, int32_t *idx, uint32_t flags);
int ftk_match_hamming_nearby(ftk_context *ctx, const uint32_t *ref, int32_t n_ref, const uint32_t *cur, int32_t n_cur, int32_t words,
                             const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx,
                             uint32_t flags);
/* Many independent ForceMatch / NearbyMatch problems in one call (one per frame pair, like the pair batches of ftk_klt_track): pair p
 * matches ref descriptors ref_offsets[p] .. ref_offsets[p+1]-1 against cur descriptors cur_offsets[p] .. cur_offsets[p+1]-1 (offsets
 * are HOST arrays of n_pairs + 1 entries starting at 0).  pred_uv == NULL: ForceMatch; otherwise NearbyMatch with pred_uv [n_ref_total][2]
 * and cur_uv [n_cur_total][2].  idx [n_ref_total] is in/out like the single-pair calls; a match is the index INSIDE the pair's cur set.
 * A pair without cur descriptors leaves its idx entries untouched (the reference returns false for it, descriptor_matcher.h:58). */
int ftk_match_hamming_pairs(ftk_context *ctx, const uint32_t *ref, const uint32_t *cur, int32_t words, int32_t n_pairs, const int32_t *ref_offsets,
                            const int32_t *cur_offsets, const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist,
                            int32_t *idx, uint32_t flags);
/* The same for float descriptors (row-major n x dim, dim <= 1024; distance 0.5 - 0.5 * cos as in ftk_match_cosine_*): every pair is an
 * independent ForceMatch (pred_uv == NULL) or NearbyMatch of the reference, evaluated with its own sequential fp32 arithmetic. */
int ftk_match_cosine_pairs(ftk_context *ctx, const float *ref, const float *cur, int32_t dim, int32_t n_pairs, const int32_t *ref_offsets,
                           const int32_t *cur_offsets, const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist,
                           int32_t *idx, uint32_t flags);
/* Float descriptors, row-major n x dim, distance 0.5 - 0.5 * cos(ref, cur). */
int ftk_match_cosine_force(ftk_context *ctx, const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, float max_dist,
                           int32_t *idx, uint32_t flags);
int ftk_match_cosine_nearby(ftk_context *ctx, const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim,
                            const float *pred_uv, const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx,
                            uint32_t flags);
/* Diagnostics: number of reference rows the last ftk_match_cosine_force call had to hand to the exact scan because the tensor-core
 * pass could not separate their candidates or their norm is abnormal (0 when the screening decided every row; -1 when the call did
 * not use the tensor-core path).  Synchronises the context.
 * Bit-exactness contract: indices equal the reference's sequential fp32 evaluation for ALL inputs, including descriptors whose sum of
 * squares overflows, underflows or is NaN (those rows / columns are evaluated with the exact arithmetic only). */
int ftk_last_cosine_exact_scan_items(ftk_context *ctx);
/* ---- mutual arg-max of a score matrix (SURVEY 8(f) "next" row; replaces the score-matrix post-processing of
 *      NNFeatureMatcher::Match, src/nn_feature_matcher/nn_feature_matcher.cpp:180-216) ---------------------------------
 * scores = n_ref x n_cur row-major floats (LightGlue log-assignment).  idx[i] = j when j is the FIRST maximum of row i
 * (`scores(j) > max_score`, :205), scores[i][j] is not < min_score (kMinValidMatchScore, :210) and the first maximum of
 * column j is row i (:211); -1 otherwise.  Every idx entry is written.  Matched positions / status then follow from
 * ftk_fill_matched.  NaN entries behave as in the reference's comparisons. */
int ftk_match_mutual_scores(ftk_context *ctx, const float *scores, int32_t n_ref, int32_t n_cur, float min_score, int32_t *idx, uint32_t flags);
/* Cross-check filter for the descriptor matchers: idx_ref_to_cur[i] (from a Force / NearbyMatch ref -> cur) is reset to -1
 * unless idx_cur_to_ref[idx_ref_to_cur[i]] == i (from the same call with the two sets swapped). */
int ftk_match_cross_check(ftk_context *ctx, int32_t *idx_ref_to_cur, int32_t n_ref, const int32_t *idx_cur_to_ref, int32_t n_cur, uint32_t flags);
/* Host-side helper mirroring FillMatchedPixelByPairIndices (descriptor_matcher.h:135-157); status_valid == 0
 * behaves as status.size() != idx.size(). */
int ftk_fill_matched(const int32_t *idx, int32_t n_ref, const float *cur_uv, int32_t n_cur, float *matched_uv, uint8_t *status,
                     int32_t status_valid);

#ifdef __cplusplus
}
#endif
#endif /* FTK_C_H_ */
