/* ORACLE (test infrastructure, never shipped or measured as product).
 * Plain-C parameter block shared by the two CPU checkers in this directory:
 *   - oracle/_ref/libftk_ref.so   : the reference's own .cpp files compiled in place against oracle/shim/
 *   - oracle/_build/libftk_oracle.so : the plain-C restatement in ftk_oracle.c
 * Field meaning follows the reference's OpticalFlowOptions (src/optical_flow_tracker/optical_flow.h:20-28),
 * OpticalFlowMethod (:12-18) and the subclass extras (affine_klt/optical_flow_affine_klt.h:18,
 * lssd_klt/optical_flow_lssd_klt.h:18-19).  The layout is deliberately identical to `ftk_klt_params` in
 * include/ftk_c.h so one ctypes.Structure serves all three libraries. */
#ifndef FTK_ORACLE_TYPES_H_
#define FTK_ORACLE_TYPES_H_
#include <stdint.h>

typedef struct ftko_klt_params {
    int32_t variant;                   /* 0 basic, 1 affine, 2 lssd */
    int32_t method;                    /* 0 kInverse, 1 kDirect, 2 kFast (3,4 = kSse/kNeon fall to kFast) */
    uint32_t max_track_points;         /* kMaxTrackPointsNumber */
    uint32_t max_iteration;            /* kMaxIteration */
    uint32_t max_tolerance_large_step; /* kMaxToleranceLargeStep */
    int32_t patch_row_half;            /* kPatchRowHalfSize */
    int32_t patch_col_half;            /* kPatchColHalfSize */
    float max_converge_step;           /* kMaxConvergeStep (on the SQUARED step) */
    float predict[4];                  /* row-major 2x2: predict_affine_ (affine) / predict_R_cr_ (lssd) */
    int32_t consider_patch_luminance;  /* lssd fast only */
} ftko_klt_params;

/* DirectMethodOptions (src/direct_method_tracker/direct_method_tracker.h:20-28); same layout as ftk_direct_params. */
typedef struct ftko_direct_params {
    uint32_t max_track_points;   /* kMaxTrackPointsNumber = 500 */
    uint32_t max_iteration;      /* kMaxIteration = 15 */
    int32_t patch_row_half;      /* kPatchRowHalfSize = 6 */
    int32_t patch_col_half;      /* kPatchColHalfSize = 6 */
    float max_converge_step;     /* kMaxConvergeStep = 1e-6 (on the squared step) */
    float max_converge_residual; /* kMaxConvergeResidual = 2.0 (unused by the reference) */
    int32_t method;              /* DirectMethodMethod: 0 kInverse, 1 kDirect (default), 2 kFast; only kDirect does work upstream */
} ftko_direct_params;

/* DenseOpticalFlow::Options (src/dense_optical_flow_tracker/dense_optical_flow.h:15-20); same layout as ftk_dense_flow_params. */
typedef struct ftko_dense_flow_params {
    int32_t max_iteration;     /* kMaxIteration = 10 */
    int32_t half_patch_size;   /* kHalfPatchSize = 2 */
    float max_converge_step;   /* kMaxConvergeStep = 1e-6 */
    float max_delta_flow_step; /* kMaxDeltaFlowStep = 1.0 */
} ftko_dense_flow_params;

/* Feature detector options.  Feature_Detector (the sibling repository that owns FeaturePointHarrisDetector / BriefDescriptor) is
 * absent from /root/reference: the option NAMES come from the call sites (test/test_descriptor_matcher_brief.cpp:59-71), the
 * arithmetic is the published Harris / Shi-Tomasi definition frozen in oracle/ftk_oracle.c.  Same layout as ftk_detector_params. */
typedef struct ftko_detector_params {
    int32_t kind;         /* 0 Harris (det - k trace^2), 1 Shi-Tomasi (smaller eigenvalue) */
    int32_t half_patch;   /* structure-tensor window half size, 1..3 */
    float harris_k;       /* 0.04 */
    float min_response;   /* kMinValidResponse */
    int32_t min_distance; /* kMinFeatureDistance */
} ftko_detector_params;

#endif
