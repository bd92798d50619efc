/* ORACLE (test infrastructure -- never shipped, never measured as the product).
 *
 * Plain-C, single-threaded CPU restatement of the reference hot path (pyramid, 3 KLT trackers x 3 methods,
 * descriptor matching).  Every function in ftk_oracle.c cites the reference file:line it follows.
 *
 * Pinning: the reference has no golden vectors or assertions for this path, so this
 * restatement is pinned against outputs of the reference itself: oracle/_ref/libftk_ref.so (the reference's
 * own .cpp compiled in place against oracle/shim/) must agree BIT-FOR-BIT with this file on the bundled EuRoC
 * pair and on seeded synthetic inputs (tests/test_oracle_vs_ref.py), and both must reproduce the committed
 * golden dumps under tests/golden/.  The external dependencies the reference leans on (Eigen LDLT,
 * Slam_Utility GrayImage / ImagePyramid) are absent from /root/reference; their semantics are frozen in
 * oracle/shim/  and restated here identically.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library. */
#ifndef FTK_ORACLE_H_
#define FTK_ORACLE_H_
#include <stdint.h>

#include "ftk_oracle_types.h"

#ifdef __cplusplus
extern "C" {
#endif

int ftko_pyramid_build(const uint8_t *image, int32_t rows, int32_t cols, int32_t levels, uint8_t *out);

int ftko_klt_track(const ftko_klt_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                   const int32_t *rows, const int32_t *cols, int32_t n, const float *ref_uv, float *cur_uv, int32_t cur_uv_count, uint8_t *status,
                   int32_t status_count, int32_t single_level);

/* Same as ftko_klt_track, and additionally reports, per feature, how many Gauss-Newton iterations (calls of
 * the per-patch accumulation routine) were executed over all levels: the "patch-iteration" unit of SURVEY.md
 * section 8(d).  iterations may be NULL. */
int ftko_klt_track_traced(const ftko_klt_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                          const int32_t *rows, const int32_t *cols, int32_t n, const float *ref_uv, float *cur_uv, int32_t cur_uv_count,
                          uint8_t *status, int32_t status_count, int32_t single_level, int32_t *iterations);

int ftko_pyramid_and_track(const ftko_klt_params *params, int32_t levels, const uint8_t *ref_image, const uint8_t *cur_image, int32_t rows,
                           int32_t cols, int32_t n, const float *ref_uv, float *cur_uv, uint8_t *status);

int ftko_match_brief_force(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, float max_dist, int32_t *idx,
                           int32_t idx_count);
int ftko_match_brief_nearby(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *pred_uv,
                            const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx, int32_t idx_count);
int ftko_match_cosine_force(const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, float max_dist, int32_t *idx,
                            int32_t idx_count);
int ftko_match_cosine_nearby(const float *ref, int32_t n_ref, const float *cur, int32_t n_cur, int32_t dim, const float *pred_uv, const float *cur_uv,
                             int32_t max_drow, int32_t max_dcol, float max_dist, int32_t *idx, int32_t idx_count);
int ftko_match_brief_nearby_uv(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *pred_uv,
                               const float *cur_uv, int32_t max_drow, int32_t max_dcol, float max_dist, float *matched_uv, uint8_t *status,
                               int32_t status_count);
int ftko_match_brief_force_uv(const uint8_t *ref_bits, int32_t n_ref, const uint8_t *cur_bits, int32_t n_cur, int32_t len, const float *cur_uv,
                              float max_dist, float *matched_uv, uint8_t *status, int32_t status_count);

/* Exposed for unit tests of the LDLT restatement: solves A x = b for n in {2,3,6}; A row-major n*n. */
/* nn_feature_matcher.cpp:180-216: mutual row / column arg-max of a score matrix (restatement only; needs no network). */
int ftko_mutual_scores(const float *scores, int32_t n_ref, int32_t n_cur, float min_score, int32_t *idx);

/* direct_method_tracker.cpp:41-95 (camera-frame TrackFeatures) + :115-192.  q_rc = (w, x, y, z), in/out like p_rc. */
int ftko_direct_method_track(const ftko_direct_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                             const int32_t *rows, const int32_t *cols, const float *K, int32_t n, const float *p_c_in_ref, const float *ref_uv,
                             float *cur_uv, int32_t cur_uv_count, float *q_rc, float *p_rc, uint8_t *status, int32_t status_count);

/* dense_optical_flow.cpp:7-85 DenseOpticalFlow::Track; flow_* = rows[0] x cols[0] floats (row / column flow component). */
int ftko_dense_flow_track(const ftko_dense_flow_params *params, int32_t levels, const uint8_t *const *ref_levels, const uint8_t *const *cur_levels,
                          const int32_t *rows, const int32_t *cols, int32_t single_level, int32_t flow_valid, float *flow_row, float *flow_col);

/* SURVEY 8(f) rank 1 -- PARITY UNPINNED (Feature_Detector's sources are absent; see the block comment in ftk_oracle.c).
 * Response map (rows x cols floats, -inf where undefined), greedy min-distance selection in response order, BRIEF bits. */
int ftko_detect_response(const ftko_detector_params *params, const uint8_t *image, int32_t rows, int32_t cols, float *response);
int ftko_detect_features(const ftko_detector_params *params, const uint8_t *image, int32_t rows, int32_t cols, const float *existing_uv,
                         int32_t n_existing, int32_t needed, float *out_uv, float *out_response);
void ftko_brief_pattern(int32_t n_bits, int32_t half_patch, uint32_t seed, int8_t *pattern);
int ftko_describe_brief(const uint8_t *image, int32_t rows, int32_t cols, const float *uv, int32_t n, const int8_t *pattern, int32_t n_bits,
                        int32_t half_patch, uint32_t *desc, uint8_t *valid);

/* Diagnostics: number of unchecked samples whose base pixel lay outside the image since the last reset (reference UB). */
long long ftko_outside_reads(int32_t reset);

void ftko_ldlt_solve(int32_t n, const float *a, const float *b, float *x);

#ifdef __cplusplus
}
#endif
#endif
