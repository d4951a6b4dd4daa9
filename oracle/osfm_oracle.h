/*
 * osfm_oracle.h -- CPU restatement of the OrthoSfM / MVE exhaustive matcher.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (orthosfm_b200/,
 * include/, the C-ABI library) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md section 8c), so the restatement is pinned against the reference's
 * own sources compiled in place (oracle/_ref, built by oracle/Makefile) on
 * seeded random, adversarial and real-image inputs: tests/test_oracle.py, and
 * the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py
 * from oracle/_ref).
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Descriptors are held as one byte per element
 * (the reference widens the same values to 16-bit lanes, exhaustive_matching.h:47-48).
 */
#ifndef OSFM_ORACLE_H
#define OSFM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* NearestNeighbor<T>::Result, src/mve/sfm/nearest_neighbor.h:50-56.  The
 * distances are the *converted* squared distances (find() post-processing). */
typedef struct {
    float dist_1st_best;   /* exact for the integer paths (values < 2^16) */
    float dist_2nd_best;
    int index_1st_best;
    int index_2nd_best;
} osfm_oracle_nn_result;

/* convert_descriptor, src/mve/sfm/exhaustive_matching.cc:18-39 */
void osfm_oracle_quantize_sift(const float* desc, int n, uint8_t* out);
void osfm_oracle_quantize_surf(const float* desc, int n, int8_t* out);

/* NearestNeighbor<T>::find, src/mve/sfm/nearest_neighbor.cc:216-289 */
void osfm_oracle_nn_u8(const uint8_t* query, const uint8_t* elements,
    int num_elements, int dim, osfm_oracle_nn_result* result);
void osfm_oracle_nn_s8(const int8_t* query, const int8_t* elements,
    int num_elements, int dim, osfm_oracle_nn_result* result);
/* sse3_order != 0: 4-lane partial sums + two hadd (nearest_neighbor.cc:155-166),
 * else the sequential scalar loop (:186-191). */
void osfm_oracle_nn_f32(const float* query, const float* elements,
    int num_elements, int dim, int sse3_order, osfm_oracle_nn_result* result);

/* Matching::oneway_match<T>, src/mve/sfm/matching.h:114-146.
 * result has set_1_size entries, -1 = no match. */
void osfm_oracle_oneway_u8(const uint8_t* set_1, int set_1_size,
    const uint8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int* result);
void osfm_oracle_oneway_s8(const int8_t* set_1, int set_1_size,
    const int8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int* result);
void osfm_oracle_oneway_f32(const float* set_1, int set_1_size,
    const float* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int sse3_order,
    int* result);

/* Matching::twoway_match<T>, src/mve/sfm/matching.h:148-159 */
void osfm_oracle_twoway_u8(const uint8_t* set_1, int set_1_size,
    const uint8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold,
    int* matches_1_2, int* matches_2_1);
void osfm_oracle_twoway_s8(const int8_t* set_1, int set_1_size,
    const int8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold,
    int* matches_1_2, int* matches_2_1);
void osfm_oracle_twoway_f32(const float* set_1, int set_1_size,
    const float* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int sse3_order,
    int* matches_1_2, int* matches_2_1);

/* Matching::remove_inconsistent_matches, src/mve/sfm/matching.cc:19-36 */
void osfm_oracle_remove_inconsistent(int* matches_1_2, int n1,
    int* matches_2_1, int n2);
/* Matching::count_consistent_matches, src/mve/sfm/matching.cc:39-47 */
int osfm_oracle_count_consistent(const int* matches_1_2, int n1,
    const int* matches_2_1, int n2);
/* Matching::combine_results, src/mve/sfm/matching.cc:50-89.
 * out_1_2 has n1_sift + n1_surf entries, out_2_1 has n2_sift + n2_surf. */
void osfm_oracle_combine_results(
    const int* sift_1_2, int n1_sift, const int* sift_2_1, int n2_sift,
    const int* surf_1_2, int n1_surf, const int* surf_2_1, int n2_surf,
    int* out_1_2, int* out_2_1);

/* ExhaustiveMatching::pairwise_match, src/mve/sfm/exhaustive_matching.cc:115-144.
 * SIFT ratio 0.8, SURF ratio 0.7, no distance threshold
 * (src/mve/sfm/matching_base.h:27-30).  Output sizes as combine_results. */
void osfm_oracle_pairwise_match(
    const uint8_t* sift_1, int n1_sift, const uint8_t* sift_2, int n2_sift,
    const int8_t* surf_1, int n1_surf, const int8_t* surf_2, int n2_surf,
    int* matches_1_2, int* matches_2_1);
/* ExhaustiveMatching::pairwise_match_lowres, exhaustive_matching.cc:147-180 */
int osfm_oracle_pairwise_match_lowres(
    const uint8_t* sift_1, int n1_sift, const uint8_t* sift_2, int n2_sift,
    const int8_t* surf_1, int n1_surf, const int8_t* surf_2, int n2_surf,
    int num_features);

/* Number of OpenMP threads the oracle will use (1 if built without OpenMP). */
int osfm_oracle_num_threads(void);
void osfm_oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif /* OSFM_ORACLE_H */
