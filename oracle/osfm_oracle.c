/*
 * osfm_oracle.c -- CPU restatement of the MVE exhaustive matcher used by
 * OrthoSfM.  TEST INFRASTRUCTURE ONLY (see osfm_oracle.h).  Parity: PINNED
 * against the reference compiled from its own sources (oracle/_ref).
 *
 * Build with -ffp-contract=off and without -ffast-math: the float path must
 * round exactly like the reference's SSE code (separate mul and add).
 *
 * All file:line citations are relative to /root/reference.
 */
#include "osfm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* math::round, src/mve/math/functions.h:70-73; math::clamp, :204-207  */

static float
mve_round (float x)
{
    return x > 0.0f ? floorf(x + 0.5f) : ceilf(x - 0.5f);
}

static float
mve_clamp (float v, float lo, float hi)
{
    return (v < lo ? lo : (v > hi ? hi : v));
}

/* convert_descriptor(Sift::Descriptor), exhaustive_matching.cc:18-27 */
void
osfm_oracle_quantize_sift (const float* desc, int n, uint8_t* out)
{
    for (long i = 0; i < (long)n * 128; ++i)
    {
        float value = mve_clamp(desc[i], 0.0f, 1.0f);
        value = mve_round(value * 255.0f);
        out[i] = (unsigned char)value;
    }
}

/* convert_descriptor(Surf::Descriptor), exhaustive_matching.cc:30-39 */
void
osfm_oracle_quantize_surf (const float* desc, int n, int8_t* out)
{
    for (long i = 0; i < (long)n * 64; ++i)
    {
        float value = mve_clamp(desc[i], -1.0f, 1.0f);
        value = mve_round(value * 127.0f);
        out[i] = (signed char)value;
    }
}

/* ------------------------------------------------------------------ */
/* short_inner_prod<unsigned short>, nearest_neighbor.cc:62-102 (SSE2 branch).
 * Eight 16-bit lanes; lane k accumulates elements k, k+8, k+16, ... with
 * pmullw (low 16 bits of the product) and paddw (wraps mod 2^16), :75-81.
 * The lanes are then added as int (:83-84).  best / second best are kept in
 * the Result's unsigned-short fields, so they are truncated on store (:94,:99)
 * while the comparison uses the untruncated int (:87,:89). */
static void
scan_u8 (const uint8_t* q, const uint8_t* el, int n, int dim,
    uint16_t* b1, uint16_t* b2, int* i1, int* i2)
{
    int const dim_8 = dim / 8;
    for (int j = 0; j < n; ++j, el += dim)
    {
        uint16_t lane[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int i = 0; i < dim_8; ++i)
            for (int k = 0; k < 8; ++k)
                lane[k] = (uint16_t)(lane[k]
                    + (uint16_t)((unsigned)q[8 * i + k] * (unsigned)el[8 * i + k]));
        int ip = 0;
        for (int k = 0; k < 8; ++k)
            ip += (int)lane[k];

        if (ip >= (int)*b2)
        {
            if (ip >= (int)*b1)
            {
                *i2 = *i1;
                *b2 = *b1;
                *i1 = j;
                *b1 = (uint16_t)ip;
            }
            else
            {
                *i2 = j;
                *b2 = (uint16_t)ip;
            }
        }
    }
}

/* short_inner_prod<short>: same loop with signed 16-bit lanes. */
static void
scan_s8 (const int8_t* q, const int8_t* el, int n, int dim,
    int16_t* b1, int16_t* b2, int* i1, int* i2)
{
    int const dim_8 = dim / 8;
    for (int j = 0; j < n; ++j, el += dim)
    {
        uint16_t lane[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int i = 0; i < dim_8; ++i)
            for (int k = 0; k < 8; ++k)
            {
                int prod = (int)q[8 * i + k] * (int)el[8 * i + k];
                lane[k] = (uint16_t)(lane[k] + (uint16_t)prod);
            }
        int ip = 0;
        for (int k = 0; k < 8; ++k)
            ip += (int)(int16_t)lane[k];

        if (ip >= (int)*b2)
        {
            if (ip >= (int)*b1)
            {
                *i2 = *i1;
                *b2 = *b1;
                *i1 = j;
                *b1 = (int16_t)(uint16_t)ip;
            }
            else
            {
                *i2 = j;
                *b2 = (int16_t)(uint16_t)ip;
            }
        }
    }
}

/* float_inner_prod, nearest_neighbor.cc:141-210 */
static void
scan_f32 (const float* q, const float* el, int n, int dim, int sse3_order,
    float* b1, float* b2, int* i1, int* i2)
{
    for (int j = 0; j < n; ++j, el += dim)
    {
        float ip;
        if (sse3_order)
        {
            /* :155-166 -- four partial sums, then hadd twice. */
            int const dim_4 = dim / 4;
            float s[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
            for (int i = 0; i < dim_4; ++i)
                for (int k = 0; k < 4; ++k)
                {
                    float prod = q[4 * i + k] * el[4 * i + k];
                    s[k] = s[k] + prod;
                }
            float h01 = s[0] + s[1];
            float h23 = s[2] + s[3];
            ip = h01 + h23;
        }
        else
        {
            /* :186-191 -- sequential scalar loop. */
            ip = 0.0f;
            for (int i = 0; i < dim; ++i)
            {
                float prod = q[i] * el[i];
                ip = ip + prod;
            }
        }

        if (ip >= *b2)
        {
            if (ip >= *b1)
            {
                *i2 = *i1;
                *b2 = *b1;
                *i1 = j;
                *b1 = ip;
            }
            else
            {
                *i2 = j;
                *b2 = ip;
            }
        }
    }
}

/* NearestNeighbor<unsigned short>::find, nearest_neighbor.cc:242-268 */
void
osfm_oracle_nn_u8 (const uint8_t* query, const uint8_t* elements,
    int num_elements, int dim, osfm_oracle_nn_result* result)
{
    uint16_t b1 = 0, b2 = 0;
    int i1 = 0, i2 = 0;
    scan_u8(query, elements, num_elements, dim, &b1, &b2, &i1, &i2);

    int d1 = 65025 < (int)b1 ? 65025 : (int)b1;
    int d2 = 65025 < (int)b2 ? 65025 : (int)b2;
    d1 = 65025 - d1;
    d2 = 65025 - d2;
    d1 = (32767 < d1 ? 32767 : d1) * 2;
    d2 = (32767 < d2 ? 32767 : d2) * 2;
    result->dist_1st_best = (float)(uint16_t)d1;
    result->dist_2nd_best = (float)(uint16_t)d2;
    result->index_1st_best = i1;
    result->index_2nd_best = i2;
}

/* NearestNeighbor<short>::find, nearest_neighbor.cc:216-238 */
void
osfm_oracle_nn_s8 (const int8_t* query, const int8_t* elements,
    int num_elements, int dim, osfm_oracle_nn_result* result)
{
    int16_t b1 = 0, b2 = 0;
    int i1 = 0, i2 = 0;
    scan_s8(query, elements, num_elements, dim, &b1, &b2, &i1, &i2);

    int d1 = (int)b1 > 0 ? (int)b1 : 0;
    int d2 = (int)b2 > 0 ? (int)b2 : 0;
    d1 = 16129 < d1 ? 16129 : d1;
    d2 = 16129 < d2 ? 16129 : d2;
    d1 = 32258 - 2 * d1;
    d2 = 32258 - 2 * d2;
    result->dist_1st_best = (float)(int16_t)d1;
    result->dist_2nd_best = (float)(int16_t)d2;
    result->index_1st_best = i1;
    result->index_2nd_best = i2;
}

/* NearestNeighbor<float>::find, nearest_neighbor.cc:272-289 */
void
osfm_oracle_nn_f32 (const float* query, const float* elements,
    int num_elements, int dim, int sse3_order, osfm_oracle_nn_result* result)
{
    float b1 = 0.0f, b2 = 0.0f;
    int i1 = 0, i2 = 0;
    scan_f32(query, elements, num_elements, dim, sse3_order,
        &b1, &b2, &i1, &i2);
    float d1 = 2.0f - 2.0f * b1;
    float d2 = 2.0f - 2.0f * b2;
    result->dist_1st_best = 0.0f > d1 ? 0.0f : d1;  /* std::max(0.0f, x) */
    result->dist_2nd_best = 0.0f > d2 ? 0.0f : d2;
    result->index_1st_best = i1;
    result->index_2nd_best = i2;
}

/* ------------------------------------------------------------------ */
/* Matching::oneway_match<T>, matching.h:114-146.  The two tests are
 *   dist_1st > distance_threshold^2            -> no match   (:138)
 *   (float)dist_1st / (float)dist_2nd > ratio^2 -> no match  (:140-143)
 * both in float; 0/0 = NaN compares false and therefore accepts. */
static int
accept (osfm_oracle_nn_result const* r, float sq_lowe, float sq_dist)
{
    if (r->dist_1st_best > sq_dist)
        return 0;
    volatile float ratio = r->dist_1st_best / r->dist_2nd_best;
    if (ratio > sq_lowe)
        return 0;
    return 1;
}

#define ONEWAY_BODY(NN_CALL)                                              \
    for (int i = 0; i < set_1_size; ++i)                                  \
        result[i] = -1;                                                   \
    if (set_1_size == 0 || set_2_size == 0)                               \
        return;                                                           \
    float const sq_lowe = lowe_ratio_threshold * lowe_ratio_threshold;    \
    float const sq_dist = distance_threshold * distance_threshold;        \
    _Pragma("omp parallel for schedule(dynamic, 16)")                     \
    for (int i = 0; i < set_1_size; ++i)                                  \
    {                                                                     \
        osfm_oracle_nn_result r;                                          \
        NN_CALL;                                                          \
        if (accept(&r, sq_lowe, sq_dist))                                 \
            result[i] = r.index_1st_best;                                 \
    }

void
osfm_oracle_oneway_u8 (const uint8_t* set_1, int set_1_size,
    const uint8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int* result)
{
    ONEWAY_BODY(osfm_oracle_nn_u8(set_1 + (long)i * dim, set_2,
        set_2_size, dim, &r))
}

void
osfm_oracle_oneway_s8 (const int8_t* set_1, int set_1_size,
    const int8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int* result)
{
    ONEWAY_BODY(osfm_oracle_nn_s8(set_1 + (long)i * dim, set_2,
        set_2_size, dim, &r))
}

void
osfm_oracle_oneway_f32 (const float* set_1, int set_1_size,
    const float* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int sse3_order,
    int* result)
{
    ONEWAY_BODY(osfm_oracle_nn_f32(set_1 + (long)i * dim, set_2,
        set_2_size, dim, sse3_order, &r))
}

/* Matching::twoway_match<T>, matching.h:148-159 */
void
osfm_oracle_twoway_u8 (const uint8_t* set_1, int set_1_size,
    const uint8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold,
    int* matches_1_2, int* matches_2_1)
{
    osfm_oracle_oneway_u8(set_1, set_1_size, set_2, set_2_size, dim,
        lowe_ratio_threshold, distance_threshold, matches_1_2);
    osfm_oracle_oneway_u8(set_2, set_2_size, set_1, set_1_size, dim,
        lowe_ratio_threshold, distance_threshold, matches_2_1);
}

void
osfm_oracle_twoway_s8 (const int8_t* set_1, int set_1_size,
    const int8_t* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold,
    int* matches_1_2, int* matches_2_1)
{
    osfm_oracle_oneway_s8(set_1, set_1_size, set_2, set_2_size, dim,
        lowe_ratio_threshold, distance_threshold, matches_1_2);
    osfm_oracle_oneway_s8(set_2, set_2_size, set_1, set_1_size, dim,
        lowe_ratio_threshold, distance_threshold, matches_2_1);
}

void
osfm_oracle_twoway_f32 (const float* set_1, int set_1_size,
    const float* set_2, int set_2_size, int dim,
    float lowe_ratio_threshold, float distance_threshold, int sse3_order,
    int* matches_1_2, int* matches_2_1)
{
    osfm_oracle_oneway_f32(set_1, set_1_size, set_2, set_2_size, dim,
        lowe_ratio_threshold, distance_threshold, sse3_order, matches_1_2);
    osfm_oracle_oneway_f32(set_2, set_2_size, set_1, set_1_size, dim,
        lowe_ratio_threshold, distance_threshold, sse3_order, matches_2_1);
}

/* ------------------------------------------------------------------ */
/* Matching::remove_inconsistent_matches, matching.cc:19-36.  Note the
 * second loop reads the already-updated matches_1_2. */
void
osfm_oracle_remove_inconsistent (int* matches_1_2, int n1,
    int* matches_2_1, int n2)
{
    for (int i = 0; i < n1; ++i)
    {
        if (matches_1_2[i] < 0)
            continue;
        if (matches_2_1[matches_1_2[i]] != i)
            matches_1_2[i] = -1;
    }
    for (int i = 0; i < n2; ++i)
    {
        if (matches_2_1[i] < 0)
            continue;
        if (matches_1_2[matches_2_1[i]] != i)
            matches_2_1[i] = -1;
    }
}

/* Matching::count_consistent_matches, matching.cc:39-47 */
int
osfm_oracle_count_consistent (const int* matches_1_2, int n1,
    const int* matches_2_1, int n2)
{
    (void)n2;
    int counter = 0;
    for (int i = 0; i < n1; ++i)
        if (matches_1_2[i] != -1 && matches_2_1[matches_1_2[i]] == i)
            counter++;
    return counter;
}

/* Matching::combine_results, matching.cc:50-89 */
void
osfm_oracle_combine_results (
    const int* sift_1_2, int n1_sift, const int* sift_2_1, int n2_sift,
    const int* surf_1_2, int n1_surf, const int* surf_2_1, int n2_surf,
    int* out_1_2, int* out_2_1)
{
    if (n1_sift > 0) memcpy(out_1_2, sift_1_2, sizeof(int) * n1_sift);
    if (n1_surf > 0) memcpy(out_1_2 + n1_sift, surf_1_2, sizeof(int) * n1_surf);
    if (n2_sift > 0) memcpy(out_2_1, sift_2_1, sizeof(int) * n2_sift);
    if (n2_surf > 0) memcpy(out_2_1 + n2_sift, surf_2_1, sizeof(int) * n2_surf);

    /* "Fix offsets", :78-88: SURF indices are shifted past the other view's
     * SIFT block. */
    int const surf_offset_1 = n1_sift;
    int const surf_offset_2 = n2_sift;
    if (surf_offset_2 > 0)
        for (int i = surf_offset_1; i < n1_sift + n1_surf; ++i)
            if (out_1_2[i] >= 0)
                out_1_2[i] += surf_offset_2;
    if (surf_offset_1 > 0)
        for (int i = surf_offset_2; i < n2_sift + n2_surf; ++i)
            if (out_2_1[i] >= 0)
                out_2_1[i] += surf_offset_1;
}

/* ------------------------------------------------------------------ */
/* MatchingBase::Options defaults, matching_base.h:27-30 */
#define SIFT_RATIO 0.8f
#define SURF_RATIO 0.7f
#define NO_DIST_THRES 3.402823466e+38f /* std::numeric_limits<float>::max() */

/* ExhaustiveMatching::pairwise_match, exhaustive_matching.cc:115-144.
 * A feature type is matched only if view 1 has descriptors of it (:123,:134);
 * otherwise its Result stays empty (both vectors of size 0!), which is what
 * combine_results then concatenates. */
void
osfm_oracle_pairwise_match (
    const uint8_t* sift_1, int n1_sift, const uint8_t* sift_2, int n2_sift,
    const int8_t* surf_1, int n1_surf, const int8_t* surf_2, int n2_surf,
    int* matches_1_2, int* matches_2_1)
{
    int r1_sift = 0, r2_sift = 0, r1_surf = 0, r2_surf = 0;
    int* s12 = NULL; int* s21 = NULL; int* f12 = NULL; int* f21 = NULL;

    if (n1_sift > 0)
    {
        r1_sift = n1_sift; r2_sift = n2_sift;
        s12 = (int*)malloc(sizeof(int) * (r1_sift + 1));
        s21 = (int*)malloc(sizeof(int) * (r2_sift + 1));
        osfm_oracle_twoway_u8(sift_1, n1_sift, sift_2, n2_sift, 128,
            SIFT_RATIO, NO_DIST_THRES, s12, s21);
        osfm_oracle_remove_inconsistent(s12, r1_sift, s21, r2_sift);
    }
    if (n1_surf > 0)
    {
        r1_surf = n1_surf; r2_surf = n2_surf;
        f12 = (int*)malloc(sizeof(int) * (r1_surf + 1));
        f21 = (int*)malloc(sizeof(int) * (r2_surf + 1));
        osfm_oracle_twoway_s8(surf_1, n1_surf, surf_2, n2_surf, 64,
            SURF_RATIO, NO_DIST_THRES, f12, f21);
        osfm_oracle_remove_inconsistent(f12, r1_surf, f21, r2_surf);
    }
    osfm_oracle_combine_results(s12, r1_sift, s21, r2_sift,
        f12, r1_surf, f21, r2_surf, matches_1_2, matches_2_1);
    free(s12); free(s21); free(f12); free(f21);
}

/* ExhaustiveMatching::pairwise_match_lowres, exhaustive_matching.cc:147-180:
 * SIFT only if view 1 has SIFT descriptors, else SURF, else 0. */
int
osfm_oracle_pairwise_match_lowres (
    const uint8_t* sift_1, int n1_sift, const uint8_t* sift_2, int n2_sift,
    const int8_t* surf_1, int n1_surf, const int8_t* surf_2, int n2_surf,
    int num_features)
{
    if (n1_sift > 0)
    {
        int a = num_features < n1_sift ? num_features : n1_sift;
        int b = num_features < n2_sift ? num_features : n2_sift;
        int* m12 = (int*)malloc(sizeof(int) * (a + 1));
        int* m21 = (int*)malloc(sizeof(int) * (b + 1));
        osfm_oracle_twoway_u8(sift_1, a, sift_2, b, 128,
            SIFT_RATIO, NO_DIST_THRES, m12, m21);
        int c = osfm_oracle_count_consistent(m12, a, m21, b);
        free(m12); free(m21);
        return c;
    }
    if (n1_surf > 0)
    {
        int a = num_features < n1_surf ? num_features : n1_surf;
        int b = num_features < n2_surf ? num_features : n2_surf;
        int* m12 = (int*)malloc(sizeof(int) * (a + 1));
        int* m21 = (int*)malloc(sizeof(int) * (b + 1));
        osfm_oracle_twoway_s8(surf_1, a, surf_2, b, 64,
            SURF_RATIO, NO_DIST_THRES, m12, m21);
        int c = osfm_oracle_count_consistent(m12, a, m21, b);
        free(m12); free(m21);
        return c;
    }
    return 0;
}

void
osfm_oracle_set_num_threads (int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int
osfm_oracle_num_threads (void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
