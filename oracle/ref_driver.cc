/*
 * ref_driver.cc -- thin extern "C" driver around the UNMODIFIED reference
 * matcher.  TEST INFRASTRUCTURE ONLY.
 *
 * This file contains no reference code.  It is compiled together with the
 * reference's own sources where they lie under /root/reference (see
 * oracle/Makefile, target _ref) into oracle/_ref/libosfm_ref.so, which is used
 * (a) to pin the C restatement in osfm_oracle.c, (b) to generate the golden
 * fixtures in tests/golden/, and (c) as the "reference" CPU baseline timed by
 * bench.py.  oracle/_ref/ is git-ignored; the prebuilt .so travels to the GPU
 * box with the gpurun snapshot.
 *
 * Reference entry points bound here (paths relative to /root/reference):
 *   sfm::Matching::twoway_match<T>            src/mve/sfm/matching.h:148-159
 *   sfm::Matching::oneway_match<T>            src/mve/sfm/matching.h:114-146
 *   sfm::Matching::remove_inconsistent_matches src/mve/sfm/matching.cc:19-36
 *   sfm::Matching::count_consistent_matches   src/mve/sfm/matching.cc:39-47
 *   sfm::Matching::combine_results            src/mve/sfm/matching.cc:50-89
 *   sfm::NearestNeighbor<T>::find             src/mve/sfm/nearest_neighbor.cc:216-289
 *   sfm::ExhaustiveMatching::{init,pairwise_match,pairwise_match_lowres}
 *                                             src/mve/sfm/exhaustive_matching.cc:56,115,147
 *   sfm::Sift::process                        src/mve/sfm/sift.cc (fixture producer only)
 *   sfm::bundler::Tracks::compute             src/mve/sfm/bundler_tracks.cc:47-146
 *   sfm::fundamental_8_point, enforce_fundamental_constraints, sampson_distance
 *                                             src/mve/sfm/fundamental.cc:78-126, 225-247
 *   sfm::RansacFundamental::estimate          src/mve/sfm/ransac_fundamental.cc:26-105
 *   sfm::bundler::Matching::init / compute    src/mve/sfm/bundler_matching.cc:45-220
 *   math::matrix_svd                          src/mve/math/matrix_svd.h
 *   sfm::bundler::save_prebundle_to_file / load_prebundle_from_file
 *                                             src/mve/sfm/bundler_common.cc:56-190
 */
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "sfm/bundler_common.h"
#include "sfm/bundler_matching.h"
#include "sfm/bundler_tracks.h"
#include "sfm/exhaustive_matching.h"
#include "sfm/fundamental.h"
#include "sfm/ransac_fundamental.h"
#include "math/matrix_svd.h"
#include "sfm/matching.h"
#include "sfm/nearest_neighbor.h"
#include "sfm/sift.h"
#include "util/aligned_memory.h"

namespace
{
    template <typename T, typename S>
    util::AlignedMemory<T, 16>
    widen (S const* src, std::size_t count)
    {
        util::AlignedMemory<T, 16> out(count + 8);
        for (std::size_t i = 0; i < count; ++i)
            out[i] = static_cast<T>(src[i]);
        return out;
    }

    void
    copy_out (std::vector<int> const& v, int* dst)
    {
        if (!v.empty())
            std::memcpy(dst, v.data(), sizeof(int) * v.size());
    }

    sfm::Matching::Options
    make_opts (int dim, float ratio, float dist)
    {
        sfm::Matching::Options o;
        o.descriptor_length = dim;
        o.lowe_ratio_threshold = ratio;
        o.distance_threshold = dist;
        return o;
    }
}

extern "C" {

struct osfm_ref_nn_result
{
    float dist_1st_best;
    float dist_2nd_best;
    int index_1st_best;
    int index_2nd_best;
};

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use the cores it is given */
void
osfm_ref_set_num_threads (int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int
osfm_ref_num_threads (void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- NearestNeighbor<T>::find ------------------------------------- */

void
osfm_ref_nn_u8 (const uint8_t* query, const uint8_t* elements, int n, int dim,
    osfm_ref_nn_result* out)
{
    auto q = widen<unsigned short>(query, dim);
    auto e = widen<unsigned short>(elements, (std::size_t)n * dim);
    sfm::NearestNeighbor<unsigned short> nn;
    nn.set_elements(e.data());
    nn.set_num_elements(n);
    nn.set_element_dimensions(dim);
    sfm::NearestNeighbor<unsigned short>::Result r;
    nn.find(q.data(), &r);
    out->dist_1st_best = static_cast<float>(r.dist_1st_best);
    out->dist_2nd_best = static_cast<float>(r.dist_2nd_best);
    out->index_1st_best = r.index_1st_best;
    out->index_2nd_best = r.index_2nd_best;
}

void
osfm_ref_nn_s8 (const int8_t* query, const int8_t* elements, int n, int dim,
    osfm_ref_nn_result* out)
{
    auto q = widen<short>(query, dim);
    auto e = widen<short>(elements, (std::size_t)n * dim);
    sfm::NearestNeighbor<short> nn;
    nn.set_elements(e.data());
    nn.set_num_elements(n);
    nn.set_element_dimensions(dim);
    sfm::NearestNeighbor<short>::Result r;
    nn.find(q.data(), &r);
    out->dist_1st_best = static_cast<float>(r.dist_1st_best);
    out->dist_2nd_best = static_cast<float>(r.dist_2nd_best);
    out->index_1st_best = r.index_1st_best;
    out->index_2nd_best = r.index_2nd_best;
}

void
osfm_ref_nn_f32 (const float* query, const float* elements, int n, int dim,
    osfm_ref_nn_result* out)
{
    auto q = widen<float>(query, dim);
    auto e = widen<float>(elements, (std::size_t)n * dim);
    sfm::NearestNeighbor<float> nn;
    nn.set_elements(e.data());
    nn.set_num_elements(n);
    nn.set_element_dimensions(dim);
    sfm::NearestNeighbor<float>::Result r;
    nn.find(q.data(), &r);
    out->dist_1st_best = r.dist_1st_best;
    out->dist_2nd_best = r.dist_2nd_best;
    out->index_1st_best = r.index_1st_best;
    out->index_2nd_best = r.index_2nd_best;
}

/* ---- Matching::twoway_match<T> ------------------------------------ */

void
osfm_ref_twoway_u8 (const uint8_t* set_1, int n1, const uint8_t* set_2, int n2,
    int dim, float ratio, float dist, int* m12, int* m21)
{
    auto a = widen<unsigned short>(set_1, (std::size_t)n1 * dim);
    auto b = widen<unsigned short>(set_2, (std::size_t)n2 * dim);
    sfm::Matching::Result r;
    sfm::Matching::twoway_match(make_opts(dim, ratio, dist),
        a.data(), n1, b.data(), n2, &r);
    copy_out(r.matches_1_2, m12);
    copy_out(r.matches_2_1, m21);
}

void
osfm_ref_twoway_s8 (const int8_t* set_1, int n1, const int8_t* set_2, int n2,
    int dim, float ratio, float dist, int* m12, int* m21)
{
    auto a = widen<short>(set_1, (std::size_t)n1 * dim);
    auto b = widen<short>(set_2, (std::size_t)n2 * dim);
    sfm::Matching::Result r;
    sfm::Matching::twoway_match(make_opts(dim, ratio, dist),
        a.data(), n1, b.data(), n2, &r);
    copy_out(r.matches_1_2, m12);
    copy_out(r.matches_2_1, m21);
}

void
osfm_ref_twoway_f32 (const float* set_1, int n1, const float* set_2, int n2,
    int dim, float ratio, float dist, int* m12, int* m21)
{
    auto a = widen<float>(set_1, (std::size_t)n1 * dim);
    auto b = widen<float>(set_2, (std::size_t)n2 * dim);
    sfm::Matching::Result r;
    sfm::Matching::twoway_match(make_opts(dim, ratio, dist),
        a.data(), n1, b.data(), n2, &r);
    copy_out(r.matches_1_2, m12);
    copy_out(r.matches_2_1, m21);
}

/* One LARGE pair on all cores: the reference's own Matching::oneway_match (matching.h:114-146)
 * called on chunks of the query rows of each direction -- a query's scan of the other set does not
 * depend on the other queries (matching.h:128-145), so twoway_match (matching.h:148-159) is the
 * concatenation of the chunks' results -- then remove_inconsistent_matches.  Returns the number of
 * consistent matches; m12 / m21 receive the filtered vectors, *digest the FNV-1a digest of the
 * correspondence list as osfm_ref_match_pairs_u8_digest computes it. */
long
osfm_ref_match_large_pair_u8 (const uint8_t* set_1, int n1, const uint8_t* set_2, int n2,
    float ratio, int* m12, int* m21, unsigned long long* digest)
{
    auto a = widen<unsigned short>(set_1, (std::size_t)n1 * 128);
    auto b = widen<unsigned short>(set_2, (std::size_t)n2 * 128);
    sfm::Matching::Options const opts = make_opts(128, ratio, std::numeric_limits<float>::max());
    sfm::Matching::Result r;
    r.matches_1_2.assign(n1, -1);
    r.matches_2_1.assign(n2, -1);
    int const chunk = 512;
    int const c1 = (n1 + chunk - 1) / chunk, c2 = (n2 + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic)
    for (int c = 0; c < c1 + c2; ++c)
    {
        bool const fwd = c < c1;
        int const row0 = (fwd ? c : c - c1) * chunk;
        int const rows = std::min(chunk, (fwd ? n1 : n2) - row0);
        std::vector<int> part;
        if (fwd)
            sfm::Matching::oneway_match(opts, a.data() + (std::size_t)row0 * 128, rows, b.data(), n2, &part);
        else
            sfm::Matching::oneway_match(opts, b.data() + (std::size_t)row0 * 128, rows, a.data(), n1, &part);
        std::copy(part.begin(), part.end(), (fwd ? r.matches_1_2 : r.matches_2_1).begin() + row0);
    }
    sfm::Matching::remove_inconsistent_matches(&r);
    unsigned long long h = 1469598103934665603ull;
    long c = 0;
    for (std::size_t i = 0; i < r.matches_1_2.size(); ++i)
    {
        if (r.matches_1_2[i] < 0)
            continue;
        int const rec[2] = { (int)i, r.matches_1_2[i] };
        unsigned char const* bytes = reinterpret_cast<unsigned char const*>(rec);
        for (int k = 0; k < 8; ++k) { h ^= bytes[k]; h *= 1099511628211ull; }
        ++c;
    }
    if (m12 != nullptr) copy_out(r.matches_1_2, m12);
    if (m21 != nullptr) copy_out(r.matches_2_1, m21);
    if (digest != nullptr) *digest = h;
    return c;
}

/* ---- filters ------------------------------------------------------ */

void
osfm_ref_remove_inconsistent (int* m12, int n1, int* m21, int n2)
{
    sfm::Matching::Result r;
    r.matches_1_2.assign(m12, m12 + n1);
    r.matches_2_1.assign(m21, m21 + n2);
    sfm::Matching::remove_inconsistent_matches(&r);
    copy_out(r.matches_1_2, m12);
    copy_out(r.matches_2_1, m21);
}

int
osfm_ref_count_consistent (const int* m12, int n1, const int* m21, int n2)
{
    sfm::Matching::Result r;
    r.matches_1_2.assign(m12, m12 + n1);
    r.matches_2_1.assign(m21, m21 + n2);
    return sfm::Matching::count_consistent_matches(r);
}

void
osfm_ref_combine_results (
    const int* sift_1_2, int n1_sift, const int* sift_2_1, int n2_sift,
    const int* surf_1_2, int n1_surf, const int* surf_2_1, int n2_surf,
    int* out_1_2, int* out_2_1)
{
    sfm::Matching::Result a, b, c;
    a.matches_1_2.assign(sift_1_2, sift_1_2 + n1_sift);
    a.matches_2_1.assign(sift_2_1, sift_2_1 + n2_sift);
    b.matches_1_2.assign(surf_1_2, surf_1_2 + n1_surf);
    b.matches_2_1.assign(surf_2_1, surf_2_1 + n2_surf);
    sfm::Matching::combine_results(a, b, &c);
    copy_out(c.matches_1_2, out_1_2);
    copy_out(c.matches_2_1, out_2_1);
}

/* ---- the reference's timed unit: twoway + remove_inconsistent over a
 *      list of pairs, OpenMP over pairs like bundler_matching.cc:74 ---- */

/* views: `num_views` pointers to n_i x 128 byte descriptors.  pairs: 2*npairs
 * view ids.  Returns the total number of consistent matches; if counts != NULL
 * it receives the per-pair consistent-match count. */
long
osfm_ref_match_pairs_u8 (const uint8_t* const* views, const int* sizes,
    int num_views, const int* pairs, int npairs, float ratio, int* counts)
{
    std::vector<util::AlignedMemory<unsigned short, 16>> wide(num_views);
    std::vector<char> used(num_views, 0);
    for (int p = 0; p < 2 * npairs; ++p)
        used[pairs[p]] = 1;
    for (int v = 0; v < num_views; ++v)
        if (used[v])
            wide[v] = widen<unsigned short>(views[v], (std::size_t)sizes[v] * 128);

    long total = 0;
#pragma omp parallel for schedule(dynamic) reduction(+:total)
    for (int p = 0; p < npairs; ++p)
    {
        int const v1 = pairs[2 * p + 0];
        int const v2 = pairs[2 * p + 1];
        sfm::Matching::Result r;
        sfm::Matching::twoway_match(make_opts(128, ratio,
            std::numeric_limits<float>::max()),
            wide[v1].data(), sizes[v1], wide[v2].data(), sizes[v2], &r);
        sfm::Matching::remove_inconsistent_matches(&r);
        int const c = sfm::Matching::count_consistent_matches(r);
        if (counts != nullptr)
            counts[p] = c;
        total += c;
    }
    return total;
}

/* Same, returning a digest of every pair's filtered result instead of just its count:
 * digest[p] = FNV-1a (64 bit) over the int32 pairs (i, matches_1_2[i]) of the surviving entries
 * in ascending i -- i.e. over the correspondence list bundler_matching.cc:178-192 builds.  Lets a
 * test hold ALL lists of a large pair set against the reference without moving them. */
long
osfm_ref_match_pairs_u8_digest (const uint8_t* const* views, const int* sizes,
    int num_views, const int* pairs, int npairs, float ratio, int* counts, unsigned long long* digest)
{
    std::vector<util::AlignedMemory<unsigned short, 16>> wide(num_views);
    std::vector<char> used(num_views, 0);
    for (int p = 0; p < 2 * npairs; ++p)
        used[pairs[p]] = 1;
    for (int v = 0; v < num_views; ++v)
        if (used[v])
            wide[v] = widen<unsigned short>(views[v], (std::size_t)sizes[v] * 128);

    long total = 0;
#pragma omp parallel for schedule(dynamic) reduction(+:total)
    for (int p = 0; p < npairs; ++p)
    {
        int const v1 = pairs[2 * p + 0];
        int const v2 = pairs[2 * p + 1];
        sfm::Matching::Result r;
        sfm::Matching::twoway_match(make_opts(128, ratio,
            std::numeric_limits<float>::max()),
            wide[v1].data(), sizes[v1], wide[v2].data(), sizes[v2], &r);
        sfm::Matching::remove_inconsistent_matches(&r);
        unsigned long long h = 1469598103934665603ull;
        int c = 0;
        for (std::size_t i = 0; i < r.matches_1_2.size(); ++i)
        {
            if (r.matches_1_2[i] < 0)
                continue;
            int const rec[2] = { (int)i, r.matches_1_2[i] };
            unsigned char const* b = reinterpret_cast<unsigned char const*>(rec);
            for (int k = 0; k < 8; ++k) { h ^= b[k]; h *= 1099511628211ull; }
            ++c;
        }
        if (counts != nullptr)
            counts[p] = c;
        if (digest != nullptr)
            digest[p] = h;
        total += c;
    }
    return total;
}

/* ---- ExhaustiveMatching (the MatchingBase plugin) ------------------ */

struct osfm_ref_exhaustive
{
    sfm::ExhaustiveMatching matcher;
    sfm::bundler::ViewportList viewports;
};

osfm_ref_exhaustive*
osfm_ref_exhaustive_create (int num_views)
{
    osfm_ref_exhaustive* h = new osfm_ref_exhaustive();
    h->viewports.resize(num_views);
    return h;
}

/* float descriptors exactly as Sift/Surf produce them (n x 128 / n x 64) */
void
osfm_ref_exhaustive_set_view (osfm_ref_exhaustive* h, int view,
    const float* sift, int n_sift, const float* surf, int n_surf)
{
    sfm::FeatureSet& fs = h->viewports[view].features;
    fs.sift_descriptors.resize(n_sift);
    for (int i = 0; i < n_sift; ++i)
    {
        sfm::Sift::Descriptor& d = fs.sift_descriptors[i];
        d.x = d.y = d.scale = d.orientation = 0.0f;
        std::copy(sift + (std::size_t)i * 128, sift + (std::size_t)(i + 1) * 128,
            d.data.begin());
    }
    fs.surf_descriptors.resize(n_surf);
    for (int i = 0; i < n_surf; ++i)
    {
        sfm::Surf::Descriptor& d = fs.surf_descriptors[i];
        d.x = d.y = d.scale = d.orientation = 0.0f;
        std::copy(surf + (std::size_t)i * 64, surf + (std::size_t)(i + 1) * 64,
            d.data.begin());
    }
}

void
osfm_ref_exhaustive_init (osfm_ref_exhaustive* h)
{
    h->matcher.init(&h->viewports);
}

/* Returns sizes through n12/n21; buffers must hold the combined lengths. */
void
osfm_ref_exhaustive_pairwise_match (osfm_ref_exhaustive* h, int v1, int v2,
    int* m12, int* n12, int* m21, int* n21)
{
    sfm::Matching::Result r;
    h->matcher.pairwise_match(v1, v2, &r);
    *n12 = (int)r.matches_1_2.size();
    *n21 = (int)r.matches_2_1.size();
    copy_out(r.matches_1_2, m12);
    copy_out(r.matches_2_1, m21);
}

int
osfm_ref_exhaustive_pairwise_match_lowres (osfm_ref_exhaustive* h,
    int v1, int v2, int num_features)
{
    return h->matcher.pairwise_match_lowres(v1, v2, (std::size_t)num_features);
}

void
osfm_ref_exhaustive_destroy (osfm_ref_exhaustive* h)
{
    delete h;
}

/* ---- fixture producer: the reference's own SIFT on a P5 PGM image --- */

/* Runs sfm::Sift with default options on an 8-bit grey image, sorts by scale
 * descending like FeatureSet::compute_sift (src/mve/sfm/feature_set.cc:58-78)
 * and writes up to max_desc descriptors (128 floats each).  Returns the count. */
int
osfm_ref_sift_gray8 (const uint8_t* pixels, int width, int height,
    float* out_desc, int max_desc)
{
    mve::ByteImage::Ptr img = mve::ByteImage::create(width, height, 1);
    std::memcpy(img->get_data_pointer(), pixels, (std::size_t)width * height);
    sfm::Sift::Options opts;
    sfm::Sift sift(opts);
    sift.set_image(img);
    sift.process();
    sfm::Sift::Descriptors descr = sift.get_descriptors();
    std::sort(descr.begin(), descr.end(),
        [] (sfm::Sift::Descriptor const& a, sfm::Sift::Descriptor const& b)
        { return a.scale > b.scale; });
    int const n = std::min<int>((int)descr.size(), max_desc);
    for (int i = 0; i < n; ++i)
        std::copy(descr[i].data.begin(), descr[i].data.end(),
            out_desc + (std::size_t)i * 128);
    return n;
}

/* ---- bundler::Tracks::compute (the consumer of the pairwise match lists) ------ */

/* features[v] = number of features of view v; pair p = views (pair_views[2p],
 * pair_views[2p+1]) with the correspondences ij[2*off[p]] .. ij[2*off[p+1]).
 * track_ids receives, view after view, Viewport::track_ids as compute() leaves them
 * (bundler_tracks.cc:58,148-203).  Returns the number of tracks. */
int
osfm_ref_tracks_compute (int num_views, const int* features, int npairs,
    const int* pair_views, const long long* off, const int* ij, int* track_ids)
{
    sfm::bundler::ViewportList viewports(num_views);
    for (int v = 0; v < num_views; ++v)
    {
        viewports[v].features.positions.resize(features[v]);
        viewports[v].features.colors.resize(features[v], math::Vec3uc(0, 0, 0));
    }
    sfm::bundler::PairwiseMatching matching(npairs);
    for (int p = 0; p < npairs; ++p)
    {
        matching[p].view_1_id = pair_views[2 * p + 0];
        matching[p].view_2_id = pair_views[2 * p + 1];
        for (long long k = off[p]; k < off[p + 1]; ++k)
            matching[p].matches.push_back(std::make_pair(ij[2 * k], ij[2 * k + 1]));
    }
    sfm::bundler::TrackList tracks;
    sfm::bundler::Tracks::Options topts;
    sfm::bundler::Tracks computer(topts);
    computer.compute(matching, &viewports, &tracks);
    std::size_t at = 0;
    for (int v = 0; v < num_views; ++v)
        for (int f = 0; f < features[v]; ++f)
            track_ids[at++] = viewports[v].track_ids[f];
    return static_cast<int>(tracks.size());
}

/* ---- prebundle file through the reference's own writer / reader -------------- */

/* Writes `path` with save_prebundle_to_file.  positions: 2 floats, colors: 3 bytes per
 * feature, concatenated over the views. */
int
osfm_ref_save_prebundle (const char* path, int num_views, const int* features,
    const float* positions, const unsigned char* colors, int npairs,
    const int* pair_views, const long long* off, const int* ij)
{
    sfm::bundler::ViewportList viewports(num_views);
    std::size_t at = 0;
    for (int v = 0; v < num_views; ++v)
    {
        for (int f = 0; f < features[v]; ++f, ++at)
        {
            viewports[v].features.positions.push_back(
                math::Vec2f(positions[2 * at], positions[2 * at + 1]));
            viewports[v].features.colors.push_back(
                math::Vec3uc(colors[3 * at], colors[3 * at + 1], colors[3 * at + 2]));
        }
    }
    sfm::bundler::PairwiseMatching matching(npairs);
    for (int p = 0; p < npairs; ++p)
    {
        matching[p].view_1_id = pair_views[2 * p + 0];
        matching[p].view_2_id = pair_views[2 * p + 1];
        for (long long k = off[p]; k < off[p + 1]; ++k)
            matching[p].matches.push_back(std::make_pair(ij[2 * k], ij[2 * k + 1]));
    }
    try { sfm::bundler::save_prebundle_to_file(viewports, matching, path); }
    catch (...) { return -1; }
    return 0;
}

/* Reads `path` with load_prebundle_from_file and returns a digest the test compares:
 * counts[0..3] = views, features (positions), pairs, matches; sums[0..2] = sum of the
 * positions, of the color bytes, of i*31 + j*17 + view ids over the matches. */
int
osfm_ref_load_prebundle_digest (const char* path, long long* counts, double* sums)
{
    sfm::bundler::ViewportList viewports;
    sfm::bundler::PairwiseMatching matching;
    try { sfm::bundler::load_prebundle_from_file(path, &viewports, &matching); }
    catch (...) { return -1; }
    counts[0] = viewports.size(); counts[1] = 0; counts[2] = matching.size(); counts[3] = 0;
    sums[0] = sums[1] = sums[2] = 0.0;
    for (std::size_t v = 0; v < viewports.size(); ++v)
    {
        counts[1] += viewports[v].features.positions.size();
        for (std::size_t f = 0; f < viewports[v].features.positions.size(); ++f)
            sums[0] += viewports[v].features.positions[f][0] + 2.0 * viewports[v].features.positions[f][1];
        for (std::size_t f = 0; f < viewports[v].features.colors.size(); ++f)
            sums[1] += viewports[v].features.colors[f][0] + 3.0 * viewports[v].features.colors[f][1]
                + 5.0 * viewports[v].features.colors[f][2];
    }
    for (std::size_t p = 0; p < matching.size(); ++p)
    {
        counts[3] += matching[p].matches.size();
        sums[2] += 1000.0 * matching[p].view_1_id + 7.0 * matching[p].view_2_id;
        for (std::size_t k = 0; k < matching[p].matches.size(); ++k)
            sums[2] += 31.0 * matching[p].matches[k].first + 17.0 * matching[p].matches[k].second;
    }
    return 0;
}

/* ---- RANSAC for the fundamental matrix ---------------------------------------- */

/* p1 / p2: 8 points x0 y0 x1 y1 ...; F row-major. */
void
osfm_ref_fundamental (const double* p1, const double* p2, double* F)
{
    sfm::Eight2DPoints a, b;
    for (int i = 0; i < 8; ++i)
    {
        a(0, i) = p1[2 * i]; a(1, i) = p1[2 * i + 1]; a(2, i) = 1.0;
        b(0, i) = p2[2 * i]; b(1, i) = p2[2 * i + 1]; b(2, i) = 1.0;
    }
    sfm::FundamentalMatrix f;
    sfm::fundamental_8_point(a, b, &f);
    sfm::enforce_fundamental_constraints(&f);
    std::copy(f.begin(), f.end(), F);
}

double
osfm_ref_sampson (const double* F, const double* m)
{
    sfm::FundamentalMatrix f;
    std::copy(F, F + 9, f.begin());
    sfm::Correspondence2D2D c;
    c.p1[0] = m[0]; c.p1[1] = m[1]; c.p2[0] = m[2]; c.p2[1] = m[3];
    return sfm::sampson_distance(f, c);
}

void
osfm_ref_svd9 (const double* a, double* s, double* v)
{
    math::matrix_svd<double>(a, 9, 9, nullptr, s, v);
}

void
osfm_ref_svd3 (const double* a, double* u, double* s, double* v)
{
    math::matrix_svd<double>(a, 3, 3, u, s, v);
}

/* RansacFundamental::estimate on n matches (x1 y1 x2 y2 as doubles) after std::srand(seed);
 * seed < 0 leaves the sequence where it is.  Returns the number of inliers. */
int
osfm_ref_ransac (const double* matches, int n, int iterations, double threshold, int seed,
    int* inliers, double* F)
{
    sfm::Correspondences2D2D m(n);
    for (int i = 0; i < n; ++i)
    {
        m[i].p1[0] = matches[4 * i + 0]; m[i].p1[1] = matches[4 * i + 1];
        m[i].p2[0] = matches[4 * i + 2]; m[i].p2[1] = matches[4 * i + 3];
    }
    if (seed >= 0)
        std::srand(seed);
    sfm::RansacFundamental::Options opts;
    opts.max_iterations = iterations;
    opts.threshold = threshold;
    sfm::RansacFundamental ransac(opts);
    sfm::RansacFundamental::Result result;
    ransac.estimate(m, &result);
    std::copy(result.inliers.begin(), result.inliers.end(), inliers);
    if (!result.inliers.empty())
        std::copy(result.fundamental.begin(), result.fundamental.end(), F);
    return static_cast<int>(result.inliers.size());
}

/* ---- bundler::Matching: the whole two-view stage ------------------------------ */

/* Runs bundler::Matching (exhaustive matcher) over the viewports of h -- descriptors as set
 * with osfm_ref_exhaustive_set_view, positions (2 floats per feature, sift then surf, view
 * after view) given here -- single-threaded, as the reference's shipped build does (no
 * OpenMP), after std::srand(seed) when seed >= 0.  opts: use_lowres_matching,
 * num_lowres_features, min_lowres_matches, min_feature_matches, min_matching_inliers,
 * match_num_previous_frames, ransac max_iterations; threshold separately.
 * out_pairs receives (view_1, view_2) per accepted pair in the order compute() appended them,
 * out_off the list offsets, out_ij the lists.  Returns the number of accepted pairs, or -1
 * when a capacity is too small. */
int
osfm_ref_bundler_compute (osfm_ref_exhaustive* h, const float* positions, const int* opts,
    double threshold, int seed, int* out_pairs, int cap_pairs, long long* out_off,
    int* out_ij, long long cap_ij)
{
    std::size_t at = 0;
    for (std::size_t v = 0; v < h->viewports.size(); ++v)
    {
        sfm::FeatureSet& fs = h->viewports[v].features;
        std::size_t const n = fs.sift_descriptors.size() + fs.surf_descriptors.size();
        fs.positions.resize(n);
        for (std::size_t i = 0; i < n; ++i, ++at)
            fs.positions[i] = math::Vec2f(positions[2 * at], positions[2 * at + 1]);
    }
    sfm::bundler::Matching::Options o;
    o.use_lowres_matching = opts[0] != 0;
    o.num_lowres_features = opts[1];
    o.min_lowres_matches = opts[2];
    o.min_feature_matches = opts[3];
    o.min_matching_inliers = opts[4];
    o.match_num_previous_frames = opts[5];
    o.ransac_opts.max_iterations = opts[6];
    o.ransac_opts.threshold = threshold;
    o.ransac_opts.verbose_output = false;
    o.matcher_type = sfm::bundler::Matching::MATCHER_EXHAUSTIVE;

    int const threads = omp_get_max_threads();
    omp_set_num_threads(1);
    if (seed >= 0)
        std::srand(seed);
    sfm::bundler::PairwiseMatching result;
    {
        /* compute() reports progress on std::cout; keep the test output clean */
        std::streambuf* old = std::cout.rdbuf(nullptr);
        sfm::bundler::Matching matching(o);
        matching.init(&h->viewports);
        matching.compute(&result);
        std::cout.rdbuf(old);
    }
    omp_set_num_threads(threads);

    if ((int)result.size() > cap_pairs)
        return -1;
    out_off[0] = 0;
    for (std::size_t p = 0; p < result.size(); ++p)
    {
        out_pairs[2 * p] = result[p].view_1_id;
        out_pairs[2 * p + 1] = result[p].view_2_id;
        long long const k = (long long)result[p].matches.size();
        if (out_off[p] + k > cap_ij)
            return -1;
        for (long long i = 0; i < k; ++i)
        {
            out_ij[2 * (out_off[p] + i)] = result[p].matches[i].first;
            out_ij[2 * (out_off[p] + i) + 1] = result[p].matches[i].second;
        }
        out_off[p + 1] = out_off[p] + k;
    }
    return (int)result.size();
}

} /* extern "C" */
