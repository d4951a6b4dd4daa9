/*
 * osfm_match.h -- C ABI of the B200-native exhaustive pairwise feature matcher
 * for OrthoSfM (libosfm_match.so).
 *
 * This is the drop-in boundary for the one hot path this project replaces: the
 * work done behind sfm::MatchingBase (reference: src/mve/sfm/matching_base.h:22-55)
 * by sfm::ExhaustiveMatching (src/mve/sfm/exhaustive_matching.{h,cc}).  A C++
 * subclass of MatchingBase that forwards to these entry points is shipped as
 * orthosfm_b200/csrc/gpu_exhaustive_matching.h; INTEGRATION.md shows the two-line
 * change in src/mve/sfm/bundler_matching.cc:31-41 that selects it.
 *
 * Conventions
 *   - plain pointers and sizes only, no C++ or torch types;
 *   - every function returns 0 on success or a negative osfm_status; it never
 *     throws and never exits the process (unlike CudaSift's safeCall,
 *     src/cuda_sift/cudautils.h:15-21).  osfm_match_last_error() gives the text;
 *   - a handle is internally serialised by a mutex, so the const, re-entrant
 *     pairwise_match() contract of MatchingBase holds when the reference calls
 *     it from its OpenMP pair loop (src/mve/sfm/bundler_matching.cc:74);
 *   - there is NO CPU fallback: without a CUDA device of compute capability 10.x
 *     osfm_match_create() fails with OSFM_ERR_NO_DEVICE.
 *
 * All file:line citations are relative to the reference tree (/root/reference).
 */
#ifndef OSFM_MATCH_H
#define OSFM_MATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSFM_MATCH_ABI_VERSION 2

typedef enum {
    OSFM_OK = 0,
    OSFM_ERR_INVALID_ARGUMENT = -1, /* std::invalid_argument in the reference  */
    OSFM_ERR_NO_DEVICE = -2,        /* no sm_100 device / CUDA unavailable       */
    OSFM_ERR_CUDA = -3,             /* a CUDA call failed, see last_error        */
    OSFM_ERR_STATE = -4,            /* call order violated (e.g. match before commit) */
    OSFM_ERR_OUT_OF_MEMORY = -5,
    OSFM_ERR_INTERNAL = -6,         /* self-check failed inside a kernel          */
    OSFM_ERR_IO = -7                /* a file could not be read or written        */
} osfm_status;

/* Descriptor kinds (reference: exhaustive_matching.h:46-48). */
typedef enum {
    OSFM_KIND_SIFT_U8 = 0, /* 128 x unsigned byte, Matching::twoway_match<unsigned short> */
    OSFM_KIND_SURF_S8 = 1  /*  64 x signed byte,   Matching::twoway_match<short>          */
} osfm_kind;

/* Mirrors MatchingBase::Options (matching_base.h:25-31): per feature type the
 * Lowe ratio and the distance threshold of Matching::Options (matching.h:29-50).
 * The descriptor lengths are fixed (128 / 64). */
typedef struct {
    int   device;                   /* CUDA device ordinal                      */
    float sift_lowe_ratio;          /* 0.8f                                     */
    float sift_distance_threshold;  /* FLT_MAX                                  */
    float surf_lowe_ratio;          /* 0.7f                                     */
    float surf_distance_threshold;  /* FLT_MAX                                  */
    int   reserved[4];              /* must be zero                             */
} osfm_match_config;

typedef struct osfm_matcher osfm_matcher;

/* Fills cfg with the reference defaults (matching_base.h:27-30), device 0. */
void osfm_match_default_config(osfm_match_config* cfg);

int osfm_match_abi_version(void);

/* Replaces `new ExhaustiveMatching()` (bundler_matching.cc:34). */
int osfm_match_create(const osfm_match_config* cfg, osfm_matcher** out);

/* One matcher over several GPUs of the box, in one process: what the reference's single call
 * bundler::Matching::compute (bundler_matching.cc:57-133, reached from
 * src/matching/matching_mve.cpp:413-415) needs to use more than one device.  devices[0] is the
 * primary (cfg->device is ignored): views are staged there, and osfm_match_commit /
 * osfm_match_commit_device replicate the descriptor pool to the other devices with one
 * ncclBroadcast per feature kind over NVLink (libnccl.so.2 is loaded at run time; a
 * single-device handle never needs it).  The batched entry points -- osfm_match_pairs,
 * osfm_match_pairs_compact, osfm_match_two_view_candidates, osfm_match_two_view -- cut the
 * caller's pair list into one contiguous range per device by cost sum(n1 * n2) (the unit the
 * reference's OpenMP loop parallelises over, bundler_matching.cc:74), run every range on its
 * device from a host thread of its own, and deliver the results into the caller's host
 * buffers in pair order; there is no collective on the data path.  Everything else (single
 * pairs, RANSAC, tracks) runs on the primary.  Results are identical to a single-device
 * handle's.  Overlapped staging is accepted but commit waits for the copies. */
int osfm_match_create_multi(const osfm_match_config* cfg, const int* devices, int num_devices,
    osfm_matcher** out);
/* 1 for osfm_match_create handles. */
int osfm_match_num_devices(const osfm_matcher* m);
void osfm_match_destroy(osfm_matcher* m);
const char* osfm_match_last_error(const osfm_matcher* m);

/* ---- staging: replaces ExhaustiveMatching::init (exhaustive_matching.cc:56-112).
 * begin / set_view x N / commit together are one init() call.  The host-to-device copies
 * are asynchronous: the buffers handed to set_view must stay valid until
 * osfm_match_commit() returns, after which the caller may free them -- exactly where
 * bundler::Matching::init frees the float descriptors, right after matcher->init
 * (bundler_matching.cc:53-55).  Views staged in increasing view_id order land directly in
 * their final place in the pool (no second copy). ------------------------------------ */

/* Declares the number of views (ViewportList::size()).  Resets the handle. */
int osfm_match_begin(osfm_matcher* m, int num_views);

/* Float descriptors exactly as Sift / Surf produce them; quantised on the device
 * like convert_descriptor (exhaustive_matching.cc:18-39).  `sift` is
 * n_sift x 128 floats, row stride sift_stride floats (>= 128; the reference's
 * Sift::Descriptor is 132 floats, sift.h:137-149, so pass &descr[0].data[0] and
 * stride 132).  `surf` is n_surf x 64 floats, stride >= 64 (68 for
 * Surf::Descriptor, surf.h:83-95).  Either may be NULL with n = 0.  The floats are read
 * at osfm_match_commit (by a few host threads at once, through page-locked buffers: a single
 * pageable copy moves a fifth of what the link can). */
int osfm_match_set_view_f32(osfm_matcher* m, int view_id,
    const float* sift, int n_sift, int sift_stride,
    const float* surf, int n_surf, int surf_stride);

/* Already quantised descriptors: n_sift x 128 unsigned bytes, n_surf x 64
 * signed bytes, densely packed. */
int osfm_match_set_view_q8(osfm_matcher* m, int view_id,
    const uint8_t* sift, int n_sift, const int8_t* surf, int n_surf);

/* Copies all staged views into one HBM-resident pool and builds the TMA
 * descriptors.  After commit the views are immutable. */
int osfm_match_commit(osfm_matcher* m);

/* Overlapped staging.  Same cycle as begin / set_view_q8 / commit, but the host-to-device
 * copies run on their own stream and osfm_match_commit returns WITHOUT waiting for them:
 * the buffers handed to osfm_match_set_view_q8 must stay valid and unchanged until the first
 * call that returns results (osfm_match_pairs*, osfm_match_pair*, osfm_match_two_view*) or
 * osfm_match_wait_staged has returned.  The first batched call then launches its filter pass
 * in several pieces -- the pairs grouped by the highest view they need, view ranges growing
 * geometrically -- each piece waiting on the device only for its own views, so the filter works
 * on the early pairs of a list in the reference's order (view_1 ascending,
 * bundler_matching.cc:92-93) while the later views are still being copied; nothing else of the
 * batch is split.  Views must be staged in ascending order (otherwise commit waits, as
 * the plain one does); quantised descriptors only.  Results are the same as with
 * osfm_match_begin. */
/* osfm_match_set_view_q8 for `count` consecutive views in one call (sift / surf: one pointer
 * per view, or NULL for "no features of this type"). */
int osfm_match_set_views_q8(osfm_matcher* m, int first_view, int count,
    const uint8_t* const* sift, const int32_t* n_sift, const int8_t* const* surf, const int32_t* n_surf);
int osfm_match_begin_overlapped(osfm_matcher* m, int num_views);
/* Waits until everything staged has arrived and is ready. */
int osfm_match_wait_staged(osfm_matcher* m);

/* Device-resident variant for the multi-GPU path: adopts a descriptor pool that
 * already lives in this device's memory (e.g. the target of an NCCL broadcast).
 * sift_pool is sum(n_sift) x 128 bytes with view v at row sift_row_offset[v];
 * it must be 128-byte aligned and followed by at least 256 readable rows of
 * padding.  surf_pool likewise with 64-byte rows (may be NULL).  The memory
 * stays owned by the caller and must outlive the handle. */
int osfm_match_commit_device(osfm_matcher* m, int num_views,
    const void* sift_pool, const int64_t* sift_row_offset, const int32_t* n_sift,
    int64_t sift_pool_rows,
    const void* surf_pool, const int64_t* surf_row_offset, const int32_t* n_surf,
    int64_t surf_pool_rows);

int osfm_match_num_views(const osfm_matcher* m);
/* Number of SIFT / SURF features of a view; their sum is the length of that
 * view's side of a Matching::Result. */
int osfm_match_view_size(const osfm_matcher* m, int view_id, int* n_sift, int* n_surf);

/* ---- matching -------------------------------------------------------------- */

/* Replaces ExhaustiveMatching::pairwise_match (exhaustive_matching.cc:115-144):
 * SIFT two-way match + remove_inconsistent_matches, the same for SURF, then
 * combine_results.  matches_1_2 receives *len_1_2 ints (index into view_2's
 * features or -1), matches_2_1 receives *len_2_1 ints.  The lengths follow the
 * reference exactly (a feature type contributes only if view_1 has descriptors
 * of it, exhaustive_matching.cc:123,134); the buffers must hold
 * n_sift+n_surf ints of the respective view.  n_consistent (may be NULL)
 * receives Matching::count_consistent_matches (matching.cc:39-47). */
int osfm_match_pair(osfm_matcher* m, int view_1_id, int view_2_id,
    int32_t* matches_1_2, int* len_1_2, int32_t* matches_2_1, int* len_2_1,
    int* n_consistent);

/* Look-ahead for callers that ask pair by pair, as the unchanged reference does: the loop of
 * bundler::Matching::compute (bundler_matching.cc:74-132) calls pairwise_match_lowres /
 * pairwise_match once per pair, in the order i -> (view_1, view_2) of bundler_matching.cc:92-93.
 * With max_pairs > 1, an osfm_match_pair(view_1 > view_2) call whose pair is not cached
 * matches that pair AND the max_pairs - 1 pairs that follow it in this order in one batched pass
 * (sharded over the devices of a multi-device handle; the correspondence lists are kept in pinned
 * host memory, at most 1 GiB, and a pair's dense vectors are rebuilt from its list); the calls
 * that follow are served from it.  osfm_match_pair_lowres does the same with a window of 16 x max_pairs.  Results are
 * identical; pairs the caller skips (its low-res gate) were matched for nothing.  0 (default): off.
 * The cache is dropped by osfm_match_begin* / osfm_match_commit_device. */
int osfm_match_set_lookahead(osfm_matcher* m, int max_pairs);

/* Replaces Matching::twoway_match<T> (matching.h:148-159) for one feature kind:
 * both one-way results WITHOUT the mutual filter.  matches_1_2 has n1 entries
 * of that kind, matches_2_1 n2. */
int osfm_match_pair_twoway(osfm_matcher* m, int kind, int view_1_id, int view_2_id,
    int32_t* matches_1_2, int32_t* matches_2_1);

/* Replaces Matching::twoway_match<float> (matching.h:148-159) with
 * NearestNeighbor<float>::find (nearest_neighbor.cc:272-289): the float descriptor path,
 * which the reference compiles out of ExhaustiveMatching (DISCRETIZE_DESCRIPTORS 1) but
 * keeps as a static template.  Stateless with respect to the staged views: set_1 is
 * n1 x dim floats, set_2 n2 x dim floats (host pointers, densely packed, dim a multiple of
 * 4 and <= 128).  Large pairs first go through a tensor-core filter (tf32 hi/lo split, fp32
 * accumulate) that decides every row whose outcome is clear within its error bound; the other
 * rows -- and small pairs altogether -- are evaluated with inner products formed in the
 * summation order of the reference's SSE3 build, so the results are bit-identical to it, not
 * merely within its tie tolerance.
 * matches_1_2 receives n1 ints, matches_2_1 n2 ints (no mutual filter). */
int osfm_match_twoway_f32(osfm_matcher* m, const float* set_1, int n1, const float* set_2, int n2,
    int dim, float lowe_ratio_threshold, float distance_threshold,
    int32_t* matches_1_2, int32_t* matches_2_1);

/* Replaces ExhaustiveMatching::pairwise_match_lowres (exhaustive_matching.cc:147-180). */
int osfm_match_pair_lowres(osfm_matcher* m, int view_1_id, int view_2_id,
    size_t num_features, int* n_consistent);

/* Batched form of osfm_match_pair for SIFT-only or SIFT+SURF views: all pairs in
 * one pass of the persistent kernel.  pairs = 2*npairs view ids (view_1, view_2).
 * Results are written densely: pair p's matches_1_2 starts at
 * matches[offsets[2p]] and its matches_2_1 at matches[offsets[2p+1]];
 * offsets has 2*npairs+1 entries (the last one is the total length) and is
 * filled by the call.  `matches` (host) must hold osfm_match_pairs_result_size()
 * ints.  n_consistent (may be NULL) gets npairs counts. */
int64_t osfm_match_pairs_result_size(osfm_matcher* m, const int32_t* pairs, int npairs);
int osfm_match_pairs(osfm_matcher* m, const int32_t* pairs, int npairs,
    int32_t* matches, int64_t* offsets, int32_t* n_consistent);

/* Device-resident batched form (no host<->device copies of descriptors or dense
 * results): the filtered matches of every pair are compacted, ordered by the
 * view_1 feature index, into (i, j) int32 pairs.  d_match_ij (device) holds
 * capacity_ij pairs; pair p's list starts at list_offset[p] (host, npairs+1
 * entries, filled by the call; list_offset[npairs] = total).  Returns
 * OSFM_ERR_OUT_OF_MEMORY if capacity_ij is too small (list_offset[npairs] then
 * holds the required capacity).  SIFT only. */
int osfm_match_pairs_compact_device(osfm_matcher* m, const int32_t* pairs, int npairs,
    int32_t* d_match_ij, int64_t capacity_ij, int64_t* list_offset);

/* The same with a HOST result buffer: match_ij receives the (i, j) pairs (2 ints
 * each) of every pair's surviving matches, ordered by i -- the correspondence list
 * bundler::Matching::two_view_matching builds from the Matching::Result
 * (bundler_matching.cc:178-192) -- about an eighth of the bytes of the dense vectors. */
int osfm_match_pairs_compact(osfm_matcher* m, const int32_t* pairs, int npairs,
    int32_t* match_ij, int64_t capacity_ij, int64_t* list_offset);

/* ---- two-view gates --------------------------------------------------------
 * bundler::Matching::two_view_matching (bundler_matching.cc:139-192) up to the point
 * where RANSAC starts, for a whole list of pairs: the pair rules of compute()
 * (:92-100), the low-resolution gate (:146-158; all eligible pairs in one batch, and
 * pairs that fail it are never matched in full), the full match, the match-count
 * threshold max(8, min_feature_matches) (:161-169) and the correspondence list
 * (i, matches_1_2[i]) in ascending i (:171-192) in the combined SIFT+SURF index space.
 * Option defaults are the reference's (bundler_matching.h:58-76). */
typedef struct {
    int use_lowres_matching;
    int num_lowres_features;
    int min_lowres_matches;
    int min_feature_matches;
    int match_num_previous_frames;
    int reserved[3];
} osfm_two_view_options;
void osfm_match_two_view_default_options(osfm_two_view_options* o);

enum {
    OSFM_TWO_VIEW_OK = 0,                /* list filled; count = consistent matches            */
    OSFM_TWO_VIEW_SKIPPED = 1,           /* previous-frames rule, or a view without features   */
    OSFM_TWO_VIEW_LOWRES_REJECTED = 2,   /* count = low-res matches < min_lowres_matches       */
    OSFM_TWO_VIEW_TOO_FEW_MATCHES = 3,   /* count = consistent matches below the threshold     */
    OSFM_TWO_VIEW_TOO_FEW_INLIERS = 4    /* osfm_match_two_view: count = RANSAC inliers below the threshold */
};

/* match_ij (host) receives the lists back to back; pair p's list is
 * match_ij[2*list_offset[p] .. 2*list_offset[p+1]) (empty unless status[p] == OK);
 * list_offset has npairs+1 entries; status and count npairs each.  On
 * OSFM_ERR_OUT_OF_MEMORY list_offset[npairs] holds the required capacity. */
int osfm_match_two_view_candidates(osfm_matcher* m, const osfm_two_view_options* opts,
    const int32_t* pairs, int npairs, int32_t* match_ij, int64_t capacity_ij,
    int64_t* list_offset, int32_t* status, int32_t* count);

/* The whole two-view stage, bundler::Matching::compute / two_view_matching
 * (src/mve/sfm/bundler_matching.cc:57-220), for a list of pairs: the candidates above,
 * then RANSAC for the fundamental matrix on every pair that passed (samples drawn from
 * std::rand() in pair order, as the reference), then the inlier threshold
 * max(8, min_matching_inliers).  positions: FeatureSet::positions of every view (2 floats
 * per feature, SIFT features then SURF features), view after view.  Outputs as for
 * osfm_match_two_view_candidates, the lists holding the inlier matches: for the pairs in the
 * order bundler::Matching::compute visits them, the accepted ones are its PairwiseMatching. */
typedef struct osfm_ransac_options {
    int max_iterations;            /* 1000  */
    int min_matching_inliers;      /* 12    */
    double threshold;              /* 0.0015, in normalised image coordinates */
    int reserved[4];
} osfm_ransac_options;
void osfm_match_ransac_default_options(osfm_ransac_options* o);
int osfm_match_two_view(osfm_matcher* m, const osfm_two_view_options* opts,
    const osfm_ransac_options* ransac, const float* positions, const int32_t* pairs, int npairs,
    int32_t* match_ij, int64_t capacity_ij, int64_t* list_offset, int32_t* status, int32_t* count);

/* ---- RANSAC for the fundamental matrix ---------------------------------------
 * Replaces sfm::RansacFundamental::estimate (src/mve/sfm/ransac_fundamental.cc:26-105) as
 * bundler::Matching::two_view_matching runs it on every pair that passed the match-count
 * gates (src/mve/sfm/bundler_matching.cc:176-220), for all those pairs in one call.
 *
 * The reference draws each 8-match sample from std::rand(), one process-wide sequence
 * consumed pair after pair.  osfm_ransac_draw_samples makes exactly those draws (it calls
 * std::rand() itself: seed it, or not, as the reference's caller does), in the order of
 * the pairs given, max_iterations samples of eight ascending match indices per pair;
 * every pair needs at least 8 matches (the reference throws below that).
 * samples: 8 * max_iterations * npairs ints. */
int osfm_ransac_draw_samples(int npairs, const int64_t* list_offset, int max_iterations,
    int32_t* samples);

/* Fits a fundamental matrix to every sample (normalised 8-point + rank-2 enforcement,
 * src/mve/sfm/fundamental.cc:78-126), counts the inliers of each (Sampson distance <
 * threshold^2, fundamental.cc:225-247) and keeps, per pair, the first sample with the most
 * inliers -- in the reference's double arithmetic, so the inlier lists are the reference's.
 * positions: 2 floats per feature (FeatureSet::positions), concatenated over the views;
 * pair p joins views pair_views[2p], pair_views[2p+1] with the matches
 * match_ij[2*list_offset[p] .. 2*list_offset[p+1]).  inlier_ij (capacity
 * 2*list_offset[npairs] ints) receives the inlier matches of the pairs back to back,
 * inlier_offset npairs+1 offsets; fundamental (may be NULL) 9 doubles per pair, row-major
 * (zeros for a pair without inliers).  samples == NULL: the library draws them itself, as
 * osfm_ransac_draw_samples would, chunk of pairs by chunk while the device works on the
 * chunk before (the draws are the longer leg).  The thresholds on the inlier count stay
 * with the caller (bundler_matching.cc:203-210). */
int osfm_ransac_fundamental(osfm_matcher* m, int num_views, const int32_t* features_per_view,
    const float* positions, const int32_t* pair_views, const int64_t* list_offset,
    const int32_t* match_ij, int npairs, const int32_t* samples, int max_iterations,
    double threshold, int32_t* inlier_ij, int64_t* inlier_offset, double* fundamental);

/* ---- tracks ------------------------------------------------------------------
 * Replaces sfm::bundler::Tracks::compute incl. remove_invalid_tracks
 * (src/mve/sfm/bundler_tracks.cc:47-203) up to the numbering of the tracks: the tracks
 * are the connected components of the match graph (nodes = (view, feature), edges =
 * the matches of every pair) that do not hold two features of one view.
 * features_per_view: num_views counts; pair p joins views pair_views[2p] and
 * pair_views[2p+1] with the matches match_ij[2*list_offset[p] .. 2*list_offset[p+1])
 * (host, (feature in view 1, feature in view 2)).  track_of_feature (host, sum of
 * features_per_view ints, view after view) receives Viewport::track_ids: the track of
 * every feature or -1; tracks are numbered in ascending order of their first feature.
 * num_conflicting (may be NULL): components dropped because of a view conflict (the
 * reference's num_invalid_tracks).  The handle only provides the device and stream. */
int osfm_tracks_compute(osfm_matcher* m, int num_views, const int32_t* features_per_view,
    const int32_t* pair_views, const int64_t* list_offset, const int32_t* match_ij, int npairs,
    int32_t* track_of_feature, int32_t* num_tracks, int32_t* num_conflicting);

/* ---- on-disk format --------------------------------------------------------
 * The MVE "prebundle" file (sfm::bundler::save_prebundle_to_file /
 * load_prebundle_from_file, src/mve/sfm/bundler_common.cc:56-190): feature
 * positions and colors per view plus the pairwise match lists -- the file through
 * which MVE's own tools pick up a matching result.  Byte-identical to the
 * reference's writer for the same data.  positions (2 floats per feature) and colors
 * (3 bytes per feature) are concatenated over the views and may be NULL (written as
 * empty vectors).  Host code only; no handle needed. */
int osfm_io_save_prebundle(const char* path, int num_views, const int32_t* features_per_view,
    const float* positions, const uint8_t* colors, int npairs, const int32_t* pair_views,
    const int64_t* list_offset, const int32_t* match_ij);

typedef struct osfm_prebundle osfm_prebundle;
/* Reads a file; the counts size the buffers for osfm_io_prebundle_get (any of which
 * may be NULL): n_positions / n_colors [num_views], positions [2*num_positions],
 * colors [3*num_colors], pair_views [2*npairs], list_offset [npairs+1],
 * match_ij [2*num_matches]. */
int osfm_io_load_prebundle(const char* path, osfm_prebundle** out, int* num_views,
    int64_t* num_positions, int64_t* num_colors, int* npairs, int64_t* num_matches);
int osfm_io_prebundle_get(const osfm_prebundle* h, int32_t* n_positions, int32_t* n_colors,
    float* positions, uint8_t* colors, int32_t* pair_views, int64_t* list_offset, int32_t* match_ij);
void osfm_io_prebundle_free(osfm_prebundle* h);

/* tracks.txt (orthosfm::saveTracksToFile / loadTracksFromFile,
 * src/matching/matching_io.cpp:16-95; replayed with --calculated-tracks,
 * src/sfm/reconstruct.cpp:70-78) written from osfm_tracks_compute's result: one line per
 * track, "count;{viewID;localID;globalID;x;y;r;g;b}*", the Feature fields as the MVE bridge
 * fills them (src/matching/matching_mve.cpp:455-466): globalID = 32768*view + feature,
 * x|y = image_width * (position + 0.5).  positions: normalised MVE feature positions,
 * 2 floats per feature, concatenated over the views; colors (r, g, b bytes) may be NULL
 * (zeros, what the bridge writes before propagateColorsToTracks).  Tracks come in ascending
 * id, features inside a track in ascending (view, feature). */
int osfm_io_save_tracks(const char* path, int num_views, const int32_t* features_per_view,
    const int32_t* track_of_feature, int num_tracks, const float* positions, double image_width,
    const uint8_t* colors);

/* The AAA_BBB.txt files of orthosfm::saveTracksToPairwiseFiles
 * (src/matching/matching_io.cpp:97-140): for every two views sharing a track, the lines
 * "x1 y1 x2 y2" of the tracks seen by both.  View ids are the view indices. */
int osfm_io_save_pairwise_tracks(const char* folder, int num_views, const int32_t* features_per_view,
    const int32_t* track_of_feature, int num_tracks, const float* positions, double image_width,
    int* files_written);

typedef struct osfm_track_table osfm_track_table;
/* Reads a tracks.txt; osfm_io_track_table_get fills (any may be NULL) track_offset
 * [num_tracks+1], ids [3*num_features] (view, local id, global id), xy [2*num_features],
 * rgb [3*num_features]. */
int osfm_io_load_tracks(const char* path, osfm_track_table** out, int64_t* num_tracks,
    int64_t* num_features);
int osfm_io_track_table_get(const osfm_track_table* h, int64_t* track_offset, uint32_t* ids,
    float* xy, uint32_t* rgb);
void osfm_io_track_table_free(osfm_track_table* h);

/* ---- introspection --------------------------------------------------------- */

typedef struct {
    int64_t kernel_launches;     /* launches of this library's kernels so far      */
    int64_t scan_items;          /* (job, 128-row block) work items processed       */
    int64_t candidate_rows;      /* rows that passed the scan kernel's filter (re-run exactly) */
    int64_t slow_rows;           /* rows without the 16-bit norm certificate (skip the filter) */
    int64_t self_check_failures; /* must stay 0                                      */
    double  last_scan_ms;        /* CUDA-event time of the last scan kernel launch(es) */
    double  last_total_ms;       /* CUDA-event time of the last batched call, device part */
    int64_t last_comparisons;    /* sum n1*n2 of the last batched call (unique)      */
    int64_t exact_rows;          /* rows re-run by the EXACT pass (survivors + uncertified) */
    int64_t last_scan_sm_cycles; /* SM cycles (clock64) of the last filter scan launch, CTA 0 */
    int64_t last_scan_ns;        /* its duration in ns (globaltimer): cycles/ns = SM clock in GHz */
    int64_t claimed_rows;        /* rows of the reverse direction of a pair that some forward match claims
                                    (the only rows of that direction that are evaluated), cumulative */
    double  last_phase_ms[8];    /* CUDA-event time of the last batched call per phase, summed over its
                                    batches: [0] filter pass, [1] classify + certify, [2] RESOLVE / EXACT of
                                    the forward direction, [3] claims + targets, [4] RESOLVE / EXACT of the
                                    claimed rows, [5] mutual filter, [6] list compaction, [7] unused */
    int64_t reverse_restricted_pairs; /* (pair, kind) jobs whose reverse direction was held against a subset of the
                                    other view only: the rows whose own best similarity can matter; cumulative */
    int64_t reverse_candidate_rows;   /* rows in those subsets, cumulative */
    int64_t exact_wide_rows;     /* of exact_rows: replayed from CUDA-core inner products spread over the device
                                    (few rows against large views) instead of by the scan pass; cumulative */
    int64_t float_filter_rows;   /* float path: rows that went through the tensor-core filter first, cumulative */
    int64_t float_exact_rows;    /* ... of which the filter could not decide (evaluated by the exact kernel) */
} osfm_match_stats;

int osfm_match_get_stats(const osfm_matcher* m, osfm_match_stats* out);

/* Test / profiling hooks (not part of the reference surface). */
/* mode 0 = normal; 1 = scan kernel skips the epilogue reduction (MMA+TMA only);
 * 2 = epilogue reads TMEM but does not reduce.  Results are invalid for != 0. */
int osfm_match_debug_set_scan_mode(osfm_matcher* m, int mode);
/* on = 1: the filtered entry points run BOTH directions of every pair through the filter pass
 * (Matching::twoway_match as the reference executes it, matching.h:155-158) instead of one
 * direction plus the claimed rows of the other.  on = 2: one direction plus the claimed rows, but
 * those are held against the whole other view instead of the subset of its rows that can matter.
 * on = 0: the default.  Same results in every mode; for A/B tests and timing. */
int osfm_match_debug_set_both_directions(osfm_matcher* m, int on);
/* Float path (osfm_match_twoway_f32).  mode 0 (default): pairs of at least 2^20 comparisons go
 * through the tensor-core filter (tf32 hi/lo split, fp32 accumulate) and only the rows it cannot
 * decide within its error bound through the exact CUDA-core kernel; smaller pairs through the exact
 * kernel alone.  mode 1: the exact kernel alone.  mode 2: always the filter first.  The match
 * vectors are the same in every mode. */
int osfm_match_debug_set_float_path(osfm_matcher* m, int mode);
/* The filter's view of one pair: per row of set_1 (then of set_2) its largest and second largest
 * similarity as the tensor cores computed them and the index of the largest.  s1 / s2 / j1:
 * n1 + n2 entries each. */
int osfm_match_debug_float_filter(osfm_matcher* m, const float* set_1, int n1, const float* set_2, int n2, int dim,
    float* s1, float* s2, int32_t* j1);
/* Rows whose similarities reach 2^16 are replayed exactly.  mode 0 (default): when their inner
 * products fit a scratch buffer they are computed on CUDA cores across the whole device and
 * replayed one warp per row; otherwise by the tensor-core scan pass.  mode 1: always the scan
 * pass.  Same results either way; for A/B tests and timing. */
int osfm_match_debug_set_exact_path(osfm_matcher* m, int mode);
/* Writes the raw int32 similarity matrix of one (query view, candidate view)
 * SIFT job as computed by the tensor-core kernel: out is n_q x ld ints,
 * ld = 256 * ceil(n_c / 256). */
int osfm_match_debug_dump_similarity(osfm_matcher* m, int kind, int view_q, int view_c,
    int32_t* out, int64_t out_ints);

/* Writes the same matrix as the scan kernel's filter reads it from tensor memory:
 * 16-bit packed (tcgen05.ld ... .pack::16b), word k of a row = columns 2k (bits 0-15) and
 * 2k+1 (bits 16-31), truncated to 16 bits; out is n_q x ld/2 words. */
int osfm_match_debug_dump_packed(osfm_matcher* m, int kind, int view_q, int view_c,
    uint32_t* out, int64_t out_words);

/* Runs the scan kernel over both directions of the given SIFT pairs with clock64()
 * time stamps of CTA 0's pipeline events: out receives 20 warps x 256 events x 4
 * int64 (warps 0-1, the MMA issuers: wait start, accumulator free, issued, half;
 * warp 2, the TMA producer: wait start, ring slot free, tile; warps 4-19, the filter
 * epilogue: wait start, accumulator ready, handed back, maxima done). */
int osfm_match_debug_trace(osfm_matcher* m, const int32_t* pairs, int npairs,
    int64_t* out, int64_t out_words);

#ifdef __cplusplus
}
#endif
#endif /* OSFM_MATCH_H */
