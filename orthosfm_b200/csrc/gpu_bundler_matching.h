/*
 * gpu_bundler_matching.h -- the reference-side binding for the whole two-view stage: a
 * drop-in for sfm::bundler::Matching (src/mve/sfm/bundler_matching.h:49-112) that hands all
 * pairs to the GPU at once instead of looping over them.
 *
 * Same interface as the reference class -- Matching(Options, Progress*), init(ViewportList*),
 * compute(PairwiseMatching*) -- and the same result: the pairs are visited in compute()'s own
 * order (bundler_matching.cc:92-93), the RANSAC samples come from the same std::rand()
 * sequence, the arithmetic is the reference's, so `pairwise_matching` receives what the
 * reference's single-threaded compute() appends (the build the reference ships has no OpenMP).
 * A caller such as calculateTracksUsingMVE (src/matching/matching_mve.cpp:405-415) switches
 * by changing the type of its `matching` object.
 *
 * Compiled INSIDE the reference tree; header-only; link with -losfm_match.
 */
#ifndef OSFM_GPU_BUNDLER_MATCHING_HEADER
#define OSFM_GPU_BUNDLER_MATCHING_HEADER

#include <stdexcept>
#include <vector>

#include "sfm/bundler_common.h"
#include "sfm/bundler_matching.h"
#include "sfm/defines.h"

#include "gpu_exhaustive_matching.h"
#include "osfm_match.h"

SFM_NAMESPACE_BEGIN
SFM_BUNDLER_NAMESPACE_BEGIN

class GpuMatching
{
public:
    typedef Matching::Options Options;
    typedef Matching::Progress Progress;

    explicit GpuMatching (Options const& options, Progress* progress = nullptr, int device = 0)
        : opts(options), progress(progress), matcher(device), viewports(nullptr)
    {
        if (this->opts.matcher_type != Matching::MATCHER_EXHAUSTIVE)
            throw std::runtime_error("GpuMatching replaces the exhaustive matcher only");
    }

    /** Several GPUs of this box: the pairs of compute() are sharded over them. */
    GpuMatching (Options const& options, Progress* progress, std::vector<int> const& devices)
        : opts(options), progress(progress), matcher(devices), viewports(nullptr)
    {
        if (this->opts.matcher_type != Matching::MATCHER_EXHAUSTIVE)
            throw std::runtime_error("GpuMatching replaces the exhaustive matcher only");
    }

    /** Stages the descriptors on the device and frees them in the viewports, as
     *  bundler::Matching::init does (bundler_matching.cc:45-56); the positions stay. */
    void init (ViewportList* viewports)
    {
        if (viewports == nullptr)
            throw std::invalid_argument("Viewports must not be null");
        this->viewports = viewports;
        this->matcher.init(viewports);
        for (std::size_t i = 0; i < viewports->size(); i++)
            viewports->at(i).features.clear_descriptors();
    }

    /** bundler_matching.cc:57-133 for all pairs in one call. */
    void compute (PairwiseMatching* pairwise_matching)
    {
        if (this->viewports == nullptr)
            throw std::runtime_error("Viewports must not be null");
        std::size_t const num_viewports = this->viewports->size();
        std::size_t const num_pairs = num_viewports * (num_viewports - 1) / 2;
        if (this->progress != nullptr)
        {
            this->progress->num_total = num_pairs;
            this->progress->num_done = 0;
        }
        if (num_pairs == 0)
            return;

        /* i -> (view_1, view_2) in the reference's order; positions back to back */
        std::vector<int32_t> pairs;
        pairs.reserve(2 * num_pairs);
        for (std::size_t v1 = 1; v1 < num_viewports; ++v1)
            for (std::size_t v2 = 0; v2 < v1; ++v2)
            {
                pairs.push_back(static_cast<int32_t>(v1));
                pairs.push_back(static_cast<int32_t>(v2));
            }
        std::vector<float> positions;
        std::size_t capacity = 1;
        for (std::size_t v = 0; v < num_viewports; ++v)
        {
            FeatureSet const& fs = this->viewports->at(v).features;
            for (std::size_t f = 0; f < fs.positions.size(); ++f)
            {
                positions.push_back(fs.positions[f][0]);
                positions.push_back(fs.positions[f][1]);
            }
        }
        for (std::size_t p = 0; p < num_pairs; ++p)
            capacity += std::min(this->viewports->at(pairs[2 * p]).features.positions.size(),
                this->viewports->at(pairs[2 * p + 1]).features.positions.size());

        osfm_two_view_options two;
        osfm_match_two_view_default_options(&two);
        two.use_lowres_matching = this->opts.use_lowres_matching ? 1 : 0;
        two.num_lowres_features = this->opts.num_lowres_features;
        two.min_lowres_matches = this->opts.min_lowres_matches;
        two.min_feature_matches = this->opts.min_feature_matches;
        two.match_num_previous_frames = this->opts.match_num_previous_frames;
        osfm_ransac_options ransac;
        osfm_match_ransac_default_options(&ransac);
        ransac.max_iterations = this->opts.ransac_opts.max_iterations;
        ransac.threshold = this->opts.ransac_opts.threshold;
        ransac.min_matching_inliers = this->opts.min_matching_inliers;

        std::vector<int32_t> ij(2 * capacity);
        std::vector<int64_t> offsets(num_pairs + 1);
        std::vector<int32_t> status(num_pairs), count(num_pairs);
        this->matcher.two_view(&two, &ransac, positions.data(), pairs.data(), static_cast<int>(num_pairs),
            ij.data(), static_cast<int64_t>(capacity), offsets.data(), status.data(), count.data());

        for (std::size_t p = 0; p < num_pairs; ++p)
        {
            if (this->progress != nullptr)
                this->progress->num_done += 1;
            if (status[p] != OSFM_TWO_VIEW_OK)
                continue;
            TwoViewMatching matching;
            matching.view_1_id = pairs[2 * p];
            matching.view_2_id = pairs[2 * p + 1];
            matching.matches.reserve(static_cast<std::size_t>(offsets[p + 1] - offsets[p]));
            for (int64_t k = offsets[p]; k < offsets[p + 1]; ++k)
                matching.matches.push_back(std::make_pair(ij[2 * k], ij[2 * k + 1]));
            pairwise_matching->push_back(matching);
        }
    }

private:
    Options opts;
    Progress* progress;
    GpuExhaustiveMatching matcher;
    ViewportList const* viewports;
};

SFM_BUNDLER_NAMESPACE_END
SFM_NAMESPACE_END

#endif /* OSFM_GPU_BUNDLER_MATCHING_HEADER */
