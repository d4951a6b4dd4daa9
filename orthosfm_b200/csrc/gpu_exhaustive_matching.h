/*
 * gpu_exhaustive_matching.h -- the reference-side binding: a third sfm::MatchingBase
 * implementation that forwards to the C ABI of libosfm_match.so (include/osfm_match.h).
 *
 * It is compiled INSIDE the reference tree (it includes the reference's own headers), next
 * to sfm::ExhaustiveMatching (src/mve/sfm/exhaustive_matching.h) and
 * sfm::CascadeHashing (src/mve/sfm/cascade_hashing.h), and is selected in the
 * bundler::Matching constructor switch (src/mve/sfm/bundler_matching.cc:31-41) -- see
 * INTEGRATION.md.  Nothing above bundler::Matching changes.
 *
 * Contract mirrored from the reference:
 *   - init() copies everything it needs: bundler::Matching::init frees the descriptors
 *     right after matcher->init() returns (bundler_matching.cc:53-55);
 *   - pairwise_match() / pairwise_match_lowres() are const and may be called from the
 *     OpenMP pair loop (bundler_matching.cc:74): the handle serialises internally;
 *   - errors surface as std::runtime_error / std::invalid_argument like the rest of MVE
 *     (bundler_matching.cc:40,48,62); an empty result is not an error.
 *
 * Header-only; link with -losfm_match.
 */
#ifndef OSFM_GPU_EXHAUSTIVE_MATCHING_HEADER
#define OSFM_GPU_EXHAUSTIVE_MATCHING_HEADER

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "sfm/bundler_common.h"
#include "sfm/defines.h"
#include "sfm/matching_base.h"

#include "osfm_match.h"

SFM_NAMESPACE_BEGIN

class GpuExhaustiveMatching : public MatchingBase
{
public:
    /** One GPU.  The environment variable OSFM_DEVICES ("0,1,2,3" or "all") overrides the
     *  choice, so that an unchanged binary can be given the whole box. */
    explicit GpuExhaustiveMatching (int device = 0) : devices(1, device), lookahead(4096), handle(nullptr)
    {
        this->devices_from_environment();
    }

    /** Several GPUs of this box behind one matcher (osfm_match_create_multi): the descriptor
     *  pool is replicated with an NCCL broadcast in init(), batched calls are sharded. */
    explicit GpuExhaustiveMatching (std::vector<int> const& devices)
        : devices(devices.empty() ? std::vector<int>(1, 0) : devices), lookahead(4096), handle(nullptr) {}

    /** Pairs matched ahead of a pairwise_match() / pairwise_match_lowres() call that follows
     *  the enumeration of bundler::Matching::compute (osfm_match_set_lookahead); 0 = none.
     *  Takes effect at the next init(). */
    void set_lookahead (int max_pairs) { this->lookahead = max_pairs; }

    ~GpuExhaustiveMatching (void) override
    {
        if (this->handle != nullptr)
            osfm_match_destroy(this->handle);
    }

    GpuExhaustiveMatching (GpuExhaustiveMatching const&) = delete;
    GpuExhaustiveMatching& operator= (GpuExhaustiveMatching const&) = delete;

    /** Stages the SIFT / SURF descriptors of every viewport into HBM (quantised on the
     *  device like convert_descriptor, exhaustive_matching.cc:18-39). */
    void init (bundler::ViewportList* viewports) override
    {
        if (viewports == nullptr)
            throw std::invalid_argument("Viewports must not be null");
        if (this->handle == nullptr)
        {
            /* The options may have been edited through MatchingBase::opts after construction,
             * so the handle is created here, not in the constructor. */
            osfm_match_config cfg;
            osfm_match_default_config(&cfg);
            cfg.device = this->devices[0];
            cfg.sift_lowe_ratio = this->opts.sift_matching_opts.lowe_ratio_threshold;
            cfg.sift_distance_threshold = this->opts.sift_matching_opts.distance_threshold;
            cfg.surf_lowe_ratio = this->opts.surf_matching_opts.lowe_ratio_threshold;
            cfg.surf_distance_threshold = this->opts.surf_matching_opts.distance_threshold;
            int const rc = this->devices.size() > 1
                ? osfm_match_create_multi(&cfg, this->devices.data(), static_cast<int>(this->devices.size()), &this->handle)
                : osfm_match_create(&cfg, &this->handle);
            if (rc != OSFM_OK)
            {
                std::string msg = this->handle ? osfm_match_last_error(this->handle) : "allocation failed";
                if (this->handle) { osfm_match_destroy(this->handle); this->handle = nullptr; }
                throw std::runtime_error("GpuExhaustiveMatching: " + msg);
            }
        }
        this->check(osfm_match_begin(this->handle, static_cast<int>(viewports->size())));
        for (std::size_t i = 0; i < viewports->size(); ++i)
        {
            FeatureSet const& fs = (*viewports)[i].features;
            /* Sift::Descriptor / Surf::Descriptor are {x, y, scale, orientation, data[]}:
             * the float data of consecutive descriptors is sizeof(Descriptor)/4 floats apart. */
            float const* sift = fs.sift_descriptors.empty() ? nullptr : fs.sift_descriptors[0].data.begin();
            float const* surf = fs.surf_descriptors.empty() ? nullptr : fs.surf_descriptors[0].data.begin();
            this->check(osfm_match_set_view_f32(this->handle, static_cast<int>(i),
                sift, static_cast<int>(fs.sift_descriptors.size()),
                static_cast<int>(sizeof(Sift::Descriptor) / sizeof(float)),
                surf, static_cast<int>(fs.surf_descriptors.size()),
                static_cast<int>(sizeof(Surf::Descriptor) / sizeof(float))));
        }
        this->check(osfm_match_commit(this->handle));
        /* The unchanged loop of bundler::Matching::compute asks pair by pair
         * (bundler_matching.cc:74-132): serve it from batched passes over the pairs that follow. */
        this->check(osfm_match_set_lookahead(this->handle, this->lookahead));
    }

    /** Matches all feature types yielding a single matching result. */
    void pairwise_match (int view_1_id, int view_2_id, Matching::Result* result) const override
    {
        this->require_init();
        int n1s = 0, n1f = 0, n2s = 0, n2f = 0;
        this->check(osfm_match_view_size(this->handle, view_1_id, &n1s, &n1f));
        this->check(osfm_match_view_size(this->handle, view_2_id, &n2s, &n2f));
        result->matches_1_2.assign(static_cast<std::size_t>(n1s + n1f) + 1, -1);
        result->matches_2_1.assign(static_cast<std::size_t>(n2s + n2f) + 1, -1);
        int len12 = 0, len21 = 0;
        this->check(osfm_match_pair(this->handle, view_1_id, view_2_id,
            result->matches_1_2.data(), &len12, result->matches_2_1.data(), &len21, nullptr));
        result->matches_1_2.resize(len12);
        result->matches_2_1.resize(len21);
    }

    /** Matches the N lowest resolution features and returns the number of matches. */
    int pairwise_match_lowres (int view_1_id, int view_2_id, std::size_t num_features) const override
    {
        this->require_init();
        int count = 0;
        this->check(osfm_match_pair_lowres(this->handle, view_1_id, view_2_id, num_features, &count));
        return count;
    }

    /** Batched form for callers that know all pairs up front (one persistent-kernel pass):
     *  results[p] is what pairwise_match(pairs[p].first, pairs[p].second) would return. */
    void pairwise_match_all (std::vector<std::pair<int, int> > const& pairs,
        std::vector<Matching::Result>* results) const
    {
        this->require_init();
        std::vector<int32_t> flat;
        flat.reserve(pairs.size() * 2);
        for (std::size_t p = 0; p < pairs.size(); ++p)
        {
            flat.push_back(pairs[p].first);
            flat.push_back(pairs[p].second);
        }
        int const npairs = static_cast<int>(pairs.size());
        int64_t const total = osfm_match_pairs_result_size(this->handle, flat.data(), npairs);
        if (total < 0)
            this->check(static_cast<int>(total));
        std::vector<int32_t> dense(static_cast<std::size_t>(total) + 1);
        std::vector<int64_t> offsets(2 * pairs.size() + 1);
        this->check(osfm_match_pairs(this->handle, flat.data(), npairs, dense.data(), offsets.data(), nullptr));
        results->resize(pairs.size());
        for (std::size_t p = 0; p < pairs.size(); ++p)
        {
            (*results)[p].matches_1_2.assign(dense.begin() + offsets[2 * p], dense.begin() + offsets[2 * p + 1]);
            (*results)[p].matches_2_1.assign(dense.begin() + offsets[2 * p + 1], dense.begin() + offsets[2 * p + 2]);
        }
    }

    /** The whole two-view stage for a list of pairs (osfm_match_two_view); used by
     *  bundler::GpuMatching (gpu_bundler_matching.h). */
    void two_view (osfm_two_view_options const* two, osfm_ransac_options const* ransac, float const* positions,
        int32_t const* pairs, int npairs, int32_t* match_ij, int64_t capacity_ij, int64_t* list_offset,
        int32_t* status, int32_t* count) const
    {
        this->require_init();
        this->check(osfm_match_two_view(this->handle, two, ransac, positions, pairs, npairs, match_ij,
            capacity_ij, list_offset, status, count));
    }

private:
    void require_init (void) const
    {
        if (this->handle == nullptr)
            throw std::runtime_error("GpuExhaustiveMatching: init() has not been called");
    }

    void check (int rc) const
    {
        if (rc == OSFM_OK)
            return;
        std::string const msg = std::string("GpuExhaustiveMatching: ")
            + (this->handle ? osfm_match_last_error(this->handle) : "no handle");
        if (rc == OSFM_ERR_INVALID_ARGUMENT)
            throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }

    void devices_from_environment (void)
    {
        char const* env = std::getenv("OSFM_DEVICES");
        if (env == nullptr || *env == '\0')
            return;
        std::vector<int> list;
        if (std::string(env) == "all")
        {
            /* the library reports a bad ordinal; 64 is more than any box has */
            for (int d = 0; d < 64; ++d)
            {
                osfm_match_config cfg;
                osfm_match_default_config(&cfg);
                cfg.device = d;
                osfm_matcher* probe = nullptr;
                int const rc = osfm_match_create(&cfg, &probe);
                if (probe != nullptr) osfm_match_destroy(probe);
                if (rc != OSFM_OK) break;
                list.push_back(d);
            }
        }
        else
        {
            std::string tok;
            for (char const* c = env; ; ++c)
            {
                if (*c == ',' || *c == '\0')
                {
                    if (!tok.empty()) list.push_back(std::atoi(tok.c_str()));
                    tok.clear();
                    if (*c == '\0') break;
                }
                else tok.push_back(*c);
            }
        }
        if (!list.empty())
            this->devices = list;
    }

    std::vector<int> devices;
    int lookahead;
    osfm_matcher* handle;
};

SFM_NAMESPACE_END

#endif /* OSFM_GPU_EXHAUSTIVE_MATCHING_HEADER */
