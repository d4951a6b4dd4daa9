/*
 * shim_check.cc -- drop-in proof.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiled (oracle/Makefile, target ref) against the reference's own headers and linked
 * with the reference's own matcher objects AND libosfm_match.so.  It drives both
 * sfm::ExhaustiveMatching (the reference, CPU) and sfm::GpuExhaustiveMatching (the C++
 * binding over the C ABI, orthosfm_b200/csrc/gpu_exhaustive_matching.h) through the same
 * sfm::MatchingBase pointer, on the same bundler::ViewportList, the way
 * bundler::Matching does (src/mve/sfm/bundler_matching.cc:45-56, 139-162), and compares
 * every Matching::Result element for element.  It then runs the reference's whole two-view
 * stage, sfm::bundler::Matching (init + compute), against sfm::bundler::GpuMatching
 * (orthosfm_b200/csrc/gpu_bundler_matching.h) on a scene with real two-view geometry and
 * the same std::rand() seed, and compares the PairwiseMatching.
 *
 * The binary lands in oracle/_ref/ (git-ignored, travels to the GPU box);
 * tests/test_gpu_parity.py::test_reference_side_binding runs it.
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "sfm/bundler_common.h"
#include "sfm/exhaustive_matching.h"
#include "sfm/matching_base.h"

#include "sfm/bundler_matching.h"

#include <omp.h>

#include "gpu_exhaustive_matching.h"
#include "gpu_bundler_matching.h"

namespace
{
    /* OSFM_SHIM_DEVICES="0,1": run every GPU object of this check over those devices
     * (osfm_match_create_multi); default: device 0. */
    std::vector<int>
    shim_devices (void)
    {
        std::vector<int> out;
        char const* env = std::getenv("OSFM_SHIM_DEVICES");
        std::string tok;
        for (char const* c = env ? env : ""; ; ++c)
        {
            if (*c == ',' || *c == '\0')
            {
                if (!tok.empty()) out.push_back(std::atoi(tok.c_str()));
                tok.clear();
                if (*c == '\0') break;
            }
            else tok.push_back(*c);
        }
        if (out.empty()) out.push_back(0);
        return out;
    }

    /* A small multi-view scene: 3-D points with a SIFT descriptor each, seen by every view
     * from its own pose (positions = projections, MVE's normalised coordinates), plus
     * private features at random positions. */
    void
    fill_scene (sfm::bundler::ViewportList* viewports, int num_views, int n, int scene_points, unsigned seed)
    {
        std::mt19937 rng(seed);
        std::normal_distribution<float> gauss(0.0f, 1.0f);
        std::uniform_real_distribution<float> uni(-1.0f, 1.0f);
        auto make = [&] (std::vector<float>& d)
        {
            float nrm = 0.0f;
            for (float x : d) nrm += x * x;
            nrm = std::sqrt(nrm);
            for (float& x : d) x = std::min(x / nrm, 0.2f);
            nrm = 0.0f;
            for (float x : d) nrm += x * x;
            nrm = std::sqrt(nrm);
            for (float& x : d) x /= nrm;
        };
        std::vector<std::vector<float>> pool(scene_points, std::vector<float>(128));
        std::vector<float> X(3 * scene_points);
        for (int i = 0; i < scene_points; ++i)
        {
            for (float& x : pool[i]) x = std::fabs(gauss(rng));
            make(pool[i]);
            X[3 * i] = uni(rng); X[3 * i + 1] = uni(rng); X[3 * i + 2] = 5.0f + uni(rng);
        }
        viewports->resize(num_views);
        for (int v = 0; v < num_views; ++v)
        {
            float const a = 0.2f * uni(rng), b = 0.2f * uni(rng), tx = 0.5f * uni(rng), ty = 0.5f * uni(rng);
            sfm::FeatureSet& fs = (*viewports)[v].features;
            fs.sift_descriptors.resize(n);
            fs.positions.resize(n);
            fs.colors.resize(n, math::Vec3uc(0, 0, 0));
            for (int i = 0; i < n; ++i)
            {
                std::vector<float> d(128);
                float px, py;
                if (i < scene_points && (rng() % 2 == 0))
                {
                    d = pool[i];
                    for (float& x : d) x = std::max(0.0f, x + 0.004f * gauss(rng));
                    /* rotation about y by a, about x by b, then a shift */
                    float const x0 = X[3 * i], y0 = X[3 * i + 1], z0 = X[3 * i + 2];
                    float const x1 = std::cos(a) * x0 + std::sin(a) * z0, z1 = -std::sin(a) * x0 + std::cos(a) * z0;
                    float const y2 = std::cos(b) * y0 - std::sin(b) * z1, z2 = std::sin(b) * y0 + std::cos(b) * z1;
                    px = 1.2f * (x1 + tx) / z2 + 0.0002f * gauss(rng);
                    py = 1.2f * (y2 + ty) / z2 + 0.0002f * gauss(rng);
                }
                else
                {
                    for (float& x : d) x = std::fabs(gauss(rng));
                    px = 0.5f * uni(rng);
                    py = 0.5f * uni(rng);
                }
                make(d);
                sfm::Sift::Descriptor& out = fs.sift_descriptors[i];
                out.x = out.y = out.scale = out.orientation = 0.0f;
                for (int k = 0; k < 128; ++k) out.data[k] = d[k];
                fs.positions[i] = math::Vec2f(px, py);
            }
        }
    }

    /* bundler::Matching (the reference, one thread, as shipped) against bundler::GpuMatching
     * on the same viewports and the same std::rand() seed. */
    int
    check_bundler (void)
    {
        sfm::bundler::ViewportList scene_ref, scene_gpu;
        fill_scene(&scene_ref, 6, 900, 700, 77u);
        scene_gpu = scene_ref;
        sfm::bundler::Matching::Options opts;
        opts.min_feature_matches = 40;
        opts.min_matching_inliers = 25;
        opts.use_lowres_matching = true;
        opts.num_lowres_features = 300;
        opts.min_lowres_matches = 8;
        opts.ransac_opts.max_iterations = 200;
        opts.ransac_opts.threshold = 0.0015;
        opts.ransac_opts.verbose_output = false;

        sfm::bundler::PairwiseMatching want, got;
        try
        {
            std::FILE* quiet = std::freopen("/dev/null", "w", stdout);      /* compute() prints progress */
            (void)quiet;
            omp_set_num_threads(1);
            std::srand(5);
            sfm::bundler::Matching ref(opts);
            ref.init(&scene_ref);
            ref.compute(&want);
            std::srand(5);
            sfm::bundler::GpuMatching gpu(opts, nullptr, shim_devices());
            gpu.init(&scene_gpu);
            gpu.compute(&got);
            std::freopen("/dev/tty", "w", stdout);
        }
        catch (std::exception const& e)
        {
            std::fprintf(stderr, "BUNDLER_CHECK ERROR %s\n", e.what());
            return 1;
        }
        int bad = want.size() == got.size() ? 0 : 1;
        long inliers = 0;
        for (std::size_t p = 0; p < std::min(want.size(), got.size()); ++p)
        {
            inliers += (long)want[p].matches.size();
            if (want[p].view_1_id != got[p].view_1_id || want[p].view_2_id != got[p].view_2_id
                || want[p].matches != got[p].matches)
                ++bad;
        }
        std::fprintf(stderr, "BUNDLER_CHECK %s pairs=%zu/%zu inliers=%ld mismatches=%d\n",
            bad == 0 && !want.empty() ? "PASS" : "FAIL", got.size(), want.size(), inliers, bad);
        return bad == 0 && !want.empty() ? 0 : 1;
    }

    void
    fill_views (sfm::bundler::ViewportList* viewports, std::vector<int> const& n_sift,
        std::vector<int> const& n_surf, unsigned seed)
    {
        std::mt19937 rng(seed);
        std::normal_distribution<float> gauss(0.0f, 1.0f);
        /* shared "scene" descriptors so that some pairs really match */
        int const pool = 400;
        std::vector<std::vector<float>> sift_pool(pool, std::vector<float>(128));
        std::vector<std::vector<float>> surf_pool(pool, std::vector<float>(64));
        auto normalise = [] (std::vector<float>& v, bool clampit)
        {
            float n = 0.0f;
            for (float x : v) n += x * x;
            n = std::sqrt(n);
            for (float& x : v) x /= n;
            if (!clampit) return;
            for (float& x : v) x = std::min(x, 0.2f);
            n = 0.0f;
            for (float x : v) n += x * x;
            n = std::sqrt(n);
            for (float& x : v) x /= n;
        };
        for (int i = 0; i < pool; ++i)
        {
            for (float& x : sift_pool[i]) x = std::fabs(gauss(rng));
            normalise(sift_pool[i], true);
            for (float& x : surf_pool[i]) x = gauss(rng);
            normalise(surf_pool[i], false);
        }
        viewports->resize(n_sift.size());
        for (std::size_t v = 0; v < n_sift.size(); ++v)
        {
            sfm::FeatureSet& fs = (*viewports)[v].features;
            fs.sift_descriptors.resize(n_sift[v]);
            for (int i = 0; i < n_sift[v]; ++i)
            {
                std::vector<float> d(128);
                if (rng() % 3 == 0)
                {
                    d = sift_pool[rng() % pool];
                    for (float& x : d) x = std::max(0.0f, x + 0.004f * gauss(rng));
                }
                else
                    for (float& x : d) x = std::fabs(gauss(rng));
                normalise(d, true);
                sfm::Sift::Descriptor& out = fs.sift_descriptors[i];
                out.x = out.y = out.scale = out.orientation = 0.0f;
                for (int k = 0; k < 128; ++k) out.data[k] = d[k];
            }
            fs.surf_descriptors.resize(n_surf[v]);
            for (int i = 0; i < n_surf[v]; ++i)
            {
                std::vector<float> d(64);
                if (rng() % 3 == 0)
                {
                    d = surf_pool[rng() % pool];
                    for (float& x : d) x += 0.01f * gauss(rng);
                }
                else
                    for (float& x : d) x = gauss(rng);
                normalise(d, false);
                sfm::Surf::Descriptor& out = fs.surf_descriptors[i];
                out.x = out.y = out.scale = out.orientation = 0.0f;
                for (int k = 0; k < 64; ++k) out.data[k] = d[k];
            }
        }
    }
}

int
main (void)
{
    std::vector<int> const n_sift = { 900, 1300, 0, 257, 640 };
    std::vector<int> const n_surf = { 300, 0, 410, 129, 0 };
    sfm::bundler::ViewportList viewports;
    fill_views(&viewports, n_sift, n_surf, 1234u);

    std::unique_ptr<sfm::MatchingBase> ref(new sfm::ExhaustiveMatching());
    std::unique_ptr<sfm::MatchingBase> gpu;
    try
    {
        gpu.reset(new sfm::GpuExhaustiveMatching(shim_devices()));
        ref->init(&viewports);
        gpu->init(&viewports);
    }
    catch (std::exception const& e)
    {
        std::printf("SHIM_CHECK ERROR %s\n", e.what());
        return 2;
    }
    /* bundler::Matching::init frees the descriptors here (bundler_matching.cc:53-55) */
    for (std::size_t i = 0; i < viewports.size(); ++i)
    {
        /* same effect as FeatureSet::clear_descriptors() (not linked here: feature_set.cc
         * would pull in the SURF extractor) */
        sfm::Sift::Descriptors().swap(viewports[i].features.sift_descriptors);
        sfm::Surf::Descriptors().swap(viewports[i].features.surf_descriptors);
    }

    int bad = 0, pairs = 0;
    long consistent = 0;
    for (int v1 = 0; v1 < (int)viewports.size(); ++v1)
        for (int v2 = 0; v2 < (int)viewports.size(); ++v2)
        {
            if (v1 == v2) continue;
            sfm::Matching::Result a, b;
            ref->pairwise_match(v1, v2, &a);
            gpu->pairwise_match(v1, v2, &b);
            bool const same = a.matches_1_2 == b.matches_1_2 && a.matches_2_1 == b.matches_2_1;
            int const la = ref->pairwise_match_lowres(v1, v2, 200);
            int const lb = gpu->pairwise_match_lowres(v1, v2, 200);
            consistent += sfm::Matching::count_consistent_matches(a);
            if (!same || la != lb)
            {
                ++bad;
                std::printf("MISMATCH pair (%d,%d): sizes %zu/%zu vs %zu/%zu lowres %d vs %d\n", v1, v2,
                    a.matches_1_2.size(), a.matches_2_1.size(), b.matches_1_2.size(), b.matches_2_1.size(), la, lb);
            }
            ++pairs;
        }

    /* the batched entry point must give the same as the per-pair one */
    std::vector<std::pair<int, int>> list;
    for (int v1 = 1; v1 < (int)viewports.size(); ++v1)
        for (int v2 = 0; v2 < v1; ++v2)
            list.push_back(std::make_pair(v1, v2));
    std::vector<sfm::Matching::Result> all;
    static_cast<sfm::GpuExhaustiveMatching*>(gpu.get())->pairwise_match_all(list, &all);
    for (std::size_t p = 0; p < list.size(); ++p)
    {
        sfm::Matching::Result a;
        ref->pairwise_match(list[p].first, list[p].second, &a);
        if (!(a.matches_1_2 == all[p].matches_1_2 && a.matches_2_1 == all[p].matches_2_1))
        {
            ++bad;
            std::printf("MISMATCH batched pair (%d,%d)\n", list[p].first, list[p].second);
        }
    }

    /* the loop of bundler::Matching::compute (bundler_matching.cc:74-132, order of :92-93) asks
     * pair by pair; the shim serves it from batched look-ahead passes (window 4096 by default,
     * which covers this whole list, and a window of 3 that has to be refilled) */
    for (int window : { 4096, 3 })
    {
        sfm::GpuExhaustiveMatching looped(shim_devices());
        looped.set_lookahead(window);
        sfm::bundler::ViewportList again;
        fill_views(&again, n_sift, n_surf, 1234u);
        looped.init(&again);
        for (std::size_t p = 0; p < list.size(); ++p)
        {
            sfm::Matching::Result a, b;
            ref->pairwise_match(list[p].first, list[p].second, &a);
            int const la = ref->pairwise_match_lowres(list[p].first, list[p].second, 200);
            int const lb = looped.pairwise_match_lowres(list[p].first, list[p].second, 200);
            looped.pairwise_match(list[p].first, list[p].second, &b);
            if (!(a.matches_1_2 == b.matches_1_2 && a.matches_2_1 == b.matches_2_1) || la != lb)
            {
                ++bad;
                std::printf("MISMATCH look-ahead %d pair (%d,%d)\n", window, list[p].first, list[p].second);
            }
        }
    }

    /* error convention: exceptions, like the rest of MVE */
    bool threw = false;
    try { sfm::Matching::Result r; gpu->pairwise_match(0, 99, &r); }
    catch (std::invalid_argument const&) { threw = true; }
    if (!threw) { ++bad; std::printf("MISMATCH: bad view id did not throw std::invalid_argument\n"); }

    std::printf("SHIM_CHECK %s pairs=%d consistent=%ld mismatches=%d\n", bad == 0 ? "PASS" : "FAIL",
        pairs, consistent, bad);
    std::fflush(stdout);
    int const bundler_bad = check_bundler();       /* reports on stderr */
    return bad == 0 && bundler_bad == 0 ? 0 : 1;
}
