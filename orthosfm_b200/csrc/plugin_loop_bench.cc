// plugin_loop_bench.cc -- the reference's plugin call pattern against the C ABI, natively.
//
// What the unchanged reference does around its matcher (paths relative to /root/reference):
//   bundler::Matching::init      src/mve/sfm/bundler_matching.cc:45-56   matcher->init(viewports) with the FLOAT
//                                                                         descriptors of every view (Sift::Descriptor
//                                                                         records: x, y, scale, orientation, data[128])
//   bundler::Matching::compute   src/mve/sfm/bundler_matching.cc:74-132  for every pair, in the order of :92-93,
//                                                                         matcher->pairwise_match(view_1, view_2, &result)
// This program does exactly that through include/osfm_match.h the way csrc/gpu_exhaustive_matching.h does
// (osfm_match_set_view_f32 from pageable memory with the record stride, then osfm_match_pair into
// std::vector<int> results), with the look-ahead off and on, and the batched osfm_match_pairs beside it.
// bench.py runs it for the `e2e.plugin` figures; no Python in the timed loops.
//
//   plugin_loop_bench <views.u8> <num_views> <descriptors_per_view> [repetitions]
// views.u8: num_views * n * 128 quantised descriptor bytes (the bench's synthetic views).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/osfm_match.h"

namespace {

struct Record { float x, y, scale, orientation, data[128]; };   // sfm::Sift::Descriptor (sift.h:137-149)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc__ = (call);                                                                       \
        if (rc__ != OSFM_OK) {                                                                   \
            std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, osfm_match_last_error(m)); \
            return 2;                                                                            \
        }                                                                                        \
    } while (0)

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <views.u8> <num_views> <n> [repetitions]\n", argv[0]);
        return 1;
    }
    int const nv = std::atoi(argv[2]), n = std::atoi(argv[3]);
    int const reps = argc > 4 ? std::atoi(argv[4]) : 3;
    std::vector<uint8_t> bytes(static_cast<size_t>(nv) * n * 128);
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(bytes.data(), 1, bytes.size(), f) != bytes.size()) {
        std::fprintf(stderr, "cannot read %zu bytes from %s\n", bytes.size(), argv[1]);
        return 1;
    }
    std::fclose(f);
    // the viewports' float descriptors, as the feature extractor leaves them
    std::vector<std::vector<Record>> views(nv, std::vector<Record>(n));
    for (int v = 0; v < nv; ++v)
        for (int i = 0; i < n; ++i) {
            Record& r = views[v][i];
            r.x = r.y = r.scale = r.orientation = 0.0f;
            for (int k = 0; k < 128; ++k) r.data[k] = bytes[(static_cast<size_t>(v) * n + i) * 128 + k] / 255.0f;
        }
    std::vector<int32_t> pairs;
    for (int v1 = 1; v1 < nv; ++v1)
        for (int v2 = 0; v2 < v1; ++v2) { pairs.push_back(v1); pairs.push_back(v2); }
    int const npairs = static_cast<int>(pairs.size() / 2);

    osfm_match_config cfg;
    osfm_match_default_config(&cfg);
    osfm_matcher* m = nullptr;
    CHECK(osfm_match_create(&cfg, &m));

    auto init = [&]() -> int {
        CHECK(osfm_match_begin(m, nv));
        for (int v = 0; v < nv; ++v)
            CHECK(osfm_match_set_view_f32(m, v, views[v][0].data, n, static_cast<int>(sizeof(Record) / sizeof(float)),
                                          nullptr, 0, 0));
        CHECK(osfm_match_commit(m));
        return 0;
    };
    long matches = 0;
    auto loop = [&](int lookahead) -> int {
        CHECK(osfm_match_set_lookahead(m, lookahead));
        matches = 0;
        std::vector<int32_t> m12, m21;
        for (int p = 0; p < npairs; ++p) {
            int const v1 = pairs[2 * p], v2 = pairs[2 * p + 1];
            // as gpu_exhaustive_matching.h: size for the worst case, call, shrink
            m12.resize(static_cast<size_t>(n) + 1);
            m21.resize(static_cast<size_t>(n) + 1);
            int l12 = 0, l21 = 0, cnt = 0;
            CHECK(osfm_match_pair(m, v1, v2, m12.data(), &l12, m21.data(), &l21, &cnt));
            m12.resize(l12);
            m21.resize(l21);
            matches += cnt;
        }
        return 0;
    };
    std::vector<int32_t> dense;
    std::vector<int64_t> offsets(2 * static_cast<size_t>(npairs) + 1);
    std::vector<int32_t> counts(npairs);
    auto batch = [&]() -> int {
        int64_t const total = osfm_match_pairs_result_size(m, pairs.data(), npairs);
        if (total < 0) return 2;
        dense.resize(static_cast<size_t>(total) + 1);
        CHECK(osfm_match_pairs(m, pairs.data(), npairs, dense.data(), offsets.data(), counts.data()));
        matches = 0;
        for (int c : counts) matches += c;
        return 0;
    };

    double t_init = 0, t_pair = 0, t_look = 0, t_batch = 0;
    long m_pair = 0, m_look = 0, m_batch = 0;
    for (int it = 0; it < reps + 1; ++it) {
        double t0 = now_ms();
        if (init()) return 2;
        double t1 = now_ms();
        if (loop(npairs)) return 2;
        double t2 = now_ms();
        m_look = matches;
        if (batch()) return 2;
        double t3 = now_ms();
        m_batch = matches;
        if (it > 0) { t_init += t1 - t0; t_look += t2 - t1; t_batch += t3 - t2; }
    }
    {
        double t0 = now_ms();
        if (loop(0)) return 2;            // pair by pair, every call its own launch sequence: once is enough
        t_pair = now_ms() - t0;
        m_pair = matches;
    }
    osfm_match_destroy(m);
    std::printf("{\"views\": %d, \"n\": %d, \"pairs\": %d, \"repetitions\": %d, \"init_f32_ms\": %.3f, "
                "\"loop_lookahead_ms\": %.3f, \"batch_dense_ms\": %.3f, \"loop_per_pair_ms\": %.3f, "
                "\"matches\": [%ld, %ld, %ld], \"h2d_bytes\": %zu}\n",
                nv, n, npairs, reps, t_init / reps, t_look / reps, t_batch / reps, t_pair, m_look, m_batch, m_pair,
                static_cast<size_t>(nv) * n * sizeof(Record));
    return (m_look == m_batch && m_batch == m_pair) ? 0 : 3;
}
