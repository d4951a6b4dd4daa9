/*
 * cudasift_driver.cu -- extern "C" driver around the reference's own GPU matcher.
 * TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference vendors CudaSift; its matcher is MatchSiftData -> FindMaxCorr10
 * (/root/reference/src/cuda_sift/matching.cu:1090-1206, 301-397): plain FP32 CUDA-core code
 * built for sm_35 there.  oracle/Makefile compiles the UNMODIFIED matching.cu and cudaImage.cu
 * where they lie, for sm_100, and links them with this driver into
 * oracle/_ref/libcudasift_ref.so, so that bench.py can time "the reference's GPU kernel for
 * this path" on the same B200 beside the tcgen05 matcher (the kernel to beat).
 *
 * It is a different algorithm from the parity target (one-way, float, strict '>', no ratio
 * threshold, no cross-check, tail n2 mod 32 skipped: SURVEY Appendix A), so only its time and its
 * arg-max on rows without ties are looked at.
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "cudaSift.h"

namespace {

int fill(SiftData& d, const uint8_t* desc, int n)
{
    d.numPts = n;
    d.maxPts = n;
    d.h_data = static_cast<SiftPoint*>(calloc(static_cast<size_t>(n), sizeof(SiftPoint)));
    if (!d.h_data) return 1;
    for (int i = 0; i < n; ++i) {
        double nrm = 0.0;
        for (int k = 0; k < 128; ++k) nrm += static_cast<double>(desc[i * 128 + k]) * desc[i * 128 + k];
        float const inv = nrm > 0.0 ? static_cast<float>(1.0 / std::sqrt(nrm)) : 0.0f;
        for (int k = 0; k < 128; ++k) d.h_data[i].data[k] = desc[i * 128 + k] * inv;   // unit norm, like CudaSift's own
        d.h_data[i].xpos = static_cast<float>(i);
        d.h_data[i].ypos = 0.0f;
    }
    if (cudaMalloc(reinterpret_cast<void**>(&d.d_data), sizeof(SiftPoint) * static_cast<size_t>(n)) != cudaSuccess) return 2;
    if (cudaMemcpy(d.d_data, d.h_data, sizeof(SiftPoint) * static_cast<size_t>(n), cudaMemcpyHostToDevice) != cudaSuccess) return 3;
    return 0;
}

void release(SiftData& d)
{
    if (d.d_data) cudaFree(d.d_data);
    free(d.h_data);
    d.d_data = nullptr;
    d.h_data = nullptr;
}

}  // namespace

extern "C" {

/* set_1 / set_2: n x 128 quantised descriptors (host).  Runs MatchSiftData(set_1 -> set_2) `reps`
 * times after one warm-up call; ms[0] = mean, ms[1] = min of the times MatchSiftData itself
 * reports (its own cudaEvent timer: CleanMatches + FindMaxCorr10 + the read-back of 5 floats per
 * point).  match_out (n1, may be NULL) receives SiftPoint::match, score_out / ambiguity_out the
 * two floats.  Returns 0 or a small positive error code. */
int osfm_cudasift_match(const uint8_t* set_1, int n1, const uint8_t* set_2, int n2, int reps, double* ms,
                        int32_t* match_out, float* score_out, float* ambiguity_out)
{
    if (!set_1 || !set_2 || n1 <= 0 || n2 <= 0 || reps <= 0 || !ms) return 10;
    SiftData a, b;
    memset(&a, 0, sizeof a);
    memset(&b, 0, sizeof b);
    int rc = fill(a, set_1, n1);
    if (rc == 0) rc = fill(b, set_2, n2);
    if (rc == 0) {
        MatchSiftData(a, b);
        double sum = 0.0, best = 1e30;
        for (int r = 0; r < reps; ++r) {
            double const t = MatchSiftData(a, b);
            sum += t;
            best = t < best ? t : best;
        }
        ms[0] = sum / reps;
        ms[1] = best;
        for (int i = 0; i < n1; ++i) {
            if (match_out) match_out[i] = a.h_data[i].match;
            if (score_out) score_out[i] = a.h_data[i].score;
            if (ambiguity_out) ambiguity_out[i] = a.h_data[i].ambiguity;
        }
    }
    release(a);
    release(b);
    return rc;
}

}  // extern "C"
