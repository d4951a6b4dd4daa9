// ransac_kernels.cuh -- RANSAC for the fundamental matrix of every candidate pair, batched
// (SURVEY.md section 8, row f2).
//
// What it replaces (reference, paths relative to /root/reference/src/mve):
//   sfm::RansacFundamental::estimate / estimate_8_point / find_inliers
//   (sfm/ransac_fundamental.cc:26-105) as called per pair by
//   bundler::Matching::two_view_matching (sfm/bundler_matching.cc:176-220).
//
// The reference draws its samples from the process-wide std::rand() sequence, pair after
// pair, so the draws are made on the host in that order (osfm_ransac_draw_samples) and come
// here as a table of eight ascending match indices per (pair, iteration).  Everything else is
// independent across (pair, iteration) and runs here:
//   fit     9 x 9 and 3 x 3 SVD per (pair, iteration) in three stages (ransac_math.cuh):
//           bidiagonalise / iterate with work fetching / rank 2
//   count   one warp per (pair, iteration): Sampson distance of every match of the pair
//   select  one CTA per pair: the first iteration with the most inliers wins (the reference
//           replaces its best only on a strictly larger count); its inliers, in match order,
//           become the pair's filtered list
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "ransac_math.cuh"

namespace osfm {

// xy[e] = (x1, y1, x2, y2) of match e: the positions of its two features.
__global__ void __launch_bounds__(256) ransac_gather_kernel(const int32_t* __restrict__ pair_views,
                                                            const int64_t* __restrict__ list_offset, int npairs,
                                                            const int2* __restrict__ ij, int64_t nmatches,
                                                            const int64_t* __restrict__ view_base,
                                                            const int32_t* __restrict__ view_n,
                                                            const float2* __restrict__ positions,
                                                            float4* __restrict__ xy, int* __restrict__ bad)
{
    int64_t const e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= nmatches) return;
    int lo = 0, hi = npairs;
    while (hi - lo > 1) {
        int const mid = (lo + hi) >> 1;
        if (list_offset[mid] <= e) lo = mid; else hi = mid;
    }
    int const v1 = pair_views[2 * lo], v2 = pair_views[2 * lo + 1];
    int2 const m = ij[e];
    if (m.x < 0 || m.x >= view_n[v1] || m.y < 0 || m.y >= view_n[v2]) {
        atomicAdd(bad, 1);
        xy[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    float2 const a = positions[view_base[v1] + m.x];
    float2 const b = positions[view_base[v2] + m.y];
    xy[e] = make_float4(a.x, a.y, b.x, b.y);
}

// ---- the fit, in three stages --------------------------------------------------------------------
// Two thirds of the design matrices never meet the reference's convergence test (their zero
// singular value is not deflated) and would run all 81 trips of its loop; nearly all of those
// sit on a fixed point after about 30 trips, where the loop may stop without changing anything
// (ransac_math.cuh).  One thread per fit from start to end cannot use that: a warp waits for its
// slowest lane (measured: 7.7 of 32 lanes active on average).  So the loop gets a kernel of its
// own in which a lane that has finished a fit fetches the next one.

constexpr int kBidiagDoubles = 17;     // diagonal (9) and superdiagonal (8)
constexpr int kGkThreads = 128;
constexpr int kGkBeginBatch = 24;      // lanes that must wait at a trip boundary before the trips start

// Stage 1, one thread per (pair, iteration): design matrix and Householder bidiagonalisation.
// samples: 8 ascending match indices (relative to the pair's list) per thread.  bd / vv:
// element-major ([element][fit]) so that neighbouring threads write neighbouring words.
__global__ void __launch_bounds__(128) ransac_bidiag_kernel(const int64_t* __restrict__ list_offset, int npairs,
                                                            int iterations, const int32_t* __restrict__ samples,
                                                            const float4* __restrict__ xy, double* __restrict__ bd,
                                                            double* __restrict__ vv, int* __restrict__ bad)
{
    int64_t const total = static_cast<int64_t>(npairs) * iterations;
    int64_t const t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int const pair = static_cast<int>(t / iterations);
    int64_t const begin = list_offset[pair];
    int64_t const count = list_offset[pair + 1] - begin;
    double p1[16], p2[16];
    int prev = -1;
    bool ok = true;
    for (int k = 0; k < 8; ++k) {
        int const s = samples[t * 8 + k];
        ok = ok && s > prev && s < count;
        prev = s;
        float4 const m = ok ? xy[begin + s] : make_float4(0.f, 0.f, 0.f, 0.f);
        p1[2 * k] = m.x; p1[2 * k + 1] = m.y;
        p2[2 * k] = m.z; p2[2 * k + 1] = m.w;
    }
    if (!ok) {
        atomicAdd(bad, 1);
        for (int k = 0; k < 16; ++k) { p1[k] = 0.0; p2[k] = 0.0; }
    }
    fmath::SquareSvd<9, false> svd;
    fmath::design_matrix(p1, p2, svd.bm);
    svd.bidiagonalize(fmath::kSvdEpsilon);
    for (int i = 0; i < 9; ++i) bd[i * total + t] = svd.B(i, i);
    for (int i = 0; i < 8; ++i) bd[(9 + i) * total + t] = svd.B(i, i + 1);
    for (int i = 0; i < 81; ++i) vv[i * total + t] = svd.vm.at(i);
}

// Stage 2, persistent.  Measured on the first version of this kernel (one trip of the loop per
// pass): a trip's sweep has 1 to 8 rotation steps, most matrices are down to one step while
// some lane of the warp still has eight, and 4.6 of 32 lanes were active in the rotations.  So
// the unit of work per pass is ONE rotation step: a lane at a trip boundary runs the trip's
// start (zero tests, block choice, shift), every lane with steps pending runs one, and a lane
// whose fit has ended takes the next fit from `next`.  B lives in shared memory (81 doubles
// per lane, strided by the CTA size: conflict-free whatever element a lane touches), V in local
// memory.  fvec: the singular vector of the smallest singular value, 9 doubles per fit.
__global__ void __launch_bounds__(kGkThreads) ransac_gk_kernel(int64_t total, const double* __restrict__ bd,
                                                               const double* __restrict__ vv,
                                                               unsigned long long* __restrict__ next,
                                                               double* __restrict__ fvec, int begin_batch)
{
    extern __shared__ double gk_shared[];
    fmath::SquareSvd<9, false, fmath::StridedMatrix> svd;
    svd.bm.p = gk_shared + threadIdx.x;
    svd.bm.stride = static_cast<int>(blockDim.x);
    svd.sweep_k = svd.sweep_end = 0;
    int64_t t = -1;
    int trips = 0;
    bool active = false, exhausted = false;
    while (true) {
        if (!active && !exhausted) {
            t = static_cast<int64_t>(atomicAdd(next, 1ull));
            if (t < total) {
                for (int i = 0; i < 81; ++i) svd.bm.at(i) = 0.0;
                for (int i = 0; i < 9; ++i) svd.B(i, i) = bd[i * total + t];
                for (int i = 0; i < 8; ++i) svd.B(i, i + 1) = bd[(9 + i) * total + t];
                for (int i = 0; i < 81; ++i) svd.vm.at(i) = vv[i * total + t];
                trips = 0;
                svd.sweep_k = svd.sweep_end = 0;
                active = true;
            } else {
                exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        // A trip's start costs three rotation steps' worth of instructions and only the lanes at
        // a trip boundary take part (measured: 6 of 32 when it ran in every pass).  So it runs
        // only when enough lanes wait at a boundary, or when no lane has a rotation left.
        bool const boundary = active && !svd.sweep_pending();
        int const waiting = __popc(__ballot_sync(0xffffffffu, boundary));
        bool const rotating = __any_sync(0xffffffffu, active && svd.sweep_pending());
        bool const start_trips = waiting >= begin_batch || !rotating;
        bool done = false;
        if (boundary && start_trips) {
            ++trips;
            done = svd.trip_begin(fmath::kSvdEpsilon);
            if (!done && !svd.sweep_pending())           // a trip without rotation steps
                done = !svd.changed() || trips >= 81;
        }
        if (active && !done && svd.sweep_pending()) {
            svd.sweep_rotate(fmath::kSvdEpsilon);
            if (!svd.sweep_pending())                    // the trip is complete
                done = !svd.changed() || trips >= 81;
        }
        if (done) {
            svd.finish(fmath::kSvdEpsilon);
            for (int r = 0; r < 9; ++r) fvec[t * 9 + r] = svd.V(r, 8);
            active = false;
        }
    }
}

// Stage 3, one thread per fit: rank 2 enforced (3 x 3 SVD with U), in place.
__global__ void __launch_bounds__(128) ransac_rank2_kernel(int64_t total, double* __restrict__ F)
{
    int64_t const t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    double f[9];
    for (int k = 0; k < 9; ++k) f[k] = F[t * 9 + k];
    fmath::enforce_rank2<true>(f);
    for (int k = 0; k < 9; ++k) F[t * 9 + k] = f[k];
}

// One warp per (pair, iteration): inliers[t] = #{ matches with sampson < thr2 }.
__global__ void __launch_bounds__(256) ransac_count_kernel(const int64_t* __restrict__ list_offset, int npairs,
                                                           int iterations, const float4* __restrict__ xy,
                                                           const double* __restrict__ F, double thr2,
                                                           int* __restrict__ inliers)
{
    int64_t const t = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int const lane = threadIdx.x & 31;
    if (t >= static_cast<int64_t>(npairs) * iterations) return;
    int const pair = static_cast<int>(t / iterations);
    int64_t const begin = list_offset[pair];
    int const count = static_cast<int>(list_offset[pair + 1] - begin);
    double f[9];
    for (int k = 0; k < 9; ++k) f[k] = F[t * 9 + k];
    int n = 0;
    for (int i = lane; i < count; i += 32) {
        float4 const m = xy[begin + i];
        n += fmath::sampson_distance(f, m.x, m.y, m.z, m.w) < thr2 ? 1 : 0;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) inliers[t] = n;
}

// One CTA (256 threads) per pair.
__global__ void __launch_bounds__(256) ransac_select_kernel(const int64_t* __restrict__ list_offset, int iterations,
                                                            const float4* __restrict__ xy, const int2* __restrict__ ij,
                                                            const double* __restrict__ F,
                                                            const int* __restrict__ inliers, double thr2,
                                                            int2* __restrict__ out_ij, int* __restrict__ out_count,
                                                            double* __restrict__ out_F)
{
    __shared__ unsigned long long s_best[8];
    __shared__ int s_warp[8];
    __shared__ int s_base;
    int const pair = blockIdx.x;
    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t const begin = list_offset[pair];
    int const count = static_cast<int>(list_offset[pair + 1] - begin);
    // most inliers, earliest iteration: max of (inliers << 32 | ~iteration)
    unsigned long long best = 0;
    for (int it = tid; it < iterations; it += 256) {
        unsigned long long const key = (static_cast<unsigned long long>(inliers[static_cast<int64_t>(pair) * iterations + it]) << 32) |
                                       static_cast<uint32_t>(~static_cast<uint32_t>(it));
        best = key > best ? key : best;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long const other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    best = s_best[0];
    for (int w = 1; w < 8; ++w) best = s_best[w] > best ? s_best[w] : best;
    int const best_count = static_cast<int>(best >> 32);
    int const best_it = static_cast<int>(~static_cast<uint32_t>(best & 0xffffffffu));
    if (best_count == 0) {       // the reference keeps an empty result
        if (tid == 0) out_count[pair] = 0;
        if (tid < 9) out_F[static_cast<int64_t>(pair) * 9 + tid] = 0.0;
        return;
    }
    double f[9];
    for (int k = 0; k < 9; ++k) f[k] = F[(static_cast<int64_t>(pair) * iterations + best_it) * 9 + k];
    if (tid < 9) out_F[static_cast<int64_t>(pair) * 9 + tid] = f[tid];
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < count; i0 += 256) {
        int const i = i0 + tid;
        bool inl = false;
        if (i < count) {
            float4 const m = xy[begin + i];
            inl = fmath::sampson_distance(f, m.x, m.y, m.z, m.w) < thr2;
        }
        unsigned const ballot = __ballot_sync(0xffffffffu, inl);
        if (lane == 0) s_warp[warp] = __popc(ballot);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (inl) out_ij[begin + before + __popc(ballot & ((1u << lane) - 1u))] = ij[begin + i];
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w = 0; w < 8; ++w) total += s_warp[w];
            s_base += total;
        }
        __syncthreads();
    }
    if (tid == 0) out_count[pair] = s_base;
}

}  // namespace osfm
