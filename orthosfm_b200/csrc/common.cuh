// common.cuh -- constants, the job record and the small exact-arithmetic helpers shared by
// the scan kernel and the post-processing kernels.
#pragma once

#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

namespace osfm {

constexpr int kRowBytes = 128;   // descriptor row pitch in the pool (SIFT 128 B; SURF zero-padded)
constexpr int kHalfM = 128;      // rows per tcgen05.mma (M) = TMEM lanes
constexpr int kItemM = 256;      // query rows per work item (two M halves share each candidate tile)
constexpr int kBlockN = 256;     // candidate rows per tile (= TMEM columns per accumulator)
constexpr int kChunk = 32;       // candidates per tcgen05.ld.x32
constexpr int kChunksPerTile = kBlockN / kChunk;

constexpr int kMasked = -(1 << 24);   // similarity of a column past the end of the view: below
                                      // any real value (|s| < 2^23) and kMasked * 8 fits an int

// One direction of one image pair.  `item_start` is the exclusive prefix sum of
// ceil(q_n / 256) over the job list; the list carries one sentinel entry at the end.
struct ScanJob {
    int32_t q_row;       // first row of the query set in the query pool
    int32_t q_n;         // number of query descriptors
    int32_t c_row;       // first pool row of the candidate view
    int32_t c_n;         // number of candidate descriptors
    int64_t out_row;     // first index of this job's rows in rowres[] / oneway[] (or xrow_map[])
    int32_t item_start;  // first work item of this job
    int32_t c_view;      // candidate view id (indexes the per-view largest squared norm)
};

__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }

__device__ __forceinline__ int find_job(const ScanJob* __restrict__ jobs, int njobs, int64_t g) {
    int lo = 0, hi = njobs;
    while (hi - lo > 1) {
        int const mid = (lo + hi) >> 1;
        if (jobs[mid].out_row <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// nearest_neighbor.cc:262-267 (unsigned) and :234-237 (signed): inner product -> distance.
template <bool SIGNED>
__device__ __forceinline__ int ip_to_dist(int ip) {
    if (SIGNED) {
        int const x = min(16129, max(0, ip));
        return 32258 - 2 * x;
    } else {
        int const x = 65025 - min(65025, ip);
        return min(32767, x) * 2;
    }
}

// matching.h:138-143.  The quotient is an IEEE float division; 0/0 = NaN compares false and
// therefore accepts.
__device__ __forceinline__ bool passes_tests(int d1, int d2, float sq_lowe, float sq_dist) {
    float const f1 = static_cast<float>(d1);
    float const f2 = static_cast<float>(d2);
    if (f1 > sq_dist) return false;
    if (__fdiv_rn(f1, f2) > sq_lowe) return false;
    return true;
}

// Exact integer dot product of two pool rows.
template <bool SIGNED>
__device__ __forceinline__ int dot_row(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b) {
    const uint4* pa = reinterpret_cast<const uint4*>(a);
    const uint4* pb = reinterpret_cast<const uint4*>(b);
    int acc = 0;
#pragma unroll
    for (int i = 0; i < kRowBytes / 16; ++i) {
        uint4 const x = __ldg(pa + i);
        uint4 const y = __ldg(pb + i);
        if (SIGNED) {
            acc = __dp4a(static_cast<int>(x.x), static_cast<int>(y.x), acc);
            acc = __dp4a(static_cast<int>(x.y), static_cast<int>(y.y), acc);
            acc = __dp4a(static_cast<int>(x.z), static_cast<int>(y.z), acc);
            acc = __dp4a(static_cast<int>(x.w), static_cast<int>(y.w), acc);
        } else {
            unsigned u = static_cast<unsigned>(acc);
            u = __dp4a(x.x, y.x, u);
            u = __dp4a(x.y, y.y, u);
            u = __dp4a(x.z, y.z, u);
            u = __dp4a(x.w, y.w, u);
            acc = static_cast<int>(u);
        }
    }
    return acc;
}

// The reference's inner product as its SSE2 loop computes it (nearest_neighbor.cc:75-84):
// eight 16-bit lanes, lane k summing elements k, k+8, ... modulo 2^16, then added as int.
template <bool SIGNED>
__device__ __noinline__ int wrapped_ip(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b) {
    // all 16 loads are issued before the first use: one memory latency, not sixteen
    uint4 x[kRowBytes / 16], y[kRowBytes / 16];
#pragma unroll
    for (int i = 0; i < kRowBytes / 16; ++i) {
        x[i] = __ldg(reinterpret_cast<const uint4*>(a) + i);
        y[i] = __ldg(reinterpret_cast<const uint4*>(b) + i);
    }
    unsigned s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < kRowBytes / 16; ++i) {
        unsigned const xa[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
        unsigned const ya[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
#pragma unroll
        for (int e = 0; e < 16; ++e) {   // byte e of this 16-byte group belongs to lane e % 8
            unsigned const xb = (xa[e >> 2] >> (8 * (e & 3))) & 0xffu;
            unsigned const yb = (ya[e >> 2] >> (8 * (e & 3))) & 0xffu;
            if (SIGNED)
                s[e & 7] += static_cast<unsigned>(static_cast<int>(static_cast<signed char>(xb)) *
                                                  static_cast<int>(static_cast<signed char>(yb)));
            else
                s[e & 7] += xb * yb;
        }
    }
    int ip = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        ip += SIGNED ? static_cast<int>(static_cast<short>(s[k] & 0xffffu))
                     : static_cast<int>(s[k] & 0xffffu);
    return ip;
}

// One step of the reference's sequential best / second-best scan including the truncating
// 16-bit stores (nearest_neighbor.cc:87-100).  x is the (wrapped) inner product as int.
template <bool SIGNED>
__device__ __forceinline__ void ref_scan_step(int x, int index, int& b1, int& b2, int& i1) {
    if (x >= b2) {
        int const stored = SIGNED ? static_cast<int>(static_cast<short>(x & 0xffff)) : (x & 0xffff);
        if (x >= b1) { b2 = b1; b1 = stored; i1 = index; }
        else         { b2 = stored; }
    }
}

}  // namespace osfm
