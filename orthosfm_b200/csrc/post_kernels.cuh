// post_kernels.cuh -- the small kernels either side of scan_kernel: descriptor
// quantisation, exact finalisation of the per-row top-2 (distance map, Lowe ratio test),
// the bit-exact 16-bit-wrap emulation for the rare rows that need it, the mutual
// (cross-check) filter and the ordered compaction of match lists.
//
// Reference (paths relative to /root/reference):
//   convert_descriptor                      src/mve/sfm/exhaustive_matching.cc:18-39
//   NearestNeighbor<T>::find (distance map) src/mve/sfm/nearest_neighbor.cc:216-268
//   Matching::oneway_match (thresholds)     src/mve/sfm/matching.h:126-144
//   Matching::remove_inconsistent_matches   src/mve/sfm/matching.cc:19-36
//   Matching::count_consistent_matches      src/mve/sfm/matching.cc:39-47
//   Matching::combine_results (offsets)     src/mve/sfm/matching.cc:74-88
#pragma once

#include <cstdint>
#include <climits>
#include <cuda_runtime.h>

#include "scan_kernel.cuh"

namespace osfm {

// ---------------------------------------------------------------- quantisation

// math::round, src/mve/math/functions.h:70-73 (explicit _rn ops: no FMA contraction,
// so the rounding is the host's).
__device__ __forceinline__ float mve_round(float x) {
    return x > 0.0f ? floorf(__fadd_rn(x, 0.5f)) : ceilf(__fadd_rn(x, -0.5f));
}

// in: n x dim floats with row stride `stride`; out: rows of kRowBytes bytes (the tail
// of a SURF row is zero).  SIGNED=false: SIFT (clamp 0..1, x255, unsigned char);
// SIGNED=true: SURF (clamp -1..1, x127, signed char).
template <bool SIGNED>
__global__ void quantize_kernel(const float* __restrict__ in, int n, int dim, int stride,
                                uint8_t* __restrict__ out)
{
    int64_t const idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(n) * kRowBytes) return;
    int const r = static_cast<int>(idx / kRowBytes);
    int const c = static_cast<int>(idx % kRowBytes);
    uint8_t q = 0;
    if (c < dim) {
        float v = in[static_cast<int64_t>(r) * stride + c];
        if (SIGNED) {
            v = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
            v = mve_round(__fmul_rn(v, 127.0f));
            q = static_cast<uint8_t>(static_cast<signed char>(static_cast<int>(v)));
        } else {
            v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            v = mve_round(__fmul_rn(v, 255.0f));
            q = static_cast<uint8_t>(static_cast<int>(v));
        }
    }
    out[idx] = q;
}

// Squared norm of every pool row and its maximum per view (signed kind only: they
// certify that no 16-bit lane of the reference's SSE loop can wrap, see finalize).
__global__ void rownorm_kernel(const uint8_t* __restrict__ pool, int64_t rows,
                               const int32_t* __restrict__ row_view, int32_t* __restrict__ norm2,
                               int32_t* __restrict__ view_max)
{
    int64_t const r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint4* p = reinterpret_cast<const uint4*>(pool + r * kRowBytes);
    int acc = 0;
#pragma unroll
    for (int i = 0; i < kRowBytes / 16; ++i) {
        uint4 const x = __ldg(p + i);
        acc = __dp4a(static_cast<int>(x.x), static_cast<int>(x.x), acc);
        acc = __dp4a(static_cast<int>(x.y), static_cast<int>(x.y), acc);
        acc = __dp4a(static_cast<int>(x.z), static_cast<int>(x.z), acc);
        acc = __dp4a(static_cast<int>(x.w), static_cast<int>(x.w), acc);
    }
    norm2[r] = acc;
    atomicMax(view_max + row_view[r], acc);
}

// ---------------------------------------------------------------- finalisation

struct PostParams {
    const uint8_t* pool;
    const ScanJob* jobs;        // njobs + 1 entries (sentinel: out_row = total_rows)
    int njobs;
    int64_t total_rows;
    const int4* rowres;         // (v1, pos, v2 lower bound) per job row
    int32_t* oneway;            // out: index of the match in the candidate view or -1
    float sq_lowe;              // lowe_ratio_threshold^2   (matching.h:126)
    float sq_dist;              // distance_threshold^2     (matching.h:127)
    int64_t* slow_list;         // rows that need the wrap emulation
    unsigned long long* counters;  // [0] slow rows of this batch, [1] candidate rows,
                                   // [2] self-check failures, [3] slow rows (cumulative)
    const int32_t* norm2;       // signed kind: squared norm per pool row
};

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

// matching.h:138-143.  The quotient is an IEEE float division; 0/0 = NaN compares
// false and therefore accepts.
__device__ __forceinline__ bool passes_tests(int d1, int d2, float sq_lowe, float sq_dist) {
    float const f1 = static_cast<float>(d1);
    float const f2 = static_cast<float>(d2);
    if (f1 > sq_dist) return false;
    if (__fdiv_rn(f1, f2) > sq_lowe) return false;
    return true;
}

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

// One thread per job row.  Rows whose ratio test fails even with the lower bound on
// the second-best similarity are final.  The others are re-examined by the whole warp:
// lane l recomputes the similarity with candidate pos*32 + l, which yields the exact
// arg-max (highest index on ties) and the exact second best.
template <bool SIGNED>
__global__ void __launch_bounds__(256) finalize_kernel(PostParams p)
{
    int64_t const g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int const lane = threadIdx.x & 31;
    bool const valid = g < p.total_rows;

    int q_prow = 0, c_row = 0, c_n = 0, v1 = 0, pos = 0, v2 = 0, result = -1;
    bool cand = false, slow = false;
    if (valid) {
        ScanJob const job = p.jobs[find_job(p.jobs, p.njobs, g)];
        q_prow = job.q_row + static_cast<int>(g - job.out_row);
        c_row = job.c_row;
        c_n = job.c_n;
        int4 const rr = p.rowres[g];
        v1 = rr.x; pos = rr.y; v2 = rr.z;
        if (SIGNED) {
            // No 16-bit lane can wrap if |a||b| < 2^15 (Cauchy-Schwarz per lane).
            slow = static_cast<int64_t>(p.norm2[q_prow]) * static_cast<int64_t>(job.c_maxnorm2)
                   >= (1ll << 30);
            if (!slow) {
                if (v1 < 0) {
                    // no candidate reached the initial best of 0: index stays 0
                    int const d = ip_to_dist<true>(0);
                    result = passes_tests(d, d, p.sq_lowe, p.sq_dist) ? 0 : -1;
                } else {
                    cand = passes_tests(ip_to_dist<true>(v1), ip_to_dist<true>(v2), p.sq_lowe, p.sq_dist);
                }
            }
        } else {
            // Any similarity >= 2^16 makes the reference's 16-bit lanes / stores wrap.
            slow = v1 >= 65536;
            if (!slow)
                cand = passes_tests(ip_to_dist<false>(v1), ip_to_dist<false>(v2), p.sq_lowe, p.sq_dist);
        }
        if (slow) {
            unsigned long long const k = atomicAdd(p.counters + 0, 1ull);
            p.slow_list[k] = g;
        }
    }

    unsigned cmask = __ballot_sync(0xffffffffu, cand);
    if (lane == 0 && cmask != 0) atomicAdd(p.counters + 1, static_cast<unsigned long long>(__popc(cmask)));
    while (cmask != 0) {
        int const src = __ffs(cmask) - 1;
        cmask &= cmask - 1;
        int const b_q = __shfl_sync(0xffffffffu, q_prow, src);
        int const b_crow = __shfl_sync(0xffffffffu, c_row, src);
        int const b_cn = __shfl_sync(0xffffffffu, c_n, src);
        int const b_v1 = __shfl_sync(0xffffffffu, v1, src);
        int const b_pos = __shfl_sync(0xffffffffu, pos, src);
        int const b_v2 = __shfl_sync(0xffffffffu, v2, src);

        int const col = b_pos * kChunk + lane;
        int dot = INT_MIN / 2;
        if (col < b_cn)
            dot = dot_row<SIGNED>(p.pool + static_cast<int64_t>(b_q) * kRowBytes,
                                  p.pool + (static_cast<int64_t>(b_crow) + col) * kRowBytes);
        unsigned const eq = __ballot_sync(0xffffffffu, dot == b_v1);
        int const jl = 31 - __clz(eq);  // highest index wins ties (nearest_neighbor.cc:89)
        int const second = __reduce_max_sync(0xffffffffu, lane == jl ? INT_MIN / 2 : dot);
        if (lane == src) {
            if (eq == 0) {
                atomicAdd(p.counters + 2, 1ull);  // the scan and the refine disagree: a bug
                result = -1;
            } else {
                int const s2 = max(b_v2, second);
                result = passes_tests(ip_to_dist<SIGNED>(b_v1), ip_to_dist<SIGNED>(s2), p.sq_lowe, p.sq_dist)
                             ? b_pos * kChunk + jl : -1;
            }
        }
    }
    if (valid && !slow) p.oneway[g] = result;
}

// ---------------------------------------------------------------- wrap emulation

// The reference's inner product as its SSE2 loop computes it (nearest_neighbor.cc:75-84):
// eight 16-bit lanes, lane k summing elements k, k+8, ... modulo 2^16, then added as int.
template <bool SIGNED>
__device__ __forceinline__ int wrapped_ip(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b) {
    unsigned s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint2* pa = reinterpret_cast<const uint2*>(a);
    const uint2* pb = reinterpret_cast<const uint2*>(b);
    for (int t = 0; t < kRowBytes / 8; ++t) {
        uint2 const x = __ldg(pa + t);
        uint2 const y = __ldg(pb + t);
        unsigned const xa[2] = {x.x, x.y};
        unsigned const ya[2] = {y.x, y.y};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned const xb = (xa[k >> 2] >> (8 * (k & 3))) & 0xffu;
            unsigned const yb = (ya[k >> 2] >> (8 * (k & 3))) & 0xffu;
            if (SIGNED)
                s[k] += static_cast<unsigned>(static_cast<int>(static_cast<signed char>(xb)) *
                                              static_cast<int>(static_cast<signed char>(yb)));
            else
                s[k] += xb * yb;
        }
    }
    int ip = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        ip += SIGNED ? static_cast<int>(static_cast<short>(s[k] & 0xffffu))
                     : static_cast<int>(s[k] & 0xffffu);
    return ip;
}

// One warp per flagged row: replays the reference's sequential scan including the
// truncating 16-bit stores of best / second best (nearest_neighbor.cc:87-100).
template <bool SIGNED>
__global__ void __launch_bounds__(256) slow_rows_kernel(PostParams p)
{
    // The list length is only known on the device (written by finalize_kernel, which
    // precedes this launch in stream order); the grid strides over it.
    int64_t const nslow = static_cast<int64_t>(*reinterpret_cast<volatile unsigned long long*>(p.counters + 0));
    if (blockIdx.x == 0 && threadIdx.x == 0 && nslow > 0)
        atomicAdd(p.counters + 3, static_cast<unsigned long long>(nslow));
    int const lane = threadIdx.x & 31;
    int64_t const nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < nslow; w += nwarps) {
        int64_t const g = p.slow_list[w];
        ScanJob const job = p.jobs[find_job(p.jobs, p.njobs, g)];
        const uint8_t* q = p.pool + (static_cast<int64_t>(job.q_row) + (g - job.out_row)) * kRowBytes;

        int b1 = 0, b2 = 0, i1 = 0;
        for (int base = 0; base < job.c_n; base += 32) {
            int const col = base + lane;
            int ip = 0;
            if (col < job.c_n)
                ip = wrapped_ip<SIGNED>(q, p.pool + (static_cast<int64_t>(job.c_row) + col) * kRowBytes);
            int const lim = min(32, job.c_n - base);
            for (int l = 0; l < lim; ++l) {
                int const x = __shfl_sync(0xffffffffu, ip, l);
                if (x >= b2) {
                    int const stored = SIGNED ? static_cast<int>(static_cast<short>(x & 0xffff))
                                              : (x & 0xffff);
                    if (x >= b1) { b2 = b1; b1 = stored; i1 = base + l; }
                    else         { b2 = stored; }
                }
            }
        }
        if (lane == 0)
            p.oneway[g] = passes_tests(ip_to_dist<SIGNED>(b1), ip_to_dist<SIGNED>(b2), p.sq_lowe, p.sq_dist)
                              ? i1 : -1;
    }
}

// ---------------------------------------------------------------- mutual filter

// One feature kind of one image pair.  in12 / in21 index oneway[] (or -1 when that
// direction was not run because a set is empty: every entry is then -1,
// matching.h:121-124); out12 / out21 index the caller-visible dense result; add12 / add21
// are combine_results' index shifts (matching.cc:78-88).
struct PairPart {
    int64_t in12, in21;
    int64_t out12, out21;
    int32_t n1, n2;
    int32_t add12, add21;
    int32_t pair;
    int32_t pad;
};

constexpr int kMutualChunk = 4096;

__global__ void __launch_bounds__(256) mutual_kernel(const PairPart* __restrict__ parts,
                                                     const int32_t* __restrict__ oneway,
                                                     int32_t* __restrict__ out,
                                                     int32_t* __restrict__ counts)
{
    PairPart const pp = parts[blockIdx.x];
    int const lo = blockIdx.y * kMutualChunk;
    int const lane = threadIdx.x & 31;
    // 1 -> 2, counting consistent matches (warp-aggregated atomic)
    int kept = 0;
    for (int i = lo + threadIdx.x; i < min(pp.n1, lo + kMutualChunk); i += blockDim.x) {
        int m = pp.in12 >= 0 ? oneway[pp.in12 + i] : -1;
        if (m >= 0 && oneway[pp.in21 + m] != i) m = -1;
        out[pp.out12 + i] = m >= 0 ? m + pp.add12 : -1;
        kept += m >= 0;
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    if (lane == 0 && kept > 0 && counts != nullptr) atomicAdd(counts + pp.pair, kept);
    // 2 -> 1
    for (int i = lo + threadIdx.x; i < min(pp.n2, lo + kMutualChunk); i += blockDim.x) {
        int m = pp.in21 >= 0 ? oneway[pp.in21 + i] : -1;
        if (m >= 0 && oneway[pp.in12 + m] != i) m = -1;
        out[pp.out21 + i] = m >= 0 ? m + pp.add21 : -1;
    }
}

// Copies the unfiltered one-way results (Matching::twoway_match) to the dense output.
__global__ void __launch_bounds__(256) copy_twoway_kernel(const PairPart* __restrict__ parts,
                                                          const int32_t* __restrict__ oneway,
                                                          int32_t* __restrict__ out)
{
    PairPart const pp = parts[blockIdx.x];
    int const lo = blockIdx.y * kMutualChunk;
    for (int i = lo + threadIdx.x; i < min(pp.n1, lo + kMutualChunk); i += blockDim.x)
        out[pp.out12 + i] = pp.in12 >= 0 ? oneway[pp.in12 + i] : -1;
    for (int i = lo + threadIdx.x; i < min(pp.n2, lo + kMutualChunk); i += blockDim.x)
        out[pp.out21 + i] = pp.in21 >= 0 ? oneway[pp.in21 + i] : -1;
}

// ---------------------------------------------------------------- compaction

// One CTA per pair part: the surviving (i, j) of matches_1_2 in ascending i -- the
// order in which the reference builds its correspondence list
// (src/mve/sfm/bundler_matching.cc:178-192) -- written to list[list_offset[pair] ...].
// Ranks inside a warp come from a ballot, warp totals from a shared-memory scan.
__global__ void __launch_bounds__(1024) compact_kernel(const PairPart* __restrict__ parts,
                                                       const int32_t* __restrict__ dense,
                                                       const int64_t* __restrict__ list_offset,
                                                       int2* __restrict__ list)
{
    __shared__ int warp_tot[32];
    __shared__ int running;
    PairPart const pp = parts[blockIdx.x];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t const base = list_offset[pp.pair];
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int start = 0; start < pp.n1; start += blockDim.x) {
        int const i = start + threadIdx.x;
        int const m = i < pp.n1 ? dense[pp.out12 + i] : -1;
        unsigned const b = __ballot_sync(0xffffffffu, m >= 0);
        if (lane == 0) warp_tot[warp] = __popc(b);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        int total = 0;
        if (threadIdx.x == 0)
            for (int w = 0; w < 32; ++w) total += warp_tot[w];
        int const run = running;
        if (m >= 0) {
            int const rank = run + before + __popc(b & ((1u << lane) - 1u));
            list[base + rank] = make_int2(i, m);
        }
        __syncthreads();
        if (threadIdx.x == 0) running = run + total;
        __syncthreads();
    }
}

}  // namespace osfm
