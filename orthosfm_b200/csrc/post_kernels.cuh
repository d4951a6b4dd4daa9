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
    const int4* rowres;         // (v1, pos, v2 lower bound, job index) per job row
    int32_t* oneway;            // out: index of the match in the candidate view or -1
    float sq_lowe;              // lowe_ratio_threshold^2   (matching.h:126)
    float sq_dist;              // distance_threshold^2     (matching.h:127)
    int64_t* slow_list;         // rows that need the wrap emulation
    unsigned long long* counters;  // [0] slow rows of this batch (signed kind), [1] candidate
                                   // rows (cumulative), [2] self-check failures, [3] slow rows
                                   // (cumulative), [4] candidate rows of this batch
    const int32_t* norm2;       // signed kind: squared norm per pool row
    int* slow_cnt;              // unsigned kind: slow rows per job; the rows of job j are
                                // listed at slow_list[jobs[j].out_row + 0 .. slow_cnt[j])
    int64_t* cand_list;         // rows that need the exact second best; length = counters[4]
};

// classify_kernel: one thread per job row.  Rows whose ratio test fails even with the lower
// bound on the second-best similarity are final (-1).  Rows that pass become *candidates*
// and are appended to a list (warp-aggregated atomic); rows whose best similarity reached
// 2^16 (unsigned) or whose norms cannot exclude a 16-bit lane wrap (signed) go to the slow
// list instead.
template <bool SIGNED>
__global__ void __launch_bounds__(256) classify_kernel(PostParams p)
{
    int64_t const g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int const lane = threadIdx.x & 31;
    bool cand = false;
    if (g < p.total_rows) {
        int4 const rr = p.rowres[g];
        int const v1 = rr.x, v2 = rr.z;
        int const ji = rr.w;                 // the scan kernel recorded the row's job
        ScanJob const job = p.jobs[ji];
        int result = -1;
        bool slow;
        if (SIGNED) {
            // No 16-bit lane can wrap if |a||b| < 2^15 (Cauchy-Schwarz per lane).
            int const q_prow = job.q_row + static_cast<int>(g - job.out_row);
            slow = static_cast<int64_t>(p.norm2[q_prow]) * static_cast<int64_t>(job.c_maxnorm2) >= (1ll << 30);
            if (!slow) {
                if (v1 < 0) {
                    // no candidate reached the initial best of 0: index stays 0
                    int const d = ip_to_dist<true>(0);
                    result = passes_tests(d, d, p.sq_lowe, p.sq_dist) ? 0 : -1;
                } else {
                    cand = passes_tests(ip_to_dist<true>(v1), ip_to_dist<true>(v2), p.sq_lowe, p.sq_dist);
                }
            }
            if (slow) p.slow_list[atomicAdd(p.counters + 0, 1ull)] = g;
        } else {
            // Any similarity >= 2^16 makes the reference's 16-bit lanes / stores wrap.
            slow = v1 >= 65536;
            if (!slow)
                cand = passes_tests(ip_to_dist<false>(v1), ip_to_dist<false>(v2), p.sq_lowe, p.sq_dist);
            else
                p.slow_list[job.out_row + atomicAdd(p.slow_cnt + ji, 1)] = g;
        }
        if (!slow && !cand) p.oneway[g] = result;
    }
    unsigned const cmask = __ballot_sync(0xffffffffu, cand);
    if (cmask != 0) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.counters + 4, static_cast<unsigned long long>(__popc(cmask)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) p.cand_list[base + __popc(cmask & ((1u << lane) - 1u))] = g;
    }
}

// refine_kernel: half a warp per candidate row.  Lane l recomputes the similarity with
// candidate pos*16 + l, which yields the exact arg-max (highest index on ties) and the exact
// second best; then the reference's tests decide.  Grid-strides over the candidate list,
// whose length is only known on the device.
template <bool SIGNED>
__global__ void __launch_bounds__(256) refine_kernel(PostParams p)
{
    unsigned long long const ncand = *reinterpret_cast<volatile unsigned long long*>(p.counters + 4);
    if (blockIdx.x == 0 && threadIdx.x == 0 && ncand > 0) atomicAdd(p.counters + 1, ncand);
    int const lane = threadIdx.x & 31;
    int const sub = lane & 15;
    unsigned const hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
    int const hshift = lane & 16;
    unsigned long long const nhalf = (static_cast<unsigned long long>(gridDim.x) * blockDim.x) >> 4;
    unsigned long long k = (static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 4;
    // both halves of a warp must iterate the same number of times (full-mask ballots)
    unsigned long long const kmax = ((ncand + 1) >> 1) << 1;
    for (; k < kmax; k += nhalf) {
        bool const active = k < ncand;
        int dot = INT_MIN / 2, v1 = 0, v2 = 0, pos = 0;
        int64_t g = 0;
        if (active) {
            g = p.cand_list[k];
            int4 const rr = p.rowres[g];
            ScanJob const job = p.jobs[rr.w];
            v1 = rr.x; pos = rr.y; v2 = rr.z;
            int const col = pos * kSub + sub;
            if (col < job.c_n)
                dot = dot_row<SIGNED>(p.pool + (static_cast<int64_t>(job.q_row) + (g - job.out_row)) * kRowBytes,
                                      p.pool + (static_cast<int64_t>(job.c_row) + col) * kRowBytes);
        }
        unsigned const eq = (__ballot_sync(0xffffffffu, active && dot == v1) & hmask) >> hshift;
        int const jl = 31 - __clz(eq);   // highest index wins ties (nearest_neighbor.cc:89); -1 if none
        int const second = __reduce_max_sync(hmask, sub == jl ? INT_MIN / 2 : dot);
        if (active && sub == 0) {
            if (eq == 0) {
                atomicAdd(p.counters + 2, 1ull);   // the scan and the refine disagree: a bug
                p.oneway[g] = -1;
            } else {
                int const s2 = max(v2, second);
                p.oneway[g] = passes_tests(ip_to_dist<SIGNED>(v1), ip_to_dist<SIGNED>(s2), p.sq_lowe, p.sq_dist)
                                  ? pos * kSub + jl : -1;
            }
        }
    }
}

// ---------------------------------------------------------------- wrap emulation

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
                ref_scan_step<SIGNED>(x, base + l, b1, b2, i1);
            }
        }
        if (lane == 0)
            p.oneway[g] = passes_tests(ip_to_dist<SIGNED>(b1), ip_to_dist<SIGNED>(b2), p.sq_lowe, p.sq_dist)
                              ? i1 : -1;
    }
}

// ---------------------------------------------------------------- exact pass set-up

// Single CTA.  Turns the per-job slow-row counts into the job list of the EXACT scan pass.
// Jobs arrive ordered by candidate view; a *segment* is a run of jobs with the same
// candidate set, and the slow rows of all its jobs are gathered back to back so that they
// form full 256-row work items against that candidate view.  One thread per segment.
// meta[0] = work items, meta[1] = exact jobs, meta[2] = gathered rows.
__global__ void __launch_bounds__(1024) exact_plan_kernel(const ScanJob* __restrict__ jobs,
                                                          const int32_t* __restrict__ seg_first, int nseg,
                                                          const int* __restrict__ slow_cnt,
                                                          ScanJob* __restrict__ xjobs, int* __restrict__ job_xrow,
                                                          int* __restrict__ meta,
                                                          unsigned long long* __restrict__ counters)
{
    __shared__ int wsum[3][32];
    __shared__ int run[3];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3) run[threadIdx.x] = 0;
    __syncthreads();
    for (int base = 0; base < nseg; base += blockDim.x) {
        int const sgi = base + threadIdx.x;
        int cnt = 0;
        if (sgi < nseg)
            for (int j = seg_first[sgi]; j < seg_first[sgi + 1]; ++j) cnt += slow_cnt[j];
        int const val[3] = {cnt, (cnt + kItemM - 1) / kItemM, cnt > 0 ? 1 : 0};
        int incl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int x = val[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int const y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            incl[k] = x;
            if (lane == 31) wsum[k][warp] = x;
        }
        __syncthreads();
        int pre[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int before = run[k];
            for (int w = 0; w < warp; ++w) before += wsum[k][w];
            pre[k] = before + incl[k] - val[k];   // exclusive prefix
        }
        if (cnt > 0) {
            ScanJob const first = jobs[seg_first[sgi]];
            ScanJob x;
            x.q_row = pre[0];
            x.q_n = cnt;
            x.c_row = first.c_row;
            x.c_n = first.c_n;
            x.out_row = pre[0];
            x.item_start = pre[1];
            x.c_maxnorm2 = 0;
            xjobs[pre[2]] = x;
            int at = pre[0];
            for (int j = seg_first[sgi]; j < seg_first[sgi + 1]; ++j) {
                job_xrow[j] = at;
                at += slow_cnt[j];
            }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) run[k] = pre[k] + val[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ScanJob s;
        s.q_row = 0; s.q_n = 0; s.c_row = 0; s.c_n = 0; s.c_maxnorm2 = 0;
        s.out_row = run[0];
        s.item_start = run[1];
        xjobs[run[2]] = s;
        meta[0] = run[1];
        meta[1] = run[2];
        meta[2] = run[0];
        if (run[0] > 0) atomicAdd(counters + 3, static_cast<unsigned long long>(run[0]));
    }
}

// Copies the slow rows' descriptors into the scratch query pool and records where their
// result goes.  Grid-stride over jobs; 8 threads move one 128-byte row.
__global__ void __launch_bounds__(256) exact_gather_kernel(const ScanJob* __restrict__ jobs, int njobs,
                                                           const int* __restrict__ slow_cnt,
                                                           const int* __restrict__ job_xrow,
                                                           const int64_t* __restrict__ slow_list,
                                                           const uint8_t* __restrict__ pool,
                                                           uint8_t* __restrict__ xpool,
                                                           int64_t* __restrict__ xrow_map)
{
    for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
        int const cnt = slow_cnt[j];
        if (cnt == 0) continue;
        ScanJob const job = jobs[j];
        int const x0 = job_xrow[j];
        for (int e = threadIdx.x; e < cnt * 8; e += blockDim.x) {
            int const s = e >> 3, part = e & 7;
            int64_t const g = slow_list[job.out_row + s];
            int64_t const src_row = static_cast<int64_t>(job.q_row) + (g - job.out_row);
            reinterpret_cast<uint4*>(xpool + (static_cast<int64_t>(x0) + s) * kRowBytes)[part] =
                __ldg(reinterpret_cast<const uint4*>(pool + src_row * kRowBytes) + part);
            if (part == 0) xrow_map[x0 + s] = g;
        }
    }
}

// Certifies the EXACT pass: for every big candidate it met, the value it used (the true
// inner product) must equal the reference's lane-wise 16-bit sum.  Where it does not, the row
// is queued for the CUDA-core replay.  One thread per record; grid-strides over the list.
__global__ void __launch_bounds__(256) verify_big_kernel(PostParams p, const int4* __restrict__ big_list,
                                                         const unsigned long long* __restrict__ big_count,
                                                         int64_t* __restrict__ replay_list,
                                                         unsigned long long* __restrict__ replay_count)
{
    unsigned long long const n = *reinterpret_cast<const volatile unsigned long long*>(big_count);
    unsigned long long const stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long k = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
        int4 const rec = big_list[k];
        int64_t const g = (static_cast<int64_t>(rec.y) << 32) | static_cast<unsigned int>(rec.x);
        ScanJob const job = p.jobs[p.rowres[g].w];
        const uint8_t* q = p.pool + (static_cast<int64_t>(job.q_row) + (g - job.out_row)) * kRowBytes;
        const uint8_t* c = p.pool + (static_cast<int64_t>(job.c_row) + rec.z) * kRowBytes;
        if (wrapped_ip<false>(q, c) != rec.w)
            replay_list[atomicAdd(replay_count, 1ull)] = g;
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
