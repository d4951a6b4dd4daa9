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

// Squared norm of every pool row, and its maximum per view: the Cauchy-Schwarz certificate
// that lets the scan kernel's filter work on 16-bit packed similarities (scan_kernel.cuh).
template <bool SIGNED>
__global__ void __launch_bounds__(256) rownorm_kernel(const uint8_t* __restrict__ pool, int64_t rows,
                                                      int32_t* __restrict__ norm2)
{
    int64_t const r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint4* p = reinterpret_cast<const uint4*>(pool + r * kRowBytes);
    int acc = 0;
#pragma unroll
    for (int i = 0; i < kRowBytes / 16; ++i) {
        uint4 const x = __ldg(p + i);
        if (SIGNED) {
            acc = __dp4a(static_cast<int>(x.x), static_cast<int>(x.x), acc);
            acc = __dp4a(static_cast<int>(x.y), static_cast<int>(x.y), acc);
            acc = __dp4a(static_cast<int>(x.z), static_cast<int>(x.z), acc);
            acc = __dp4a(static_cast<int>(x.w), static_cast<int>(x.w), acc);
        } else {
            unsigned u = static_cast<unsigned>(acc);
            u = __dp4a(x.x, x.x, u);
            u = __dp4a(x.y, x.y, u);
            u = __dp4a(x.z, x.z, u);
            u = __dp4a(x.w, x.w, u);
            acc = static_cast<int>(u);   // <= 128 * 255^2 < 2^23
        }
    }
    norm2[r] = acc;
}

constexpr int kViewMaxChunk = 8192;   // rows per CTA

// grid = (views, chunks): view_max[v] = max of norm2 over the view's rows (views may overlap
// or leave gaps in a caller-provided device pool).  view_max must be zeroed beforehand.
__global__ void __launch_bounds__(256) viewmax_kernel(const int32_t* __restrict__ norm2,
                                                      const int64_t* __restrict__ view_off,
                                                      const int32_t* __restrict__ view_n,
                                                      int32_t* __restrict__ view_max)
{
    int const v = blockIdx.x;
    int const n = view_n[v];
    int const lo = blockIdx.y * kViewMaxChunk;
    if (lo >= n) return;
    const int32_t* p = norm2 + view_off[v];
    int m = 0;
    for (int i = lo + threadIdx.x; i < min(n, lo + kViewMaxChunk); i += blockDim.x) m = max(m, p[i]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(view_max + v, m);
}

constexpr int kDangerCap = 1024;    // high-norm candidates listed per view
constexpr int kDangerBins = 4096;   // histogram range below the view's largest squared norm

// One CTA (kDangerCap threads) per view.  A query row a whose norm product with the view's
// largest norm reaches 2^32 has no certificate from the norms alone -- but only the view's
// few highest-norm rows can actually take a similarity to 2^16.  This kernel lists the view's
// (up to) kDangerCap highest-norm rows, sorted by norm, descending, as (row in view, squared
// norm) pairs; danger_floor[v] is the largest squared norm among the rows that are NOT listed.
// certify_kernel computes the few similarities a doubtful query row has with the head of
// that list and certifies the row after the fact if none reaches 2^16.
// Selection: histogram of (largest norm - norm) over kDangerBins unit bins, the largest
// distance whose cumulative count still fits the list, then collect and sort (bitonic).
__global__ void __launch_bounds__(kDangerCap) danger_kernel(const int32_t* __restrict__ norm2,
                                                            const int64_t* __restrict__ view_off,
                                                            const int32_t* __restrict__ view_n,
                                                            const int32_t* __restrict__ view_max,
                                                            int2* __restrict__ danger, int32_t* __restrict__ danger_cnt,
                                                            int32_t* __restrict__ danger_floor)
{
    __shared__ int hist[kDangerBins];
    __shared__ int2 ent[kDangerCap];
    __shared__ int count, dstar;
    int const v = blockIdx.x;
    int const n = view_n[v];
    int const vmax = view_max[v];
    const int32_t* p = norm2 + view_off[v];
    for (int i = threadIdx.x; i < kDangerBins; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) { count = 0; dstar = -1; }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int const d = vmax - p[i];
        if (d < kDangerBins) atomicAdd(&hist[d], 1);
    }
    __syncthreads();
    // inclusive prefix sums of the histogram, four bins per thread
    int local[kDangerBins / kDangerCap];
    int sum = 0;
#pragma unroll
    for (int q = 0; q < kDangerBins / kDangerCap; ++q) { sum += hist[threadIdx.x * (kDangerBins / kDangerCap) + q]; local[q] = sum; }
    __shared__ int part[kDangerCap];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < kDangerCap; off <<= 1) {
        int const add = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += add;
        __syncthreads();
    }
    int const before = part[threadIdx.x] - sum;
#pragma unroll
    for (int q = 0; q < kDangerBins / kDangerCap; ++q)
        if (before + local[q] <= kDangerCap) atomicMax(&dstar, threadIdx.x * (kDangerBins / kDangerCap) + q);
    __syncthreads();
    int const dmax = dstar;            // rows with vmax - norm <= dmax are listed (-1: none fits)
    int rest = -1;                     // the largest squared norm that is not listed
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int const x = p[i];
        if (vmax - x <= dmax) ent[atomicAdd(&count, 1)] = make_int2(i, x);
        else rest = max(rest, x);
    }
    rest = __reduce_max_sync(0xffffffffu, rest);
    __shared__ int floor_s;
    if (threadIdx.x == 0) floor_s = -1;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && rest >= 0) atomicMax(&floor_s, rest);
    __syncthreads();
    int const cnt = count;
    // bitonic sort of kDangerCap slots (unused ones hold norm -1), descending by norm
    if (threadIdx.x >= cnt) ent[threadIdx.x] = make_int2(0, -1);
    __syncthreads();
    for (int k = 2; k <= kDangerCap; k <<= 1) {
        for (int jx = k >> 1; jx > 0; jx >>= 1) {
            int const i = threadIdx.x, l = i ^ jx;
            if (l > i) {
                int2 const a = ent[i], b = ent[l];
                bool const desc = (i & k) == 0;
                if (desc ? (a.y < b.y) : (a.y > b.y)) { ent[i] = b; ent[l] = a; }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x < cnt) danger[static_cast<int64_t>(v) * kDangerCap + threadIdx.x] = ent[threadIdx.x];
    if (threadIdx.x == 0) {
        danger_cnt[v] = cnt;
        danger_floor[v] = floor_s;      // -1: every row is listed
    }
}

// ---------------------------------------------------------------- finalisation

struct PostParams {
    const uint8_t* pool;
    const ScanJob* jobs;        // njobs + 1 entries (sentinel: out_row = total_rows)
    int njobs;
    int32_t* oneway;            // out: index of the match in the candidate view or -1
    float sq_lowe;              // lowe_ratio_threshold^2   (matching.h:126)
    float sq_dist;              // distance_threshold^2     (matching.h:127)
    int64_t* slow_list;         // rows slow_rows_kernel replays
    unsigned long long* counters;  // [0] length of slow_list
    const unsigned long long* slow_len;   // length of slow_list (normally counters + 0)
};

// ---------------------------------------------------------------- filter pass cut along the candidates

// The filter pass of a batch with few work items runs on `split` column segments per job (run_jobs in
// osfm_match.cu): part[s * rows + g] is segment s's record of row g, its job field counting the
// segment jobs (s * jobs + j).  The fold is the one the filter's epilogue applies to its two column
// halves: best = the larger best, bound on the second best = the largest of the smaller best and
// the two bounds.  All segments of a row were scanned the same way (the choice depends on the
// row and the candidate view only); a row that wraps in one segment wraps.  One thread per row.
template <bool SIGNED>
__global__ void __launch_bounds__(256) merge_split_kernel(const int2* __restrict__ part, int split, int64_t rows, int jobs,
                                                          int2* __restrict__ rowres)
{
    int64_t const g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= rows) return;
    int v1 = 0, v2 = 0, flag = 0, job = 0;
    bool any = false;
    for (int s = 0; s < split; ++s) {
        int2 const rr = part[static_cast<int64_t>(s) * rows + g];
        if (rr.y == -1) continue;      // a segment without columns (a small view in the batch) wrote nothing
        int const a1 = SIGNED ? static_cast<int>(static_cast<short>(rr.x & 0xffff)) : (rr.x & 0xffff);
        int const a2 = SIGNED ? (rr.x >> 16) : static_cast<int>(static_cast<uint32_t>(rr.x) >> 16);
        int const f = static_cast<int>(static_cast<uint32_t>(rr.y) >> kRowFlagShift);
        if (!any) {
            v1 = a1; v2 = a2; flag = f; job = (rr.y & kRowJobMask) % jobs;
            any = true;
        } else {
            v2 = max(max(min(v1, a1), v2), a2);
            v1 = max(v1, a1);
            flag = max(flag, f);
        }
    }
    rowres[g] = pack_rowres(v1, v2, job, flag);
}

// ---------------------------------------------------------------- filter decision

struct ClassifyParams {
    const ScanJob* jobs;
    int64_t total_rows;
    const int2* rowres;          // pack_rowres(v1, v2 lower bound, job) per job row
    const int32_t* norm2;        // squared norm of every pool row
    const int32_t* viewmax;      // largest squared norm per view (index: ScanJob::c_view)
    const uint8_t* pool;
    const int2* danger;          // per view: its highest-norm rows (danger_kernel)
    const int32_t* danger_cnt;
    const int32_t* danger_floor; // upper bound of the squared norms NOT in the view's list (-1: none)
    int32_t* oneway;             // out: -1 for rows the filter rejects
    int64_t* surv_list;          // certified survivors of job j (RESOLVE pass):
    int* surv_cnt;               //   surv_list[jobs[j].out_row + 0 .. surv_cnt[j])
    int64_t* exact_list;         // unsigned kind: rows without certificate (EXACT pass), same layout
    int* exact_cnt;
    int64_t* uncert_list;        // rows without certificate from the norms alone, flat: signed kind:
                                 // CUDA-core replay; unsigned kind: certify_kernel's input
    unsigned long long* counters;  // [1] certified survivors (cumulative),
                                   // [3] rows that stay without certificate (cumulative)
    unsigned long long* uncert_len;   // length of uncert_list
    float sq_lowe, sq_dist;
};

// One thread per job row: applies the reference's tests (matching.h:126-144) to the scan
// kernel's (best, lower bound of second best).  A row that fails them is final (-1): the
// ratio test is monotone in the second best.  A row that passes is a *survivor* and is
// queued, per job, for the RESOLVE pass.  Both statements need the row's similarities to have
// fitted the filter's 16 bits, which the norm certificate guarantees; rows without it are
// queued for the EXACT pass (unsigned) or go to the CUDA-core replay (signed: the reference's
// 16-bit lanes may wrap as well).
template <bool SIGNED>
__global__ void __launch_bounds__(256) classify_kernel(ClassifyParams p)
{
    int64_t const g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int const lane = threadIdx.x & 31;
    bool certified = false, survive = false, wraps = false, live = g < p.total_rows;
    int ji = -1, v1 = 0;
    int64_t out_row = 0;
    if (live) {
        int2 const rr = p.rowres[g];
        v1 = SIGNED ? static_cast<int>(static_cast<short>(rr.x & 0xffff)) : (rr.x & 0xffff);
        int const v2 = SIGNED ? (rr.x >> 16) : static_cast<int>(static_cast<uint32_t>(rr.x) >> 16);
        ji = rr.y & kRowJobMask;
        int const flag = static_cast<int>(static_cast<uint32_t>(rr.y) >> kRowFlagShift);
        ScanJob const job = p.jobs[ji];
        out_row = job.out_row;
        int64_t const limit = SIGNED ? (1ll << 30) : (1ll << 32);
        int const q_prow = job.q_row + static_cast<int>(g - job.out_row);
        int64_t const qn2 = p.norm2[q_prow];
        // scanned with 32-bit loads: the best itself says whether anything left the 16-bit range
        wraps = flag == kRowWideWraps;
        certified = flag == kRowWideOk || (!wraps && qn2 * static_cast<int64_t>(p.viewmax[job.c_view]) < limit);
        // signed: a row whose best is the initial 0 may have no candidate >= 0 at all; the
        // EXACT pass sorts that out (the index then stays 0)
        survive = certified && ((SIGNED && v1 == 0) ||
                                passes_tests(ip_to_dist<SIGNED>(v1), ip_to_dist<SIGNED>(v2), p.sq_lowe, p.sq_dist));
        if (certified && !survive) p.oneway[g] = -1;
    }
    // survivors -> the job's RESOLVE list, rows known to wrap -> its EXACT list; one atomic per
    // job present in the warp (a warp spans at most a few jobs)
#pragma unroll
    for (int which = 0; which < (SIGNED ? 1 : 2); ++which) {
        bool const mine = which == 0 ? survive : wraps;
        unsigned const xm = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            unsigned const peers = __match_any_sync(xm, ji);
            int const leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd((which == 0 ? p.surv_cnt : p.exact_cnt) + ji, __popc(peers));
            base = __shfl_sync(peers, base, leader);
            (which == 0 ? p.surv_list : p.exact_list)[out_row + base + __popc(peers & ((1u << lane) - 1u))] =
                surv_entry(g, v1, which == 0);
        }
    }
    // rows without certificate -> flat list (signed: replayed on CUDA cores; unsigned: looked
    // at again by certify_kernel)
    bool const doubtful = live && !certified && !wraps;
    unsigned const um = __ballot_sync(0xffffffffu, doubtful);
    unsigned const sm = __ballot_sync(0xffffffffu, survive);
    unsigned const wm = __ballot_sync(0xffffffffu, wraps);
    if (um != 0) {
        unsigned long long ub = 0;
        if (lane == 0) ub = atomicAdd(p.uncert_len, static_cast<unsigned long long>(__popc(um)));
        ub = __shfl_sync(0xffffffffu, ub, 0);
        if (doubtful) p.uncert_list[ub + __popc(um & ((1u << lane) - 1u))] = g;
    }
    // statistics: one atomic per CTA
    __shared__ unsigned s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) {
        if (sm) atomicAdd(&s_cnt[0], __popc(sm));
        if (SIGNED && um) atomicAdd(&s_cnt[1], __popc(um));   // unsigned: certify_kernel counts the rest
        if (wm) atomicAdd(&s_cnt[1], __popc(wm));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt[0]) atomicAdd(p.counters + 1, static_cast<unsigned long long>(s_cnt[0]));
    if (threadIdx.x == 1 && s_cnt[1]) atomicAdd(p.counters + 3, static_cast<unsigned long long>(s_cnt[1]));
}

// Unsigned kind, one warp per row that has no certificate from the norms alone.  Only the
// candidate view's few highest-norm rows (danger_kernel's list, sorted) can take a similarity
// of this row to 2^16: the lanes compute those similarities exactly.  If none reaches 2^16, no
// similarity of the row left the 16-bit range, the filter saw the row exactly, and it is
// treated like any certified row (rejected, or queued for the RESOLVE pass); otherwise it is
// queued for the EXACT pass.  Grid-strides over the list, whose length is only known on the
// device.
// TARGETS: the rows are claimed rows of a reverse job (claim_kernel): rowres holds (V, job),
// and a row certified after the fact is queued for the RESOLVE pass whatever its value.
template <bool TARGETS>
__global__ void __launch_bounds__(256) certify_kernel(ClassifyParams p)
{
    int64_t const n = static_cast<int64_t>(*reinterpret_cast<volatile unsigned long long*>(p.uncert_len));
    int const lane = threadIdx.x & 31;
    int64_t const nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    unsigned n_exact = 0, n_surv = 0;
    for (int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n; w += nwarps) {
        int64_t const g = p.uncert_list[w];
        int2 const rr = p.rowres[g];
        int const v1 = rr.x & 0xffff;
        int const v2 = static_cast<int>(static_cast<uint32_t>(rr.x) >> 16);
        int const ji = rr.y & kRowJobMask;
        ScanJob const job = p.jobs[ji];
        int const q_prow = job.q_row + static_cast<int>(g - job.out_row);
        int64_t const qn2 = p.norm2[q_prow];
        int const dc = p.danger_cnt[job.c_view];
        // rows that are not listed are harmless if even the largest of them is
        bool wraps = qn2 * static_cast<int64_t>(p.danger_floor[job.c_view]) >= (1ll << 32);
        const int2* dl = p.danger + static_cast<int64_t>(job.c_view) * kDangerCap;
        for (int base = 0; base < dc && !wraps; base += 32) {
            int const i = base + lane;
            int2 const e = i < dc ? dl[i] : make_int2(0, 0);
            bool const risky = qn2 * static_cast<int64_t>(e.y) >= (1ll << 32);
            bool hit = false;
            if (risky && e.x < job.c_n)
                hit = dot_row<false>(p.pool + static_cast<int64_t>(q_prow) * kRowBytes,
                                     p.pool + (static_cast<int64_t>(job.c_row) + e.x) * kRowBytes) >= 65536;
            wraps = __any_sync(0xffffffffu, hit);
            if (!__all_sync(0xffffffffu, risky)) break;      // sorted by norm: the rest is harmless
        }
        if (lane == 0) {
            if (wraps) {
                p.exact_list[job.out_row + atomicAdd(p.exact_cnt + ji, 1)] = surv_entry(g, v1, false);
                ++n_exact;
            } else if (TARGETS || passes_tests(ip_to_dist<false>(v1), ip_to_dist<false>(v2), p.sq_lowe, p.sq_dist)) {
                p.surv_list[job.out_row + atomicAdd(p.surv_cnt + ji, 1)] = surv_entry(g, v1, true);
                ++n_surv;
            } else {
                p.oneway[g] = -1;
            }
        }
    }
    if (lane == 0) {
        if (n_exact) atomicAdd(p.counters + 3, static_cast<unsigned long long>(n_exact));
        if (n_surv) atomicAdd(p.counters + 1, static_cast<unsigned long long>(n_surv));
    }
}

// ---------------------------------------------------------------- reverse direction of a pair
//
// Matching::twoway_match (matching.h:148-159) scans both directions of a pair, and
// remove_inconsistent_matches (matching.cc:19-36) then keeps i -> j only if j -> i.  After the
// mutual filter a row j of the second view therefore matters only if some row i of the first
// view *claims* it (oneway_12[i] == j), and all that matters about it is whether its own nearest
// neighbour, under the reference's rules, is that i.  So only the forward direction of a pair
// goes through the filter pass (one tensor-core product per pair instead of two); the reverse
// direction is evaluated for the claimed rows alone:
//   claim_kernel    V[j] = the largest similarity any claimant has with j (clamped at the
//                   reference's initial 0), kept in the (otherwise unused) rowres slot of j; the
//                   first claimant also queues j like classify_kernel queues survivors: rows
//                   with the 16-bit norm certificate for the RESOLVE pass in verify mode (a
//                   similarity above V: some row that is no claimant is nearer, the result is -1
//                   for the mutual filter; otherwise V is the row's best and RESOLVE's result is
//                   the reference's), the others for certify_kernel / the EXACT pass / the
//                   CUDA-core replay, which need no V.
// Rows nobody claims keep the -1 the host wrote beforehand.

struct ClaimParams {
    const ScanJob* jobs;
    const int32_t* rev_of;       // per job: its pair's reverse job (or -1)
    // the forward rows whose result may be a match: a second pass's gathered row list (xrow_map,
    // length *n_int) or the flat list of rows replayed on CUDA cores (length *n_ull)
    const int64_t* rows;
    const int* n_int;
    const unsigned long long* n_ull;
    const uint8_t* pool;
    const int32_t* oneway;
    int2* rowres;                // forward rows: the filter's record; reverse rows: (V, job), preset to -1
    const int32_t* norm2;        // the norm certificate's inputs (as in ClassifyParams)
    const int32_t* viewmax;
    int64_t* surv_list;          // RESOLVE queue of the reverse jobs (certified rows), per job at out_row
    int* surv_cnt;
    int64_t* uncert_list;        // claimed rows without certificate, flat
    unsigned long long* uncert_len;
    unsigned long long* counters;   // [3] signed rows without certificate, [4] claimed rows
    int32_t* smin;               // per reverse job: the smallest claimed value (nullptr: not wanted)
};

// The largest similarity t <= s for which a best of s and a second best of t still pass the
// reference's tests (matching.h:138-143); -1 if even a second best of 0 fails.  The tests are
// monotone in t (the distance falls as t grows, the quotient rises), so a bisection finds it.
template <bool SIGNED>
__device__ __forceinline__ int ratio_limit(int s, float sq_lowe, float sq_dist) {
    int const d1 = ip_to_dist<SIGNED>(s);
    if (!passes_tests(d1, ip_to_dist<SIGNED>(0), sq_lowe, sq_dist)) return -1;
    int lo = 0, hi = s;              // passes at lo; find the largest passing value in [lo, hi]
    while (lo < hi) {
        int const mid = lo + (hi - lo + 1) / 2;
        if (passes_tests(d1, ip_to_dist<SIGNED>(mid), sq_lowe, sq_dist)) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// One thread per listed forward row (grid-stride: the list length is only known on the device).
// TRUSTED: the list is the RESOLVE pass's, whose entries carry the row's best similarity -- which
// IS the similarity of the row with its match; otherwise (EXACT pass, replay) the product is
// computed again.  The FIRST claimant of a row (the atomic maximum returns the preset -1) also
// queues it; the value the queue entry carries is filled in later, from the slot, once every claim
// has landed (gather_rows_kernel).
template <bool SIGNED, bool TRUSTED>
__global__ void __launch_bounds__(256) claim_kernel(ClaimParams p)
{
    int64_t const n = p.n_int != nullptr ? static_cast<int64_t>(*p.n_int)
                                         : static_cast<int64_t>(*reinterpret_cast<const volatile unsigned long long*>(p.n_ull));
    int const lane = threadIdx.x & 31;
    int64_t const stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    // (whole warps leave the loop together: the bound is rounded up to the warp)
    for (int64_t k0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + (threadIdx.x & ~31); k0 < n; k0 += stride) {
        int64_t const k = k0 + lane;
        bool first = false, certified = false;
        int rj = -1, claimed_s = -1;
        int64_t target = 0, out_row = 0;
        if (k < n) {
            int64_t const entry = p.rows[k];
            int64_t const g = surv_row(entry);
            int const mt = p.oneway[g];
            if (mt >= 0) {
                int const ji = p.rowres[g].y & kRowJobMask;
                rj = p.rev_of[ji];
                if (rj >= 0) {
                    int s;
                    if (TRUSTED) {
                        uint32_t const v16 = static_cast<uint32_t>((static_cast<uint64_t>(entry) >> kSurvRowBits) & 0xffffu);
                        s = SIGNED ? static_cast<int>(static_cast<short>(v16)) : static_cast<int>(v16);
                    } else {
                        ScanJob const job = p.jobs[ji];
                        s = dot_row<SIGNED>(p.pool + (static_cast<int64_t>(job.q_row) + (g - job.out_row)) * kRowBytes,
                                            p.pool + (static_cast<int64_t>(job.c_row) + mt) * kRowBytes);
                    }
                    s = max(s, 0);   // the reference's best starts at 0 (nearest_neighbor.cc:221-224, 246-249)
                    claimed_s = s;
                    ScanJob const rjob = p.jobs[rj];
                    out_row = rjob.out_row;
                    target = out_row + mt;
                    int2* const slot = p.rowres + target;
                    // (the certificate's inputs are loaded before the atomic: one round trip less
                    // in a thread that is a chain of dependent memory operations as it is)
                    int64_t const limit = SIGNED ? (1ll << 30) : (1ll << 32);
                    int64_t const qn2 = p.norm2[rjob.q_row + mt];
                    int64_t const vm = p.viewmax[rjob.c_view];
                    first = atomicMax(&slot->x, s) < 0;
                    if (first) {
                        slot->y = rj;
                        certified = qn2 * vm < limit;
                    }
                }
            }
        }
        // the smallest claimed value per job: one atomic per job present in the warp (the lists are
        // ordered by job, so unaggregated atomics would queue up on a handful of addresses)
        if (p.smin != nullptr) {
            unsigned const cm = __ballot_sync(0xffffffffu, claimed_s >= 0);
            if (claimed_s >= 0) {
                unsigned const peers = __match_any_sync(cm, rj);
                int const mn = __reduce_min_sync(peers, claimed_s);
                if (lane == __ffs(peers) - 1) atomicMin(p.smin + rj, mn);
            }
        }
        // certified rows -> the reverse job's RESOLVE queue, one atomic per job present in the warp
        bool const queue = first && certified;
        unsigned const qm = __ballot_sync(0xffffffffu, queue);
        if (queue) {
            unsigned const peers = __match_any_sync(qm, rj);
            int const leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(p.surv_cnt + rj, __popc(peers));
            base = __shfl_sync(peers, base, leader);
            p.surv_list[out_row + base + __popc(peers & ((1u << lane) - 1u))] = surv_entry(target, 0, true);
        }
        // without certificate: signed -> CUDA-core replay; unsigned -> certify_kernel<true> looks at
        // the few candidates that could take the row to 2^16
        bool const doubtful = first && !certified;
        unsigned const um = __ballot_sync(0xffffffffu, doubtful);
        if (um != 0) {
            unsigned long long ub = 0;
            if (lane == 0) ub = atomicAdd(p.uncert_len, static_cast<unsigned long long>(__popc(um)));
            ub = __shfl_sync(0xffffffffu, ub, 0);
            if (doubtful) p.uncert_list[ub + __popc(um & ((1u << lane) - 1u))] = target;
        }
        unsigned const fm = __ballot_sync(0xffffffffu, first);
        if (lane == 0) {
            if (fm) atomicAdd(p.counters + 4, static_cast<unsigned long long>(__popc(fm)));      // claimed rows (cumulative)
            if (SIGNED && um) atomicAdd(p.counters + 3, static_cast<unsigned long long>(__popc(um)));   // unsigned: certify_kernel counts
        }
    }
}

// Ordered compaction over a CTA of 1024 threads, each holding kRowsPerThread consecutive rows:
// given the number of rows a thread keeps, returns the rank of its first kept row among all rows
// kept so far and advances the running total (the same in every thread).  Two barriers a round.
constexpr int kRowsPerThread = 8;
struct BlockRank {
    int warp_tot[32];
    int warp_pre[32];
    int total;
};
__device__ __forceinline__ int block_rank(int my_count, int& running, BlockRank& sm)
{
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = my_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int const t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int const t = sm.warp_tot[lane];
        int inc2 = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int const u = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += u;
        }
        sm.warp_pre[lane] = running + inc2 - t;
        if (lane == 31) sm.total = running + inc2;
    }
    __syncthreads();
    // (the next round writes warp_pre / total only after its first barrier: these reads are safe)
    int const rank = sm.warp_pre[warp] + incl - my_count;
    running = sm.total;
    return rank;
}

// Restricted candidate sets for the reverse pass.  One CTA per forward job (pair).  The reverse
// job of the pair holds its claimed rows against the forward job's query rows; of those only the
// rows whose own best similarity exceeds tau (below) -- or whose record is not exact -- can
// matter, and the filter pass has that best similarity for every row already.  The kernel compacts
// them, in order (ties go to the highest index: the order must survive), into the gathered
// candidate pool at the forward job's own row offset, records the column -> row map, and points the
// reverse job (in the copy of the job list the reverse RESOLVE pass plans from) at the subset, flagged
// by a negative c_view.  Pairs where the subset would not pay (more than half of the rows) keep
// the whole view.
struct SelectParams {
    const ScanJob* jobs;
    ScanJob* jobs_rev;           // copy of jobs; reverse entries are redirected here
    const int32_t* rev_of;
    int fwd_jobs;                // jobs [0, fwd_jobs) are forward jobs
    const int2* rowres;
    const int32_t* norm2;
    const int32_t* viewmax;
    const int32_t* smin;         // per reverse job: the smallest claimed value (claim_kernel)
    float sq_lowe, sq_dist;
    const int* surv_cnt;         // claimed rows per reverse job queued for the RESOLVE pass
    const uint8_t* pool;
    uint8_t* cand_pool;
    int32_t* cand_map;
    int32_t* cand_cnt;           // per forward job: rows kept if the pair is restricted (preset to -1)
    unsigned long long* counters;   // [9] candidate rows kept (cumulative), [10] pairs restricted (cumulative)
};

template <bool SIGNED>
__global__ void __launch_bounds__(1024) select_candidates_kernel(SelectParams p)
{
    __shared__ BlockRank sm;
    int const ji = blockIdx.x;
    int const rj = p.rev_of[ji];
    if (rj < 0 || p.surv_cnt[rj] == 0) return;          // nothing claimed: the reverse job has no work
    ScanJob const job = p.jobs[ji];
    // Rows of this view whose own best similarity is at most tau cannot influence what the reverse
    // pass decides about any row this pair claims: with s the smallest claimed value, they lie below
    // every claimed value (so they neither beat nor tie one) and at most at the largest second best
    // that still passes the ratio test beside a best of s -- and beside any larger best, the limit
    // being monotone in the best (so they cannot turn an accept into a reject).
    int const sm_claim = p.smin[rj];
    int const tau = min(ratio_limit<SIGNED>(sm_claim, p.sq_lowe, p.sq_dist), sm_claim - 1);
    int64_t const limit = SIGNED ? (1ll << 30) : (1ll << 32);
    int64_t const vmax = p.viewmax[job.c_view];
    int running = 0;
    for (int start = 0; start < job.q_n; start += blockDim.x * kRowsPerThread) {
        int const r0 = start + threadIdx.x * kRowsPerThread;
        unsigned keep = 0;
#pragma unroll
        for (int k = 0; k < kRowsPerThread; ++k) {
            int const r = r0 + k;
            if (r < job.q_n) {
                int2 const rr = p.rowres[job.out_row + r];
                int const flag = static_cast<int>(static_cast<uint32_t>(rr.y) >> kRowFlagShift);
                bool const trusted = flag == kRowWideOk ||
                    (flag == kRowPacked && static_cast<int64_t>(p.norm2[job.q_row + r]) * vmax < limit);
                int const v1 = SIGNED ? static_cast<int>(static_cast<short>(rr.x & 0xffff)) : (rr.x & 0xffff);
                if (!trusted || v1 > tau) keep |= 1u << k;
            }
        }
        int rank = block_rank(__popc(keep), running, sm);
#pragma unroll
        for (int k = 0; k < kRowsPerThread; ++k)
            if ((keep >> k) & 1u) p.cand_map[job.out_row + rank++] = r0 + k;
    }
    int const cnt = running;
    // not worth it for small views or when most rows stay: the reverse job keeps the whole view
    if (job.q_n < 1024 || 2 * cnt > job.q_n) return;
    if (threadIdx.x == 0) {
        // gather_candidates_kernel copies the rows; padded <= q_n / 2 + 255 <= q_n: inside the job's slots
        ScanJob rjob = p.jobs[rj];
        rjob.c_row = static_cast<int32_t>(job.out_row);
        rjob.c_n = (cnt + kBlockN - 1) / kBlockN * kBlockN;
        rjob.c_view = -1 - rjob.c_view;
        p.jobs_rev[rj] = rjob;
        p.cand_cnt[ji] = cnt;
        atomicAdd(p.counters + 9, static_cast<unsigned long long>(cnt));
        atomicAdd(p.counters + 10, 1ull);
    }
}

// The rows select_candidates_kernel kept (8 threads move one 128-byte row), then zero rows up to a
// whole number of candidate tiles, so that the RESOLVE pass never meets a ragged tile.  A zero row
// has similarity 0 with every row, below every claimed value of a restricted job (a claimed value
// of 0 gives tau = -1, which keeps every row and so the whole view) and no more than the
// reference's initial second best of 0.  blockIdx.x: forward job; blockIdx.y: slices of its rows.
__global__ void __launch_bounds__(256) gather_candidates_kernel(const ScanJob* __restrict__ jobs,
                                                                 const int32_t* __restrict__ cand_cnt,
                                                                 const int32_t* __restrict__ cand_map,
                                                                 const uint8_t* __restrict__ pool,
                                                                 uint8_t* __restrict__ cand_pool)
{
    int const cnt = cand_cnt[blockIdx.x];
    if (cnt < 0) return;
    ScanJob const job = jobs[blockIdx.x];
    int const padded = (cnt + kBlockN - 1) / kBlockN * kBlockN;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < padded * 8; e += gridDim.y * blockDim.x) {
        int const sidx = e >> 3, part = e & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (sidx < cnt) {
            int const r = cand_map[job.out_row + sidx];
            v = __ldg(reinterpret_cast<const uint4*>(pool + (static_cast<int64_t>(job.q_row) + r) * kRowBytes) + part);
        }
        reinterpret_cast<uint4*>(cand_pool + (job.out_row + sidx) * kRowBytes)[part] = v;
    }
}

// ---------------------------------------------------------------- wrap emulation

// One warp per flagged row: replays the reference's sequential scan including the
// truncating 16-bit stores of best / second best (nearest_neighbor.cc:87-100).
template <bool SIGNED>
__global__ void __launch_bounds__(256) slow_rows_kernel(PostParams p)
{
    // The list length is only known on the device (written by a kernel that precedes this
    // launch in stream order); the grid strides over it.
    int64_t const nslow = static_cast<int64_t>(*reinterpret_cast<const volatile unsigned long long*>(p.slow_len));
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

// item_job[] of a host-built job list (the filter pass): one thread per job.
__global__ void __launch_bounds__(256) fill_item_job_kernel(const ScanJob* __restrict__ jobs, int njobs,
                                                            int32_t* __restrict__ item_job)
{
    int const j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    int const first = jobs[j].item_start, last = jobs[j + 1].item_start;
    for (int it = first; it < last; ++it) item_job[it] = j;
}

// ---------------------------------------------------------------- exact pass, few rows

// The EXACT scan pass walks a row group's candidate tiles one after the other, which for a
// handful of rows against a very large view (one pair of 200 000 x 200 000) is one CTA working
// through hundreds of tiles while the others idle.  When the rows' inner products fit the scratch
// buffer they are computed here instead, spread over the whole device -- exact_dots_kernel: every
// inner product of every gathered row (dp4a), plus the largest one per block of 256 columns --
// and exact_replay_kernel then replays the reference's sequential scan (nearest_neighbor.cc:87-100)
// one warp per row, skipping the column blocks (and, inside a block, the groups of 32 columns) that
// hold nothing at or above the row's current second best.  A skipped stretch cannot change the
// state, and every value that is looked at goes through the same ref_scan_step() and big-candidate
// bookkeeping as in the scan pass, so both paths give the same rows the same results.
constexpr int kWideRowBatch = 32;      // gathered rows one CTA holds in shared memory
constexpr int kWideColBlock = 256;     // columns per block: one per thread

struct ExactWideParams {
    const ScanJob* xjobs;          // plan_rows_kernel's gathered jobs (rows in xpool, candidates in pool)
    const int* xmeta;              // plan_rows_kernel: [0] work items, [1] gathered jobs, [2] gathered rows
    const uint8_t* xpool;
    const uint8_t* pool;
    const int64_t* xrow_map;
    int64_t* x_off;                // per gathered job: offset of its inner products / block maxima
    int64_t* xm_off;
    int* unit_first;               // per gathered job: its first (row batch, column block) unit
    int* meta;                     // [0] 1: this path is taken, [1] units, [4] work items left to the scan pass
    int32_t* x;
    int32_t* xmax;
    int64_t x_cap, xm_cap;
    int mode;                      // 0: by capacity, 1: always the scan pass
    int max_jobs;
    // results, as in the scan pass
    int32_t* oneway;
    float sq_lowe, sq_dist;
    int4* big_list;
    unsigned long long* big_count;
    int64_t* replay_list;
    unsigned long long* replay_count;
    uint32_t* replay_flags;
    unsigned long long* self_check;
    unsigned long long* wide_rows;   // rows that took this path (cumulative)
};

// One thread (the tail of plan_rows_kernel): sizes the scratch layout and decides which path the
// rows take.
__device__ __forceinline__ void exact_wide_plan(ExactWideParams const& p, const ScanJob* xjobs, int items, int nx, int rows)
{
    bool ok = p.mode == 0 && nx > 0 && nx <= p.max_jobs;
    int64_t xo = 0, mo = 0, units = 0;
    for (int s = 0; ok && s < nx; ++s) {
        ScanJob const xj = xjobs[s];
        int64_t const ncb = (xj.c_n + kWideColBlock - 1) / kWideColBlock;
        p.x_off[s] = xo;
        p.xm_off[s] = mo;
        p.unit_first[s] = static_cast<int>(units);
        xo += static_cast<int64_t>(xj.q_n) * xj.c_n;
        mo += static_cast<int64_t>(xj.q_n) * ncb;
        units += static_cast<int64_t>((xj.q_n + kWideRowBatch - 1) / kWideRowBatch) * ncb;
        ok = xo <= p.x_cap && mo <= p.xm_cap && units < (1ll << 30);
    }
    if (ok) p.unit_first[nx] = static_cast<int>(units);
    p.meta[0] = ok ? 1 : 0;
    p.meta[1] = ok ? static_cast<int>(units) : 0;
    p.meta[4] = ok ? 0 : items;
    if (ok) atomicAdd(p.wide_rows, static_cast<unsigned long long>(rows));
}

// ---------------------------------------------------------------- exact pass set-up

// Single CTA.  Turns the per-job slow-row counts into the job list of the EXACT scan pass.
// Jobs arrive ordered by candidate view; a *segment* is a run of jobs with the same
// candidate set, and the slow rows of all its jobs are gathered back to back so that they
// form full 256-row work items against that candidate view.  One thread per segment.
// meta[0] = work items, meta[1] = exact jobs, meta[2] = gathered rows.
__global__ void __launch_bounds__(1024) plan_rows_kernel(const ScanJob* __restrict__ jobs,
                                                          const int32_t* __restrict__ seg_first, int nseg,
                                                          const int* __restrict__ slow_cnt,
                                                          ScanJob* __restrict__ xjobs, int* __restrict__ job_xrow,
                                                          int* __restrict__ meta,
                                                          unsigned long long* __restrict__ rows_total,
                                                          int32_t* __restrict__ item_job,
                                                          ExactWideParams wide)
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
            x.c_view = first.c_view;
            xjobs[pre[2]] = x;
            for (int q = 0; q < val[1]; ++q) item_job[pre[1] + q] = pre[2];
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
        s.q_row = 0; s.q_n = 0; s.c_row = 0; s.c_n = 0; s.c_view = 0;
        s.out_row = run[0];
        s.item_start = run[1];
        xjobs[run[2]] = s;
        meta[0] = run[1];
        meta[1] = run[2];
        meta[2] = run[0];
        if (run[0] > 0 && rows_total != nullptr) atomicAdd(rows_total, static_cast<unsigned long long>(run[0]));   // cumulative
        if (wide.meta != nullptr) exact_wide_plan(wide, xjobs, run[1], run[2], run[0]);      // EXACT pass: which way do the rows go?
    }
}

// Copies the slow rows' descriptors into the scratch query pool and records where their
// result goes.  Grid-stride over jobs; 8 threads move one 128-byte row.
__global__ void __launch_bounds__(256) gather_rows_kernel(const ScanJob* __restrict__ jobs, int njobs,
                                                           const int* __restrict__ slow_cnt,
                                                           const int* __restrict__ job_xrow,
                                                           const int64_t* __restrict__ slow_list,
                                                           const uint8_t* __restrict__ pool,
                                                           uint8_t* __restrict__ xpool,
                                                           int64_t* __restrict__ xrow_map,
                                                           const int2* __restrict__ claimed_v)
{
    for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
        int const cnt = slow_cnt[j];
        if (cnt == 0) continue;
        ScanJob const job = jobs[j];
        int const x0 = job_xrow[j];
        // blockIdx.y: slices of one job's list (a single pair of large views has few jobs)
        for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < cnt * 8; e += gridDim.y * blockDim.x) {
            int const s = e >> 3, part = e & 7;
            int64_t const entry = slow_list[job.out_row + s];
            int64_t const src_row = static_cast<int64_t>(job.q_row) + (surv_row(entry) - job.out_row);
            reinterpret_cast<uint4*>(xpool + (static_cast<int64_t>(x0) + s) * kRowBytes)[part] =
                __ldg(reinterpret_cast<const uint4*>(pool + src_row * kRowBytes) + part);
            // reverse pass: the value a claimed row is held against is final only now
            if (part == 0)
                xrow_map[x0 + s] = claimed_v == nullptr ? entry
                    : surv_entry(surv_row(entry), claimed_v[surv_row(entry)].x, (static_cast<uint64_t>(entry) & kSurvCertified) != 0);
        }
    }
}

// ---------------------------------------------------------------- exact pass, few rows (kernels)

// largest s with first[s] <= v (first[] ascending, first[0] = 0, n >= 1)
__device__ __forceinline__ int last_not_above(const int* first, int n, int v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int const mid = (lo + hi + 1) >> 1;
        if (first[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(kWideColBlock) exact_dots_kernel(ExactWideParams p)
{
    if (p.meta[0] == 0) return;
    __shared__ uint4 rows[kWideRowBatch][8];
    __shared__ int wmax[kWideRowBatch][kWideColBlock / 32];
    int const nx = p.xmeta[1], units = p.meta[1];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int const s = last_not_above(p.unit_first, nx, u);
        ScanJob const xj = p.xjobs[s];
        int const ncb = (xj.c_n + kWideColBlock - 1) / kWideColBlock;
        int const rb = (u - p.unit_first[s]) / ncb, cb = (u - p.unit_first[s]) % ncb;
        int const r0 = rb * kWideRowBatch;
        int const nr = min(kWideRowBatch, xj.q_n - r0);
        __syncthreads();                       // the previous unit's rows are no longer read
        {
            int const r = threadIdx.x >> 3, part = threadIdx.x & 7;      // 256 threads: 32 rows x 8 parts
            if (r < nr)
                rows[r][part] = __ldg(reinterpret_cast<const uint4*>(p.xpool + (static_cast<int64_t>(xj.q_row) + r0 + r) * kRowBytes) + part);
        }
        int const col = cb * kWideColBlock + threadIdx.x;
        bool const live = col < xj.c_n;
        uint4 c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            c[k] = live ? __ldg(reinterpret_cast<const uint4*>(p.pool + (static_cast<int64_t>(xj.c_row) + col) * kRowBytes) + k)
                        : make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        int32_t* const xr = p.x + p.x_off[s] + static_cast<int64_t>(r0) * xj.c_n + col;
        for (int r = 0; r < nr; ++r) {
            unsigned acc = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint4 const q = rows[r][k];
                acc = __dp4a(q.x, c[k].x, acc);
                acc = __dp4a(q.y, c[k].y, acc);
                acc = __dp4a(q.z, c[k].z, acc);
                acc = __dp4a(q.w, c[k].w, acc);
            }
            int const v = live ? static_cast<int>(acc) : INT_MIN;
            if (live) xr[static_cast<int64_t>(r) * xj.c_n] = v;
            int const m = __reduce_max_sync(0xffffffffu, v);
            if (lane == 0) wmax[r][warp] = m;
        }
        __syncthreads();
        if (threadIdx.x < nr) {
            int m = INT_MIN;
#pragma unroll
            for (int w = 0; w < kWideColBlock / 32; ++w) m = max(m, wmax[threadIdx.x][w]);
            p.xmax[p.xm_off[s] + static_cast<int64_t>(r0 + threadIdx.x) * ncb + cb] = m;
        }
    }
}

__global__ void __launch_bounds__(256) exact_replay_kernel(ExactWideParams p)
{
    if (p.meta[0] == 0) return;
    int const nx = p.xmeta[1], total_rows = p.xmeta[2];
    int const lane = threadIdx.x & 31;
    int const nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total_rows; w += nwarps) {
        // the gathered job of row w: xjobs[].out_row ascends, xjobs[nx] is the end marker
        int lo = 0, hi = nx - 1;
        while (lo < hi) {
            int const mid = (lo + hi + 1) >> 1;
            if (p.xjobs[mid].out_row <= w) lo = mid; else hi = mid - 1;
        }
        ScanJob const xj = p.xjobs[lo];
        int const r = w - static_cast<int>(xj.out_row);
        int const ncb = (xj.c_n + kWideColBlock - 1) / kWideColBlock;
        const int32_t* const xr = p.x + p.x_off[lo] + static_cast<int64_t>(r) * xj.c_n;
        const int32_t* const xm = p.xmax + p.xm_off[lo] + static_cast<int64_t>(r) * ncb;
        int64_t const entry = p.xrow_map[w];
        int64_t const g = surv_row(entry);

        int b1 = 0, b2 = 0, i1 = 0, nbig = 0;     // warp-uniform (nearest_neighbor.cc:246-249)
        for (int cb0 = 0; cb0 < ncb; cb0 += 32) {
            int const mx = cb0 + lane < ncb ? xm[cb0 + lane] : INT_MIN;
            if (!__any_sync(0xffffffffu, mx >= b2)) continue;      // nothing in 32 blocks can enter
            int const nb = min(32, ncb - cb0);
            for (int l = 0; l < nb; ++l) {
                // against the second best as it is now (a wrapped store can lower it)
                if (__shfl_sync(0xffffffffu, mx, l) < b2) continue;
                int const col0 = (cb0 + l) * kWideColBlock;
                int x[kWideColBlock / 32];
#pragma unroll
                for (int q = 0; q < kWideColBlock / 32; ++q) {
                    int const col = col0 + q * 32 + lane;
                    x[q] = col < xj.c_n ? xr[col] : INT_MIN;
                }
#pragma unroll
                for (int q = 0; q < kWideColBlock / 32; ++q) {
                    if (!__any_sync(0xffffffffu, x[q] >= b2)) continue;
                    for (int t = 0; t < 32; ++t) {
                        int const v = __shfl_sync(0xffffffffu, x[q], t);
                        if (v >= b2) {
                            int const col = col0 + q * 32 + t;
                            if (v >= 65536) {
                                if (nbig < kMaxBigPerRow && lane == 0)
                                    p.big_list[atomicAdd(p.big_count, 1ull)] =
                                        make_int4(static_cast<int>(g), static_cast<int>(g >> 32), col, v);
                                ++nbig;
                            }
                            ref_scan_step<false>(v, col, b1, b2, i1);
                        }
                    }
                }
            }
        }
        if (lane == 0) {
            bool const ok = passes_tests(ip_to_dist<false>(b1), ip_to_dist<false>(b2), p.sq_lowe, p.sq_dist);
            p.oneway[g] = ok ? i1 : -1;
            if (nbig > kMaxBigPerRow && mark_once(p.replay_flags, g))   // cannot be certified by verify_big_kernel
                p.replay_list[atomicAdd(p.replay_count, 1ull)] = g;
            if ((static_cast<uint64_t>(entry) & kSurvCertified) != 0 &&
                (b1 & 0xffff) != static_cast<int>((static_cast<uint64_t>(entry) >> kSurvRowBits) & 0xffffu))
                atomicAdd(p.self_check, 1ull);
        }
    }
}

// Certifies the EXACT pass: for every big candidate it met, the value it used (the true
// inner product) must equal the reference's lane-wise 16-bit sum.  Where it does not, the row
// is queued for the CUDA-core replay.  One thread per record; grid-strides over the list.
__global__ void __launch_bounds__(256) verify_big_kernel(PostParams p, const int4* __restrict__ big_list,
                                                         const unsigned long long* __restrict__ big_count,
                                                         int64_t* __restrict__ replay_list,
                                                         unsigned long long* __restrict__ replay_count,
                                                         uint32_t* __restrict__ replay_flags)
{
    unsigned long long const n = *reinterpret_cast<const volatile unsigned long long*>(big_count);
    unsigned long long const stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    for (unsigned long long k = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
        int4 const rec = big_list[k];
        int64_t const g = (static_cast<int64_t>(rec.y) << 32) | static_cast<unsigned int>(rec.x);
        ScanJob const job = p.jobs[find_job(p.jobs, p.njobs, g)];
        const uint8_t* q = p.pool + (static_cast<int64_t>(job.q_row) + (g - job.out_row)) * kRowBytes;
        const uint8_t* c = p.pool + (static_cast<int64_t>(job.c_row) + rec.z) * kRowBytes;
        // a row has up to kMaxBigPerRow records: it enters the list once
        if (wrapped_ip<false>(q, c) != rec.w && mark_once(replay_flags, g))
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
// Each thread holds eight consecutive rows; ranks come from block_rank().
__global__ void __launch_bounds__(1024) compact_kernel(const PairPart* __restrict__ parts,
                                                       const int32_t* __restrict__ dense,
                                                       const int64_t* __restrict__ list_offset,
                                                       int2* __restrict__ list)
{
    __shared__ BlockRank sm;
    PairPart const pp = parts[blockIdx.x];
    int64_t const base = list_offset[pp.pair];
    int running = 0;
    for (int start = 0; start < pp.n1; start += blockDim.x * kRowsPerThread) {
        int const i0 = start + threadIdx.x * kRowsPerThread;
        int mt[kRowsPerThread];
        int kept = 0;
#pragma unroll
        for (int k = 0; k < kRowsPerThread; ++k) {
            mt[k] = i0 + k < pp.n1 ? dense[pp.out12 + i0 + k] : -1;
            kept += mt[k] >= 0;
        }
        int rank = block_rank(kept, running, sm);
#pragma unroll
        for (int k = 0; k < kRowsPerThread; ++k)
            if (mt[k] >= 0) list[base + rank++] = make_int2(i0 + k, mt[k]);
    }
}

}  // namespace osfm
