// float_tc_kernels.cuh -- a tensor-core filter in front of the float descriptor path.
//
// Reference (paths relative to /root/reference):
//   float_inner_prod, SSE3 branch            src/mve/sfm/nearest_neighbor.cc:141-176
//   NearestNeighbor<float>::find             src/mve/sfm/nearest_neighbor.cc:272-289
//   Matching::oneway_match<float>            src/mve/sfm/matching.h:114-146
//
// float_kernels.cuh forms every inner product in the reference's own summation order on CUDA
// cores; its results are bit-identical to the reference's, and it costs 2.4 ms for a pair of
// 8192 x 8192 x 128 floats.  Nearly all of that work decides nothing: a row's match is its best
// candidate unless the best and the second best are almost equal, and the ratio test's outcome is
// clear unless the ratio sits next to the threshold.  So, as on the integer path, the bulk goes
// through the tensor cores as a filter and only the rows it cannot decide are evaluated exactly:
//
//  1. float_split_kernel: every descriptor a is written as a_hi + a_lo, both representable in
//     tf32 (10-bit mantissa): a_hi = rna(a), a_lo = rna(a - a_hi); and the rows' Euclidean norms.
//  2. float_filter_kernel (tcgen05.mma kind::tf32, TMA, TMEM): S = Ahi*Bhi + Alo*Bhi + Ahi*Blo
//     accumulated in fp32, 128 query rows x 256 candidates per accumulator; the epilogue keeps
//     each row's two largest similarities and the index of the largest.  Both directions of the
//     pair are work items of one launch.
//  3. float_decide_kernel: S differs from the reference's own fp32 value by at most
//     eps = kFtEpsRel * |a| * max|b| (below).  If the best and the second best are more than
//     2 eps apart the best index is the reference's; if the ratio test has the same outcome for
//     every pair of values within eps of the filter's, that outcome is the reference's.  Rows
//     for which either fails go to a list ...
//  4. ... which float_oneway_kernel evaluates exactly as before.  The match vectors are therefore
//     still the reference's bit for bit, not merely within its tie tolerance.
//
// The error bound.  Dropped by the split: a_lo*b_lo and the two rounding residuals, at most
// 3 * 2^-22 |a_k b_k| per term.  Accumulation in the tensor core: 48 instructions of 8 products
// each onto an fp32 accumulator; with truncation after alignment to the largest exponent each
// instruction loses at most 9 ulp of the largest magnitude involved, so at most 48 * 9 * 2^-23
// of sum |a_k b_k| in all = 5.2e-5.  The reference's own evaluation (four partial sums of 32
// separately rounded products and adds, then two adds): at most 35 * 2^-24 = 2.1e-6.  With
// sum |a_k b_k| <= |a| |b| (Cauchy-Schwarz) all of it is below 6e-5 |a| |b|; what is observed
// is around 2e-7 (tests/test_gpu_parity.py::test_float_filter_error_is_far_below_the_bound).
// Non-finite input makes eps non-finite and every row of the pair goes to the exact kernel.
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "float_kernels.cuh"
#include "ptx.cuh"

namespace osfm {

constexpr int kFtM = 128;                          // query rows per work item (one accumulator)
constexpr int kFtN = 256;                          // candidates per tile
constexpr int kFtChunks = kFDim * 4 / 128;         // a row is four 128-byte chunks of 32 floats
constexpr int kFtStages = 3;                       // candidate-chunk ring
constexpr int kFtQChunkBytes = kFtM * 128;         // 16 KB
constexpr int kFtCStageBytes = kFtN * 128;         // 32 KB
constexpr int kFtSmemQhi = 0;
constexpr int kFtSmemQlo = kFtChunks * kFtQChunkBytes;
constexpr int kFtSmemC = 2 * kFtChunks * kFtQChunkBytes;
constexpr int kFtSmemBar = kFtSmemC + kFtStages * kFtCStageBytes;
constexpr int kFtSmemTmemPtr = kFtSmemBar + 128;
constexpr int kFtSmemBytes = kFtSmemBar + 256 + 1024;   // + slack to align the base to 1024
constexpr int kFtThreads = 192;                    // TMA producer, MMA issuer, four epilogue warps
constexpr int kFtTmemCols = 512;                   // two accumulators of 256 fp32 columns
constexpr float kFtEpsRel = 6.0e-5f;

static_assert(kFtSmemBytes <= 232448, "shared memory budget of one CTA");

// tcgen05 instruction descriptor for kind::tf32 (dense, K-major A and B, fp32 D).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4)                               // D format: F32
           | (2u << 7)                             // A format: TF32
           | (2u << 10)                            // B format: TF32
           | (static_cast<uint32_t>(n >> 3) << 17)
           | (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, tf32 operands (fp32 words, low 13 mantissa bits ignored),
// fp32 accumulate.  K = 8 per instruction.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ float round_to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// One warp per row of a zero-padded n_pad x 128 float matrix: hi / lo parts and the row's norm
// (rounded up); *maxnorm_bits receives the largest norm as an int (norms are non-negative, so
// their bit patterns order like the values; a NaN ends up above everything).  Rows >= n of hi and
// lo are zero.
__global__ void __launch_bounds__(256) float_split_kernel(const float* __restrict__ src, int n, int n_pad,
                                                          float* __restrict__ hi, float* __restrict__ lo,
                                                          float* __restrict__ norm, int* __restrict__ maxnorm_bits)
{
    __shared__ int bmax;
    if (threadIdx.x == 0) bmax = 0;
    __syncthreads();
    int const lane = threadIdx.x & 31;
    int const row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row < n_pad) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < n) a = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(row) * kFDim) + lane);
        float4 h, l;
        h.x = round_to_tf32(a.x); l.x = round_to_tf32(a.x - h.x);
        h.y = round_to_tf32(a.y); l.y = round_to_tf32(a.y - h.y);
        h.z = round_to_tf32(a.z); l.z = round_to_tf32(a.z - h.z);
        h.w = round_to_tf32(a.w); l.w = round_to_tf32(a.w - h.w);
        reinterpret_cast<float4*>(hi + static_cast<int64_t>(row) * kFDim)[lane] = h;
        reinterpret_cast<float4*>(lo + static_cast<int64_t>(row) * kFDim)[lane] = l;
        if (row < n) {
            float s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            float const nr = fabsf(sqrtf(s) * 1.0001f);
            if (lane == 0) {
                norm[row] = nr;
                atomicMax(&bmax, __float_as_int(nr));
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && bmax != 0) atomicMax(maxnorm_bits, bmax);
}

// What the filter knows about a row: its two largest similarities and where the largest is.
struct FloatTopRow {
    float s1, s2;
    int j1, pad;
};

// Both directions of one pair: work items [0, items_1) are 128-row blocks of set 1 against set 2,
// the rest 128-row blocks of set 2 against set 1.  top: n_1 records, then n_2.
__global__ void __launch_bounds__(kFtThreads, 1)
float_filter_kernel(const __grid_constant__ CUtensorMap tmap_hi1, const __grid_constant__ CUtensorMap tmap_lo1,
                    const __grid_constant__ CUtensorMap tmap_hi2, const __grid_constant__ CUtensorMap tmap_lo2,
                    int n_1, int n_2, FloatTopRow* __restrict__ top)
{
    extern __shared__ uint8_t ft_smem_raw[];
    uint32_t const smem_base = (smem_u32(ft_smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = ft_smem_raw + (smem_base - smem_u32(ft_smem_raw));
    uint32_t const bar = smem_base + kFtSmemBar;
    uint32_t const q_full = bar, q_empty = bar + 8;
    auto c_full = [&](int i) { return bar + 16u + 8u * i; };
    auto c_empty = [&](int i) { return bar + 16u + 8u * (kFtStages + i); };
    auto acc_full = [&](int i) { return bar + 16u + 8u * (2 * kFtStages + i); };
    auto acc_empty = [&](int i) { return bar + 16u + 8u * (2 * kFtStages + 2 + i); };

    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const items_1 = (n_1 + kFtM - 1) / kFtM, items_2 = (n_2 + kFtM - 1) / kFtM;
    int const nitems = items_1 + items_2;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < kFtStages; ++i) { mbar_init(c_full(i), 1); mbar_init(c_empty(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_hi1); prefetch_tensormap(&tmap_lo1);
        prefetch_tensormap(&tmap_hi2); prefetch_tensormap(&tmap_lo2);
    }
    if (warp == 1) {
        tmem_alloc(smem_base + kFtSmemTmemPtr, kFtTmemCols);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    uint32_t const tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + kFtSmemTmemPtr);

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            uint32_t ic = 0, cnt = 0;
            for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++ic) {
                bool const fwd = it < items_1;
                int const q_row = (fwd ? it : it - items_1) * kFtM;
                int const n_c = fwd ? n_2 : n_1;
                const CUtensorMap* const qh = fwd ? &tmap_hi1 : &tmap_hi2;
                const CUtensorMap* const ql = fwd ? &tmap_lo1 : &tmap_lo2;
                const CUtensorMap* const ch = fwd ? &tmap_hi2 : &tmap_hi1;
                const CUtensorMap* const cl = fwd ? &tmap_lo2 : &tmap_lo1;
                mbar_wait(q_empty, (ic & 1) ^ 1, 101, ic);
                mbar_arrive_expect_tx(q_full, 2 * kFtChunks * kFtQChunkBytes);
                for (int kc = 0; kc < kFtChunks; ++kc) {
                    tma_load_2d(smem_base + kFtSmemQhi + kc * kFtQChunkBytes, qh, q_full, kc * 128, q_row);
                    tma_load_2d(smem_base + kFtSmemQlo + kc * kFtQChunkBytes, ql, q_full, kc * 128, q_row);
                }
                int const ntiles = (n_c + kFtN - 1) / kFtN;
                for (int t = 0; t < ntiles; ++t)
                    for (int kc = 0; kc < kFtChunks; ++kc)
                        for (int part = 0; part < 2; ++part, ++cnt) {
                            int const s = cnt % kFtStages;
                            mbar_wait(c_empty(s), ((cnt / kFtStages) & 1) ^ 1, 102, cnt);
                            mbar_arrive_expect_tx(c_full(s), kFtCStageBytes);
                            uint32_t const dst = smem_base + kFtSmemC + s * kFtCStageBytes;
                            const CUtensorMap* const tm = part == 0 ? ch : cl;
                            tma_load_2d(dst, tm, c_full(s), kc * 128, t * kFtN);
                            tma_load_2d(dst + kFtCStageBytes / 2, tm, c_full(s), kc * 128, t * kFtN + kFtN / 2);
                        }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (the warp stays converged; one elected lane issues) ====
        uint32_t const idesc = make_idesc_tf32(kFtM, kFtN);
        uint32_t ic = 0, cnt = 0, tc = 0;
        for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++ic) {
            int const n_c = it < items_1 ? n_2 : n_1;
            int const ntiles = (n_c + kFtN - 1) / kFtN;
            mbar_wait(q_full, ic & 1, 103, ic);
            for (int t = 0; t < ntiles; ++t, ++tc) {
                int const acc = tc & 1;
                mbar_wait(acc_empty(acc), ((tc >> 1) & 1) ^ 1, 104, tc);
                uint32_t const d_tmem = tmem_base + acc * kFtN;
                for (int kc = 0; kc < kFtChunks; ++kc) {
                    uint64_t const ah = make_smem_desc_sw128(smem_base + kFtSmemQhi + kc * kFtQChunkBytes);
                    uint64_t const al = make_smem_desc_sw128(smem_base + kFtSmemQlo + kc * kFtQChunkBytes);
                    // candidate hi chunk: Ahi * Bhi + Alo * Bhi
                    int s = cnt % kFtStages;
                    mbar_wait(c_full(s), (cnt / kFtStages) & 1, 105, cnt);
                    tc_fence_after_sync();
                    if (elect_one_sync()) {
                        uint64_t const b = make_smem_desc_sw128(smem_base + kFtSmemC + s * kFtCStageBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)      // +2 in the start-address field = 32 bytes = 8 floats along K
                            mma_tf32_ss(d_tmem, ah + 2 * k, b + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_tf32_ss(d_tmem, al + 2 * k, b + 2 * k, idesc, 1u);
                        mma_commit(c_empty(s));
                    }
                    __syncwarp();
                    ++cnt;
                    // candidate lo chunk: Ahi * Blo
                    s = cnt % kFtStages;
                    mbar_wait(c_full(s), (cnt / kFtStages) & 1, 106, cnt);
                    tc_fence_after_sync();
                    if (elect_one_sync()) {
                        uint64_t const b = make_smem_desc_sw128(smem_base + kFtSmemC + s * kFtCStageBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_tf32_ss(d_tmem, ah + 2 * k, b + 2 * k, idesc, 1u);
                        mma_commit(c_empty(s));
                        if (kc == kFtChunks - 1) mma_commit(acc_full(acc));
                    }
                    __syncwarp();
                    ++cnt;
                }
            }
            if (elect_one_sync()) mma_commit(q_empty);      // the query block may be replaced
            __syncwarp();
        }
    } else {
        // ===================== epilogue: warp w reads TMEM lanes 32 (w mod 4) ... =====================
        int const quad = warp & 3;
        int const row = quad * 32 + lane;
        uint32_t tc = 0;
        float const ninf = __int_as_float(0xff800000);
        for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
            bool const fwd = it < items_1;
            int const n_q = fwd ? n_1 : n_2, n_c = fwd ? n_2 : n_1;
            int const q_row = (fwd ? it : it - items_1) * kFtM + row;
            int const ntiles = (n_c + kFtN - 1) / kFtN;
            float s1 = ninf, s2 = ninf;
            int j1 = 0;
            for (int t = 0; t < ntiles; ++t, ++tc) {
                int const acc = tc & 1;
                mbar_wait(acc_full(acc), (tc >> 1) & 1, 107, tc);
                tc_fence_after_sync();
                uint32_t const taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kFtN;
                bool const ragged = (t + 1) * kFtN > n_c;
#pragma unroll 1
                for (int c = 0; c < kFtN / 32; ++c) {
                    int32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    if (c == kFtN / 32 - 1) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(acc));
                    }
                    int const col0 = t * kFtN + c * 32;
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        float x = __int_as_float(v[q]);
                        if (ragged && col0 + q >= n_c) x = ninf;
                        bool const g1 = x > s1;
                        s2 = g1 ? s1 : fmaxf(s2, x);
                        j1 = g1 ? col0 + q : j1;
                        s1 = g1 ? x : s1;
                    }
                }
            }
            if (q_row < n_q) {
                FloatTopRow r;
                r.s1 = s1; r.s2 = s2; r.j1 = j1; r.pad = 0;
                top[(fwd ? 0 : n_1) + q_row] = r;
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, kFtTmemCols);
    }
}

// One thread per row of either direction: the rows the filter decides get their result, the
// others are listed for float_oneway_kernel.  lists: n_1 + n_2 slots, direction 0's rows from the
// front (count[0]), direction 1's from slot n_1 (count[1]).
__global__ void __launch_bounds__(256) float_decide_kernel(const FloatTopRow* __restrict__ top,
                                                           const float* __restrict__ norm_1, const float* __restrict__ norm_2,
                                                           const int* __restrict__ maxnorm_bits,   // [0]: set 1, [1]: set 2
                                                           int n_1, int n_2, float sq_lowe, float sq_dist,
                                                           int32_t* __restrict__ out, int32_t* __restrict__ lists,
                                                           int* __restrict__ count)
{
    int const g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_1 + n_2) return;
    bool const fwd = g < n_1;
    FloatTopRow const r = top[g];
    float const nq = fwd ? norm_1[g] : norm_2[g - n_1];
    float const nc = __int_as_float(maxnorm_bits[fwd ? 1 : 0]);
    // (double: the interval arithmetic below must not add rounding of its own)
    double const eps = static_cast<double>(kFtEpsRel) * nq * nc + 1e-30;
    double const s1 = r.s1, s2 = r.s2;
    // the reference starts from best = second = 0 and ignores negative similarities
    // (nearest_neighbor.cc:276-284): stay away from that corner as well
    bool decided = (s1 - s2 > 2.0 * eps) && (s2 > 2.0 * eps);
    int result = -1;
    if (decided) {
        // d = max(0, 2 - 2 ip) as the reference rounds it: within 2 eps + 5e-7 of 2 - 2 s
        double const ed = 2.0 * eps + 5e-7;
        double const d1 = 2.0 - 2.0 * s1, d2 = 2.0 - 2.0 * s2;
        double const d1lo = fmax(0.0, d1 - ed), d1hi = fmax(0.0, d1 + ed);
        double const d2lo = d2 - ed, d2hi = d2 + ed;
        double const tl = sq_lowe, td = sq_dist;
        if (!(d2lo > 0.0)) decided = false;                       // 0 / 0 and friends: exact
        else if (d1lo > td * (1.0 + 1e-6)) result = -1;           // matching.h:138
        else if (!(d1hi < td * (1.0 - 1e-6)) && !(td > 3.0e38)) decided = false;
        else if (d1lo / d2hi > tl * (1.0 + 1e-6)) result = -1;    // :140-143
        else if (d1hi / d2lo < tl * (1.0 - 1e-6)) result = r.j1;
        else decided = false;
    }
    if (decided) {
        out[g] = result;
    } else {
        int const slot = atomicAdd(count + (fwd ? 0 : 1), 1);
        lists[(fwd ? 0 : n_1) + slot] = fwd ? g : g - n_1;
    }
}

}  // namespace osfm
