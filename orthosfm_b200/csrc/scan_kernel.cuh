// scan_kernel.cuh -- the hot kernel: all-pairs 8-bit descriptor similarity on tcgen05
// tensor cores with a fused running top-2 epilogue.  sm_100a only.
//
// What it replaces (reference, paths relative to /root/reference):
//   short_inner_prod<T> + the top-2 scan of NearestNeighbor<T>::find
//   (src/mve/sfm/nearest_neighbor.cc:62-129, 216-268), called once per query by
//   Matching::oneway_match<T> (src/mve/sfm/matching.h:114-146).
//
// A *job* is one direction of one image pair: every descriptor of a query view
// against every descriptor of a candidate view.  A *work item* is a 128-row block of a
// job's queries.  For each item a persistent CTA
//   - TMA-loads the 128 x 128 B query tile once (SWIZZLE_128B, K-major),
//   - streams the candidate view through a ring of 256 x 128 B tiles,
//   - issues tcgen05.mma kind::i8 (M=128, N=256, 4 x K=32) into one of two 256-column
//     TMEM accumulator stages (u8 x u8 or s8 x s8 -> s32, exact),
//   - two epilogue warp-groups (one per TMEM stage) read the accumulators with
//     tcgen05.ld and reduce each row on the fly; the similarity matrix never leaves
//     the SM.
//
// Per row the epilogue keeps, over "chunks" of 32 consecutive candidates:
//   v1  = the largest similarity (exact),
//   pos = index of the LAST chunk that contains v1 (the reference's ">=" makes the
//         highest index win ties, nearest_neighbor.cc:87-100),
//   v2  = the second largest chunk maximum, clamped below at 0 -- a lower bound on the
//         reference's second-best inner product, exact unless best and second best
//         share a chunk.
// That costs 16 three-input integer max instructions per 32 similarities instead of a
// 3-instruction top-2 update per similarity.  finalize_kernel (post_kernels.cuh) turns
// (v1, pos, v2) into the exact reference result: rows whose ratio test already fails
// with the lower bound are rejected for good (the test is monotone in v2); only the
// remaining candidate rows re-evaluate their 32-candidate chunk exactly.
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace osfm {

constexpr int kBlockM = 128;        // query rows per work item
constexpr int kBlockN = 256;        // candidate rows per tile (= TMEM columns per stage)
constexpr int kRowBytes = 128;      // descriptor row pitch in the pool (SIFT 128 B; SURF zero-padded)
constexpr int kChunk = 32;          // candidates per epilogue chunk
constexpr int kChunksPerTile = kBlockN / kChunk;  // 8 -> 3 key bits
constexpr int kStages = 5;          // candidate-tile ring depth
constexpr int kEpilogueWarps = 8;   // two groups of four (one group per TMEM stage)
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kScanThreads = 320;
constexpr int kTmemCols = 512;

constexpr int kATileBytes = kBlockM * kRowBytes;   // 16 KB
constexpr int kBTileBytes = kBlockN * kRowBytes;   // 32 KB
constexpr int kSmemA = 0;
constexpr int kSmemB = 2 * kATileBytes;
constexpr int kSmemBar = kSmemB + kStages * kBTileBytes;
constexpr int kNumBars = 2 + 2 + 2 * kStages + 2 + 2;
constexpr int kSmemTmemPtr = kSmemBar + kNumBars * 8;
constexpr int kSmemMerge = (kSmemTmemPtr + 4 + 15) & ~15;
constexpr int kSmemTotal = kSmemMerge + kBlockM * 16;
constexpr int kScanSmemBytes = kSmemTotal + 1024;  // slack for manual 1024-byte alignment

constexpr int kInitV1 = -(1 << 30);   // "nothing seen yet" (never multiplied)
constexpr int kMasked = -(1 << 24);   // similarity of a column past the end of the view;
                                      // below any real value (|s| < 2^23), and
                                      // kMasked * 8 still fits an int

// One direction of one image pair.  `item_start` is the exclusive prefix sum of
// ceil(q_n / 128) over the job list; the list carries one sentinel entry at the end.
struct ScanJob {
    int32_t q_row;       // first pool row of the query view
    int32_t q_n;         // number of query descriptors
    int32_t c_row;       // first pool row of the candidate view
    int32_t c_n;         // number of candidate descriptors
    int64_t out_row;     // first index of this job's rows in rowres[] / oneway[]
    int32_t item_start;  // first work item of this job
    int32_t c_maxnorm2;  // signed kind: largest squared norm in the candidate view
};

// Hang-report codes (see ptx.cuh).
enum : uint32_t {
    kWaitAEmpty = 1, kWaitBEmpty = 2, kWaitAFull = 3, kWaitAccEmpty = 4,
    kWaitBFull = 5, kWaitAccFull = 6
};

__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }

// Ties the 32 registers to the completion of the tcgen05.ld that produced them, so the
// compiler cannot schedule their consumers above the wait.
__device__ __forceinline__ void tmem_ld_wait_regs(int32_t (&v)[32]) {
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
          "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),
          "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]),
          "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
          "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
        :
        : "memory");
}

// Maximum of 32 values: 16 three-input max instructions, depth 4.
__device__ __forceinline__ int max32(const int32_t (&v)[32]) {
    int a0 = max3(v[0], v[1], v[2]);
    int a1 = max3(v[3], v[4], v[5]);
    int a2 = max3(v[6], v[7], v[8]);
    int a3 = max3(v[9], v[10], v[11]);
    int a4 = max3(v[12], v[13], v[14]);
    int a5 = max3(v[15], v[16], v[17]);
    int a6 = max3(v[18], v[19], v[20]);
    int a7 = max3(v[21], v[22], v[23]);
    int a8 = max3(v[24], v[25], v[26]);
    int a9 = max3(v[27], v[28], v[29]);
    int b0 = max3(a0, a1, a2);
    int b1 = max3(a3, a4, a5);
    int b2 = max3(a6, a7, a8);
    int b3 = max3(a9, v[30], v[31]);
    return max(max3(b0, b1, b2), b3);
}

// Running top-2 over chunk keys.  key = (chunk maximum << 3) | chunk-in-tile, so the
// later chunk wins ties.
__device__ __forceinline__ void push_chunk(int cmax, int c, int& t1key, int& t2key) {
    int const ckey = cmax * kChunksPerTile + c;
    t2key = max(t2key, min(t1key, ckey));
    t1key = max(t1key, ckey);
}

template <bool MASKED>
__device__ __forceinline__ void reduce_chunk(int32_t (&v)[32], int c, int ncols, int& t1key,
                                             int& t2key) {
    if (MASKED) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c * kChunk + j >= ncols) v[j] = kMasked;
    }
    push_chunk(max32(v), c, t1key, t2key);
}

// MODE 0: normal.  1: epilogue only hands the accumulator back (MMA/TMA ceiling).
// 2: epilogue reads TMEM but reduces nothing (TMEM-read ceiling).  3: dump the raw
// similarity tile to `dump` (row-major, leading dimension dump_ld) -- debug only.
template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap, const ScanJob* __restrict__ jobs,
            int total_items, int4* __restrict__ rowres, uint32_t idesc,
            int32_t* __restrict__ dump, int64_t dump_ld)
{
    extern __shared__ uint8_t smem_raw[];
    uint32_t const smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    uint32_t const bar_base = smem_base + kSmemBar;
    auto a_full = [&](int i) { return bar_base + 8u * (0 + i); };
    auto a_empty = [&](int i) { return bar_base + 8u * (2 + i); };
    auto b_full = [&](int i) { return bar_base + 8u * (4 + i); };
    auto b_empty = [&](int i) { return bar_base + 8u * (4 + kStages + i); };
    auto acc_full = [&](int i) { return bar_base + 8u * (4 + 2 * kStages + i); };
    auto acc_empty = [&](int i) { return bar_base + 8u * (6 + 2 * kStages + i); };

    int const warp = threadIdx.x >> 5;
    int const lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full(i), 1);
            mbar_init(a_empty(i), 1);
            mbar_init(acc_full(i), 1);
            mbar_init(acc_empty(i), 4);  // one arrive per warp of the owning group
        }
        for (int i = 0; i < kStages; ++i) {
            mbar_init(b_full(i), 1);
            mbar_init(b_empty(i), 1);
        }
        fence_barrier_init();
    }
    if (warp == kProducerWarp && lane == 0) prefetch_tensormap(&tmap);
    if (warp == kMmaWarp) {
        tmem_alloc(smem_base + kSmemTmemPtr, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    uint32_t const tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + kSmemTmemPtr);

    if (warp == kProducerWarp) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int j = 0;
            uint32_t bcnt = 0, ic = 0;
            for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
                while (it >= jobs[j + 1].item_start) ++j;
                ScanJob const job = jobs[j];
                int const rb = it - job.item_start;
                int const abuf = ic & 1;
                mbar_wait(a_empty(abuf), ((ic >> 1) & 1) ^ 1, kWaitAEmpty, ic);
                mbar_arrive_expect_tx(a_full(abuf), kATileBytes);
                tma_load_2d(smem_base + kSmemA + abuf * kATileBytes, &tmap, a_full(abuf), 0,
                            job.q_row + rb * kBlockM);
                int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;
                for (int t = 0; t < ntiles; ++t, ++bcnt) {
                    int const s = bcnt % kStages;
                    mbar_wait(b_empty(s), ((bcnt / kStages) & 1) ^ 1, kWaitBEmpty, bcnt);
                    mbar_arrive_expect_tx(b_full(s), kBTileBytes);
                    uint32_t const dst = smem_base + kSmemB + s * kBTileBytes;
                    int const row = job.c_row + t * kBlockN;
                    tma_load_2d(dst, &tmap, b_full(s), 0, row);
                    tma_load_2d(dst + kBTileBytes / 2, &tmap, b_full(s), 0, row + kBlockN / 2);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            int j = 0;
            uint32_t bcnt = 0, tcnt = 0, ic = 0;
            for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
                while (it >= jobs[j + 1].item_start) ++j;
                int const c_n = jobs[j].c_n;
                int const abuf = ic & 1;
                mbar_wait(a_full(abuf), (ic >> 1) & 1, kWaitAFull, ic);
                uint64_t const adesc = make_smem_desc_sw128(smem_base + kSmemA + abuf * kATileBytes);
                int const ntiles = (c_n + kBlockN - 1) / kBlockN;
                for (int t = 0; t < ntiles; ++t, ++bcnt, ++tcnt) {
                    int const s = bcnt % kStages;
                    int const as = tcnt & 1;
                    mbar_wait(acc_empty(as), ((tcnt >> 1) & 1) ^ 1, kWaitAccEmpty, tcnt);
                    mbar_wait(b_full(s), (bcnt / kStages) & 1, kWaitBFull, bcnt);
                    tc_fence_after_sync();
                    uint64_t const bdesc = make_smem_desc_sw128(smem_base + kSmemB + s * kBTileBytes);
                    uint32_t const d_tmem = tmem_base + as * kBlockN;
#pragma unroll
                    for (int k = 0; k < kRowBytes / 32; ++k) {
                        // +2 in the start-address field = 32 bytes along K inside the swizzle span
                        mma_i8_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
                    }
                    mma_commit(b_empty(s));     // candidate stage may be refilled
                    mma_commit(acc_full(as));   // accumulator ready for its epilogue group
                }
                mma_commit(a_empty(abuf));      // query tile may be overwritten
            }
        }
    } else {
        // ===================== epilogue: 2 groups x 4 warps =====================
        int const g = warp >> 2;        // group = TMEM stage it owns
        int const quad = warp & 3;      // TMEM lane quadrant this warp may access
        int const row = quad * 32 + lane;
        uint32_t const taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + g * kBlockN;
        int4* const merge = reinterpret_cast<int4*>(smem_gen + kSmemMerge);

        int j = 0;
        uint32_t tcnt = 0, ecnt = 0;
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            while (it >= jobs[j + 1].item_start) ++j;
            ScanJob const job = jobs[j];
            int const rb = it - job.item_start;
            int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;

            int r1val = kInitV1, r1pos = 0, r2val = 0;
            for (int t = 0; t < ntiles; ++t, ++tcnt) {
                if (static_cast<int>(tcnt & 1) != g) continue;
                mbar_wait(acc_full(g), ecnt & 1, kWaitAccFull, ecnt);
                ++ecnt;
                tc_fence_after_sync();

                int t1key = kInitV1, t2key = kInitV1;
                if (MODE == 0 || MODE == 3) {
                    int const ncols = job.c_n - t * kBlockN;
                    int32_t va[32], vb[32];
                    tmem_ld_32x32b_x32(taddr, va);
#pragma unroll
                    for (int c = 0; c < kChunksPerTile; c += 2) {
                        tmem_ld_wait_regs(va);
                        tmem_ld_32x32b_x32(taddr + (c + 1) * kChunk, vb);
                        if (MODE == 3) {
                            int64_t const r = static_cast<int64_t>(rb) * kBlockM + row;
                            if (r < job.q_n) {
#pragma unroll
                                for (int q = 0; q < 32; ++q)
                                    dump[r * dump_ld + t * kBlockN + c * kChunk + q] = va[q];
                            }
                        }
                        if (ncols >= kBlockN) reduce_chunk<false>(va, c, ncols, t1key, t2key);
                        else                  reduce_chunk<true>(va, c, ncols, t1key, t2key);
                        tmem_ld_wait_regs(vb);
                        if (c + 2 < kChunksPerTile) tmem_ld_32x32b_x32(taddr + (c + 2) * kChunk, va);
                        if (MODE == 3) {
                            int64_t const r = static_cast<int64_t>(rb) * kBlockM + row;
                            if (r < job.q_n) {
#pragma unroll
                                for (int q = 0; q < 32; ++q)
                                    dump[r * dump_ld + t * kBlockN + (c + 1) * kChunk + q] = vb[q];
                            }
                        }
                        if (ncols >= kBlockN) reduce_chunk<false>(vb, c + 1, ncols, t1key, t2key);
                        else                  reduce_chunk<true>(vb, c + 1, ncols, t1key, t2key);
                    }
                } else if (MODE == 2) {
                    int32_t va[32];
                    int acc = 0;
#pragma unroll
                    for (int c = 0; c < kChunksPerTile; ++c) {
                        tmem_ld_32x32b_x32(taddr + c * kChunk, va);
                        tmem_ld_wait_regs(va);
                        acc |= va[c];
                    }
                    t1key = acc;
                }
                // hand the accumulator stage back to the MMA warp
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(g));

                // fold the tile into the row state; ">=" lets the later tile win ties
                int const v1 = t1key >> 3;
                int const v2 = t2key >> 3;
                r2val = max3(min(r1val, v1), r2val, v2);
                if (v1 >= r1val) {
                    r1val = v1;
                    r1pos = t * kChunksPerTile + (t1key & (kChunksPerTile - 1));
                }
            }

            // combine the two groups' partial row states (group 1 -> smem -> group 0)
            if (g == 1) merge[row] = make_int4(r1val, r1pos, r2val, 0);
            named_barrier_sync(1, kEpilogueWarps * 32);
            if (g == 0) {
                int4 const o = merge[row];
                int const s2 = max3(min(r1val, o.x), r2val, o.z);
                if (o.x > r1val || (o.x == r1val && o.y > r1pos)) {
                    r1val = o.x;
                    r1pos = o.y;
                }
                int64_t const r = static_cast<int64_t>(rb) * kBlockM + row;
                if (r < job.q_n) rowres[job.out_row + r] = make_int4(r1val, r1pos, s2, 0);
            }
            named_barrier_sync(2, kEpilogueWarps * 32);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        __syncwarp();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace osfm
