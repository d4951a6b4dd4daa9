// scan_kernel.cuh -- the hot kernel: all-pairs 8-bit descriptor similarity on tcgen05
// tensor cores with a fused running top-2 epilogue.  sm_100a only.
//
// What it replaces (reference, paths relative to /root/reference):
//   short_inner_prod<T> + the top-2 scan of NearestNeighbor<T>::find
//   (src/mve/sfm/nearest_neighbor.cc:62-129, 216-268), called once per query by
//   Matching::oneway_match<T> (src/mve/sfm/matching.h:114-146).
//
// A *job* is one direction of one image pair: every descriptor of a query set against
// every descriptor of a candidate view.  A *work item* is a 256-row block of a job's
// queries.  For each item a persistent CTA
//   - TMA-loads the 256 x 128 B query tile once (two 128-row halves, SWIZZLE_128B, K-major),
//   - streams the candidate view through a ring of 256 x 128 B tiles; each candidate tile
//     feeds two tcgen05.mma groups (one per query half), which halves the L2 -> SMEM
//     traffic per similarity,
//   - issues tcgen05.mma kind::i8 (M=128, N=256, 4 x K=32; u8 x u8 or s8 x s8 -> s32,
//     exact) into the TMEM accumulator of that half (2 x 256 columns = all of TMEM),
//   - 16 epilogue warps (4 per TMEM lane quadrant, 64 columns each) pull the accumulator
//     into registers with tcgen05.ld, hand the TMEM stage straight back to the MMA warp,
//     and then reduce from registers.  The similarity matrix never leaves the SM.
//
// Per row the epilogue keeps, over "sub-chunks" of 16 consecutive candidates:
//   v1  = the largest similarity (exact),
//   pos = index of the LAST sub-chunk that contains v1 (the reference's ">=" makes the
//         highest index win ties, nearest_neighbor.cc:87-100),
//   v2  = the second largest sub-chunk maximum, clamped below at 0 -- a lower bound on the
//         reference's second-best inner product, exact unless best and second best share
//         a sub-chunk.
// That costs 8 three-input integer max instructions per 16 similarities instead of a
// 3-instruction top-2 update per similarity.  classify_kernel / refine_kernel
// (post_kernels.cuh) turn (v1, pos, v2) into the exact reference result: rows whose ratio
// test already fails with the lower bound are rejected for good (the test is monotone in
// v2); only the remaining candidate rows re-evaluate their 16-candidate window exactly.
//
// EXACT = true is the second, rare pass over the rows whose best similarity reached 2^16:
// there the reference's 16-bit lanes and 16-bit stores wrap (nearest_neighbor.cc:75-100)
// and its result depends on the scan order.  Those query rows are gathered into a scratch
// pool; the same MMA pipeline recomputes their similarities and four epilogue warps replay
// the reference's sequential scan per row, in column order, emulating the wrapped lanes
// for the few candidates that need it.
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "ptx.cuh"

namespace osfm {

constexpr int kStages = 4;            // candidate-tile ring depth
constexpr int kColGroups = 4;         // epilogue warps per TMEM lane quadrant
constexpr int kEpilogueWarps = 4 * kColGroups;
constexpr int kColsPerWarp = kBlockN / kColGroups;   // 64 = two chunks
constexpr int kProducerWarp = kEpilogueWarps;
constexpr int kMmaWarp = kEpilogueWarps + 1;
constexpr int kScanThreads = (kEpilogueWarps + 2) * 32;   // 576
constexpr int kTmemCols = 512;

constexpr int kAHalfBytes = kHalfM * kRowBytes;    // 16 KB
constexpr int kATileBytes = kItemM * kRowBytes;    // 32 KB
constexpr int kBTileBytes = kBlockN * kRowBytes;   // 32 KB
constexpr int kSmemA = 0;
constexpr int kSmemB = 2 * kATileBytes;
constexpr int kSmemBar = kSmemB + kStages * kBTileBytes;
constexpr int kNumBars = 2 + 2 + 2 * kStages + 2 + 2;
constexpr int kSmemTmemPtr = kSmemBar + kNumBars * 8;
constexpr int kSmemMerge = (kSmemTmemPtr + 4 + 15) & ~15;
constexpr int kSmemTotal = kSmemMerge + (kColGroups - 1) * kItemM * 16;
constexpr int kScanSmemBytes = kSmemTotal + 1024;  // slack for manual 1024-byte alignment

// Hang-report codes (see ptx.cuh).
enum : uint32_t {
    kWaitAEmpty = 1, kWaitBEmpty = 2, kWaitAFull = 3, kWaitAccEmpty = 4,
    kWaitBFull = 5, kWaitAccFull = 6
};

// Extra arguments of the EXACT pass.
struct ExactParams {
    const uint8_t* qpool;        // gathered query rows (what tmap_q describes)
    const uint8_t* cpool;        // candidate pool (what tmap_c describes)
    const int64_t* xrow_map;     // gathered row -> index into oneway[]
    int32_t* oneway;
    const int* total_items_dev;  // number of work items, computed on the device
    float sq_lowe, sq_dist;
    int64_t* replay_list;        // rows that need the full replay on CUDA cores
    unsigned long long* replay_count;
    int4* big_list;              // (row lo, row hi, column, similarity) of every big candidate met
    unsigned long long* big_count;
};

constexpr int kMaxBigPerRow = 4;   // big candidates per row that verify_big_kernel will certify

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

// Maximum of 16 values: 8 three-input max instructions, depth 3.
__device__ __forceinline__ int max16(const int32_t* v) {
    int a0 = max3(v[0], v[1], v[2]);
    int a1 = max3(v[3], v[4], v[5]);
    int a2 = max3(v[6], v[7], v[8]);
    int a3 = max3(v[9], v[10], v[11]);
    int a4 = max3(v[12], v[13], v[14]);
    int b0 = max3(a0, a1, a2);
    int b1 = max3(a3, a4, v[15]);
    return max(b0, b1);
}

__device__ __forceinline__ int max32(const int32_t (&v)[32]) { return max(max16(v), max16(v + 16)); }

__device__ __forceinline__ void mask_chunk(int32_t (&v)[32], int first_col, int ncols) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
        if (first_col + j >= ncols) v[j] = kMasked;
}

// Per-row running state of the fast epilogue, over sub-chunks of 16 candidates.
struct RowState {
    int v1, pos, v2;
    __device__ __forceinline__ void init() { v1 = kInitV1; pos = 0; v2 = 0; }
    // Folds the maxima q0..q3 of four consecutive sub-chunks, the first of which has index
    // `first` (= tile * 16 + sub-chunk in tile).  ">=" lets the later sub-chunk win ties.
    __device__ __forceinline__ void fold4(int q0, int q1, int q2, int q3, int first) {
        int const a = max(q0, q1), b = min(q0, q1);
        int const c = max(q2, q3), d = min(q2, q3);
        int const top = max(a, c);
        int const sec = max3(min(a, c), b, d);
        int const idx = (c >= a) ? (q3 >= q2 ? 3 : 2) : (q1 >= q0 ? 1 : 0);
        v2 = max3(min(v1, top), v2, sec);
        if (top >= v1) { v1 = top; pos = first + idx; }
    }
    // Merges the state another warp accumulated over a disjoint set of sub-chunks.
    __device__ __forceinline__ void merge(int ov1, int opos, int ov2) {
        v2 = max3(min(v1, ov1), v2, ov2);
        if (ov1 > v1 || (ov1 == v1 && opos > pos)) { v1 = ov1; pos = opos; }
    }
};

// MODE 0: normal.  1: epilogue only hands the accumulator back (MMA/TMA ceiling).
// 2: epilogue reads TMEM but reduces nothing (TMEM-read ceiling).  3: additionally dump the
// raw similarity tile to `dump` (row-major, leading dimension dump_ld) -- debug only.
template <int MODE, bool EXACT>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
            const ScanJob* __restrict__ jobs, int total_items_host, int4* __restrict__ rowres,
            uint32_t idesc, int32_t* __restrict__ dump, int64_t dump_ld, ExactParams ex)
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
    int const total_items = EXACT ? *ex.total_items_dev : total_items_host;
    // In the EXACT pass one group of four warps per query half scans whole rows in order.
    constexpr uint32_t kAccEmptyCount = EXACT ? 4 : kEpilogueWarps;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full(i), 1);
            mbar_init(a_empty(i), 1);
            mbar_init(acc_full(i), 1);
            mbar_init(acc_empty(i), kAccEmptyCount);  // one arrive per participating warp
        }
        for (int i = 0; i < kStages; ++i) {
            mbar_init(b_full(i), 1);
            mbar_init(b_empty(i), 1);
        }
        fence_barrier_init();
    }
    if (warp == kProducerWarp && lane == 0) {
        prefetch_tensormap(&tmap_q);
        prefetch_tensormap(&tmap_c);
    }
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
                int const nh = (job.q_n - rb * kItemM > kHalfM) ? 2 : 1;
                int const abuf = ic & 1;
                mbar_wait(a_empty(abuf), ((ic >> 1) & 1) ^ 1, kWaitAEmpty, ic);
                mbar_arrive_expect_tx(a_full(abuf), nh * kAHalfBytes);
                for (int h = 0; h < nh; ++h)
                    tma_load_2d(smem_base + kSmemA + abuf * kATileBytes + h * kAHalfBytes, &tmap_q,
                                a_full(abuf), 0, job.q_row + rb * kItemM + h * kHalfM);
                int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;
                for (int t = 0; t < ntiles; ++t, ++bcnt) {
                    int const s = bcnt % kStages;
                    mbar_wait(b_empty(s), ((bcnt / kStages) & 1) ^ 1, kWaitBEmpty, bcnt);
                    mbar_arrive_expect_tx(b_full(s), kBTileBytes);
                    uint32_t const dst = smem_base + kSmemB + s * kBTileBytes;
                    int const row = job.c_row + t * kBlockN;
                    tma_load_2d(dst, &tmap_c, b_full(s), 0, row);
                    tma_load_2d(dst + kBTileBytes / 2, &tmap_c, b_full(s), 0, row + kBlockN / 2);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            int j = 0;
            uint32_t bcnt = 0, ic = 0;
            uint32_t hcnt[2] = {0, 0};  // uses of each accumulator stage so far
            for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
                while (it >= jobs[j + 1].item_start) ++j;
                int const c_n = jobs[j].c_n;
                int const rb = it - jobs[j].item_start;
                int const nh = (jobs[j].q_n - rb * kItemM > kHalfM) ? 2 : 1;
                int const abuf = ic & 1;
                mbar_wait(a_full(abuf), (ic >> 1) & 1, kWaitAFull, ic);
                uint64_t const adesc0 = make_smem_desc_sw128(smem_base + kSmemA + abuf * kATileBytes);
                int const ntiles = (c_n + kBlockN - 1) / kBlockN;
                for (int t = 0; t < ntiles; ++t, ++bcnt) {
                    int const s = bcnt % kStages;
                    mbar_wait(b_full(s), (bcnt / kStages) & 1, kWaitBFull, bcnt);
                    uint64_t const bdesc = make_smem_desc_sw128(smem_base + kSmemB + s * kBTileBytes);
                    for (int h = 0; h < nh; ++h) {
                        mbar_wait(acc_empty(h), (hcnt[h] & 1) ^ 1, kWaitAccEmpty, hcnt[h]);
                        ++hcnt[h];
                        tc_fence_after_sync();
                        uint64_t const adesc = adesc0 + static_cast<uint64_t>(h * (kAHalfBytes >> 4));
                        uint32_t const d_tmem = tmem_base + h * kBlockN;
#pragma unroll
                        for (int k = 0; k < kRowBytes / 32; ++k) {
                            // +2 in the start-address field = 32 bytes along K inside the swizzle span
                            mma_i8_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
                        }
                        mma_commit(acc_full(h));    // accumulator of this half is ready
                    }
                    mma_commit(b_empty(s));         // candidate stage may be refilled
                }
                mma_commit(a_empty(abuf));          // query tile may be overwritten
            }
        }
    } else if (!EXACT) {
        // ===================== epilogue: 16 warps, 64 columns each =====================
        int const quad = warp & 3;           // TMEM lane quadrant this warp may access
        int const cg = warp >> 2;            // column group: columns [64 cg, 64 cg + 64)
        int const row = quad * 32 + lane;    // row inside a 128-row half
        int const s0 = cg * (kColsPerWarp / kSub);   // first sub-chunk (of 16) this warp owns in a tile
        uint32_t const taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + cg * kColsPerWarp;
        int4* const merge = reinterpret_cast<int4*>(smem_gen + kSmemMerge);

        int j = 0;
        uint32_t hcnt[2] = {0, 0};
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            while (it >= jobs[j + 1].item_start) ++j;
            ScanJob const job = jobs[j];
            int const rb = it - job.item_start;
            int const nh = (job.q_n - rb * kItemM > kHalfM) ? 2 : 1;
            int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;

            RowState st[2];
            st[0].init();
            st[1].init();
            for (int t = 0; t < ntiles; ++t) {
                int const ncols = job.c_n - t * kBlockN;   // valid columns of this tile
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h >= nh) break;
                    mbar_wait(acc_full(h), hcnt[h] & 1, kWaitAccFull, hcnt[h]);
                    ++hcnt[h];
                    tc_fence_after_sync();
                    int32_t va[32], vb[32];
                    if (MODE != 1) {
                        uint32_t const taddr = taddr0 + h * kBlockN;
                        tmem_ld_32x32b_x32(taddr, va);
                        tmem_ld_32x32b_x32(taddr + kChunk, vb);
                        tmem_ld_wait_regs(va);
                        tmem_ld_wait_regs(vb);
                    }
                    // the data is in registers: hand the TMEM stage back before reducing
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(h));

                    if (MODE == 0 || MODE == 3) {
                        if (MODE == 3) {
                            int64_t const r = static_cast<int64_t>(rb) * kItemM + h * kHalfM + row;
                            if (r < job.q_n) {
                                int32_t* d = dump + r * dump_ld + t * kBlockN + cg * kColsPerWarp;
#pragma unroll
                                for (int q = 0; q < 32; ++q) { d[q] = va[q]; d[32 + q] = vb[q]; }
                            }
                        }
                        if (ncols < kBlockN) {
                            mask_chunk(va, cg * kColsPerWarp, ncols);
                            mask_chunk(vb, cg * kColsPerWarp + kChunk, ncols);
                        }
                        st[h].fold4(max16(va), max16(va + 16), max16(vb), max16(vb + 16), t * kSubsPerTile + s0);
                    } else if (MODE == 2) {
                        st[h].v1 |= va[0] | vb[0];
                    }
                }
            }

            // combine the four column groups' partial row states (groups 1..3 -> smem -> group 0)
            if (cg != 0) {
                for (int h = 0; h < nh; ++h)
                    merge[(cg - 1) * kItemM + h * kHalfM + row] = make_int4(st[h].v1, st[h].pos, st[h].v2, 0);
            }
            named_barrier_sync(1, kEpilogueWarps * 32);
            if (cg == 0) {
                for (int h = 0; h < nh; ++h) {
#pragma unroll
                    for (int g = 0; g < kColGroups - 1; ++g) {
                        int4 const o = merge[g * kItemM + h * kHalfM + row];
                        st[h].merge(o.x, o.y, o.z);
                    }
                    int64_t const r = static_cast<int64_t>(rb) * kItemM + h * kHalfM + row;
                    if (r < job.q_n) rowres[job.out_row + r] = make_int4(st[h].v1, st[h].pos, st[h].v2, j);
                }
            }
            named_barrier_sync(2, kEpilogueWarps * 32);
        }
    } else if ((warp >> 2) < 2) {
        // ===================== EXACT epilogue: 2 x 4 warps replay the reference's scan ======
        int const my_h = warp >> 2;          // the query half this group owns
        int const quad = warp & 3;
        int const row = quad * 32 + lane;
        uint32_t const taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + my_h * kBlockN;

        int j = 0;
        uint32_t hcnt = 0;
        for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
            while (it >= jobs[j + 1].item_start) ++j;
            ScanJob const job = jobs[j];
            int const rb = it - job.item_start;
            int const nh = (job.q_n - rb * kItemM > kHalfM) ? 2 : 1;
            if (my_h >= nh) continue;
            int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;
            int const r_in_job = rb * kItemM + my_h * kHalfM + row;
            // rows past the end of the job hold whatever follows in the scratch pool; they
            // must not drag the warp into the update path
            bool const live = r_in_job < job.q_n;

            // Reference state (nearest_neighbor.cc:246-249), replayed in column order.
            //
            // A candidate >= 2^16 ("big") enters with the value the tensor core computed.  That
            // equals the reference's lane-wise 16-bit sum unless one of the eight lanes itself
            // reached 2^16, which needs an adversarial descriptor; every big candidate is
            // therefore recorded and verify_big_kernel re-checks it with the lane emulation
            // afterwards.  Rows that fail the check (or have more than kMaxBigPerRow big
            // candidates) are replayed on CUDA cores by slow_rows_kernel, which overwrites the
            // result written here.
            int b1 = 0, b2 = 0, i1 = 0, nbig = 0;
            int64_t const g = live ? ex.xrow_map[job.out_row + r_in_job] : 0;
            for (int t = 0; t < ntiles; ++t) {
                int const ncols = job.c_n - t * kBlockN;
                mbar_wait(acc_full(my_h), hcnt & 1, kWaitAccFull, hcnt);
                ++hcnt;
                tc_fence_after_sync();
                int32_t v[32], vn[32];
                tmem_ld_32x32b_x32(taddr0, v);
#pragma unroll 1
                for (int c = 0; c < kChunksPerTile; ++c) {
                    tmem_ld_wait_regs(v);
                    // prefetch the next chunk while this one is processed
                    if (c + 1 < kChunksPerTile) {
                        tmem_ld_32x32b_x32(taddr0 + (c + 1) * kChunk, vn);
                    } else {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(my_h));
                    }
                    if (ncols < kBlockN) mask_chunk(v, c * kChunk, ncols);
                    // b2 <= 65535, so a big candidate always triggers
                    int const cmax = max32(v);
                    bool const trig = live && cmax >= b2;
                    if (__any_sync(0xffffffffu, trig)) {
                        int const col0 = t * kBlockN + c * kChunk;
                        if (__any_sync(0xffffffffu, trig && cmax >= 65536)) {
                            // rare: a triggered lane meets a big candidate
                            if (trig) {
#pragma unroll
                                for (int q = 0; q < 32; ++q) {
                                    int const x = v[q];
                                    if (x >= b2) {
                                        if (x >= 65536) {
                                            if (nbig < kMaxBigPerRow)
                                                ex.big_list[atomicAdd(ex.big_count, 1ull)] =
                                                    make_int4(static_cast<int>(g), static_cast<int>(g >> 32), col0 + q, x);
                                            ++nbig;
                                        }
                                        ref_scan_step<false>(x, col0 + q, b1, b2, i1);
                                    }
                                }
                            }
                        } else {
                            // common: plain sequential top-2 with the reference's tie rule,
                            // branch-free (values < 2^16 are stored untruncated)
                            int s1 = b1, s2 = b2, si = i1;
#pragma unroll
                            for (int q = 0; q < 32; ++q) {
                                int const x = v[q];
                                bool const ge2 = x >= s2;
                                bool const ge1 = ge2 && x >= s1;
                                s2 = ge1 ? s1 : (ge2 ? x : s2);
                                s1 = ge1 ? x : s1;
                                si = ge1 ? col0 + q : si;
                            }
                            if (trig) { b1 = s1; b2 = s2; i1 = si; }
                        }
                    }
                    if (c + 1 < kChunksPerTile) {
                        tmem_ld_wait_regs(vn);
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = vn[q];
                    }
                }
            }
            if (live) {
                bool const ok = passes_tests(ip_to_dist<false>(b1), ip_to_dist<false>(b2), ex.sq_lowe, ex.sq_dist);
                ex.oneway[g] = ok ? i1 : -1;
                if (nbig > kMaxBigPerRow)   // cannot be certified by verify_big_kernel
                    ex.replay_list[atomicAdd(ex.replay_count, 1ull)] = g;
            }
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
