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
// Per row the fast epilogue is a *filter*.  It reads the accumulator with
// tcgen05.ld ... .pack::16b (two adjacent columns per 32-bit register) and keeps sixteen
// packed running maxima per thread with three-input 16-bit SIMD max instructions
// (VIMNMX3.U16x2 / .S16x2): one instruction per four similarities.  At the end of a work item
// the 64 "slot" maxima of a row (16 per thread x 4 column groups) give
//   v1  = the largest similarity of the row (exact), and
//   v2  = the second largest slot maximum, clamped below at 0 -- a lower bound on the
//         reference's second-best inner product (exact unless best and second best share a
//         slot).
// The reference's ratio test is monotone in the second best, so a row that fails it with the
// lower bound fails it for good and is final (-1) right here.  The rows that pass (the
// *survivors*, typically the rows that really have a match) are appended to a per-job list and
// re-run by the EXACT pass below, which produces the reference's bytes.
//
// The 16-bit packing is only valid if no similarity of the row leaves the 16-bit range.  That
// is certified per (row, candidate view) by Cauchy-Schwarz from the squared norms computed at
// commit: |a|^2 * max|b|^2 < 2^32 (unsigned) or < 2^30 (signed; this also excludes a wrap of
// the reference's 16-bit SSE lanes).  Rows without the certificate skip the filter: unsigned
// rows join the survivors, signed rows go to the CUDA-core replay (slow_rows_kernel).
//
// EXACT = true is the second pass, over the gathered survivor rows: the same MMA pipeline
// recomputes their similarities and eight epilogue warps replay the reference's sequential
// best / second-best scan per row, in column order (nearest_neighbor.cc:87-100), including
// the 16-bit wrap of the lanes and stores for candidates that reach 2^16.
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
constexpr int kColsPerWarp = kBlockN / kColGroups;   // 64 = one packed tcgen05.ld.x32
constexpr int kSlotRegs = 4;          // packed running maxima per thread and query half (8 slots)
constexpr int kProducerWarp = kEpilogueWarps;
constexpr int kMmaWarp = kEpilogueWarps + 1;             // issuer of query half 0; half 1: the next warp
constexpr int kScanThreads = (kEpilogueWarps + 3) * 32;   // 608
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
constexpr int kMergeBufBytes = (kColGroups - 1) * kItemM * 8;
constexpr int kSmemTotal = kSmemMerge + 2 * kMergeBufBytes;   // double-buffered across items
constexpr int kScanSmemBytes = kSmemTotal + 1024;  // slack for manual 1024-byte alignment

// Hang-report codes (see ptx.cuh).
enum : uint32_t {
    kWaitAEmpty = 1, kWaitBEmpty = 2, kWaitAFull = 3, kWaitAccEmpty = 4,
    kWaitBFull = 5, kWaitAccFull = 6
};

// A survivor-list entry: the row's index into oneway[] plus what the filter knew about it,
// so that the EXACT pass can cross-check itself against the filter.
constexpr int kSurvRowBits = 40;
constexpr uint64_t kSurvRowMask = (1ull << kSurvRowBits) - 1;
constexpr uint64_t kSurvCertified = 1ull << 56;
__host__ __device__ __forceinline__ int64_t surv_entry(int64_t g, int v1, bool certified) {
    return static_cast<int64_t>(static_cast<uint64_t>(g) |
                                (static_cast<uint64_t>(static_cast<uint32_t>(v1) & 0xffffu) << kSurvRowBits) |
                                (certified ? kSurvCertified : 0ull));
}
__host__ __device__ __forceinline__ int64_t surv_row(int64_t e) {
    return static_cast<int64_t>(static_cast<uint64_t>(e) & kSurvRowMask);
}

// What the filter epilogue leaves per row for classify_kernel (post_kernels.cuh): the largest
// similarity and the lower bound on the second largest, 16 bits each, and the row's job.
__device__ __forceinline__ int2 pack_rowres(int v1, int v2, int job) {
    return make_int2(static_cast<int>((static_cast<uint32_t>(v1) & 0xffffu) | (static_cast<uint32_t>(v2) << 16)), job);
}

// Extra arguments of the EXACT pass.
struct ExactParams {
    const uint8_t* qpool;        // gathered query rows (what tmap_q describes)
    const uint8_t* cpool;        // candidate pool (what tmap_c describes)
    const int64_t* xrow_map;     // gathered row -> survivor-list entry (surv_entry)
    int32_t* oneway;
    const int* total_items_dev;  // number of work items, computed on the device
    float sq_lowe, sq_dist;
    int64_t* replay_list;        // rows that need the full replay on CUDA cores
    unsigned long long* replay_count;
    int4* big_list;              // (row lo, row hi, column, similarity) of every big candidate met
    unsigned long long* big_count;
    unsigned long long* self_check;   // filter and EXACT pass disagree on a certified row's best
};

constexpr int kTraceEvents = 256;   // per warp, MODE 5

__device__ __forceinline__ long long clock64_() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}

constexpr int kMaxBigPerRow = 4;   // big candidates per row that verify_big_kernel will certify

// Ties the 32 registers to the completion of the tcgen05.ld that produced them, so the
// compiler cannot schedule their consumers above the wait.
template <typename T>
__device__ __forceinline__ void tmem_ld_wait_regs(T (&v)[32]) {
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

// ---- 16-bit SIMD helpers (two similarities per register) ----

template <bool SIGNED>
__device__ __forceinline__ uint32_t pmax(uint32_t a, uint32_t b) {
    uint32_t r;
    if (SIGNED) asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else        asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
template <bool SIGNED>
__device__ __forceinline__ uint32_t pmin(uint32_t a, uint32_t b) {
    uint32_t r;
    if (SIGNED) asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else        asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// ptxas fuses the pair into one VIMNMX3.{U,S}16x2
template <bool SIGNED>
__device__ __forceinline__ uint32_t pmax3(uint32_t a, uint32_t b, uint32_t c) {
    return pmax<SIGNED>(pmax<SIGNED>(a, b), c);
}
template <bool SIGNED> __device__ __forceinline__ int plo(uint32_t x) {
    return SIGNED ? static_cast<int>(static_cast<short>(x & 0xffffu)) : static_cast<int>(x & 0xffffu);
}
template <bool SIGNED> __device__ __forceinline__ int phi(uint32_t x) {
    return SIGNED ? (static_cast<int>(x) >> 16) : static_cast<int>(x >> 16);
}

// Columns of a ragged last tile that lie past the end of the view hold other views' rows:
// replace them by the smallest value.  Register k of a packed load holds columns
// first_col + 2k (low half) and first_col + 2k + 1 (high half).
template <bool SIGNED>
__device__ __forceinline__ void mask_packed(uint32_t (&r)[32], int first_col, int ncols) {
    uint32_t const lo_min = SIGNED ? 0x00008000u : 0u;
    uint32_t const hi_min = SIGNED ? 0x80000000u : 0u;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        int const c = first_col + 2 * k;
        if (c >= ncols) r[k] = lo_min | hi_min;
        else if (c + 1 >= ncols) r[k] = (r[k] & 0x0000ffffu) | hi_min;
    }
}

// Folds one packed x32 load (64 columns) into the running slot registers: slot k keeps the
// maxima of the columns congruent to 2k and 2k+1 modulo 2 * kSlotRegs.
template <bool SIGNED>
__device__ __forceinline__ void fold_packed(uint32_t (&slot)[kSlotRegs], const uint32_t (&r)[32]) {
#pragma unroll
    for (int k = 0; k < kSlotRegs; ++k) {
#pragma unroll
        for (int q = 0; q < 32 / kSlotRegs; q += 2)
            slot[k] = pmax3<SIGNED>(slot[k], r[k + q * kSlotRegs], r[k + (q + 1) * kSlotRegs]);
    }
}

// Largest and second largest (with multiplicity) of the 16 slot maxima held in eight packed
// registers: a tournament per 16-bit lane -- the maximum over all "losers" of a tournament is
// its second largest entry -- followed by the merge of the two lanes.
template <bool SIGNED>
__device__ __forceinline__ void slots_top2(const uint32_t (&m)[kSlotRegs], int& v1, int& v2) {
    static_assert(kSlotRegs == 4 || kSlotRegs == 8, "tournament is written for 4 or 8 registers");
    uint32_t w, l;
    if (kSlotRegs == 8) {
        uint32_t const w01 = pmax<SIGNED>(m[0], m[1]), l01 = pmin<SIGNED>(m[0], m[1]);
        uint32_t const w23 = pmax<SIGNED>(m[2], m[3]), l23 = pmin<SIGNED>(m[2], m[3]);
        uint32_t const w45 = pmax<SIGNED>(m[4 % kSlotRegs], m[5 % kSlotRegs]), l45 = pmin<SIGNED>(m[4 % kSlotRegs], m[5 % kSlotRegs]);
        uint32_t const w67 = pmax<SIGNED>(m[6 % kSlotRegs], m[7 % kSlotRegs]), l67 = pmin<SIGNED>(m[6 % kSlotRegs], m[7 % kSlotRegs]);
        uint32_t const wa = pmax<SIGNED>(w01, w23), la = pmin<SIGNED>(w01, w23);
        uint32_t const wb = pmax<SIGNED>(w45, w67), lb = pmin<SIGNED>(w45, w67);
        w = pmax<SIGNED>(wa, wb);
        l = pmax3<SIGNED>(pmax3<SIGNED>(l01, l23, l45), pmax3<SIGNED>(l67, la, lb), pmin<SIGNED>(wa, wb));
    } else {
        uint32_t const w01 = pmax<SIGNED>(m[0], m[1]), l01 = pmin<SIGNED>(m[0], m[1]);
        uint32_t const w23 = pmax<SIGNED>(m[2], m[3]), l23 = pmin<SIGNED>(m[2], m[3]);
        w = pmax<SIGNED>(w01, w23);
        l = pmax3<SIGNED>(l01, l23, pmin<SIGNED>(w01, w23));
    }
    int const wl = plo<SIGNED>(w), wh = phi<SIGNED>(w);
    v1 = max(wl, wh);
    v2 = max3(min(wl, wh), plo<SIGNED>(l), phi<SIGNED>(l));
}

// MODE 0: normal.  1: epilogue only hands the accumulator back (MMA/TMA ceiling).
// 2: epilogue reads TMEM (packed) but reduces nothing (TMEM-read ceiling).  3: dump the raw
// 32-bit similarity tile to `dump` (row-major, leading dimension dump_ld).  4: dump the packed
// registers as the filter sees them (dump_ld/2 words per row).  5: normal epilogue, and CTA 0
// records clock64() time stamps of its pipeline events in `dump` (kTraceEvents x 4 int64 per
// warp).  Modes 1-5 produce no results.
template <int MODE, bool EXACT, bool SIGNED>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
            const ScanJob* __restrict__ jobs, int total_items_host, uint32_t idesc, int ksteps,
            int32_t* __restrict__ dump, int64_t dump_ld, ExactParams ex, int2* __restrict__ rowres)
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
            mbar_init(a_empty(i), 2);                 // one arrive per MMA issuer
            mbar_init(acc_full(i), 1);
            mbar_init(acc_empty(i), kAccEmptyCount);  // one arrive per participating warp
        }
        for (int i = 0; i < kStages; ++i) {
            mbar_init(b_full(i), 1);
            mbar_init(b_empty(i), 2);
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
                    long long const tp0 = MODE == 5 ? clock64_() : 0;
                    mbar_wait(b_empty(s), ((bcnt / kStages) & 1) ^ 1, kWaitBEmpty, bcnt);
                    if (MODE == 5 && blockIdx.x == 0 && bcnt < kTraceEvents) {
                        long long* tr = reinterpret_cast<long long*>(dump) + (static_cast<size_t>(kProducerWarp) * kTraceEvents + bcnt) * 4;
                        tr[0] = tp0; tr[1] = clock64_(); tr[2] = bcnt; tr[3] = 0;
                    }
                    mbar_arrive_expect_tx(b_full(s), kBTileBytes);
                    uint32_t const dst = smem_base + kSmemB + s * kBTileBytes;
                    int const row = job.c_row + t * kBlockN;
                    tma_load_2d(dst, &tmap_c, b_full(s), 0, row);
                    tma_load_2d(dst + kBTileBytes / 2, &tmap_c, b_full(s), 0, row + kBlockN / 2);
                }
            }
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
        // ===================== MMA issuers (one thread per query half) =====================
        // tcgen05.mma blocks its issuing thread while the tensor pipe's queue is full, and that
        // queue is short: with a single issuer, everything the thread does between two groups
        // (commits, barrier waits, the trip around the loop) shows up as tensor idle time.  Two
        // issuers, one per accumulator stage, keep the queue fed from the other thread meanwhile.
        int const h = warp - kMmaWarp;
        if (lane == 0) {
            int j = 0;
            uint32_t bcnt = 0, ic = 0, hc = 0;   // hc: uses of this issuer's accumulator stage so far
            for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
                while (it >= jobs[j + 1].item_start) ++j;
                int const c_n = jobs[j].c_n;
                int const rb = it - jobs[j].item_start;
                bool const active = (jobs[j].q_n - rb * kItemM > kHalfM) || h == 0;
                int const abuf = ic & 1;
                mbar_wait(a_full(abuf), (ic >> 1) & 1, kWaitAFull, ic);
                uint64_t const adesc = make_smem_desc_sw128(smem_base + kSmemA + abuf * kATileBytes + h * kAHalfBytes);
                uint32_t const d_tmem = tmem_base + h * kBlockN;
                int const ntiles = (c_n + kBlockN - 1) / kBlockN;
                for (int t = 0; t < ntiles; ++t, ++bcnt) {
                    int const s = bcnt % kStages;
                    long long const tb0 = MODE == 5 ? clock64_() : 0;
                    mbar_wait(b_full(s), (bcnt / kStages) & 1, kWaitBFull, bcnt);
                    if (!active) {                  // a 128-row item: nothing for the second half
                        mbar_arrive(b_empty(s));
                        continue;
                    }
                    uint64_t const bdesc = make_smem_desc_sw128(smem_base + kSmemB + s * kBTileBytes);
                    long long const tw0 = MODE == 5 ? clock64_() : 0;
                    mbar_wait(acc_empty(h), (hc & 1) ^ 1, kWaitAccEmpty, hc);
                    long long const tw1 = MODE == 5 ? clock64_() : 0;
                    ++hc;
                    tc_fence_after_sync();
#pragma unroll
                    for (int k = 0; k < kRowBytes / 32; ++k) {
                        // +2 in the start-address field = 32 bytes along K inside the swizzle span;
                        // 64-byte descriptors (SURF) are zero beyond K = 64: two steps suffice
                        if (k < ksteps)
                            mma_i8_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
                    }
                    mma_commit(acc_full(h));    // accumulator of this half is ready
                    mma_commit(b_empty(s));     // this issuer is done with the candidate stage
                    if (MODE == 5 && blockIdx.x == 0 && hc - 1 < kTraceEvents) {
                        long long* tr = reinterpret_cast<long long*>(dump) + (static_cast<size_t>(warp) * kTraceEvents + (hc - 1)) * 4;
                        tr[0] = tw0; tr[1] = tw1; tr[2] = clock64_(); tr[3] = tb0;
                    }
                }
                // this issuer is done with the query tile
                if (active) mma_commit(a_empty(abuf)); else mbar_arrive(a_empty(abuf));
            }
        }
    } else if (!EXACT) {
        // ===================== filter epilogue: 16 warps, 64 columns each =====================
        int const quad = warp & 3;           // TMEM lane quadrant this warp may access
        int const cg = warp >> 2;            // column group: columns [64 cg, 64 cg + 64)
        int const row = quad * 32 + lane;    // row inside a 128-row half
        uint32_t const taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + cg * kColsPerWarp;
        int2* const merge_base = reinterpret_cast<int2*>(smem_gen + kSmemMerge);

        int j = 0;
        uint32_t hcnt[2] = {0, 0}, ic = 0;
        for (int it = blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
            while (it >= jobs[j + 1].item_start) ++j;
            ScanJob const job = jobs[j];
            int const rb = it - job.item_start;
            int const nh = (job.q_n - rb * kItemM > kHalfM) ? 2 : 1;
            int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;

            // 0 is the reference's initial best / second best (nearest_neighbor.cc:221-224)
            uint32_t slot[2][kSlotRegs];
#pragma unroll
            for (int k = 0; k < kSlotRegs; ++k) { slot[0][k] = 0u; slot[1][k] = 0u; }
            // Software pipeline over (tile, half) stages: the accumulator of a stage is pulled
            // into registers with one packed x32 load (64 columns); while that load is in flight
            // the warp takes the maxima of the previous stage's registers, so it never sits
            // between "stage ready" and "stage handed back" with arithmetic to do.  Two
            // register buffers, indexed by the half (static after unrolling).
            uint32_t rr[2][32];
            bool pending = false;       // the previous stage's registers still await their fold
            for (int t = 0; t < ntiles; ++t) {
                int const ncols = job.c_n - t * kBlockN;   // valid columns of this tile
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h >= nh) break;
                    long long const te0 = MODE == 5 ? clock64_() : 0;
                    mbar_wait(acc_full(h), hcnt[h] & 1, kWaitAccFull, hcnt[h]);
                    long long const te1 = MODE == 5 ? clock64_() : 0;
                    ++hcnt[h];
                    tc_fence_after_sync();
                    if (MODE == 3 || MODE == 4) {
                        // debug dumps: no reduction, plain loads
                        int64_t const qr = static_cast<int64_t>(rb) * kItemM + h * kHalfM + row;
                        if (MODE == 3) {
                            int32_t va[32], vb[32];
                            uint32_t const taddr = taddr0 + h * kBlockN;
                            tmem_ld_32x32b_x32(taddr, va);
                            tmem_ld_32x32b_x32(taddr + kChunk, vb);
                            tmem_ld_wait_regs(va);
                            tmem_ld_wait_regs(vb);
                            if (qr < job.q_n) {
                                int32_t* d = dump + qr * dump_ld + t * kBlockN + cg * kColsPerWarp;
#pragma unroll
                                for (int q = 0; q < 32; ++q) { d[q] = va[q]; d[32 + q] = vb[q]; }
                            }
                        } else {
                            tmem_ld_32x32b_x32_pack16(taddr0 + h * kBlockN, rr[0]);
                            tmem_ld_wait_regs(rr[0]);
                            if (qr < job.q_n) {
                                uint32_t* d = reinterpret_cast<uint32_t*>(dump) + qr * (dump_ld / 2) +
                                              t * (kBlockN / 2) + cg * (kColsPerWarp / 2);
#pragma unroll
                                for (int q = 0; q < 32; ++q) d[q] = rr[0][q];
                            }
                        }
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(h));
                        continue;
                    }
                    if (MODE == 1) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(h));
                        continue;
                    }
                    constexpr bool kReduce = MODE == 0 || MODE == 5;
                    if (nh == 2) {
                        tmem_ld_32x32b_x32_pack16(taddr0 + h * kBlockN, rr[h]);
                        if (kReduce && pending) fold_packed<SIGNED>(slot[1 - h], rr[1 - h]);
                        tmem_ld_wait_regs(rr[h]);
                    } else {
                        // a single half: nothing to overlap with
                        if (kReduce && pending) fold_packed<SIGNED>(slot[0], rr[0]);
                        tmem_ld_32x32b_x32_pack16(taddr0, rr[0]);
                        tmem_ld_wait_regs(rr[0]);
                    }
                    // the data is in registers: hand the TMEM stage back
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(h));
                    if (kReduce) {
                        if (ncols < kBlockN) mask_packed<SIGNED>(rr[h], cg * kColsPerWarp, ncols);
                        pending = true;
                    } else {
                        slot[h][0] |= rr[h][0] | rr[h][31];
                    }
                    if (MODE == 5 && blockIdx.x == 0 && lane == 0) {
                        uint32_t const e = hcnt[0] + hcnt[1] - 1;
                        if (e < kTraceEvents) {
                            long long* tr = reinterpret_cast<long long*>(dump) + (static_cast<size_t>(warp) * kTraceEvents + e) * 4;
                            tr[0] = te0; tr[1] = te1; tr[2] = clock64_(); tr[3] = h;
                        }
                    }
                }
            }
            if ((MODE == 0 || MODE == 5) && pending) {
                if (nh == 2) fold_packed<SIGNED>(slot[1], rr[1]);
                else         fold_packed<SIGNED>(slot[0], rr[0]);
            }
            if (MODE != 0 && MODE != 5) {
                if (MODE == 2 && slot[0][0] == 0x12345678u && slot[1][0] == 0x9abcdef0u) dump[0] = 1;  // keep the loads alive
                continue;
            }

            // The row's 64 slots: 16 here, the other column groups' through shared memory.  The
            // merge area is double-buffered across items, so one barrier per item suffices: a
            // buffer is rewritten two items later, i.e. after the next item's barrier, which the
            // reading warps only reach once they are done with it.
            int2* const merge = merge_base + (ic & 1) * (kMergeBufBytes / 8);
            int v1[2], v2[2];
            for (int h = 0; h < nh; ++h) slots_top2<SIGNED>(slot[h], v1[h], v2[h]);
            if (cg != 0) {
                for (int h = 0; h < nh; ++h)
                    merge[(cg - 1) * kItemM + h * kHalfM + row] = make_int2(v1[h], v2[h]);
            }
            named_barrier_sync(1, kEpilogueWarps * 32);
            if (cg == 0) {
                for (int h = 0; h < nh; ++h) {
#pragma unroll
                    for (int g = 0; g < kColGroups - 1; ++g) {
                        int2 const o = merge[g * kItemM + h * kHalfM + row];
                        v2[h] = max3(min(v1[h], o.x), v2[h], o.y);
                        v1[h] = max(v1[h], o.x);
                    }
                    int const r_in_job = rb * kItemM + h * kHalfM + row;
                    if (r_in_job < job.q_n) rowres[job.out_row + r_in_job] = pack_rowres(v1[h], v2[h], j);
                }
            }
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
            // Unsigned: a candidate >= 2^16 ("big") enters with the value the tensor core
            // computed.  That equals the reference's lane-wise 16-bit sum unless one of the eight
            // lanes itself reached 2^16, which needs an adversarial descriptor; every big
            // candidate is therefore recorded and verify_big_kernel re-checks it with the lane
            // emulation afterwards.  Rows that fail the check (or have more than kMaxBigPerRow big
            // candidates) are replayed on CUDA cores by slow_rows_kernel, which overwrites the
            // result written here.  Signed rows only get here with the norm certificate: neither
            // a lane nor a 16-bit store can wrap.
            int b1 = 0, b2 = 0, i1 = 0, nbig = 0;
            int64_t const entry = live ? ex.xrow_map[job.out_row + r_in_job] : 0;
            int64_t const g = surv_row(entry);
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
                        if (!SIGNED && __any_sync(0xffffffffu, trig && cmax >= 65536)) {
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
                            // branch-free (values inside the 16-bit range are stored untruncated)
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
                bool const ok = passes_tests(ip_to_dist<SIGNED>(b1), ip_to_dist<SIGNED>(b2), ex.sq_lowe, ex.sq_dist);
                ex.oneway[g] = ok ? i1 : -1;
                if (nbig > kMaxBigPerRow)   // cannot be certified by verify_big_kernel
                    ex.replay_list[atomicAdd(ex.replay_count, 1ull)] = g;
                // a certified row has no big candidate: the filter's best is the true best
                if ((static_cast<uint64_t>(entry) & kSurvCertified) != 0 &&
                    (b1 & 0xffff) != static_cast<int>((static_cast<uint64_t>(entry) >> kSurvRowBits) & 0xffffu))
                    atomicAdd(ex.self_check, 1ull);
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
