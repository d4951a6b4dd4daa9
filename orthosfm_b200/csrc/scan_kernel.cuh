// scan_kernel.cuh -- the hot kernel: all-pairs 8-bit descriptor similarity on tcgen05
// tensor cores with fused top-2 epilogues.  sm_100a only.
//
// What it replaces (reference, paths relative to /root/reference):
//   short_inner_prod<T> + the top-2 scan of NearestNeighbor<T>::find
//   (src/mve/sfm/nearest_neighbor.cc:62-129, 216-268), called once per query by
//   Matching::oneway_match<T> (src/mve/sfm/matching.h:114-146).
//
// A *job* is one direction of one image pair: every descriptor of a query set against
// every descriptor of a candidate view.  A *work item* is a 256-row block of a job's
// queries; item_job[] names every item's job (a walk along the jobs' item prefix sums costs a
// dependent load per job passed, which for the short jobs of the second passes -- a handful of
// items each, 148 items between two items of a CTA -- came to more than the item's own work).  For each item a persistent CTA
//   - TMA-loads the 256 x 128 B query tile once (two 128-row halves, SWIZZLE_128B, K-major),
//   - streams the candidate view through a ring of 256 x 128 B tiles; each candidate tile
//     feeds two tcgen05.mma groups (one per query half), which halves the L2 -> SMEM
//     traffic per similarity,
//   - issues tcgen05.mma kind::i8 (M=128, N=256, 4 x K=32; u8 x u8 or s8 x s8 -> s32,
//     exact) into the TMEM accumulator of that half (2 x 256 columns = all of TMEM),
//   - 16 epilogue warps, one per (TMEM lane quadrant, query half, 128-column half), pull
//     their part of that half's accumulator into registers with tcgen05.ld, hand the
//     accumulator straight back to the MMA issuers, and then reduce from registers.  The
//     similarity matrix never leaves the SM.
// Warps 0 and 1 issue the MMAs (one per query half, taking turns), warp 2 drives TMA.
//
// A launch covers the work items [item_first, total_items) of the job list (the filter pass of a
// batch may be cut into several launches that start as the views they need arrive).
//
// Three passes share this pipeline and differ in the epilogue (template parameter PASS):
//
// FILTER (all rows).  Reads the accumulator with tcgen05.ld ... .pack::16b (two adjacent
// columns per 32-bit register) and folds it with three-input 16-bit SIMD max instructions
// (VIMNMX3.U16x2 / .S16x2, one instruction per four similarities) into one packed register
// per thread: the running maxima of the even and of the odd columns of the warp's column
// half.  At the end of a work item the four "slot" maxima of a row give
//   v1  = the largest similarity of the row (exact), and
//   v2  = the second largest slot maximum, clamped below at 0 -- a lower bound on the
//         reference's second-best inner product (exact unless best and second best share a
//         slot).
// The reference's ratio test is monotone in the second best, so a row that fails it with the
// lower bound fails it for good (classify_kernel writes -1).  The rows that pass (the
// *survivors*, typically the rows that really have a match) go to the RESOLVE pass.
//
// The 16-bit packing is only valid if no similarity of the row leaves the 16-bit range.  That
// is certified per (row, candidate view) by Cauchy-Schwarz from the squared norms computed at
// commit: |a|^2 * max|b|^2 < 2^32 (unsigned) or < 2^30 (signed; this also excludes a wrap of
// the reference's 16-bit SSE lanes).  A few rows without that certificate are looked at again
// by certify_kernel; a warp with many of them reads the item with 32-bit loads instead.
//
// RESOLVE (the gathered survivors).  Same packed loads; the row's best value V is known, and
// the epilogue finds the last column equal to V, how many there are, and the largest value
// below V -- which is all the reference's sequential scan ends with.
//
// EXACT (the gathered rows that really reach 2^16, unsigned).  32-bit loads; eight epilogue
// warps replay the reference's sequential best / second-best scan per row, in column order
// (nearest_neighbor.cc:87-100), including the 16-bit wrap of its lanes and stores.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "ptx.cuh"

namespace osfm {

constexpr int kTmemColsTotal = 512;
constexpr int kStages = 4;            // candidate-tile ring depth
constexpr int kAccStages = 2;         // TMEM accumulators: one per query half, kBlockN columns each
constexpr int kAccCols = kBlockN / 2; // columns one filter warp drains: half an accumulator
constexpr int kEpilogueWarps = 4 * 2 * kAccStages;       // one warp per (lane quadrant, half, column half)
#ifndef OSFM_SLOT_REGS
#define OSFM_SLOT_REGS 1
#endif
// Packed running maxima per thread.  ONE register (two slots: the even and the odd columns)
// on purpose: its updates form a single dependent chain, so the warp stalls on its own result
// after every instruction and the scheduler turns to other warps.  With several independent
// chains the fold issues back to back, and the greedy scheduler then keeps the MMA issuer
// warp that shares the scheduler from feeding the tensor pipe for the length of the burst
// (measured: 8 chains 11.4 M cycles per launch, 2 chains 10.4 M, no fold at all 9.6 M).
constexpr int kSlotRegs = OSFM_SLOT_REGS;
// Warp roles.  The control warps have the LOWEST ids: the warp scheduler favours the oldest
// ready warp once the one it is issuing from stalls, and a burst of epilogue arithmetic on an
// MMA issuer's scheduler otherwise keeps the issuer (were it the youngest warp) from feeding
// the tensor pipe for the length of the burst -- measured: the epilogue's ALU time added almost
// one-to-one to the tile time.
constexpr int kMmaWarp = 0;                // issuer of query half 0; half 1: the next warp
constexpr int kProducerWarp = 2;           // TMA
constexpr int kFirstEpilogueWarp = 4;      // (warp 3 idles) a multiple of 4: TMEM lane quadrant = warp % 4
constexpr int kScanThreads = (kFirstEpilogueWarp + kEpilogueWarps) * 32;   // 640
constexpr int kTmemCols = kTmemColsTotal;

constexpr int kAHalfBytes = kHalfM * kRowBytes;    // 16 KB
constexpr int kATileBytes = kItemM * kRowBytes;    // 32 KB
constexpr int kBTileBytes = kBlockN * kRowBytes;   // 32 KB
constexpr int kSmemA = 0;
constexpr int kSmemB = 2 * kATileBytes;
constexpr int kSmemBar = kSmemB + kStages * kBTileBytes;
constexpr int kNumBars = 2 + 2 + 2 * kStages + 2 * kAccStages + 2;
constexpr int kSmemTmemPtr = kSmemBar + kNumBars * 8;
constexpr int kSmemMerge = (kSmemTmemPtr + 4 + 15) & ~15;
constexpr int kMergeBufBytes = kItemM * 16;  // per row: what the upper column half's warp found
constexpr int kSmemTotal = kSmemMerge + 2 * kMergeBufBytes;   // double-buffered across items
constexpr int kScanSmemBytes = kSmemTotal + 1024;  // slack for manual 1024-byte alignment

// Hang-report codes (see ptx.cuh).
enum : uint32_t {
    kWaitAEmpty = 1, kWaitBEmpty = 2, kWaitAFull = 3, kWaitAccEmpty = 4,
    kWaitBFull = 5, kWaitAccFull = 6, kWaitTurn = 7
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
// The job index shares its word with a flag: how the row was scanned.
constexpr int kRowFlagShift = 28;
constexpr int kRowJobMask = (1 << kRowFlagShift) - 1;
constexpr int kRowPacked = 0;     // 16-bit packed loads: valid if the row is certified
constexpr int kRowWideOk = 1;     // 32-bit loads, best below 2^16: every similarity fits 16 bits
constexpr int kRowWideWraps = 2;  // 32-bit loads, best reaches 2^16: the reference's arithmetic wraps
__device__ __forceinline__ int2 pack_rowres(int v1, int v2, int job, int flag = kRowPacked) {
    return make_int2(static_cast<int>((static_cast<uint32_t>(v1) & 0xffffu) | (static_cast<uint32_t>(v2) << 16)),
                     job | (flag << kRowFlagShift));
}
// A warp of the filter takes the 32-bit route for an item if at least this many of its 32 rows
// lack the norm certificate (unit-norm descriptors: practically never; raw bytes: always).
constexpr int kWideThreshold = 8;

// Extra arguments of the EXACT pass.
struct ExactParams {
    const uint8_t* qpool;        // gathered query rows (what tmap_q describes)
    const uint8_t* cpool;        // candidate pool (what tmap_c describes)
    const int64_t* xrow_map;     // gathered row -> survivor-list entry (surv_entry)
    int32_t* oneway;
    const int* total_items_dev;  // number of work items, computed on the device
    float sq_lowe, sq_dist;
    int64_t* replay_list;        // rows that need the full replay on CUDA cores (each row once:
    unsigned long long* replay_count;   // replay_flags is a bitmap over the batch's rows)
    uint32_t* replay_flags;
    int4* big_list;              // (row lo, row hi, column, similarity) of every big candidate met
    unsigned long long* big_count;
    unsigned long long* self_check;   // filter and EXACT pass disagree on a certified row's best
    // filter pass only: the norm certificate's inputs (squared norm per pool row, largest per view)
    const int32_t* norm2;
    const int32_t* viewmax;
    // RESOLVE pass only.  0: the rows are the filter's survivors, V is the row's largest
    // similarity (anything above it is a self-check failure).  1: the rows are the *claimed*
    // rows of the reverse direction of a pair (see claim_kernel in post_kernels.cuh), V is the
    // largest similarity any claimant has with the row; a similarity above V means that the
    // row's nearest neighbour is not one of its claimants, and the row's result is -1 as far as
    // the mutual filter is concerned.
    int verify;
    // RESOLVE pass over a restricted candidate set (jobs with c_view < 0 read their candidates from
    // the second candidate tensor map, a gathered subset of the view): column -> row of the view
    const int32_t* col_map;
    // RESOLVE pass: 8 x uint4 per thread of the grid, for the one packed load per row and item
    // that holds the row's best value (ResolveStash)
    uint4* stash;
};

// Sets bit g of a bitmap; true for the caller that set it.  Keeps a row from entering the replay
// list more than once (the list is sized for one entry per row).
__device__ __forceinline__ bool mark_once(uint32_t* flags, int64_t g) {
    uint32_t const bit = 1u << (g & 31);
    return (atomicOr(flags + (g >> 5), bit) & bit) == 0u;
}

constexpr int kTraceEvents = 256;   // per warp, MODE 5

__device__ __forceinline__ long long clock64_() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}

constexpr int kPassFilter = 0, kPassExact = 1, kPassResolve = 2;

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

// Both loads of a tile at once, pairing a register of the first with one of the second.  The
// very first instruction then needs both loads, which keeps ptxas from sinking the second
// tcgen05.ld below the fold of the first (it does, to save registers, and the two load
// latencies then add up on the path that hands the accumulator back).
template <bool SIGNED>
__device__ __forceinline__ void fold_packed2(uint32_t (&slot)[kSlotRegs], const uint32_t (&a)[32], const uint32_t (&b)[32]) {
#pragma unroll
    for (int k = 0; k < kSlotRegs; ++k) {
#pragma unroll
        for (int q = 0; q < 32 / kSlotRegs; ++q)
            slot[k] = pmax3<SIGNED>(slot[k], a[k + q * kSlotRegs], b[k + q * kSlotRegs]);
    }
}

// Largest and second largest (with multiplicity) of the 16 slot maxima held in eight packed
// registers: a tournament per 16-bit lane -- the maximum over all "losers" of a tournament is
// its second largest entry -- followed by the merge of the two lanes.
template <bool SIGNED>
__device__ __forceinline__ void slots_top2(const uint32_t (&m)[kSlotRegs], int& v1, int& v2) {
    static_assert(kSlotRegs == 1 || kSlotRegs == 2 || kSlotRegs == 4 || kSlotRegs == 8, "tournament sizes");
    uint32_t w, l;
    if (kSlotRegs == 8) {
        uint32_t const w01 = pmax<SIGNED>(m[0], m[1 % kSlotRegs]), l01 = pmin<SIGNED>(m[0], m[1 % kSlotRegs]);
        uint32_t const w23 = pmax<SIGNED>(m[2 % kSlotRegs], m[3 % kSlotRegs]), l23 = pmin<SIGNED>(m[2 % kSlotRegs], m[3 % kSlotRegs]);
        uint32_t const w45 = pmax<SIGNED>(m[4 % kSlotRegs], m[5 % kSlotRegs]), l45 = pmin<SIGNED>(m[4 % kSlotRegs], m[5 % kSlotRegs]);
        uint32_t const w67 = pmax<SIGNED>(m[6 % kSlotRegs], m[7 % kSlotRegs]), l67 = pmin<SIGNED>(m[6 % kSlotRegs], m[7 % kSlotRegs]);
        uint32_t const wa = pmax<SIGNED>(w01, w23), la = pmin<SIGNED>(w01, w23);
        uint32_t const wb = pmax<SIGNED>(w45, w67), lb = pmin<SIGNED>(w45, w67);
        w = pmax<SIGNED>(wa, wb);
        l = pmax3<SIGNED>(pmax3<SIGNED>(l01, l23, l45), pmax3<SIGNED>(l67, la, lb), pmin<SIGNED>(wa, wb));
    } else if (kSlotRegs == 4) {
        uint32_t const w01 = pmax<SIGNED>(m[0], m[1 % kSlotRegs]), l01 = pmin<SIGNED>(m[0], m[1 % kSlotRegs]);
        uint32_t const w23 = pmax<SIGNED>(m[2 % kSlotRegs], m[3 % kSlotRegs]), l23 = pmin<SIGNED>(m[2 % kSlotRegs], m[3 % kSlotRegs]);
        w = pmax<SIGNED>(w01, w23);
        l = pmax3<SIGNED>(l01, l23, pmin<SIGNED>(w01, w23));
    } else if (kSlotRegs == 2) {
        w = pmax<SIGNED>(m[0], m[1 % kSlotRegs]);
        l = pmin<SIGNED>(m[0], m[1 % kSlotRegs]);
    } else {
        w = m[0];
        l = 0u;      // the reference's initial second best; never above a real lower bound
    }
    int const wl = plo<SIGNED>(w), wh = phi<SIGNED>(w);
    v1 = max(wl, wh);
    v2 = max3(min(wl, wh), plo<SIGNED>(l), phi<SIGNED>(l));
}

#ifndef OSFM_RESOLVE_CHAINS
#define OSFM_RESOLVE_CHAINS 2
#endif
constexpr int kResolveChains = OSFM_RESOLVE_CHAINS;   // dependent chains of the RESOLVE load maximum

// RESOLVE: a packed load (64 columns, 32 registers) that contains the row's best value V is set
// aside whole, in this thread's slot of a global scratch buffer, to be looked at value by value
// at the end of the work item.  Setting it aside is eight 16-byte stores: the thread is back at
// its accumulator in time (finding and keeping the 16-column group in registers, as round 1 did,
// was ~100 dependent instructions in a divergent region, on 40 % of the tile visits -- more than
// the slack a tile leaves, so one late warp of the eight stalled the accumulator's hand-back).
struct ResolveStash {
    uint4* slot;        // element q of this thread's slot is slot[q * kScanThreads]
    int col;            // first column of the load in the candidate view; -1: nothing set aside
};

// Looks at eight packed registers (16 columns from column `col`) value by value: counts the
// columns equal to V, remembers the last one and takes the maximum of the others.
// (col < c_n: a masked column of an all-zero row is no column.)
template <bool SIGNED>
__device__ __forceinline__ void resolve_scan16(const uint32_t (&r)[8], int col, int c_n, int V, int& cnt, int& idx, int& v2) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int const c0 = col + 2 * k;
        int const x0 = plo<SIGNED>(r[k]), x1 = phi<SIGNED>(r[k]);
        bool const e0 = x0 == V && c0 < c_n;
        bool const e1 = x1 == V && c0 + 1 < c_n;
        cnt += (e0 ? 1 : 0) + (e1 ? 1 : 0);
        idx = e1 ? c0 + 1 : (e0 ? c0 : idx);
        v2 = max(v2, e0 ? 0 : x0);
        v2 = max(v2, e1 ? 0 : x1);
    }
}

// Looks at the load set aside.  Normally exactly one of its four groups of 16 columns holds V:
// that group is picked with selects (every lane of the warp has its own, so no branch per group)
// and scanned; the other groups only feed v2.  A lane with V in several groups (duplicates) scans
// all four.
template <bool SIGNED>
__device__ __forceinline__ void resolve_flush(ResolveStash& st, int c_n, int V, int& cnt, int& idx, int& v2) {
    if (st.col < 0) return;
    uint32_t r[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint4 const w = st.slot[q * kScanThreads];
        r[4 * q] = w.x; r[4 * q + 1] = w.y; r[4 * q + 2] = w.z; r[4 * q + 3] = w.w;
    }
    int mg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t const a = pmax3<SIGNED>(r[8 * i], r[8 * i + 1], r[8 * i + 2]);
        uint32_t const b = pmax3<SIGNED>(r[8 * i + 3], r[8 * i + 4], r[8 * i + 5]);
        uint32_t const g = pmax<SIGNED>(pmax3<SIGNED>(a, b, r[8 * i + 6]), r[8 * i + 7]);
        mg[i] = max(plo<SIGNED>(g), phi<SIGNED>(g));
    }
    int const hits = (mg[0] >= V ? 1 : 0) + (mg[1] >= V ? 1 : 0) + (mg[2] >= V ? 1 : 0) + (mg[3] >= V ? 1 : 0);
    if (hits == 1) {
        int const gi = mg[1] >= V ? 1 : (mg[2] >= V ? 2 : (mg[3] >= V ? 3 : 0));
        uint32_t sel[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t const lo = (gi & 1) ? r[8 + k] : r[k];
            uint32_t const hi = (gi & 1) ? r[24 + k] : r[16 + k];
            sel[k] = (gi & 2) ? hi : lo;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) v2 = max(v2, i == gi ? 0 : mg[i]);
        resolve_scan16<SIGNED>(sel, st.col + 16 * gi, c_n, V, cnt, idx, v2);
    } else {
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            uint32_t grp[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t const lo = (i & 1) ? r[8 + k] : r[k];
                uint32_t const hi = (i & 1) ? r[24 + k] : r[16 + k];
                grp[k] = (i & 2) ? hi : lo;
            }
            resolve_scan16<SIGNED>(grp, st.col + 16 * i, c_n, V, cnt, idx, v2);
        }
    }
    st.col = -1;
}

// RESOLVE: one packed load (64 columns starting at column col0 of the candidate view) of a row
// whose largest similarity V is known.  A load that cannot contain V only feeds v2; one that does
// is set aside (columns are visited in ascending order).
template <bool SIGNED>
__device__ __forceinline__ void resolve_load(const uint32_t (&r)[32], int col0, int& V, bool& beaten,
                                             ResolveStash& st, bool& dup, int& v2)
{
    // the load's maximum as ONE dependent chain (like the filter's fold: a warp that stalls on
    // its own result leaves issue slots to the MMA issuers; a tree would issue back to back)
    uint32_t acc[kResolveChains];
#pragma unroll
    for (int q = 0; q < kResolveChains; ++q) acc[q] = r[q];
#pragma unroll
    for (int k = kResolveChains; k + 2 * kResolveChains <= 32; k += 2 * kResolveChains)
#pragma unroll
        for (int q = 0; q < kResolveChains; ++q)
            acc[q] = pmax3<SIGNED>(acc[q], r[k + 2 * q], r[k + 2 * q + 1]);
    // (32 - kResolveChains) is not a multiple of 2 * kResolveChains for every choice: the rest
#pragma unroll
    for (int k = kResolveChains + ((32 - kResolveChains) / (2 * kResolveChains)) * 2 * kResolveChains; k < 32; ++k)
        acc[0] = pmax<SIGNED>(acc[0], r[k]);
#pragma unroll
    for (int q = 1; q < kResolveChains; ++q) acc[0] = pmax<SIGNED>(acc[0], acc[q]);
    int const m = max(plo<SIGNED>(acc[0]), phi<SIGNED>(acc[0]));
    if (m < V) {            // the common case: nothing of interest in these 64 columns
        v2 = max(v2, m);
        return;
    }
    if (m > V) {            // reverse pass: a row that is no claimant is nearer than every claimant
        beaten = true;
        V = 0x7fffffff;     // nothing reaches this: the row stays on the fast path from here on
        return;
    }
    // This load holds the row's best value (once per row and item, unless it has duplicates).  An
    // earlier load still set aside then also holds V: the row has at least two columns equal to
    // V -- all that the result needs of the earlier one (the second best is V itself, the index
    // is the LAST column equal to V, which lies in this load or a later one) -- and it is dropped.
    dup = dup || st.col >= 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        st.slot[q * kScanThreads] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    st.col = col0;
}

// MODE 0: normal.  1: epilogue only hands the accumulator back (MMA/TMA ceiling).
// 2: epilogue reads TMEM (packed) but reduces nothing (TMEM-read ceiling).  3: dump the raw
// 32-bit similarity tile to `dump` (row-major, leading dimension dump_ld).  4: dump the packed
// registers as the filter sees them (dump_ld/2 words per row).  5: normal epilogue, and CTA 0
// records clock64() time stamps of its pipeline events in `dump` (kTraceEvents x 4 int64 per
// warp).  Modes 1-5 produce no results.
template <int MODE, int PASS, bool SIGNED>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
            const __grid_constant__ CUtensorMap tmap_c2,
            const ScanJob* __restrict__ jobs, const int32_t* __restrict__ item_job, int item_first, int total_items_host, uint32_t idesc, int ksteps,
            int32_t* __restrict__ dump, int64_t dump_ld, ExactParams ex, int2* __restrict__ rowres,
            unsigned long long* __restrict__ prof)
{
    // CTA 0 reports how many SM cycles and how many nanoseconds the kernel took: their ratio is
    // the SM clock the kernel actually ran at (it drops under sustained tensor load).
    long long prof_c0 = 0;
    uint64_t prof_t0 = 0;
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { prof_c0 = clock64_(); prof_t0 = globaltimer_ns(); }

    extern __shared__ uint8_t smem_raw[];
    uint32_t const smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    uint32_t const bar_base = smem_base + kSmemBar;
    auto a_full = [&](int i) { return bar_base + 8u * (0 + i); };
    auto a_empty = [&](int i) { return bar_base + 8u * (2 + i); };
    auto b_full = [&](int i) { return bar_base + 8u * (4 + i); };
    auto b_empty = [&](int i) { return bar_base + 8u * (4 + kStages + i); };
    auto acc_full = [&](int i) { return bar_base + 8u * (4 + 2 * kStages + i); };
    auto acc_empty = [&](int i) { return bar_base + 8u * (4 + 2 * kStages + kAccStages + i); };
    auto turn = [&](int i) { return bar_base + 8u * (4 + 2 * kStages + 2 * kAccStages + i); };

    int const warp = threadIdx.x >> 5;
    int const lane = threadIdx.x & 31;
    constexpr bool EXACT = PASS == kPassExact;
    constexpr bool RESOLVE = PASS == kPassResolve;
    int const total_items = PASS != kPassFilter ? *ex.total_items_dev : total_items_host;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full(i), 1);
            mbar_init(a_empty(i), 2);                 // one arrive per MMA issuer
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(turn(i), 1);
            mbar_init(acc_full(i), 1);
            mbar_init(acc_empty(i), EXACT ? 4 : 8);   // the warps that read it
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
        if (PASS == kPassResolve) prefetch_tensormap(&tmap_c2);
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
            for (int it = item_first + blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
                j = item_job[it];
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
                    // (RESOLVE: a job flagged by a negative c_view scans a gathered subset of its view)
                    const CUtensorMap* const tm = (PASS == kPassResolve && job.c_view < 0) ? &tmap_c2 : &tmap_c;
                    tma_load_2d(dst, tm, b_full(s), 0, row);
                    tma_load_2d(dst + kBTileBytes / 2, tm, b_full(s), 0, row + kBlockN / 2);
                }
            }
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
        // ===================== MMA issuers (one thread per query half) =====================
        // A candidate tile (256 rows) x a query half (128 rows) is one *group* of tcgen05.mma
        // (M = 128, N = 256, K = 32 each) into that half's TMEM accumulator.  (N = 128 groups
        // into four accumulators were tried: they re-read the A tile twice as often and run
        // into the shared-memory bandwidth, 8 KB per 64 cycles.)
        //
        // The tensor pipe's queue is short: a thread is held in tcgen05.mma until the previous
        // MMA has started, so it leaves the last MMA of a group one MMA time (128 cycles) before
        // the pipe runs dry, and its commits, barrier waits (70-90 cycles each even when
        // complete) and the trip around the loop do not fit into that.  With one issuer per
        // query half the other thread is already queueing its group meanwhile.  The two take
        // turns (turn[] barriers), which keeps the two accumulators in anti-phase -- one is being
        // computed while the other is being read; left alone the two streams interleave in the
        // queue, both accumulators finish together and the pipe idles while both are drained.
        // (Measured per launch of the 630-pair workload, epilogue switched off: one issuer
        // 9.7 M cycles, one issuer with the next group's waits hoisted 12.8 M, two issuers
        // free-running 12 M, two issuers taking turns 9.45 M; 8.9 M is the tensor pipe's floor.)
        // The whole warp runs this loop converged (see elect_one_sync in ptx.cuh); one elected
        // lane issues.
        int const h = warp_uniform(warp - kMmaWarp);
        uint32_t const tmem_u = warp_uniform(tmem_base);
        {
            int j = 0;
            uint32_t bcnt = 0, ic = 0, hc = 0;   // hc: tiles this issuer has issued so far
            for (int it = item_first + blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
                j = warp_uniform(item_job[it]);
                int const c_n = warp_uniform(jobs[j].c_n);
                int const rb = it - warp_uniform(jobs[j].item_start);
                bool const active = (warp_uniform(jobs[j].q_n) - rb * kItemM > kHalfM) || h == 0;
                int const abuf = ic & 1;
                mbar_wait(a_full(abuf), (ic >> 1) & 1, kWaitAFull, ic);
                uint64_t const adesc = make_smem_desc_sw128(smem_base + kSmemA + abuf * kATileBytes + h * kAHalfBytes);
                uint32_t const d_tmem = tmem_u + h * kBlockN;
                int const ntiles = (c_n + kBlockN - 1) / kBlockN;
                for (int t = 0; t < ntiles; ++t, ++bcnt) {
                    int const s = bcnt % kStages;
                    mbar_wait(b_full(s), (bcnt / kStages) & 1, kWaitBFull, bcnt);
                    uint32_t const turn_parity = h == 0 ? ((bcnt & 1) ^ 1) : (bcnt & 1);
                    if (!active) {                  // a 128-row item: nothing for the second half
                        mbar_wait(turn(h), turn_parity, kWaitTurn, bcnt);
                        if (elect_one_sync()) {
                            mbar_arrive(turn(1 - h));
                            mbar_arrive(b_empty(s));
                        }
                        __syncwarp();
                        continue;
                    }
                    uint64_t const bdesc = make_smem_desc_sw128(smem_base + kSmemB + s * kBTileBytes);
                    long long const tw0 = MODE == 5 ? clock64_() : 0;
                    mbar_wait(acc_empty(h), (hc & 1) ^ 1, kWaitAccEmpty, hc);
                    mbar_wait(turn(h), turn_parity, kWaitTurn, bcnt);
                    long long const tw1 = MODE == 5 ? clock64_() : 0;
                    tc_fence_after_sync();
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < kRowBytes / 32; ++k) {
                            // +2 in the start-address field = 32 bytes along K inside the swizzle
                            // span; 64-byte descriptors (SURF) are zero beyond K = 64: two steps
                            if (k < ksteps)
                                mma_i8_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
                        }
                        mbar_arrive(turn(1 - h));   // the other half's group may follow
                        mma_commit(acc_full(h));    // this accumulator is ready
                        mma_commit(b_empty(s));     // this issuer is done with the candidate stage
                        if (MODE == 5 && blockIdx.x == 0 && hc < kTraceEvents) {
                            long long* tr = reinterpret_cast<long long*>(dump) + (static_cast<size_t>(warp) * kTraceEvents + hc) * 4;
                            tr[0] = tw0; tr[1] = tw1; tr[2] = clock64_(); tr[3] = h;
                        }
                    }
                    __syncwarp();
                    ++hc;
                }
                // this issuer is done with the query tile
                if (elect_one_sync()) {
                    if (active) mma_commit(a_empty(abuf)); else mbar_arrive(a_empty(abuf));
                }
                __syncwarp();
            }
        }
    } else if (warp < kFirstEpilogueWarp) {
        // idle
    } else if (PASS == kPassFilter) {
        // ===================== filter epilogue: one warp per (lane quadrant, half, column half) =
        int const ew = warp - kFirstEpilogueWarp;
        int const quad = warp & 3;           // TMEM lane quadrant this warp may access
        int const h = ew >> 3;               // query half = the accumulator this warp drains
        int const c = (ew >> 2) & 1;         // column half: columns [128 c, 128 c + 128) of the tile
        int const row = quad * 32 + lane;    // row inside the 128-row half
        uint32_t const taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + h * kBlockN + c * kAccCols;
        int4* const merge_base = reinterpret_cast<int4*>(smem_gen + kSmemMerge);

        int j = 0;
        uint32_t cnt = 0, ic = 0;            // cnt: tiles of this warp's accumulator so far
        for (int it = item_first + blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
            j = item_job[it];
            ScanJob const job = jobs[j];
            int const rb = it - job.item_start;
            int const nh = (job.q_n - rb * kItemM > kHalfM) ? 2 : 1;
            if (h >= nh) continue;           // a 128-row item has no second half
            int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;
            int64_t const qr = static_cast<int64_t>(rb) * kItemM + h * kHalfM + row;   // row in the job
            int4* const merge = merge_base + (ic & 1) * (kMergeBufBytes / 16) + h * kHalfM + row;

            // Unsigned rows without the norm certificate may hold similarities beyond 16 bits.
            // A few per warp are sorted out afterwards (certify_kernel); if they are many (input
            // that is not unit-norm), this warp reads the item's accumulators as 32-bit values
            // instead: twice the loads and instructions, but exact for any input.  (Both column
            // halves see the same rows and decide alike.)  A loop of its own, so that the packed
            // loop below stays as tight as it is.
            if (!SIGNED && MODE == 0) {
                bool const doubtful = qr < job.q_n &&
                    static_cast<int64_t>(ex.norm2[job.q_row + qr]) * static_cast<int64_t>(ex.viewmax[job.c_view]) >= (1ll << 32);
                if (__popc(__ballot_sync(0xffffffffu, doubtful)) >= kWideThreshold) {
                    int w0 = 0, w1 = 0;          // running maxima of the even / odd columns
                    for (int t = 0; t < ntiles; ++t, ++cnt) {
                        int const ncols = job.c_n - t * kBlockN;
                        mbar_wait(acc_full(h), cnt & 1, kWaitAccFull, cnt);
                        tc_fence_after_sync();
                        // 128 columns as four 32-bit loads, two in flight at a time
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            int32_t va[32], vb[32];
                            tmem_ld_32x32b_x32(taddr + half * 64, va);
                            tmem_ld_32x32b_x32(taddr + half * 64 + 32, vb);
                            tmem_ld_wait_regs(va);
                            tmem_ld_wait_regs(vb);
                            if (half == 1) {
                                tc_fence_before_sync();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(acc_empty(h));
                            }
                            if (ncols < kBlockN) {
                                mask_chunk(va, c * kAccCols + half * 64, ncols);
                                mask_chunk(vb, c * kAccCols + half * 64 + 32, ncols);
                            }
#pragma unroll
                            for (int q = 0; q < 32; q += 4) {
                                w0 = max3(w0, va[q], va[q + 2]);
                                w1 = max3(w1, va[q + 1], va[q + 3]);
                            }
#pragma unroll
                            for (int q = 0; q < 32; q += 4) {
                                w0 = max3(w0, vb[q], vb[q + 2]);
                                w1 = max3(w1, vb[q + 1], vb[q + 3]);
                            }
                        }
                    }
                    int v1 = max(w0, w1), v2 = min(w0, w1);
                    if (c == 1) *merge = make_int4(v1, v2, 0, 0);
                    named_barrier_sync(1 + h * 4 + quad, 64);
                    if (c == 0) {
                        int4 const o = *merge;
                        v2 = max3(min(v1, o.x), v2, o.y);
                        v1 = max(v1, o.x);
                        if (qr < job.q_n)
                            rowres[job.out_row + qr] = pack_rowres(v1, v2, j, v1 < 65536 ? kRowWideOk : kRowWideWraps);
                    }
                    continue;
                }
            }

            // 0 is the reference's initial best / second best (nearest_neighbor.cc:221-224)
            uint32_t slot[kSlotRegs];
#pragma unroll
            for (int k = 0; k < kSlotRegs; ++k) slot[k] = 0u;
            // full tiles here; a ragged last tile (debug modes: every tile) after the loop / inside
            int const nloop = (MODE == 0 || MODE == 5) ? job.c_n / kBlockN : ntiles;
            for (int t = 0; t < nloop; ++t, ++cnt) {
                long long const te0 = MODE == 5 ? clock64_() : 0;
                mbar_wait(acc_full(h), cnt & 1, kWaitAccFull, cnt);
                long long const te1 = MODE == 5 ? clock64_() : 0;
                tc_fence_after_sync();
                if (MODE == 3) {
                    // debug: raw 32-bit similarities, 32 columns at a time
#pragma unroll 1
                    for (int q4 = 0; q4 < kAccCols / kChunk; ++q4) {
                        int32_t v[32];
                        tmem_ld_32x32b_x32(taddr + q4 * kChunk, v);
                        tmem_ld_wait_regs(v);
                        if (qr < job.q_n) {
                            int32_t* d = dump + qr * dump_ld + t * kBlockN + c * kAccCols + q4 * kChunk;
#pragma unroll
                            for (int q = 0; q < 32; ++q) d[q] = v[q];
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
                // the accumulator's 128 columns as two packed loads of 64, both in flight at once
                uint32_t ra[32], rc[32];
                tmem_ld_32x32b_x32_pack16(taddr, ra);
                tmem_ld_32x32b_x32_pack16(taddr + kAccCols / 2, rc);
                tmem_ld_wait_regs(ra);
                tmem_ld_wait_regs(rc);
                // the data is in registers: hand the accumulator back before reducing
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(h));
                long long const te2 = MODE == 5 ? clock64_() : 0;

                if (MODE == 0 || MODE == 5) {
                    fold_packed2<SIGNED>(slot, ra, rc);
                    if (MODE == 5 && blockIdx.x == 0 && lane == 0 && cnt < kTraceEvents) {
                        long long t3 = clock64_();
                        asm volatile("" : "+l"(t3) : "r"(slot[0]));
                        long long* tr = reinterpret_cast<long long*>(dump) + (static_cast<size_t>(warp) * kTraceEvents + cnt) * 4;
                        tr[0] = te0; tr[1] = te1; tr[2] = te2; tr[3] = t3;
                    }
                } else if (MODE == 2) {
                    slot[0] |= ra[0] | ra[31] | rc[0] | rc[31];
                } else if (MODE == 4) {
                    if (qr < job.q_n) {
                        uint32_t* d = reinterpret_cast<uint32_t*>(dump) + qr * (dump_ld / 2) +
                                      t * (kBlockN / 2) + c * (kAccCols / 2);
#pragma unroll
                        for (int q = 0; q < 32; ++q) { d[q] = ra[q]; d[32 + q] = rc[q]; }
                    }
                }
            }
            if ((MODE == 0 || MODE == 5) && nloop < ntiles) {
                int const t = nloop;
                int const ncols = job.c_n - t * kBlockN;   // valid columns of this tile
                mbar_wait(acc_full(h), cnt & 1, kWaitAccFull, cnt);
                tc_fence_after_sync();
                {
                    // A ragged last tile: the columns past the end of the view hold other views'
                    // rows.  Handled apart (32 columns at a time, 32-bit, valid columns only), so
                    // that the loop above carries no masking code.  (The values of a certified row
                    // fit 16 bits; an uncertified row's result is not used.)
                    int e0 = SIGNED ? -32768 : 0, e1 = e0;
#pragma unroll 1
                    for (int q4 = 0; q4 < kAccCols / kChunk; ++q4) {
                        int32_t v[32];
                        tmem_ld_32x32b_x32(taddr + q4 * kChunk, v);
                        tmem_ld_wait_regs(v);
                        int const col0 = c * kAccCols + q4 * kChunk;
#pragma unroll
                        for (int q = 0; q < 32; q += 2) {
                            if (col0 + q < ncols) e0 = max(e0, v[q]);
                            if (col0 + q + 1 < ncols) e1 = max(e1, v[q + 1]);
                        }
                    }
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(h));
                    slot[0] = pmax<SIGNED>(slot[0], (static_cast<uint32_t>(e0) & 0xffffu) | (static_cast<uint32_t>(e1) << 16));
                    ++cnt;
                }
            }
            if (MODE != 0 && MODE != 5) {
                if (MODE == 2 && slot[0] == 0x12345678u) dump[0] = 1;  // keep the loads alive
                continue;
            }

            // The row's 32 slots: 16 here, the other column half's through shared memory.  The
            // two warps of a (half, quadrant) pair meet at their own named barrier.  The merge
            // area is double-buffered across items, so that one barrier per item suffices: a
            // buffer is rewritten two items later, i.e. after the next item's barrier, which the
            // reading warp only reaches once it is done with this one.
            int v1, v2;
            slots_top2<SIGNED>(slot, v1, v2);
            if (c == 1) *merge = make_int4(v1, v2, 0, 0);
            named_barrier_sync(1 + h * 4 + quad, 64);
            if (c == 0) {
                int4 const o = *merge;
                v2 = max3(min(v1, o.x), v2, o.y);
                v1 = max(v1, o.x);
                if (qr < job.q_n) rowres[job.out_row + qr] = pack_rowres(v1, v2, j);
            }
        }
    } else if (RESOLVE) {
        // ===================== RESOLVE epilogue: exact result for the filter's survivors ======
        // Same warp layout as the filter.  Every row here carries the 16-bit norm certificate
        // and its best similarity V, which the filter found exactly, so the packed 16-bit view
        // of the accumulator is exact and nothing can exceed V.  What the reference's scan
        // (nearest_neighbor.cc:87-100) ends with follows from three facts about the row:
        //   idx = the last column whose similarity equals V (">=" lets the later one win),
        //   cnt = how many columns equal V (two or more: the second best is V itself),
        //   v2  = the largest similarity below V, not less than the initial 0.
        // Per load of 64 columns the warp takes the maximum of four groups of 16 columns; a
        // thread whose group maximum equals V sets that group's registers aside, and the groups
        // set aside are looked at value by value at the end of the item -- for all lanes at once,
        // instead of dragging the whole warp through the scan whenever one lane has a hit.
        int const ew = warp - kFirstEpilogueWarp;
        int const quad = warp & 3;
        int const h = ew >> 3;
        int const c = (ew >> 2) & 1;
        int const row = quad * 32 + lane;
        uint32_t const taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + h * kBlockN + c * kAccCols;
        int4* const merge_base = reinterpret_cast<int4*>(smem_gen + kSmemMerge);

        int j = 0;
        uint32_t cnt_tiles = 0, ic = 0;
        for (int it = item_first + blockIdx.x; it < total_items; it += gridDim.x, ++ic) {
            j = item_job[it];
            ScanJob const job = jobs[j];
            int const rb = it - job.item_start;
            int const nh = (job.q_n - rb * kItemM > kHalfM) ? 2 : 1;
            if (h >= nh) continue;
            int const ntiles = (job.c_n + kBlockN - 1) / kBlockN;
            int const r_in_job = rb * kItemM + h * kHalfM + row;
            // rows past the end of the job hold whatever follows in the scratch pool
            bool const live = r_in_job < job.q_n;
            int64_t const entry = live ? ex.xrow_map[job.out_row + r_in_job] : 0;
            uint32_t const v16 = static_cast<uint32_t>((static_cast<uint64_t>(entry) >> kSurvRowBits) & 0xffffu);
            // a dead row must never take the slow path: give it a V nothing can reach
            int const V0 = live ? (SIGNED ? static_cast<int>(static_cast<short>(v16)) : static_cast<int>(v16)) : 0x7fffffff;
            int V = V0;
            bool beaten = false;    // some similarity exceeds V (legitimate in the reverse pass only)

            int cnt = 0, idx = -1, v2 = 0;
            int cnt_r = 0, idx_r = -1, v2_r = 0;     // what a ragged last tile adds (its columns come last)
            bool dup = false;                        // V seen in more than one load of this warp
            ResolveStash pd;
            pd.slot = ex.stash + static_cast<size_t>(blockIdx.x) * 8 * kScanThreads + threadIdx.x;
            pd.col = -1;
            for (int t = 0; t < ntiles; ++t, ++cnt_tiles) {
                int const ncols = job.c_n - t * kBlockN;
                mbar_wait(acc_full(h), cnt_tiles & 1, kWaitAccFull, cnt_tiles);
                tc_fence_after_sync();
                if (ncols < kBlockN) {
                    // a ragged last tile, apart from the loop proper (see the filter): value by
                    // value over the valid columns
#pragma unroll 1
                    for (int q4 = 0; q4 < kAccCols / kChunk; ++q4) {
                        int32_t v[32];
                        tmem_ld_32x32b_x32(taddr + q4 * kChunk, v);
                        tmem_ld_wait_regs(v);
                        int const colq = t * kBlockN + c * kAccCols + q4 * kChunk;
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            bool const valid = colq + q < job.c_n;
                            bool const eq = valid && v[q] == V;
                            beaten = beaten || (valid && v[q] > V);
                            cnt_r += eq ? 1 : 0;
                            idx_r = eq ? colq + q : idx_r;
                            v2_r = max(v2_r, (valid && !eq) ? v[q] : 0);
                        }
                    }
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(h));
                    continue;
                }
                uint32_t ra[32], rc[32];
                tmem_ld_32x32b_x32_pack16(taddr, ra);
                tmem_ld_32x32b_x32_pack16(taddr + kAccCols / 2, rc);
                tmem_ld_wait_regs(ra);
                tmem_ld_wait_regs(rc);
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(h));
                int const col0 = t * kBlockN + c * kAccCols;
                resolve_load<SIGNED>(ra, col0, V, beaten, pd, dup, v2);
                resolve_load<SIGNED>(rc, col0 + kAccCols / 2, V, beaten, pd, dup, v2);
            }
            // The load set aside is looked at once, here, for all the rows of the warp; then
            // what the ragged tile found in the columns after it.
            resolve_flush<SIGNED>(pd, job.c_n, V, cnt, idx, v2);
            cnt += cnt_r + (dup ? 1 : 0);
            idx = idx_r >= 0 ? idx_r : idx;
            v2 = max(v2, v2_r);
            int4* const merge = merge_base + (ic & 1) * (kMergeBufBytes / 16) + h * kHalfM + row;
            if (c == 1) *merge = make_int4(cnt, idx, v2, beaten ? 1 : 0);
            named_barrier_sync(1 + h * 4 + quad, 64);
            if (c == 0 && live) {
                int4 const o = *merge;
                cnt += o.x;
                idx = max(idx, o.y);
                v2 = max(v2, o.z);
                beaten = beaten || o.w != 0;
                if (beaten) {
                    // reverse pass: the row's nearest neighbour is none of its claimants, so
                    // whatever it is, the mutual filter drops it.  Forward pass: cannot happen.
                    ex.oneway[surv_row(entry)] = -1;
                    if (!ex.verify) atomicAdd(ex.self_check, 1ull);
                } else {
                    int const second = cnt >= 2 ? V0 : v2;
                    bool const ok = passes_tests(ip_to_dist<SIGNED>(V0), ip_to_dist<SIGNED>(second), ex.sq_lowe, ex.sq_dist);
                    // signed: a best of 0 may be the initial value, reached by no candidate: index 0
                    int const found = (job.c_view < 0 && idx >= 0) ? ex.col_map[job.c_row + idx] : max(idx, 0);
                    ex.oneway[surv_row(entry)] = ok ? found : -1;
                    if (cnt == 0 && !(SIGNED && V0 == 0)) atomicAdd(ex.self_check, 1ull);   // the row's best was not found
                }
            }
        }
    } else if (EXACT && ((warp - kFirstEpilogueWarp) >> 2) < 2) {
        // ===================== EXACT epilogue: 2 x 4 warps replay the reference's scan ======
        int const my_h = (warp - kFirstEpilogueWarp) >> 2;   // the query half this group owns
        int const quad = warp & 3;
        int const row = quad * 32 + lane;
        uint32_t const taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + my_h * kBlockN;

        int j = 0;
        uint32_t hcnt = 0;                   // tiles this group has processed so far
        for (int it = item_first + blockIdx.x; it < total_items; it += gridDim.x) {
            j = item_job[it];
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
            for (int t = 0; t < ntiles; ++t, ++hcnt) {
                int const ncols = job.c_n - t * kBlockN;
                mbar_wait(acc_full(my_h), hcnt & 1, kWaitAccFull, hcnt);
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
                if (nbig > kMaxBigPerRow && mark_once(ex.replay_flags, g))   // cannot be certified by verify_big_kernel
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
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        prof[0] = static_cast<unsigned long long>(clock64_() - prof_c0);
        prof[1] = globaltimer_ns() - prof_t0;
    }
}

}  // namespace osfm
