// osfm_match.cu -- host side of libosfm_match.so: the C ABI declared in
// include/osfm_match.h over the sm_100a kernels in scan_kernel.cuh / post_kernels.cuh.
//
// Host-side mirror of the reference (paths relative to /root/reference):
//   ExhaustiveMatching::init                  src/mve/sfm/exhaustive_matching.cc:56-112
//   ExhaustiveMatching::pairwise_match        src/mve/sfm/exhaustive_matching.cc:115-144
//   ExhaustiveMatching::pairwise_match_lowres src/mve/sfm/exhaustive_matching.cc:147-180
//   Matching::twoway_match                    src/mve/sfm/matching.h:148-159
//   Matching::combine_results                 src/mve/sfm/matching.cc:50-89
//
// There is no CPU path in this file: every result comes from the kernels.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include <cuda.h>
#include <cuda_runtime.h>
#include <nccl.h>      // types and prototypes only: the library is resolved at run time (NcclApi)

#include "../../include/osfm_match.h"
#include "float_kernels.cuh"
#include "float_tc_kernels.cuh"
#include "io_formats.cuh"
#include "post_kernels.cuh"
#include "ransac_kernels.cuh"
#include "scan_kernel.cuh"
#include "tracks_kernels.cuh"

using namespace osfm;

namespace {

constexpr int kPadRows = 256;                  // readable zero rows after the last view
constexpr int64_t kMaxBatchRows = 16ll << 20;  // job rows per batch (bounds scratch memory)
constexpr int64_t kMaxBatchDense = 1ll << 30;  // dense result ints per batch (4 GiB)

// One direction of one image pair.  reverse = false: goes through the filter pass.  reverse = true:
// the other direction of spec `partner`, evaluated only for the rows the partner's results claim
// (post_kernels.cuh, "reverse direction of a pair").
struct JobSpec { int q_view, q_n, c_view, c_n; bool reverse; int partner; };

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t count) {
        if (count <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t const want = std::max<size_t>(count + count / 8, 1024);
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct KindPool {
    bool is_signed = false;
    int dim = 128;
    uint8_t* pool = nullptr;        // rows of kRowBytes bytes
    bool owned = false;
    int64_t rows = 0;               // rows that belong to views
    std::vector<int64_t> off;       // first row of each view
    std::vector<int32_t> n;         // descriptors per view
    // wrap certificate (scan_kernel.cuh): squared norm per pool row, largest per view
    DevBuf<int32_t> d_norm2;
    DevBuf<int32_t> d_viewmax;
    DevBuf<int64_t> d_view_off;
    DevBuf<int32_t> d_view_n;
    DevBuf<int2> d_danger;          // unsigned kind: kDangerCap highest-norm rows per view
    DevBuf<int32_t> d_danger_cnt;
    DevBuf<int32_t> d_danger_floor;
    CUtensorMap tmap;
    // Staging arena: views are appended in the order set_view is called.  When that is the
    // view-id order (the normal case) the arena simply becomes the pool at commit; it is
    // kept across begin() calls so that re-staging allocates nothing.
    uint8_t* arena = nullptr;
    int64_t arena_cap = 0;          // rows
    int64_t arena_used = 0;         // rows
    std::vector<int64_t> stage_off; // arena row of each staged view (-1: not staged)
    int last_staged = -1;
    bool in_order = true;
    int norm_views = 0;             // views [0, norm_views) have their norms (lazy commit)
    float lowe = 0.8f, dist = FLT_MAX;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// One image pair as the batched entry points see it.
struct PairPlan {
    int v1, v2;
    int n1[2], n2[2];        // effective sizes per kind (0 if the kind does not contribute)
    int64_t out12, out21;    // offsets of the combined vectors in the dense result
    int len12, len21;
};

// NCCL, bound at run time: a multi-device matcher replicates its descriptor pool with one
// ncclBroadcast over NVLink (osfm_match_create_multi).  dlopen instead of a link-time dependency
// so that a host process that already carries an NCCL (PyTorch bundles its own) is not handed a
// second copy, and a single-device user needs none.
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(lib, "ncclBroadcast"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Broadcast || !GetErrorString) {
            err = "libnccl.so.2 lacks an expected symbol";
            return false;
        }
        return true;
    }
};

// Device time per phase of a batched call: events recorded on the handle's stream between the
// phases, read after the call's synchronisations (osfm_match_stats::last_phase_ms).
enum Phase { kPhFilter = 0, kPhClassify, kPhResolveFwd, kPhClaim, kPhResolveRev, kPhMutual, kPhCompact, kPhCount };
struct PhaseTimer {
    std::vector<cudaEvent_t> pool;
    std::vector<int> phase;          // phase that starts at event i (-1: untimed)
    double acc[kPhCount] = {0, 0, 0, 0, 0, 0, 0};
    cudaError_t mark(cudaStream_t s, int ph) {
        if (phase.size() == pool.size()) {
            cudaEvent_t e = nullptr;
            cudaError_t r = cudaEventCreate(&e);
            if (r != cudaSuccess) return r;
            pool.push_back(e);
        }
        cudaError_t r = cudaEventRecord(pool[phase.size()], s);
        if (r == cudaSuccess) phase.push_back(ph);
        return r;
    }
    // all marks must have completed (call after synchronising the stream)
    void collect() {
        for (size_t i = 0; i + 1 < phase.size(); ++i) {
            if (phase[i] < 0) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pool[i], pool[i + 1]) == cudaSuccess) acc[phase[i]] += ms;
        }
        phase.clear();
    }
    void reset() { phase.clear(); for (double& a : acc) a = 0.0; }
    void release() { for (cudaEvent_t e : pool) cudaEventDestroy(e); pool.clear(); phase.clear(); }
};

}  // namespace

struct osfm_matcher {
    std::recursive_mutex mu;     // recursive: osfm_match_two_view holds it across the stages it calls
    std::string err;
    osfm_match_config cfg;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    EncodeTiledFn encode = nullptr;
    HangReport* hang_host = nullptr;

    int num_views = 0;
    bool began = false, committed = false;
    KindPool kind[2];
    // Overlapped staging (osfm_match_begin_overlapped): the host-to-device copies go to their own
    // stream, one event per view; commit returns without waiting and the first batches of a
    // pair list are matched while the later views are still on their way.
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> view_ev;
    std::vector<char> view_ev_set;
    std::vector<int64_t> view_ev_seq;   // order in which the events were recorded
    int64_t ev_seq = 0;
    bool overlap = false;      // this begin/commit cycle stages through copy_stream
    bool copies_in_flight = false;   // copy_stream may still be reading the caller's buffers
    bool lazy = false;         // committed, but norms (and maybe copies) of later views still pending

    DevBuf<ScanJob> d_jobs;
    DevBuf<int2> d_rowres;
    // few work items (one pair of large views): the filter pass runs on column segments of the jobs
    // (run_jobs), its records land in d_rowres_part and merge_split_kernel folds them
    DevBuf<ScanJob> d_fjobs;
    DevBuf<int2> d_rowres_part;
    const ScanJob* filter_jobs = nullptr;      // what launch_scan_t hands to the kernel (null: d_jobs / d_rowres)
    int2* filter_rowres = nullptr;
    DevBuf<int32_t> d_oneway;
    DevBuf<int64_t> d_cand;
    DevBuf<int64_t> d_cand_rev;          // claimed rows without the norm certificate (reverse pass)
    DevBuf<int4> d_big;
    DevBuf<PairPart> d_parts;
    DevBuf<int32_t> d_dense;
    DevBuf<int32_t> d_counts;
    DevBuf<int64_t> d_listoff;
    DevBuf<int2> d_list;
    DevBuf<int> tr_ints;                 // tracks scratch (osfm_tracks_compute)
    DevBuf<unsigned long long> tr_table;
    DevBuf<int64_t> tr_meta;
    DevBuf<int32_t> tr_meta32;
    DevBuf<float4> rs_xy;                // RANSAC scratch (osfm_ransac_fundamental)
    DevBuf<float2> rs_pos;
    DevBuf<int32_t> rs_samples;
    DevBuf<double> rs_F;
    DevBuf<int> rs_cnt;
    DevBuf<double> rs_stage1;            // bidiagonal + V of every fit of one chunk
    DevBuf<int2> rs_out;
    int32_t* rs_stage[2] = {nullptr, nullptr};   // pinned staging for samples drawn here, double-buffered
    size_t rs_stage_ints = 0;
    cudaEvent_t rs_stage_free[2] = {nullptr, nullptr};
    DevBuf<float> d_ftmp;
    // float path, tensor-core filter (float_tc_kernels.cuh): hi / lo parts, norms, the filter's row records,
    // the rows left to the exact kernel
    DevBuf<float> d_fsplit, d_fnorm;
    DevBuf<FloatTopRow> d_ftop;
    DevBuf<int32_t> d_flist;
    DevBuf<FloatRowState> d_fparts;      // listed rows x candidate slices (float_oneway_kernel / float_finish_kernel)
    int* d_fmeta = nullptr;              // [0], [1]: rows left per direction; [2], [3]: largest norm per set (float bits)
    int float_mode = 0;                  // 0: by size; 1: exact kernel only; 2: always filter first
    // Float descriptors (osfm_match_set_view_f32) are staged at commit, by a few host threads at
    // once: each copies views into its own page-locked buffers, sends them on and quantises them on
    // its own stream (a single pageable cudaMemcpy moves about 11 GB/s, a fifth of the link).
    struct FloatView { int kd; int64_t arena_row; const float* src; int n; int stride; };
    std::vector<FloatView> float_views;
    struct FloatLane {
        cudaStream_t stream = nullptr;
        float* pinned[2] = {nullptr, nullptr};
        float* dev[2] = {nullptr, nullptr};
        cudaEvent_t done[2] = {nullptr, nullptr};
        size_t cap = 0;            // floats per buffer
    };
    std::vector<FloatLane> float_lanes;
    DevBuf<int32_t> d_seg_first;
    DevBuf<int32_t> d_item_job;          // filter pass: the job of every work item
    DevBuf<uint4> d_stash;               // RESOLVE pass: one set-aside packed load per thread (ResolveStash)
    DevBuf<int32_t> d_rev_of;            // per job: the reverse job of its pair (or -1)
    // restricted candidate sets of the reverse pass (select_candidates_kernel)
    DevBuf<int32_t> d_tau;
    DevBuf<ScanJob> d_jobs_rev;
    DevBuf<int32_t> d_seg_first_rev;
    DevBuf<uint8_t> d_cand_pool;
    DevBuf<int32_t> d_cand_map;
    DevBuf<int32_t> d_cand_cnt;                // per forward job: rows kept (-1: the reverse job keeps the whole view)
    CUtensorMap cand_tmap;
    uint8_t* cand_tmap_for = nullptr;
    size_t cand_tmap_rows = 0;
    int reverse_mode = 0;                // 0: restricted candidate sets; 2: whole views (A/B switch)
    DevBuf<uint32_t> d_replay_flags;     // bitmap over a batch's rows: already in the replay list
    // EXACT rows of very few jobs against large views: inner products on CUDA cores, spread over
    // the device, then a warp-per-row replay (post_kernels.cuh, exact_dots_kernel)
    DevBuf<int32_t> d_xw_x, d_xw_xmax;
    DevBuf<int64_t> d_xw_off, d_xw_moff;
    DevBuf<int> d_xw_unit;
    int* d_xw_meta = nullptr;            // [0] path taken, [1] units, [4] work items left to the scan pass
    int exact_mode = 0;                  // 0: by capacity; 1: always the scan pass (A/B switch)
    int batch_max_cn = 0;                // largest candidate set of the batch being run
    // scratch of the two second passes over gathered rows: [0] RESOLVE (the filter's certified
    // survivors), [1] EXACT (unsigned rows without the 16-bit norm certificate)
    struct SecondPass {
        DevBuf<int64_t> list;       // per-job segments of row entries (surv_entry)
        DevBuf<int> cnt;            // rows per job
        DevBuf<int> job_xrow;
        DevBuf<ScanJob> xjobs;
        DevBuf<int32_t> item_job;
        DevBuf<uint8_t> xpool;      // gathered query rows
        DevBuf<int64_t> xrow_map;
        int* d_xmeta = nullptr;
        CUtensorMap tmap;
        uint8_t* tmap_for = nullptr;
        size_t tmap_rows = 0;
        void release() {
            list.release(); cnt.release(); job_xrow.release(); xjobs.release(); item_job.release(); xpool.release(); xrow_map.release();
            if (d_xmeta) cudaFree(d_xmeta);
            d_xmeta = nullptr;
        }
    } pass[2];
    unsigned long long* d_counters = nullptr;  // see PostParams::counters

    // Multi-device matcher (osfm_match_create_multi): this handle is the primary; peers[] are full
    // handles on the other devices that hold a replica of the pool and take a share of every
    // batched call.  comms[0] belongs to this device, comms[1 + i] to peers[i].
    std::vector<osfm_matcher*> peers;
    std::vector<ncclComm_t> comms;
    NcclApi nccl;
    double last_broadcast_ms = 0.0;

    // Look-ahead for the pair-by-pair plugin loop (osfm_match_set_lookahead): dense results / low-res
    // counts of a window of pairs in the reference's enumeration order, kept in pinned host memory.
    int lookahead = 0;
    struct PairCache {
        int64_t first = -1;               // flat pair index of the first cached pair (-1: empty)
        int count = 0;
        int num_features = 0;             // low-res cache: the feature limit it was computed with
        std::vector<int64_t> offsets;     // full cache: list offsets, count + 1
        std::vector<int32_t> counts;
        std::vector<int32_t> lens;        // full cache: lengths of the two dense vectors per pair
        int32_t* host = nullptr;          // full cache: the (i, j) lists, pinned
        size_t host_cap = 0;              // ints
        void clear() { first = -1; count = 0; }
    } cache_full, cache_lowres;

    int scan_mode = 0;
    bool both_directions = false;     // debug / A-B switch: run both directions through the filter
    PhaseTimer phases;
    osfm_match_stats stats;
};

namespace {

int fail(osfm_matcher* m, int code, const char* fmt, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m) m->err = buf;
    return code;
}

int cuda_fail(osfm_matcher* m, cudaError_t e, const char* what) {
    char extra[200] = "";
    if (m && m->hang_host && m->hang_host->flag)
        snprintf(extra, sizeof extra,
                 " [kernel watchdog: wait code %u block %u thread %u parity %u aux %u]",
                 m->hang_host->code, m->hang_host->block, m->hang_host->thread,
                 m->hang_host->parity, m->hang_host->aux);
    return fail(m, OSFM_ERR_CUDA, "%s: %s%s", what, cudaGetErrorString(e), extra);
}

#define CU_TRY(m, call)                                             \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail((m), e__, #call);  \
    } while (0)

#define OS_TRY(call)                  \
    do {                              \
        int r__ = (call);             \
        if (r__ != OSFM_OK) return r__; \
    } while (0)

// No C++ exception may cross the C ABI (a std::bad_alloc from a vector, say, would terminate the
// host process): every entry point that does more than read a field runs inside this pair.
#define OSFM_TRY_BEGIN try {
#define OSFM_TRY_END(handle, internal_code)                                                             \
    } catch (const std::bad_alloc&) {                                                                   \
        return fail((handle), OSFM_ERR_OUT_OF_MEMORY, "out of host memory");                           \
    } catch (const std::exception& e) {                                                                 \
        return fail((handle), (internal_code), "unexpected exception: %s", e.what());                  \
    } catch (...) {                                                                                     \
        return fail((handle), (internal_code), "unexpected exception");                                \
    }

// Forgets the views (keeps the staging arena's memory unless release_arena).
void reset_kind(KindPool& k, bool release_arena) {
    if (k.owned && k.pool && k.pool != k.arena) cudaFree(k.pool);
    k.pool = nullptr; k.owned = false;
    k.rows = 0; k.off.clear(); k.n.clear();
    k.stage_off.clear(); k.arena_used = 0; k.last_staged = -1; k.in_order = true;
    if (release_arena && k.arena) { cudaFree(k.arena); k.arena = nullptr; k.arena_cap = 0; }
    if (release_arena) { k.d_norm2.release(); k.d_viewmax.release(); k.d_view_off.release(); k.d_view_n.release(); k.d_danger.release(); k.d_danger_cnt.release(); k.d_danger_floor.release(); }
}

// Makes room for `rows` more rows in the staging arena (contents are preserved).
int arena_reserve(osfm_matcher* m, KindPool& k, int64_t rows);

// 2-D tensor map over a pool of 128-byte rows; one TMA box = 128 rows (a query half or half
// a candidate tile), SWIZZLE_128B so the tile lands in the layout tcgen05.mma reads.
int encode_tmap(osfm_matcher* m, CUtensorMap* map, void* base, int64_t rows_with_pad) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(kRowBytes), static_cast<cuuint64_t>(rows_with_pad)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kRowBytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kRowBytes), static_cast<cuuint32_t>(kHalfM)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = m->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(m, OSFM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return OSFM_OK;
}

// Float matrix of rows_with_pad x 128 floats (512-byte rows): boxes of 128 bytes x 128 rows, i.e.
// one 32-float chunk of 128 descriptors, laid out like the 8-bit tiles (SWIZZLE_128B).
int encode_tmap_f32(osfm_matcher* m, CUtensorMap* map, void* base, int64_t rows_with_pad) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(kFDim * 4), static_cast<cuuint64_t>(rows_with_pad)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kFDim * 4)};
    cuuint32_t box[2] = {128u, static_cast<cuuint32_t>(kFtM)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = m->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(m, OSFM_ERR_CUDA, "cuTensorMapEncodeTiled (float) failed (%d)", (int)r);
    return OSFM_OK;
}

int make_tmap(osfm_matcher* m, KindPool& k) { return encode_tmap(m, &k.tmap, k.pool, k.rows + kPadRows); }

// Per-row squared norms + per-view maxima (the filter's wrap certificate).  Asynchronous on
// the handle's stream; k.off / k.n are staged by the runtime before the calls return.
// Buffers and the per-view tables the norm kernels read.
int norms_prepare(osfm_matcher* m, KindPool& k) {
    size_t const nv = k.n.size();
    CU_TRY(m, k.d_norm2.reserve(static_cast<size_t>(k.rows + kPadRows)));
    CU_TRY(m, k.d_viewmax.reserve(std::max<size_t>(nv, 1)));
    CU_TRY(m, cudaMemsetAsync(k.d_viewmax.p, 0, sizeof(int32_t) * std::max<size_t>(nv, 1), m->stream));
    k.norm_views = 0;
    if (k.rows == 0 || nv == 0) { k.norm_views = static_cast<int>(nv); return OSFM_OK; }
    CU_TRY(m, k.d_view_off.reserve(nv));
    CU_TRY(m, k.d_view_n.reserve(nv));
    CU_TRY(m, cudaMemcpyAsync(k.d_view_off.p, k.off.data(), sizeof(int64_t) * nv, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpyAsync(k.d_view_n.p, k.n.data(), sizeof(int32_t) * nv, cudaMemcpyHostToDevice, m->stream));
    if (!k.is_signed) {
        CU_TRY(m, k.d_danger.reserve(nv * kDangerCap));
        CU_TRY(m, k.d_danger_cnt.reserve(nv));
        CU_TRY(m, k.d_danger_floor.reserve(nv));
    }
    return OSFM_OK;
}

// Norms, per-view maxima and danger lists of the views [v0, v1) (their rows are contiguous:
// off[] ascends with the view id).
int norms_run(osfm_matcher* m, KindPool& k, int v0, int v1) {
    if (v1 <= v0 || k.rows == 0) return OSFM_OK;
    int64_t const row0 = k.off[v0];
    int64_t const row1 = k.off[v1 - 1] + k.n[v1 - 1];
    int max_n = 0;
    for (int v = v0; v < v1; ++v) max_n = std::max(max_n, k.n[v]);
    if (row1 > row0) {
        int const grid = static_cast<int>((row1 - row0 + 255) / 256);
        uint8_t const* const rows = k.pool + static_cast<size_t>(row0) * kRowBytes;
        if (k.is_signed) rownorm_kernel<true><<<grid, 256, 0, m->stream>>>(rows, row1 - row0, k.d_norm2.p + row0);
        else             rownorm_kernel<false><<<grid, 256, 0, m->stream>>>(rows, row1 - row0, k.d_norm2.p + row0);
        CU_TRY(m, cudaGetLastError());
        m->stats.kernel_launches++;
    }
    dim3 const vgrid(static_cast<unsigned>(v1 - v0), static_cast<unsigned>(std::max(1, (max_n + kViewMaxChunk - 1) / kViewMaxChunk)));
    viewmax_kernel<<<vgrid, 256, 0, m->stream>>>(k.d_norm2.p, k.d_view_off.p + v0, k.d_view_n.p + v0, k.d_viewmax.p + v0);
    CU_TRY(m, cudaGetLastError());
    m->stats.kernel_launches++;
    if (!k.is_signed) {
        danger_kernel<<<static_cast<unsigned>(v1 - v0), kDangerCap, 0, m->stream>>>(
            k.d_norm2.p, k.d_view_off.p + v0, k.d_view_n.p + v0, k.d_viewmax.p + v0,
            k.d_danger.p + static_cast<size_t>(v0) * kDangerCap, k.d_danger_cnt.p + v0, k.d_danger_floor.p + v0);
        CU_TRY(m, cudaGetLastError());
        m->stats.kernel_launches++;
    }
    return OSFM_OK;
}

int compute_norms(osfm_matcher* m, KindPool& k) {
    OS_TRY(norms_prepare(m, k));
    int const nv = static_cast<int>(k.n.size());
    OS_TRY(norms_run(m, k, 0, nv));
    k.norm_views = nv;
    return OSFM_OK;
}

// Lazy commit: makes the views up to max_view usable on the main stream (their copies have
// arrived, their norms exist).
int ensure_views(osfm_matcher* m, int max_view) {
    if (!m->lazy) return OSFM_OK;
    max_view = std::min(max_view, m->num_views - 1);
    bool need = false;
    for (int kd = 0; kd < 2; ++kd) need = need || m->kind[kd].norm_views <= max_view;
    if (need) {
        // copy_stream is in order: the event recorded last among the views up to max_view
        // covers every copy those views need
        int ev = -1;
        for (int v = 0; v <= max_view; ++v)
            if (m->view_ev_set[v] && (ev < 0 || m->view_ev_seq[v] > m->view_ev_seq[ev])) ev = v;
        if (ev >= 0) CU_TRY(m, cudaStreamWaitEvent(m->stream, m->view_ev[ev], 0));
        for (int kd = 0; kd < 2; ++kd) {
            KindPool& k = m->kind[kd];
            if (k.norm_views > max_view) continue;
            OS_TRY(norms_run(m, k, k.norm_views, max_view + 1));
            k.norm_views = max_view + 1;
        }
    }
    if (m->kind[0].norm_views >= m->num_views && m->kind[1].norm_views >= m->num_views) m->lazy = false;
    return OSFM_OK;
}

int arena_reserve(osfm_matcher* m, KindPool& k, int64_t rows) {
    int64_t const need = k.arena_used + rows;
    if (need <= k.arena_cap) return OSFM_OK;
    int64_t const cap = std::max<int64_t>(need + need / 2, 4096);
    uint8_t* fresh = nullptr;
    CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&fresh), static_cast<size_t>(cap) * kRowBytes));
    if (k.arena) {
        if (m->copy_stream) CU_TRY(m, cudaStreamSynchronize(m->copy_stream));    // copies into the old arena
        if (k.arena_used > 0)
            CU_TRY(m, cudaMemcpyAsync(fresh, k.arena, static_cast<size_t>(k.arena_used) * kRowBytes,
                                      cudaMemcpyDeviceToDevice, m->stream));
        CU_TRY(m, cudaStreamSynchronize(m->stream));
        if (k.pool == k.arena) k.pool = fresh;
        cudaFree(k.arena);
    }
    k.arena = fresh;
    k.arena_cap = cap;
    return OSFM_OK;
}

int ksteps_of(const KindPool& k) { return (k.dim + 31) / 32; }

template <int MODE, bool SIGNED>
cudaError_t launch_scan_t(osfm_matcher* m, const KindPool& k, int item_first, int total_items, int32_t* dump, int64_t dump_ld) {
    cudaError_t e = cudaFuncSetAttribute(scan_kernel<MODE, kPassFilter, SIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kScanSmemBytes);
    if (e != cudaSuccess) return e;
    int const grid = std::min(m->num_sms, total_items - item_first);
    uint32_t const idesc = make_idesc_i8(kHalfM, kBlockN, SIGNED ? 1 : 0, SIGNED ? 1 : 0);
    ExactParams ex;
    memset(&ex, 0, sizeof ex);
    ex.norm2 = k.d_norm2.p;
    ex.viewmax = k.d_viewmax.p;
    scan_kernel<MODE, kPassFilter, SIGNED><<<grid, kScanThreads, kScanSmemBytes, m->stream>>>(
        k.tmap, k.tmap, k.tmap, m->filter_jobs ? m->filter_jobs : m->d_jobs.p, m->d_item_job.p, item_first, total_items, idesc,
        ksteps_of(k), dump, dump_ld, ex, m->filter_rowres ? m->filter_rowres : m->d_rowres.p,
        m->d_counters + 6);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_scan(osfm_matcher* m, const KindPool& k, int item_first, int total_items, int32_t* dump, int64_t dump_ld) {
    return k.is_signed ? launch_scan_t<MODE, true>(m, k, item_first, total_items, dump, dump_ld)
                       : launch_scan_t<MODE, false>(m, k, item_first, total_items, dump, dump_ld);
}

// A second pass over gathered rows (RESOLVE: the filter's certified survivors; EXACT: unsigned
// rows without certificate): plan, gather, and that variant of the scan kernel.  Everything is
// sized on the device; the host never learns how many rows there were until it reads the
// counters.
// The reverse RESOLVE pass over restricted candidate sets: its own copy of the job list (reverse
// jobs redirected to their subsets), one segment per reverse job, the gathered candidate pool.
struct ReverseSubset {
    const ScanJob* jobs;
    const int32_t* seg_first;
    int nseg;
    const CUtensorMap* tmap;
    const int32_t* col_map;
};

template <int PASS, bool SIGNED>
int launch_second_pass(osfm_matcher* m, const KindPool& k, int njobs, int nseg, int64_t rows, const PostParams& pp,
                       bool verify, const ReverseSubset* subset = nullptr) {
    osfm_matcher::SecondPass& sp = m->pass[PASS == kPassResolve ? 0 : 1];
    CU_TRY(m, sp.job_xrow.reserve(static_cast<size_t>(njobs)));
    CU_TRY(m, sp.xjobs.reserve(static_cast<size_t>(nseg) + 1));
    CU_TRY(m, sp.xrow_map.reserve(static_cast<size_t>(rows)));
    CU_TRY(m, sp.xpool.reserve(static_cast<size_t>(rows + kPadRows) * kRowBytes));
    if (sp.tmap_for != sp.xpool.p || sp.tmap_rows != sp.xpool.cap) {
        OS_TRY(encode_tmap(m, &sp.tmap, sp.xpool.p, static_cast<int64_t>(sp.xpool.cap / kRowBytes)));
        sp.tmap_for = sp.xpool.p;
        sp.tmap_rows = sp.xpool.cap;
    }
    if (subset) CU_TRY(m, sp.xjobs.reserve(static_cast<size_t>(subset->nseg) + 1));
    // every segment's rows make ceil(rows / 256) items: at most rows / 256 + one per segment
    CU_TRY(m, sp.item_job.reserve(static_cast<size_t>(rows / kItemM) + static_cast<size_t>(subset ? subset->nseg : nseg) + 1));
    ExactWideParams xw;
    memset(&xw, 0, sizeof xw);
    if (PASS == kPassExact) {
        CU_TRY(m, m->d_big.reserve(static_cast<size_t>(rows) * kMaxBigPerRow));
        size_t const flag_words = static_cast<size_t>(rows / 32 + 1);
        CU_TRY(m, m->d_replay_flags.reserve(flag_words));
        CU_TRY(m, cudaMemsetAsync(m->d_replay_flags.p, 0, flag_words * sizeof(uint32_t), m->stream));
        CU_TRY(m, cudaMemsetAsync(m->d_counters + 8, 0, sizeof(unsigned long long), m->stream));
        CU_TRY(m, cudaMemsetAsync(m->d_counters + 12, 0, sizeof(unsigned long long), m->stream));
        // Few rows against large views: every inner product on CUDA cores, spread over the device,
        // then a warp-per-row replay; the scan pass then finds no work items.  Decided on the device
        // (the host does not know how many rows there are, plan_rows_kernel's tail does) by what
        // fits the scratch buffer.
        if (!m->d_xw_meta) {
            CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&m->d_xw_meta), 8 * sizeof(int)));
            CU_TRY(m, cudaMemset(m->d_xw_meta, 0, 8 * sizeof(int)));
        }
        xw.x_cap = std::min<int64_t>(int64_t(1) << 26, std::max<int64_t>(int64_t(1) << 20, int64_t(512) * m->batch_max_cn));
        xw.xm_cap = xw.x_cap / 16;
        xw.max_jobs = std::min(nseg, 64);      // the layout is sized by one thread; many candidate views: scan pass
        CU_TRY(m, m->d_xw_x.reserve(static_cast<size_t>(xw.x_cap)));
        CU_TRY(m, m->d_xw_xmax.reserve(static_cast<size_t>(xw.xm_cap)));
        CU_TRY(m, m->d_xw_off.reserve(static_cast<size_t>(nseg) + 1));
        CU_TRY(m, m->d_xw_moff.reserve(static_cast<size_t>(nseg) + 1));
        CU_TRY(m, m->d_xw_unit.reserve(static_cast<size_t>(nseg) + 1));
        xw.xjobs = sp.xjobs.p;
        xw.xmeta = sp.d_xmeta;
        xw.xpool = sp.xpool.p;
        xw.pool = k.pool;
        xw.xrow_map = sp.xrow_map.p;
        xw.x_off = m->d_xw_off.p;
        xw.xm_off = m->d_xw_moff.p;
        xw.unit_first = m->d_xw_unit.p;
        xw.meta = m->d_xw_meta;
        xw.x = m->d_xw_x.p;
        xw.xmax = m->d_xw_xmax.p;
        xw.mode = m->exact_mode;
        xw.oneway = pp.oneway;
        xw.sq_lowe = pp.sq_lowe;
        xw.sq_dist = pp.sq_dist;
        xw.big_list = m->d_big.p;
        xw.big_count = m->d_counters + 12;
        xw.replay_list = m->d_cand.p;
        xw.replay_count = m->d_counters + 8;
        xw.replay_flags = m->d_replay_flags.p;
        xw.self_check = m->d_counters + 2;
        xw.wide_rows = m->d_counters + 13;
    }
    plan_rows_kernel<<<1, 1024, 0, m->stream>>>(subset ? subset->jobs : m->d_jobs.p,
                                                 subset ? subset->seg_first : m->d_seg_first.p,
                                                 subset ? subset->nseg : nseg, sp.cnt.p,
                                                 sp.xjobs.p, sp.job_xrow.p, sp.d_xmeta,
                                                 PASS == kPassExact ? m->d_counters + 5 : nullptr, sp.item_job.p, xw);
    CU_TRY(m, cudaGetLastError());
    // few jobs (one pair of two large views): several blocks share a job's list
    int const gather_x = std::min(std::max(njobs, 1), m->num_sms * 16);
    dim3 const gather_grid(gather_x, std::max(1, std::min(64, m->num_sms * 4 / gather_x)));
    gather_rows_kernel<<<gather_grid, 256, 0, m->stream>>>(m->d_jobs.p, njobs, sp.cnt.p,
                                                              sp.job_xrow.p, sp.list.p, k.pool,
                                                              sp.xpool.p, sp.xrow_map.p,
                                                              verify && PASS == kPassResolve ? m->d_rowres.p : nullptr);
    CU_TRY(m, cudaGetLastError());
    CU_TRY(m, cudaFuncSetAttribute(scan_kernel<0, PASS, SIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, kScanSmemBytes));
    ExactParams ex;
    ex.qpool = sp.xpool.p;
    ex.cpool = k.pool;
    ex.xrow_map = sp.xrow_map.p;
    ex.oneway = pp.oneway;
    ex.total_items_dev = sp.d_xmeta;
    ex.sq_lowe = pp.sq_lowe;
    ex.sq_dist = pp.sq_dist;
    ex.replay_list = m->d_cand.p;
    ex.replay_count = m->d_counters + 8;
    ex.big_list = nullptr;
    ex.big_count = m->d_counters + 12;
    ex.self_check = m->d_counters + 2;
    ex.norm2 = nullptr;
    ex.viewmax = nullptr;
    ex.verify = verify ? 1 : 0;
    ex.col_map = subset ? subset->col_map : nullptr;
    ex.stash = nullptr;
    if (PASS == kPassResolve) {
        CU_TRY(m, m->d_stash.reserve(static_cast<size_t>(m->num_sms) * 8 * kScanThreads));
        ex.stash = m->d_stash.p;
    }
    ex.replay_flags = nullptr;
    if (PASS == kPassExact) {
        ex.big_list = m->d_big.p;
        ex.replay_flags = m->d_replay_flags.p;
    }
    if (PASS == kPassExact) {
        exact_dots_kernel<<<m->num_sms * 4, kWideColBlock, 0, m->stream>>>(xw);
        CU_TRY(m, cudaGetLastError());
        exact_replay_kernel<<<m->num_sms, 256, 0, m->stream>>>(xw);
        CU_TRY(m, cudaGetLastError());
        m->stats.kernel_launches += 2;
        ex.total_items_dev = m->d_xw_meta + 4;
    }
    uint32_t const idesc = make_idesc_i8(kHalfM, kBlockN, SIGNED ? 1 : 0, SIGNED ? 1 : 0);
    scan_kernel<0, PASS, SIGNED><<<m->num_sms, kScanThreads, kScanSmemBytes, m->stream>>>(
        sp.tmap, k.tmap, subset ? *subset->tmap : k.tmap, sp.xjobs.p, sp.item_job.p, 0, 0, idesc, ksteps_of(k), nullptr, 0, ex, nullptr, nullptr);
    CU_TRY(m, cudaGetLastError());
    m->stats.kernel_launches += 3;
    if (PASS == kPassExact) {
        // certify the big candidates; rows that fail (a 16-bit lane really wrapped) get the
        // warp-per-row replay.  counters[8] is the replay list length.
        verify_big_kernel<<<m->num_sms * 2, 256, 0, m->stream>>>(pp, m->d_big.p, m->d_counters + 12,
                                                                m->d_cand.p, m->d_counters + 8, m->d_replay_flags.p);
        CU_TRY(m, cudaGetLastError());
        PostParams rp = pp;
        rp.slow_list = m->d_cand.p;
        rp.counters = m->d_counters + 8;
        rp.slow_len = m->d_counters + 8;
        slow_rows_kernel<false><<<m->num_sms * 2, 256, 0, m->stream>>>(rp);
        CU_TRY(m, cudaGetLastError());
        m->stats.kernel_launches += 2;
    }
    return OSFM_OK;
}

// Runs the filter scan + RESOLVE / EXACT passes (+ wrap emulation) for a list of jobs of one kind,
// then the reverse jobs (evaluated for their claimed rows only).  On return (in stream order)
// m->d_oneway holds, for job i, q_n one-way results starting at out_row[i] (out_row[i] = -1 if the
// job was not run because one side is empty).  A reverse job's rows hold the reference's result
// where it can survive the mutual filter and -1 elsewhere.
int run_jobs(osfm_matcher* m, int kind_id, const std::vector<JobSpec>& specs,
             std::vector<int64_t>& out_row, int32_t* dump = nullptr, int64_t dump_ld = 0, int dump_mode = 3) {
    KindPool& k = m->kind[kind_id];
    out_row.assign(specs.size(), -1);
    m->batch_max_cn = 0;
    std::vector<ScanJob> jobs;
    jobs.reserve(specs.size() + 1);
    // Forward jobs first, then the reverse jobs; each group ordered by candidate view:
    // concurrently running work items then stream the same candidate tiles (L2 locality), and the
    // second passes can merge the rows of all jobs that share a candidate view into full 256-row
    // items.
    //
    // Overlapped staging (views still on their way, m->lazy): the forward jobs are grouped into
    // buckets by the highest view they need, view ranges growing geometrically, and the filter pass
    // is launched bucket by bucket, each launch waiting (on the device) only for its views: the
    // filter works on the early pairs of a list in the reference's order while the later views
    // are being copied, and nothing else of the batch is split.
    bool const staged = m->lazy && dump == nullptr;
    std::vector<int> bucket_end;      // exclusive upper view id per bucket
    if (staged && m->num_views >= 8) {
        int b = std::max(2, (m->num_views + 5) / 6);
        while (b < m->num_views) {
            bucket_end.push_back(b);
            b = std::max(b + 1, (b * 29 + 19) / 20);      // x 1.45
        }
    }
    bucket_end.push_back(std::max(m->num_views, 1));
    auto bucket_of = [&](JobSpec const& sp) {
        if (sp.reverse || bucket_end.size() == 1) return 0;
        int const v = std::max(sp.q_view, sp.c_view);
        int b = 0;
        while (b + 1 < static_cast<int>(bucket_end.size()) && v >= bucket_end[b]) ++b;
        return b;
    };
    std::vector<uint32_t> order(specs.size());
    for (size_t i = 0; i < specs.size(); ++i) order[i] = static_cast<uint32_t>(i);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (specs[a].reverse != specs[b].reverse) return !specs[a].reverse;
        int const ba = bucket_of(specs[a]), bb = bucket_of(specs[b]);
        if (ba != bb) return ba < bb;
        if (specs[a].c_view != specs[b].c_view) return specs[a].c_view < specs[b].c_view;
        return specs[a].c_n < specs[b].c_n;
    });
    std::vector<int64_t> bucket_items(bucket_end.size(), 0);   // items up to and including each bucket
    std::vector<int32_t> seg_first;   // first job of every (direction, candidate view, c_n) segment
    std::vector<int32_t> job_of(specs.size(), -1);
    int64_t rows = 0, items = 0, fwd_rows = 0;
    int fwd_jobs = 0;
    bool prev_reverse = false;
    int prev_bucket = 0;
    for (size_t oi = 0; oi < order.size(); ++oi) {
        size_t const i = order[oi];
        JobSpec const& s = specs[i];
        if (s.q_n <= 0 || s.c_n <= 0) continue;
        int const bk = bucket_of(s);
        if (jobs.empty() || prev_reverse != s.reverse || prev_bucket != bk || jobs.back().c_view != s.c_view ||
            jobs.back().c_n != s.c_n)
            seg_first.push_back(static_cast<int32_t>(jobs.size()));
        prev_reverse = s.reverse;
        prev_bucket = bk;
        ScanJob j;
        j.q_row = static_cast<int32_t>(k.off[s.q_view]);
        j.q_n = s.q_n;
        j.c_row = static_cast<int32_t>(k.off[s.c_view]);
        j.c_n = s.c_n;
        m->batch_max_cn = std::max(m->batch_max_cn, s.c_n);
        j.out_row = rows;
        j.item_start = static_cast<int32_t>(items);   // reverse jobs: no items in the filter pass
        j.c_view = s.c_view;
        out_row[i] = rows;
        job_of[i] = static_cast<int32_t>(jobs.size());
        rows += s.q_n;
        if (!s.reverse) {
            items += (s.q_n + kItemM - 1) / kItemM;
            for (size_t b = static_cast<size_t>(bk); b < bucket_items.size(); ++b) bucket_items[b] = items;
            fwd_rows = rows;
            fwd_jobs = static_cast<int>(jobs.size()) + 1;
        }
        jobs.push_back(j);
    }
    if (jobs.empty()) return OSFM_OK;
    if (items > INT32_MAX) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "too many work items in one batch");
    if (rows > static_cast<int64_t>(kSurvRowMask) || jobs.size() > static_cast<size_t>(kRowJobMask))
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "too many rows / jobs in one batch");
    int const njobs = static_cast<int>(jobs.size());
    int const nseg = static_cast<int>(seg_first.size());
    seg_first.push_back(njobs);
    std::vector<int32_t> rev_of(njobs, -1);
    for (size_t i = 0; i < specs.size(); ++i)
        if (specs[i].reverse && job_of[i] >= 0 && specs[i].partner >= 0 && job_of[specs[i].partner] >= 0)
            rev_of[job_of[specs[i].partner]] = job_of[i];
    ScanJob sentinel;
    memset(&sentinel, 0, sizeof sentinel);
    sentinel.out_row = rows;
    sentinel.item_start = static_cast<int32_t>(items);
    jobs.push_back(sentinel);

    CU_TRY(m, m->d_jobs.reserve(jobs.size()));
    CU_TRY(m, m->d_oneway.reserve(static_cast<size_t>(rows)));
    CU_TRY(m, m->d_cand.reserve(static_cast<size_t>(rows)));
    for (int i = 0; i < (k.is_signed ? 1 : 2); ++i) {
        CU_TRY(m, m->pass[i].list.reserve(static_cast<size_t>(rows)));
        CU_TRY(m, m->pass[i].cnt.reserve(static_cast<size_t>(njobs)));
        CU_TRY(m, cudaMemsetAsync(m->pass[i].cnt.p, 0, sizeof(int) * njobs, m->stream));
    }
    CU_TRY(m, m->d_seg_first.reserve(seg_first.size()));
    CU_TRY(m, cudaMemcpyAsync(m->d_seg_first.p, seg_first.data(), sizeof(int32_t) * seg_first.size(),
                              cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpyAsync(m->d_jobs.p, jobs.data(), sizeof(ScanJob) * jobs.size(),
                              cudaMemcpyHostToDevice, m->stream));
    bool const have_reverse = rows > fwd_rows;
    if (have_reverse) {
        CU_TRY(m, m->d_rev_of.reserve(static_cast<size_t>(njobs)));
        CU_TRY(m, cudaMemcpyAsync(m->d_rev_of.p, rev_of.data(), sizeof(int32_t) * njobs, cudaMemcpyHostToDevice, m->stream));
    }
    // pageable sources: the copies are staged before the calls return, the vectors may die.
    CU_TRY(m, cudaMemsetAsync(m->d_counters, 0, sizeof(unsigned long long), m->stream));      // [0]

    CU_TRY(m, m->d_rowres.reserve(static_cast<size_t>(rows)));
    float const sq_lowe = k.lowe * k.lowe;  // MATH_POW2 in float (matching.h:126)
    float const sq_dist = k.dist * k.dist;  // FLT_MAX^2 = +inf: never rejects (matching.h:127)

    // A batch with few work items (one pair of two large views: 782 items for 200 000 rows, 5.3 waves
    // of 148 CTAs rounded up to 6) is cut along the candidates as well: every forward job becomes
    // `split` jobs over consecutive column segments, the filter's (best, bound on the second best)
    // records of the segments are folded afterwards (merge_split_kernel; the fold is the one the
    // filter's own epilogue applies to its column halves), and everything downstream sees the
    // records of whole jobs.  `split` minimises waves x tiles per item.
    int split = 1;
    if (!staged && dump == nullptr && m->scan_mode == 0 && fwd_jobs > 0 && items < 8ll * m->num_sms) {
        int max_tiles = 0;
        for (int j = 0; j < fwd_jobs; ++j) max_tiles = std::max(max_tiles, (jobs[j].c_n + kBlockN - 1) / kBlockN);
        int64_t best = (items + m->num_sms - 1) / m->num_sms * max_tiles;
        for (int sp = 2; sp <= 16 && max_tiles / sp >= 8; ++sp) {
            // (+ 1 tile per item: the query tile's load and the item's tail are not free)
            int64_t const cost = (items * sp + m->num_sms - 1) / m->num_sms * ((max_tiles + sp - 1) / sp + 1);
            if (cost < best) { best = cost; split = sp; }
        }
        if (static_cast<int64_t>(fwd_jobs) * split > kRowJobMask || items * split > INT32_MAX) split = 1;
    }
    int64_t filter_items = items;
    if (split > 1) {
        std::vector<ScanJob> fjobs;
        fjobs.reserve(static_cast<size_t>(fwd_jobs) * split + 1);
        int64_t fitems = 0;
        for (int sgm = 0; sgm < split; ++sgm)
            for (int j = 0; j < fwd_jobs; ++j) {
                ScanJob f = jobs[j];
                int const tiles = (f.c_n + kBlockN - 1) / kBlockN;
                int const t0 = static_cast<int>(static_cast<int64_t>(tiles) * sgm / split);
                int const t1 = static_cast<int>(static_cast<int64_t>(tiles) * (sgm + 1) / split);
                f.c_row = jobs[j].c_row + t0 * kBlockN;
                f.c_n = std::min(jobs[j].c_n, t1 * kBlockN) - t0 * kBlockN;      // 0: a view with fewer tiles than segments
                f.out_row = static_cast<int64_t>(sgm) * fwd_rows + jobs[j].out_row;
                f.item_start = static_cast<int32_t>(fitems);
                if (f.c_n > 0) fitems += (f.q_n + kItemM - 1) / kItemM;
                fjobs.push_back(f);
            }
        ScanJob fs;
        memset(&fs, 0, sizeof fs);
        fs.out_row = static_cast<int64_t>(split) * fwd_rows;
        fs.item_start = static_cast<int32_t>(fitems);
        fjobs.push_back(fs);
        filter_items = fitems;
        CU_TRY(m, m->d_fjobs.reserve(fjobs.size()));
        CU_TRY(m, cudaMemcpyAsync(m->d_fjobs.p, fjobs.data(), sizeof(ScanJob) * fjobs.size(), cudaMemcpyHostToDevice, m->stream));
        CU_TRY(m, m->d_rowres_part.reserve(static_cast<size_t>(split) * fwd_rows));
        // (a segment without columns writes nothing: its records must read as "nothing seen", y = -1)
        CU_TRY(m, cudaMemsetAsync(m->d_rowres_part.p, 0xff, sizeof(int2) * static_cast<size_t>(split) * fwd_rows, m->stream));
        CU_TRY(m, m->d_item_job.reserve(static_cast<size_t>(fitems) + 1));
        fill_item_job_kernel<<<(static_cast<int>(fjobs.size()) - 1 + 255) / 256, 256, 0, m->stream>>>(
            m->d_fjobs.p, static_cast<int>(fjobs.size()) - 1, m->d_item_job.p);
        m->filter_jobs = m->d_fjobs.p;
        m->filter_rowres = m->d_rowres_part.p;
        bucket_items.back() = fitems;
    } else {
        CU_TRY(m, m->d_item_job.reserve(static_cast<size_t>(items) + 1));
        fill_item_job_kernel<<<(njobs + 255) / 256, 256, 0, m->stream>>>(m->d_jobs.p, njobs, m->d_item_job.p);
    }
    CU_TRY(m, cudaGetLastError());
    m->stats.kernel_launches++;
    CU_TRY(m, m->phases.mark(m->stream, kPhFilter));
    cudaError_t e = cudaSuccess;
    int64_t launched = 0;
    for (size_t b = 0; b < bucket_end.size(); ++b) {
        // (a lazy commit: the views of this bucket have arrived and have their norms)
        OS_TRY(ensure_views(m, bucket_end[b] - 1));
        int const first = static_cast<int>(launched), last = static_cast<int>(bucket_items[b]);
        if (last <= first) continue;
        switch (dump ? dump_mode : m->scan_mode) {
            case 1: e = launch_scan<1>(m, k, first, last, nullptr, 0); break;
            case 2: e = launch_scan<2>(m, k, first, last, m->d_oneway.p, 0); break;
            case 3: e = launch_scan<3>(m, k, first, last, dump, dump_ld); break;
            case 4: e = launch_scan<4>(m, k, first, last, dump, dump_ld); break;
            case 5: e = launch_scan<5>(m, k, first, last, dump, dump_ld); break;
            default: e = launch_scan<0>(m, k, first, last, nullptr, 0); break;
        }
        if (e != cudaSuccess) return cuda_fail(m, e, "scan_kernel launch");
        m->stats.kernel_launches++;
        launched = last;
    }
    m->filter_jobs = nullptr;
    m->filter_rowres = nullptr;
    if (split > 1) {
        int const mgrid = static_cast<int>((fwd_rows + 255) / 256);
        if (k.is_signed) merge_split_kernel<true><<<mgrid, 256, 0, m->stream>>>(m->d_rowres_part.p, split, fwd_rows, fwd_jobs, m->d_rowres.p);
        else             merge_split_kernel<false><<<mgrid, 256, 0, m->stream>>>(m->d_rowres_part.p, split, fwd_rows, fwd_jobs, m->d_rowres.p);
        CU_TRY(m, cudaGetLastError());
        m->stats.kernel_launches++;
    }
    CU_TRY(m, m->phases.mark(m->stream, kPhClassify));
    m->stats.scan_items += filter_items;
    if (dump) { CU_TRY(m, m->phases.mark(m->stream, -1)); return OSFM_OK; }   // debug dumps produce no results
    if (m->scan_mode != 0) {
        // timing modes: the filter wrote no row records, so nothing downstream may look at them;
        // every row reports "no match"
        CU_TRY(m, cudaMemsetAsync(m->d_oneway.p, 0xff, sizeof(int32_t) * rows, m->stream));
        CU_TRY(m, m->phases.mark(m->stream, -1));
        return OSFM_OK;
    }

    ClassifyParams cp;
    cp.jobs = m->d_jobs.p;
    cp.total_rows = fwd_rows;
    cp.rowres = m->d_rowres.p;
    cp.norm2 = k.d_norm2.p;
    cp.viewmax = k.d_viewmax.p;
    cp.pool = k.pool;
    cp.danger = k.d_danger.p;
    cp.danger_cnt = k.d_danger_cnt.p;
    cp.danger_floor = k.d_danger_floor.p;
    cp.oneway = m->d_oneway.p;
    cp.surv_list = m->pass[0].list.p;
    cp.surv_cnt = m->pass[0].cnt.p;
    cp.exact_list = m->pass[1].list.p;
    cp.exact_cnt = m->pass[1].cnt.p;
    cp.uncert_list = m->d_cand.p;
    cp.counters = m->d_counters;
    cp.uncert_len = m->d_counters + 0;
    cp.sq_lowe = sq_lowe;
    cp.sq_dist = sq_dist;

    PostParams pp;
    pp.pool = k.pool;
    pp.jobs = m->d_jobs.p;
    pp.njobs = njobs;
    pp.oneway = m->d_oneway.p;
    pp.sq_lowe = sq_lowe;
    pp.sq_dist = sq_dist;
    pp.slow_list = m->d_cand.p;
    pp.counters = m->d_counters;
    pp.slow_len = m->d_counters + 0;

    // The rows queued by classify_kernel / claim_kernel: CUDA-core replay of the signed rows
    // without certificate, RESOLVE pass, EXACT pass.
    auto second_passes = [&](bool verify, const ReverseSubset* subset) -> int {
        if (k.is_signed) {
            // rows without the norm certificate (adversarial input only): warp-per-row emulation
            slow_rows_kernel<true><<<m->num_sms * 2, 256, 0, m->stream>>>(pp);
            CU_TRY(m, cudaGetLastError());
            m->stats.kernel_launches++;
            OS_TRY((launch_second_pass<kPassResolve, true>(m, k, njobs, nseg, rows, pp, verify, subset)));
        } else {
            OS_TRY((launch_second_pass<kPassResolve, false>(m, k, njobs, nseg, rows, pp, verify, subset)));
            OS_TRY((launch_second_pass<kPassExact, false>(m, k, njobs, nseg, rows, pp, verify)));
        }
        return OSFM_OK;
    };

    if (fwd_rows > 0) {
        int const cgrid = static_cast<int>((fwd_rows + 255) / 256);
        if (k.is_signed) classify_kernel<true><<<cgrid, 256, 0, m->stream>>>(cp);
        else             classify_kernel<false><<<cgrid, 256, 0, m->stream>>>(cp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(m, e, "classify_kernel launch");
        m->stats.kernel_launches++;
        if (!k.is_signed) {
            certify_kernel<false><<<m->num_sms * 32, 256, 0, m->stream>>>(cp);   // latency-bound: many warps
            e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(m, e, "certify_kernel launch");
            m->stats.kernel_launches++;
        }
        CU_TRY(m, m->phases.mark(m->stream, kPhResolveFwd));
        OS_TRY(second_passes(false, nullptr));
    }

    if (have_reverse) {
        // the forward results are final: which rows of the other view do they claim?
        CU_TRY(m, m->phases.mark(m->stream, kPhClaim));
        int64_t const rev_rows = rows - fwd_rows;
        CU_TRY(m, cudaMemsetAsync(m->d_oneway.p + fwd_rows, 0xff, sizeof(int32_t) * rev_rows, m->stream));
        CU_TRY(m, cudaMemsetAsync(m->d_rowres.p + fwd_rows, 0xff, sizeof(int2) * rev_rows, m->stream));
        for (int i = 0; i < (k.is_signed ? 1 : 2); ++i)
            CU_TRY(m, cudaMemsetAsync(m->pass[i].cnt.p, 0, sizeof(int) * njobs, m->stream));
        CU_TRY(m, cudaMemsetAsync(m->d_counters + 11, 0, sizeof(unsigned long long), m->stream));
        CU_TRY(m, m->d_cand_rev.reserve(static_cast<size_t>(rev_rows)));
        bool const restricted = m->reverse_mode == 0;
        if (restricted) {
            CU_TRY(m, m->d_tau.reserve(static_cast<size_t>(njobs)));
            CU_TRY(m, cudaMemsetAsync(m->d_tau.p, 0x7f, sizeof(int32_t) * njobs, m->stream));   // "no claim yet"
        }
        // Every forward row whose result can be a match sits in one of the second passes' row lists
        // (the rest was rejected by the filter): claims are made list by list.
        ClaimParams cl;
        cl.jobs = m->d_jobs.p;
        cl.rev_of = m->d_rev_of.p;
        cl.pool = k.pool;
        cl.oneway = m->d_oneway.p;
        cl.rowres = m->d_rowres.p;
        cl.norm2 = k.d_norm2.p;
        cl.viewmax = k.d_viewmax.p;
        cl.surv_list = m->pass[0].list.p;
        cl.surv_cnt = m->pass[0].cnt.p;
        cl.uncert_list = m->d_cand_rev.p;
        cl.uncert_len = m->d_counters + 11;
        cl.counters = m->d_counters;
        cl.smin = restricted ? m->d_tau.p : nullptr;
        // The lists' lengths are only known on the device.  A thread's work is one long chain of
        // dependent loads, so the grid is sized to cover the list without striding where that is
        // cheap: a quarter of the forward rows is more than the filter lets through on any
        // descriptor set worth matching (a longer list is strided over), and every block beyond the
        // list costs ~0.6 ns of launch time.
        int const cgrid = static_cast<int>(std::min<int64_t>((fwd_rows + 1023) / 1024 + 1, static_cast<int64_t>(m->num_sms) * 256));
        int const cgrid_rare = std::min(cgrid, m->num_sms * 8);       // lists that are normally (nearly) empty
        cl.rows = m->pass[0].xrow_map.p;          // RESOLVE pass: the entries carry the rows' best similarity
        cl.n_int = m->pass[0].d_xmeta + 2;
        cl.n_ull = nullptr;
        if (k.is_signed) claim_kernel<true, true><<<cgrid, 256, 0, m->stream>>>(cl);
        else             claim_kernel<false, true><<<cgrid, 256, 0, m->stream>>>(cl);
        if (k.is_signed) {
            cl.rows = m->d_cand.p;                // rows replayed on CUDA cores (classify_kernel's flat list)
            cl.n_int = nullptr;
            cl.n_ull = m->d_counters + 0;
            claim_kernel<true, false><<<cgrid_rare, 256, 0, m->stream>>>(cl);
        } else {
            cl.rows = m->pass[1].xrow_map.p;      // EXACT pass
            cl.n_int = m->pass[1].d_xmeta + 2;
            claim_kernel<false, false><<<cgrid_rare, 256, 0, m->stream>>>(cl);
        }
        CU_TRY(m, cudaGetLastError());
        m->stats.kernel_launches += 2;
        cp.total_rows = rows;
        cp.uncert_list = m->d_cand_rev.p;
        cp.uncert_len = m->d_counters + 11;
        pp.slow_list = m->d_cand_rev.p;
        pp.slow_len = m->d_counters + 11;
        if (!k.is_signed) {
            certify_kernel<true><<<m->num_sms * 32, 256, 0, m->stream>>>(cp);
            CU_TRY(m, cudaGetLastError());
            m->stats.kernel_launches++;
        }
        ReverseSubset subset;
        if (restricted) {
            // candidate subsets of the reverse jobs (post_kernels.cuh, select_candidates_kernel)
            std::vector<int32_t> seg_rev;
            for (int j = fwd_jobs; j <= njobs; ++j) seg_rev.push_back(j);        // every reverse job alone
            CU_TRY(m, m->d_seg_first_rev.reserve(seg_rev.size()));
            CU_TRY(m, cudaMemcpyAsync(m->d_seg_first_rev.p, seg_rev.data(), sizeof(int32_t) * seg_rev.size(),
                                      cudaMemcpyHostToDevice, m->stream));
            CU_TRY(m, m->d_jobs_rev.reserve(jobs.size()));
            CU_TRY(m, cudaMemcpyAsync(m->d_jobs_rev.p, m->d_jobs.p, sizeof(ScanJob) * jobs.size(),
                                      cudaMemcpyDeviceToDevice, m->stream));
            CU_TRY(m, m->d_cand_pool.reserve(static_cast<size_t>(fwd_rows + kPadRows) * kRowBytes));
            CU_TRY(m, m->d_cand_map.reserve(static_cast<size_t>(fwd_rows)));
            if (m->cand_tmap_for != m->d_cand_pool.p || m->cand_tmap_rows != m->d_cand_pool.cap) {
                OS_TRY(encode_tmap(m, &m->cand_tmap, m->d_cand_pool.p, static_cast<int64_t>(m->d_cand_pool.cap / kRowBytes)));
                m->cand_tmap_for = m->d_cand_pool.p;
                m->cand_tmap_rows = m->d_cand_pool.cap;
            }
            SelectParams sel;
            sel.jobs = m->d_jobs.p;
            sel.jobs_rev = m->d_jobs_rev.p;
            sel.rev_of = m->d_rev_of.p;
            sel.fwd_jobs = fwd_jobs;
            sel.rowres = m->d_rowres.p;
            sel.norm2 = k.d_norm2.p;
            sel.viewmax = k.d_viewmax.p;
            sel.smin = m->d_tau.p;
            sel.sq_lowe = sq_lowe;
            sel.sq_dist = sq_dist;
            sel.surv_cnt = m->pass[0].cnt.p;
            sel.pool = k.pool;
            sel.cand_pool = m->d_cand_pool.p;
            sel.cand_map = m->d_cand_map.p;
            CU_TRY(m, m->d_cand_cnt.reserve(static_cast<size_t>(fwd_jobs)));
            CU_TRY(m, cudaMemsetAsync(m->d_cand_cnt.p, 0xff, sizeof(int32_t) * fwd_jobs, m->stream));   // "whole view"
            sel.cand_cnt = m->d_cand_cnt.p;
            sel.counters = m->d_counters;
            if (k.is_signed) select_candidates_kernel<true><<<fwd_jobs, 1024, 0, m->stream>>>(sel);
            else             select_candidates_kernel<false><<<fwd_jobs, 1024, 0, m->stream>>>(sel);
            CU_TRY(m, cudaGetLastError());
            dim3 const cgrid2(fwd_jobs, std::max(1, std::min(64, m->num_sms * 4 / fwd_jobs)));
            gather_candidates_kernel<<<cgrid2, 256, 0, m->stream>>>(m->d_jobs.p, m->d_cand_cnt.p, m->d_cand_map.p, k.pool,
                                                                   m->d_cand_pool.p);
            CU_TRY(m, cudaGetLastError());
            m->stats.kernel_launches += 2;
            subset.jobs = m->d_jobs_rev.p;
            subset.seg_first = m->d_seg_first_rev.p;
            subset.nseg = njobs - fwd_jobs;
            subset.tmap = &m->cand_tmap;
            subset.col_map = m->d_cand_map.p;
        }
        CU_TRY(m, m->phases.mark(m->stream, kPhResolveRev));
        OS_TRY(second_passes(true, restricted ? &subset : nullptr));
    }
    CU_TRY(m, m->phases.mark(m->stream, -1));

    return OSFM_OK;
}

// Adds up the phase times recorded so far; the stream must have been synchronised.
int collect_scan_time(osfm_matcher* m) {
    m->phases.collect();
    return OSFM_OK;
}

void publish_phase_times(osfm_matcher* m) {
    m->phases.collect();
    for (int i = 0; i < kPhCount; ++i) m->stats.last_phase_ms[i] = m->phases.acc[i];
    m->stats.last_phase_ms[kPhCount] = 0.0;
    m->stats.last_scan_ms = m->phases.acc[kPhFilter];
}

int check_view(osfm_matcher* m, int v) {
    if (v < 0 || v >= m->num_views) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "view id %d out of range [0,%d)", v, m->num_views);
    return OSFM_OK;
}

// Effective sizes of one pair, following exhaustive_matching.cc:123,134: a feature type
// takes part only if view_1 has descriptors of it.
PairPlan plan_pair(osfm_matcher* m, int v1, int v2, int limit /* <=0: none */, bool sift_only_lowres) {
    PairPlan p;
    p.v1 = v1; p.v2 = v2;
    for (int kd = 0; kd < 2; ++kd) {
        int a = m->kind[kd].n.empty() ? 0 : m->kind[kd].n[v1];
        int b = m->kind[kd].n.empty() ? 0 : m->kind[kd].n[v2];
        if (a <= 0) { a = 0; b = 0; }
        if (limit > 0) { a = std::min(a, limit); b = std::min(b, limit); }
        p.n1[kd] = a; p.n2[kd] = b;
    }
    if (sift_only_lowres) {
        // pairwise_match_lowres: SIFT if view_1 has SIFT, else SURF (exhaustive_matching.cc:153-177)
        if (p.n1[0] > 0) { p.n1[1] = 0; p.n2[1] = 0; }
    }
    p.len12 = p.n1[0] + p.n1[1];
    p.len21 = p.n2[0] + p.n2[1];
    p.out12 = p.out21 = 0;
    return p;
}

enum OutputMode { kFiltered = 0, kTwoway = 1 };

// Device part of a batch: fills m->d_dense (layout given by plans[].out12/out21, relative
// to the batch) and m->d_counts (one per plan; only in kFiltered mode).
int run_batch(osfm_matcher* m, const std::vector<PairPlan>& plans, int64_t dense_ints, OutputMode mode,
              int only_kind /* -1: both */) {
    int const np = static_cast<int>(plans.size());
    // (a lazy commit is completed by run_jobs, bucket by bucket of its filter pass)
    CU_TRY(m, m->d_dense.reserve(static_cast<size_t>(std::max<int64_t>(dense_ints, 1))));
    CU_TRY(m, m->d_counts.reserve(static_cast<size_t>(np)));
    CU_TRY(m, cudaMemsetAsync(m->d_counts.p, 0, sizeof(int32_t) * np, m->stream));

    std::vector<JobSpec> specs;
    std::vector<int64_t> out_row;
    std::vector<PairPart> parts;
    for (int kd = 0; kd < 2; ++kd) {
        if (only_kind >= 0 && kd != only_kind) continue;
        specs.clear();
        bool any = false;
        // kFiltered: only mutual matches survive, so the second direction of a pair is evaluated
        // for the rows the first direction's results claim (one tensor-core product per pair);
        // kTwoway returns both unfiltered directions, which takes both products.
        for (PairPlan const& p : plans) {
            int const fwd = static_cast<int>(specs.size());
            specs.push_back({p.v1, p.n1[kd], p.v2, p.n2[kd], false, -1});
            specs.push_back({p.v2, p.n2[kd], p.v1, p.n1[kd], mode == kFiltered && !m->both_directions, fwd});
            any = any || p.n1[kd] > 0 || p.n2[kd] > 0;
        }
        if (!any) continue;
        // the previous kind's scan time has to be read before its events are recorded again;
        // everything else the two kinds share is ordered by the stream.  No host round trip
        // after the last kind: the caller synchronises once, after queueing its own copies.
        OS_TRY(run_jobs(m, kd, specs, out_row));
        parts.clear();
        int max_n = 0;
        for (int i = 0; i < np; ++i) {
            PairPlan const& p = plans[i];
            if (p.n1[kd] == 0 && p.n2[kd] == 0) continue;
            PairPart pt;
            pt.in12 = out_row[2 * i];
            pt.in21 = out_row[2 * i + 1];
            // combine_results: SURF block follows the SIFT block; indices into the other
            // view's combined vector are shifted by that view's SIFT count (matching.cc:74-88).
            pt.out12 = p.out12 + (kd == 1 ? p.n1[0] : 0);
            pt.out21 = p.out21 + (kd == 1 ? p.n2[0] : 0);
            pt.n1 = p.n1[kd];
            pt.n2 = p.n2[kd];
            pt.add12 = (kd == 1 && mode == kFiltered) ? p.n2[0] : 0;
            pt.add21 = (kd == 1 && mode == kFiltered) ? p.n1[0] : 0;
            pt.pair = i;
            pt.pad = 0;
            parts.push_back(pt);
            max_n = std::max(max_n, std::max(pt.n1, pt.n2));
        }
        if (parts.empty()) continue;
        CU_TRY(m, m->d_parts.reserve(parts.size()));
        CU_TRY(m, cudaMemcpyAsync(m->d_parts.p, parts.data(), sizeof(PairPart) * parts.size(),
                                  cudaMemcpyHostToDevice, m->stream));
        dim3 const grid(static_cast<unsigned>(parts.size()),
                        static_cast<unsigned>((max_n + kMutualChunk - 1) / kMutualChunk));
        CU_TRY(m, m->phases.mark(m->stream, kPhMutual));
        if (mode == kFiltered)
            mutual_kernel<<<grid, 256, 0, m->stream>>>(m->d_parts.p, m->d_oneway.p, m->d_dense.p, m->d_counts.p);
        else
            copy_twoway_kernel<<<grid, 256, 0, m->stream>>>(m->d_parts.p, m->d_oneway.p, m->d_dense.p);
        CU_TRY(m, cudaGetLastError());
        CU_TRY(m, m->phases.mark(m->stream, -1));
        m->stats.kernel_launches++;
    }
    return OSFM_OK;
}

int read_counters(osfm_matcher* m) {
    unsigned long long c[16];
    CU_TRY(m, cudaMemcpy(c, m->d_counters, sizeof c, cudaMemcpyDeviceToHost));
    m->stats.reverse_candidate_rows = static_cast<int64_t>(c[9]);
    m->stats.reverse_restricted_pairs = static_cast<int64_t>(c[10]);
    m->stats.exact_rows = static_cast<int64_t>(c[5]);
    m->stats.exact_wide_rows = static_cast<int64_t>(c[13]);
    m->stats.last_scan_sm_cycles = static_cast<int64_t>(c[6]);
    m->stats.last_scan_ns = static_cast<int64_t>(c[7]);
    m->stats.candidate_rows = static_cast<int64_t>(c[1]);
    m->stats.self_check_failures = static_cast<int64_t>(c[2]);
    m->stats.slow_rows = static_cast<int64_t>(c[3]);
    m->stats.claimed_rows = static_cast<int64_t>(c[4]);
    if (c[2] != 0) return fail(m, OSFM_ERR_INTERNAL, "kernel self-check failed %llu times (filter / EXACT pass mismatch)", c[2]);
    return OSFM_OK;
}

// Splits `plans` into batches bounded by scratch memory; calls fn(first, last, dense_ints).
// phase_views > 0 (a lazy commit is pending): a batch also ends where the pairs start to need
// views of a later "phase".  Two phases: the pairs within the first quarter of the views, then
// the rest -- every extra batch costs about 0.25 ms of fixed work and host round trips
// (measured), and matching the first sixteenth of the pairs already lasts as long as the
// remaining copies (three finer phases hid more of the copies and lost more than that), so
// that the early pairs of a list in the reference's order (view_1 ascending) are matched while
// the later views are still being copied.
template <typename Fn>
int for_each_batch(std::vector<PairPlan>& plans, Fn fn, int phase_views = 0) {
    // OSFM_OVERLAP_PHASES="a" or "a,b" (tuning knob): phase limits at 1/a (and 1/b) of the views
    int d1 = 4, d2 = 0;
    if (const char* env = phase_views > 0 ? getenv("OSFM_OVERLAP_PHASES") : nullptr) {
        if (sscanf(env, "%d,%d", &d1, &d2) < 1 || d1 < 1) { d1 = 4; d2 = 0; }
    }
    auto phase_of = [phase_views, d1, d2](PairPlan const& p) {
        int const v = std::max(p.v1, p.v2);
        if (phase_views <= 0) return 0;
        if (v < (phase_views + d1 - 1) / d1) return 0;
        return d2 > 0 && v >= (phase_views + d2 - 1) / d2 ? 2 : 1;
    };
    size_t first = 0;
    while (first < plans.size()) {
        int64_t rows = 0, dense = 0;
        size_t last = first;
        int const phase = phase_of(plans[first]);
        while (last < plans.size()) {
            PairPlan& p = plans[last];
            int64_t const r = std::max<int64_t>(p.n1[0] + p.n2[0], p.n1[1] + p.n2[1]);
            int64_t const d = static_cast<int64_t>(p.len12) + p.len21;
            if (last > first && (rows + r > kMaxBatchRows || dense + d > kMaxBatchDense)) break;
            if (last > first && phase_of(p) > phase) break;
            p.out12 = dense;
            p.out21 = dense + p.len12;
            rows += r; dense += d;
            ++last;
        }
        int r = fn(first, last, dense);
        if (r != OSFM_OK) return r;
        first = last;
    }
    return OSFM_OK;
}

// Overlapped staging: the contract lets the caller release its descriptor buffers after the
// first call that returns results.  A call that touched only some views has waited (on the
// device) for those views alone, so every result-returning call ends by waiting for whatever is
// left of the copies (normally nothing: they finished long before the matching did).
int finish_staging(osfm_matcher* m) {
    if (!m->copies_in_flight) return OSFM_OK;
    CU_TRY(m, cudaStreamSynchronize(m->copy_stream));
    m->copies_in_flight = false;
    return OSFM_OK;
}

int require_committed(osfm_matcher* m) {
    if (!m->committed) return fail(m, OSFM_ERR_STATE, "matcher not committed (call osfm_match_commit first)");
    return OSFM_OK;
}

int64_t comparisons_of(const std::vector<PairPlan>& plans) {
    int64_t c = 0;
    for (PairPlan const& p : plans)
        for (int kd = 0; kd < 2; ++kd) c += static_cast<int64_t>(p.n1[kd]) * p.n2[kd];
    return c;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================

extern "C" {

void osfm_match_default_config(osfm_match_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->device = 0;
    cfg->sift_lowe_ratio = 0.8f;            // matching_base.h:27
    cfg->sift_distance_threshold = FLT_MAX;
    cfg->surf_lowe_ratio = 0.7f;            // matching_base.h:29
    cfg->surf_distance_threshold = FLT_MAX;
}

int osfm_match_abi_version(void) { return OSFM_MATCH_ABI_VERSION; }

const char* osfm_match_last_error(const osfm_matcher* m) {
    return m ? m->err.c_str() : "null handle";
}

int osfm_match_create(const osfm_match_config* cfg, osfm_matcher** out) {
    OSFM_TRY_BEGIN
    if (!out) return OSFM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    osfm_matcher* m = new (std::nothrow) osfm_matcher();
    if (!m) return OSFM_ERR_OUT_OF_MEMORY;
    *out = m;  // returned even on failure so that last_error() is readable
    memset(&m->stats, 0, sizeof m->stats);
    if (cfg) m->cfg = *cfg; else osfm_match_default_config(&m->cfg);
    m->device = m->cfg.device;

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(m, OSFM_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (m->device < 0 || m->device >= count)
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", m->device, count);
    cudaDeviceProp prop;
    CU_TRY(m, cudaGetDeviceProperties(&prop, m->device));
    if (prop.major != 10)
        return fail(m, OSFM_ERR_NO_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only",
                    m->device, prop.major, prop.minor);
    CU_TRY(m, cudaSetDevice(m->device));
    m->num_sms = prop.multiProcessorCount;
    CU_TRY(m, cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    CU_TRY(m, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : m->ev) CU_TRY(m, cudaEventCreate(&ev));
    CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&m->d_counters), 32 * sizeof(unsigned long long)));
    CU_TRY(m, cudaMemset(m->d_counters, 0, 32 * sizeof(unsigned long long)));
    for (auto& sp : m->pass) {
        CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&sp.d_xmeta), 4 * sizeof(int)));
        CU_TRY(m, cudaMemset(sp.d_xmeta, 0, 4 * sizeof(int)));
    }

    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(m, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return fail(m, OSFM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    m->encode = reinterpret_cast<EncodeTiledFn>(fn);

    // watchdog report buffer in mapped pinned memory
    // One report buffer per device for the life of the process, shared by every handle on that
    // device (the kernels find it through a per-device symbol): a handle that goes away must not
    // take the others' watchdog report with it.
    {
        static std::mutex hang_mu;
        static HangReport* per_device[64] = {};
        std::lock_guard<std::mutex> hang_lock(hang_mu);
        if (m->device < 64 && per_device[m->device] != nullptr) {
            m->hang_host = per_device[m->device];
        } else {
            CU_TRY(m, cudaHostAlloc(reinterpret_cast<void**>(&m->hang_host), sizeof(HangReport), cudaHostAllocMapped));
            memset(m->hang_host, 0, sizeof(HangReport));
            HangReport* dptr = nullptr;
            CU_TRY(m, cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), m->hang_host, 0));
            CU_TRY(m, cudaMemcpyToSymbol(g_hang_report, &dptr, sizeof dptr));
            if (m->device < 64) per_device[m->device] = m->hang_host;
        }
    }

    m->kind[0].is_signed = false; m->kind[0].dim = 128;
    m->kind[0].lowe = m->cfg.sift_lowe_ratio; m->kind[0].dist = m->cfg.sift_distance_threshold;
    m->kind[1].is_signed = true;  m->kind[1].dim = 64;
    m->kind[1].lowe = m->cfg.surf_lowe_ratio; m->kind[1].dist = m->cfg.surf_distance_threshold;
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

void osfm_match_destroy(osfm_matcher* m) {
    if (!m) return;
    for (ncclComm_t c : m->comms)
        if (c && m->nccl.CommDestroy) m->nccl.CommDestroy(c);
    m->comms.clear();
    for (osfm_matcher* p : m->peers) osfm_match_destroy(p);
    m->peers.clear();
    cudaSetDevice(m->device);
    if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
    if (m->stream) cudaStreamSynchronize(m->stream);
    reset_kind(m->kind[0], true);
    reset_kind(m->kind[1], true);
    m->d_jobs.release(); m->d_rowres.release(); m->d_oneway.release();
    m->d_cand.release(); m->d_cand_rev.release(); m->d_big.release();
    m->d_parts.release(); m->d_dense.release(); m->d_counts.release(); m->d_listoff.release(); m->d_list.release();
    m->tr_ints.release(); m->tr_table.release(); m->tr_meta.release(); m->tr_meta32.release();
    m->rs_xy.release(); m->rs_pos.release(); m->rs_samples.release(); m->rs_F.release(); m->rs_cnt.release();
    m->rs_out.release(); m->rs_stage1.release();
    for (int k = 0; k < 2; ++k) {
        if (m->rs_stage[k]) cudaFreeHost(m->rs_stage[k]);
        if (m->rs_stage_free[k]) cudaEventDestroy(m->rs_stage_free[k]);
        m->rs_stage[k] = nullptr; m->rs_stage_free[k] = nullptr;
    }
    m->rs_stage_ints = 0;
    m->d_ftmp.release();
    for (auto& ln : m->float_lanes) {
        for (int b = 0; b < 2; ++b) {
            if (ln.pinned[b]) cudaFreeHost(ln.pinned[b]);
            if (ln.dev[b]) cudaFree(ln.dev[b]);
            if (ln.done[b]) cudaEventDestroy(ln.done[b]);
        }
        if (ln.stream) cudaStreamDestroy(ln.stream);
    }
    m->float_lanes.clear();
    m->d_seg_first.release();
    m->d_rev_of.release(); m->d_item_job.release(); m->d_stash.release();
    m->d_tau.release(); m->d_jobs_rev.release(); m->d_seg_first_rev.release();
    m->d_cand_pool.release(); m->d_cand_map.release(); m->d_cand_cnt.release();
    m->d_fjobs.release(); m->d_rowres_part.release();
    m->d_fsplit.release(); m->d_fnorm.release(); m->d_ftop.release(); m->d_flist.release(); m->d_fparts.release();
    if (m->d_fmeta) cudaFree(m->d_fmeta);
    m->d_fmeta = nullptr;
    m->d_xw_x.release(); m->d_xw_xmax.release(); m->d_xw_off.release(); m->d_xw_moff.release(); m->d_xw_unit.release();
    if (m->d_xw_meta) cudaFree(m->d_xw_meta);
    m->d_xw_meta = nullptr;
    m->d_replay_flags.release();
    for (auto& sp : m->pass) sp.release();
    if (m->d_counters) cudaFree(m->d_counters);
    m->hang_host = nullptr;      // per-device, process-lifetime (see osfm_match_create)
    m->phases.release();
    if (m->cache_full.host) cudaFreeHost(m->cache_full.host);
    for (auto& ev : m->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : m->view_ev) if (ev) cudaEventDestroy(ev);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int osfm_match_create_multi(const osfm_match_config* cfg, const int* devices, int num_devices, osfm_matcher** out) {
    OSFM_TRY_BEGIN
    if (!out) return OSFM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!devices || num_devices < 1) return OSFM_ERR_INVALID_ARGUMENT;
    osfm_match_config c;
    if (cfg) c = *cfg; else osfm_match_default_config(&c);
    c.device = devices[0];
    int rc = osfm_match_create(&c, out);
    osfm_matcher* m = *out;
    if (rc != OSFM_OK || num_devices == 1) return rc;
    for (int i = 1; i < num_devices; ++i) {
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "device %d listed twice", devices[i]);
        c.device = devices[i];
        osfm_matcher* p = nullptr;
        rc = osfm_match_create(&c, &p);
        if (rc != OSFM_OK) {
            m->err = "device " + std::to_string(devices[i]) + ": " + (p ? p->err : std::string("allocation failed"));
            if (p) osfm_match_destroy(p);
            return rc;
        }
        m->peers.push_back(p);
    }
    std::string err;
    if (!m->nccl.load(err)) return fail(m, OSFM_ERR_CUDA, "%s", err.c_str());
    m->comms.assign(static_cast<size_t>(num_devices), nullptr);
    ncclResult_t const nr = m->nccl.CommInitAll(m->comms.data(), num_devices, devices);
    if (nr != ncclSuccess) {
        m->comms.clear();
        return fail(m, OSFM_ERR_CUDA, "ncclCommInitAll: %s", m->nccl.GetErrorString(nr));
    }
    CU_TRY(m, cudaSetDevice(m->device));
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

int osfm_match_num_devices(const osfm_matcher* m) {
    return m ? 1 + static_cast<int>(m->peers.size()) : OSFM_ERR_INVALID_ARGUMENT;
}

// Replicates the committed pools of the primary to every peer with one NCCL broadcast per kind
// (NVLink), then lets every peer finish its own commit (tensor map, norms).  The primary's stream
// has been synchronised: its pools are complete.
static int replicate_to_peers(osfm_matcher* m) {
    if (m->peers.empty()) return OSFM_OK;
    cudaEvent_t t0 = m->ev[2], t1 = m->ev[3];
    CU_TRY(m, cudaSetDevice(m->device));
    CU_TRY(m, cudaEventRecord(t0, m->stream));
    for (int kd = 0; kd < 2; ++kd) {
        KindPool& k = m->kind[kd];
        size_t const bytes = static_cast<size_t>(k.rows + kPadRows) * kRowBytes;
        for (osfm_matcher* p : m->peers) {
            std::lock_guard<std::recursive_mutex> lock(p->mu);
            KindPool& pk = p->kind[kd];
            CU_TRY(p, cudaSetDevice(p->device));
            CU_TRY(p, cudaStreamSynchronize(p->copy_stream));
            CU_TRY(p, cudaStreamSynchronize(p->stream));
            reset_kind(pk, false);
            pk.n = k.n;
            pk.off = k.off;
            pk.stage_off.assign(k.n.size(), -1);
            int const rc = arena_reserve(p, pk, k.rows + kPadRows);
            if (rc != OSFM_OK) { m->err = "device " + std::to_string(p->device) + ": " + p->err; return rc; }
            pk.pool = pk.arena;
            pk.owned = true;
            pk.rows = k.rows;
            pk.arena_used = k.rows;
            pk.lowe = k.lowe; pk.dist = k.dist;
        }
        ncclResult_t nr = m->nccl.GroupStart();
        if (nr == ncclSuccess) nr = m->nccl.Broadcast(k.pool, k.pool, bytes, ncclUint8, 0, m->comms[0], m->stream);
        for (size_t i = 0; i < m->peers.size() && nr == ncclSuccess; ++i)
            nr = m->nccl.Broadcast(m->peers[i]->kind[kd].pool, m->peers[i]->kind[kd].pool, bytes, ncclUint8, 0,
                                   m->comms[1 + i], m->peers[i]->stream);
        ncclResult_t const ne = m->nccl.GroupEnd();
        if (nr == ncclSuccess) nr = ne;
        if (nr != ncclSuccess) return fail(m, OSFM_ERR_CUDA, "ncclBroadcast: %s", m->nccl.GetErrorString(nr));
    }
    CU_TRY(m, cudaSetDevice(m->device));
    CU_TRY(m, cudaEventRecord(t1, m->stream));
    for (osfm_matcher* p : m->peers) {
        std::lock_guard<std::recursive_mutex> lock(p->mu);
        CU_TRY(p, cudaSetDevice(p->device));
        for (int kd = 0; kd < 2; ++kd) {
            int rc = make_tmap(p, p->kind[kd]);
            if (rc == OSFM_OK) rc = compute_norms(p, p->kind[kd]);
            if (rc != OSFM_OK) { m->err = "device " + std::to_string(p->device) + ": " + p->err; return rc; }
        }
        p->num_views = m->num_views;
        p->began = true;
        p->committed = true;
        p->overlap = false;
        p->lazy = false;
        p->cache_full.clear();
        p->cache_lowres.clear();
    }
    for (osfm_matcher* p : m->peers) {
        CU_TRY(p, cudaSetDevice(p->device));
        CU_TRY(p, cudaStreamSynchronize(p->stream));
    }
    CU_TRY(m, cudaSetDevice(m->device));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess) m->last_broadcast_ms = ms;
    return OSFM_OK;
}

static int begin_impl(osfm_matcher* m, int num_views, bool overlap) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->stream) return fail(m, OSFM_ERR_STATE, "handle was not created successfully");
    if (num_views < 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "num_views must be >= 0");
    CU_TRY(m, cudaSetDevice(m->device));
    CU_TRY(m, cudaStreamSynchronize(m->copy_stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    m->copies_in_flight = false;
    for (int kd = 0; kd < 2; ++kd) {
        reset_kind(m->kind[kd], false);
        m->kind[kd].n.assign(num_views, 0);
        m->kind[kd].off.assign(num_views, 0);
        m->kind[kd].stage_off.assign(num_views, -1);
    }
    m->num_views = num_views;
    m->began = true;
    m->committed = false;
    m->cache_full.clear();
    m->cache_lowres.clear();
    m->float_views.clear();
    m->overlap = overlap;
    m->lazy = false;
    if (overlap) {
        while (m->view_ev.size() < static_cast<size_t>(num_views)) {
            cudaEvent_t ev = nullptr;
            CU_TRY(m, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            m->view_ev.push_back(ev);
        }
        m->view_ev_set.assign(static_cast<size_t>(num_views), 0);
        m->view_ev_seq.assign(static_cast<size_t>(num_views), 0);
    }
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_begin(osfm_matcher* m, int num_views) { return begin_impl(m, num_views, false); }

int osfm_match_begin_overlapped(osfm_matcher* m, int num_views) { return begin_impl(m, num_views, true); }

constexpr int kFloatChunkRows = 2048;      // rows of a float view one staging step (run_float_views) moves

// Stages one view of one kind at the end of the arena.  All copies are asynchronous on the
// handle's stream; the source must stay valid until osfm_match_commit() returns.
static int stage_view(osfm_matcher* m, int kd, int view, const void* src, int n, int stride, bool is_float) {
    OSFM_TRY_BEGIN
    KindPool& k = m->kind[kd];
    if (k.stage_off[view] >= 0 || view < k.last_staged) k.in_order = false;  // re-staged or out of order
    k.n[view] = 0;
    k.stage_off[view] = -1;
    k.last_staged = std::max(k.last_staged, view);    // empty views count: ids must ascend for the lazy commit
    if (n <= 0) return OSFM_OK;
    if (!src) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "null descriptor pointer with n = %d", n);
    OS_TRY(arena_reserve(m, k, n));
    uint8_t* const d = k.arena + static_cast<size_t>(k.arena_used) * kRowBytes;
    cudaStream_t const cs = m->overlap ? m->copy_stream : m->stream;
    if (is_float) {
        if (m->overlap) return fail(m, OSFM_ERR_STATE, "overlapped staging takes quantised descriptors (set_view_q8)");
        if (stride < k.dim) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "stride %d < descriptor length %d", stride, k.dim);
        // moved at commit (run_float_views); the source stays valid until then
        // (in chunks of rows: the first transfer starts after a fraction of a view has been copied
        // into page-locked memory, and the lanes' buffers stay small)
        for (int r0 = 0; r0 < n; r0 += kFloatChunkRows)
            m->float_views.push_back({kd, k.arena_used + r0, static_cast<const float*>(src) + static_cast<size_t>(r0) * stride,
                                      std::min(kFloatChunkRows, n - r0), stride});
        (void)d;
    } else if (k.dim == kRowBytes) {
        CU_TRY(m, cudaMemcpyAsync(d, src, static_cast<size_t>(n) * kRowBytes, cudaMemcpyHostToDevice, cs));
    } else {
        // 64-byte rows are zero-padded to the 128-byte pool pitch
        CU_TRY(m, cudaMemsetAsync(d, 0, static_cast<size_t>(n) * kRowBytes, cs));
        CU_TRY(m, cudaMemcpy2DAsync(d, kRowBytes, src, k.dim, k.dim, n, cudaMemcpyHostToDevice, cs));
    }
    k.stage_off[view] = k.arena_used;
    k.arena_used += n;
    k.n[view] = n;
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_set_view_f32(osfm_matcher* m, int view_id, const float* sift, int n_sift, int sift_stride,
                            const float* surf, int n_surf, int surf_stride) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->began || m->committed) return fail(m, OSFM_ERR_STATE, "set_view outside begin/commit");
    OS_TRY(check_view(m, view_id));
    CU_TRY(m, cudaSetDevice(m->device));
    OS_TRY(stage_view(m, 0, view_id, sift, n_sift, sift_stride, true));
    OS_TRY(stage_view(m, 1, view_id, surf, n_surf, surf_stride, true));
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

static int set_view_q8_locked(osfm_matcher* m, int view_id, const uint8_t* sift, int n_sift,
                              const int8_t* surf, int n_surf) {
    OS_TRY(check_view(m, view_id));
    OS_TRY(stage_view(m, 0, view_id, sift, n_sift, 128, false));
    OS_TRY(stage_view(m, 1, view_id, surf, n_surf, 64, false));
    if (m->overlap) {
        CU_TRY(m, cudaEventRecord(m->view_ev[view_id], m->copy_stream));
        m->view_ev_set[view_id] = 1;
        m->view_ev_seq[view_id] = ++m->ev_seq;
        m->copies_in_flight = true;
    }
    return OSFM_OK;
}

int osfm_match_set_view_q8(osfm_matcher* m, int view_id, const uint8_t* sift, int n_sift,
                           const int8_t* surf, int n_surf) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->began || m->committed) return fail(m, OSFM_ERR_STATE, "set_view outside begin/commit");
    CU_TRY(m, cudaSetDevice(m->device));
    return set_view_q8_locked(m, view_id, sift, n_sift, surf, n_surf);
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_set_views_q8(osfm_matcher* m, int first_view, int count, const uint8_t* const* sift,
                            const int32_t* n_sift, const int8_t* const* surf, const int32_t* n_surf) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->began || m->committed) return fail(m, OSFM_ERR_STATE, "set_views outside begin/commit");
    if (count < 0 || (count > 0 && ((sift && !n_sift) || (surf && !n_surf))))
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad argument");
    CU_TRY(m, cudaSetDevice(m->device));
    for (int i = 0; i < count; ++i)
        OS_TRY(set_view_q8_locked(m, first_view + i, sift ? sift[i] : nullptr, sift ? n_sift[i] : 0,
                                  surf ? surf[i] : nullptr, surf ? n_surf[i] : 0));
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// Stages the float views recorded by osfm_match_set_view_f32: a few host threads, each with two
// page-locked buffers, two device buffers and a stream of its own: copy a view into page-locked
// memory, send it, quantise it into its place in the arena (convert_descriptor,
// exhaustive_matching.cc:18-39, on the device), take the next view while that runs.
static int run_float_views(osfm_matcher* m) {
    std::vector<osfm_matcher::FloatView>& views = m->float_views;
    if (views.empty()) return OSFM_OK;
    size_t max_floats = 0;
    for (auto const& v : views)
        max_floats = std::max(max_floats, static_cast<size_t>(v.n - 1) * v.stride + m->kind[v.kd].dim);
    unsigned const hw = std::max(1u, std::thread::hardware_concurrency());
    size_t const lanes = std::min<size_t>(std::min<size_t>(12, std::max(1u, hw * 3 / 4)), views.size());
    CU_TRY(m, cudaStreamSynchronize(m->stream));          // the arena may just have been moved
    if (m->float_lanes.size() < lanes) m->float_lanes.resize(lanes);
    for (size_t l = 0; l < lanes; ++l) {
        osfm_matcher::FloatLane& ln = m->float_lanes[l];
        if (!ln.stream) CU_TRY(m, cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b)
            if (!ln.done[b]) CU_TRY(m, cudaEventCreateWithFlags(&ln.done[b], cudaEventDisableTiming));
        if (ln.cap < max_floats) {
            for (int b = 0; b < 2; ++b) {
                if (ln.pinned[b]) cudaFreeHost(ln.pinned[b]);
                if (ln.dev[b]) cudaFree(ln.dev[b]);
                ln.pinned[b] = nullptr; ln.dev[b] = nullptr;
            }
            ln.cap = 0;
            size_t const want = max_floats + max_floats / 8;
            for (int b = 0; b < 2; ++b) {
                CU_TRY(m, cudaHostAlloc(reinterpret_cast<void**>(&ln.pinned[b]), want * sizeof(float), cudaHostAllocDefault));
                CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&ln.dev[b]), want * sizeof(float)));
            }
            ln.cap = want;
        }
    }
    std::atomic<size_t> next(0);
    std::vector<cudaError_t> err(lanes, cudaSuccess);
    auto work = [&](size_t l) {
        osfm_matcher::FloatLane& ln = m->float_lanes[l];
        cudaError_t e = cudaSetDevice(m->device);
        bool used[2] = {false, false};
        int b = 0;
        for (size_t i = next.fetch_add(1); e == cudaSuccess && i < views.size(); i = next.fetch_add(1), b ^= 1) {
            osfm_matcher::FloatView const& v = views[i];
            KindPool const& k = m->kind[v.kd];
            size_t const count = static_cast<size_t>(v.n - 1) * v.stride + k.dim;
            if (used[b]) e = cudaEventSynchronize(ln.done[b]);       // this pair of buffers is free again
            if (e != cudaSuccess) break;
            memcpy(ln.pinned[b], v.src, count * sizeof(float));
            e = cudaMemcpyAsync(ln.dev[b], ln.pinned[b], count * sizeof(float), cudaMemcpyHostToDevice, ln.stream);
            if (e != cudaSuccess) break;
            uint8_t* const dst = k.arena + static_cast<size_t>(v.arena_row) * kRowBytes;
            int64_t const total = static_cast<int64_t>(v.n) * kRowBytes;
            int const grid = static_cast<int>((total + 255) / 256);
            if (k.is_signed) quantize_kernel<true><<<grid, 256, 0, ln.stream>>>(ln.dev[b], v.n, k.dim, v.stride, dst);
            else             quantize_kernel<false><<<grid, 256, 0, ln.stream>>>(ln.dev[b], v.n, k.dim, v.stride, dst);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaEventRecord(ln.done[b], ln.stream);
            used[b] = true;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ln.stream);
        err[l] = e;
    };
    std::vector<std::thread> threads;
    try {
        for (size_t l = 1; l < lanes; ++l) threads.emplace_back(work, l);
    } catch (...) {
        // fewer threads than planned: the lanes that did start (and this one) drain the queue
    }
    work(0);
    for (std::thread& t : threads) t.join();
    m->stats.kernel_launches += static_cast<int64_t>(views.size());
    views.clear();
    CU_TRY(m, cudaSetDevice(m->device));
    for (cudaError_t e : err)
        if (e != cudaSuccess) return cuda_fail(m, e, "staging float descriptors");
    return OSFM_OK;
}

int osfm_match_commit(osfm_matcher* m) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->began || m->committed) return fail(m, OSFM_ERR_STATE, "commit outside begin/commit");
    CU_TRY(m, cudaSetDevice(m->device));
    // Overlapped staging stays lazy only if the arena already is the pool (views staged in
    // ascending order); otherwise everything is waited for here, as in the plain commit.
    // (a multi-device matcher broadcasts the complete pool at commit: nothing stays lazy)
    OS_TRY(run_float_views(m));
    bool const lazy = m->overlap && m->kind[0].in_order && m->kind[1].in_order && m->peers.empty();
    if (m->overlap && !lazy) CU_TRY(m, cudaStreamSynchronize(m->copy_stream));
    for (int kd = 0; kd < 2; ++kd) {
        KindPool& k = m->kind[kd];
        int64_t const rows = k.arena_used;
        if (rows + kPadRows > INT32_MAX) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "descriptor pool exceeds 2^31 rows");
        k.rows = rows;
        if (k.in_order) {
            // the arena already is the pool: view v sits where it was staged
            OS_TRY(arena_reserve(m, k, kPadRows));
            k.pool = k.arena;
            k.owned = true;
            int64_t next = 0;
            for (int v = 0; v < m->num_views; ++v) {
                k.off[v] = k.stage_off[v] >= 0 ? k.stage_off[v] : next;
                next = k.off[v] + k.n[v];
            }
        } else {
            size_t const bytes = static_cast<size_t>(rows + kPadRows) * kRowBytes;
            CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&k.pool), bytes));
            k.owned = true;
            int64_t next = 0;
            for (int v = 0; v < m->num_views; ++v) {
                k.off[v] = next;
                if (k.n[v] > 0)
                    CU_TRY(m, cudaMemcpyAsync(k.pool + static_cast<size_t>(next) * kRowBytes,
                                              k.arena + static_cast<size_t>(k.stage_off[v]) * kRowBytes,
                                              static_cast<size_t>(k.n[v]) * kRowBytes, cudaMemcpyDeviceToDevice, m->stream));
                next += k.n[v];
            }
        }
        CU_TRY(m, cudaMemsetAsync(k.pool + static_cast<size_t>(rows) * kRowBytes, 0,
                                  static_cast<size_t>(kPadRows) * kRowBytes, m->stream));
        OS_TRY(make_tmap(m, k));
        if (lazy) OS_TRY(norms_prepare(m, k));      // the norms follow view by view (ensure_views)
        else      OS_TRY(compute_norms(m, k));
    }
    m->lazy = lazy;
    if (lazy) {
        // nothing to wait for: the caller keeps its buffers until the first call that returns
        // results (or osfm_match_wait_staged) has returned
        OS_TRY(ensure_views(m, -1));                // clears `lazy` at once if there is nothing to do
    } else {
        CU_TRY(m, cudaStreamSynchronize(m->stream));  // from here on the caller may free its buffers
    }
    m->copies_in_flight = m->copies_in_flight && lazy;
    OS_TRY(replicate_to_peers(m));
    m->committed = true;
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_wait_staged(osfm_matcher* m) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->stream) return fail(m, OSFM_ERR_STATE, "handle was not created successfully");
    CU_TRY(m, cudaSetDevice(m->device));
    if (m->committed) OS_TRY(ensure_views(m, m->num_views - 1));
    CU_TRY(m, cudaStreamSynchronize(m->copy_stream));
    m->copies_in_flight = false;
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_commit_device(osfm_matcher* m, int num_views,
                             const void* sift_pool, const int64_t* sift_row_offset, const int32_t* n_sift,
                             int64_t sift_pool_rows,
                             const void* surf_pool, const int64_t* surf_row_offset, const int32_t* n_surf,
                             int64_t surf_pool_rows) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->stream) return fail(m, OSFM_ERR_STATE, "handle was not created successfully");
    if (num_views < 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "num_views must be >= 0");
    CU_TRY(m, cudaSetDevice(m->device));
    CU_TRY(m, cudaStreamSynchronize(m->copy_stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    m->overlap = false;
    m->lazy = false;
    const void* pools[2] = {sift_pool, surf_pool};
    const int64_t* offs[2] = {sift_row_offset, surf_row_offset};
    const int32_t* ns[2] = {n_sift, n_surf};
    int64_t prow[2] = {sift_pool_rows, surf_pool_rows};
    for (int kd = 0; kd < 2; ++kd) {
        KindPool& k = m->kind[kd];
        reset_kind(k, false);
        k.n.assign(num_views, 0);
        k.off.assign(num_views, 0);
        k.stage_off.assign(num_views, -1);
        if (!pools[kd]) {
            // empty pool: the (possibly empty) arena provides the padding rows for the tensor map
            OS_TRY(arena_reserve(m, k, kPadRows));
            CU_TRY(m, cudaMemsetAsync(k.arena, 0, static_cast<size_t>(kPadRows) * kRowBytes, m->stream));
            k.pool = k.arena;
            k.owned = true;
            k.rows = 0;
        } else {
            if ((reinterpret_cast<uintptr_t>(pools[kd]) & 127u) != 0)
                return fail(m, OSFM_ERR_INVALID_ARGUMENT, "device pool must be 128-byte aligned");
            if (!offs[kd] || !ns[kd]) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "null offset / size array");
            if (prow[kd] + kPadRows > INT32_MAX) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "descriptor pool exceeds 2^31 rows");
            for (int v = 0; v < num_views; ++v) {
                if (ns[kd][v] < 0 || offs[kd][v] < 0 || offs[kd][v] + ns[kd][v] > prow[kd])
                    return fail(m, OSFM_ERR_INVALID_ARGUMENT, "view %d lies outside the pool", v);
                k.n[v] = ns[kd][v];
                k.off[v] = offs[kd][v];
            }
            k.pool = const_cast<uint8_t*>(static_cast<const uint8_t*>(pools[kd]));
            k.owned = false;
            k.rows = prow[kd];
        }
        OS_TRY(make_tmap(m, k));
        OS_TRY(compute_norms(m, k));
    }
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    m->num_views = num_views;
    m->began = true;
    m->cache_full.clear();
    m->cache_lowres.clear();
    OS_TRY(replicate_to_peers(m));
    m->committed = true;
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_num_views(const osfm_matcher* m) { return m ? m->num_views : OSFM_ERR_INVALID_ARGUMENT; }

int osfm_match_view_size(const osfm_matcher* m, int view_id, int* n_sift, int* n_surf) {
    if (!m || view_id < 0 || view_id >= m->num_views) return OSFM_ERR_INVALID_ARGUMENT;
    if (n_sift) *n_sift = m->kind[0].n[view_id];
    if (n_surf) *n_surf = m->kind[1].n[view_id];
    return OSFM_OK;
}

// ---- batched dense --------------------------------------------------------------------

static int build_plans(osfm_matcher* m, const int32_t* pairs, int npairs, int limit, bool lowres,
                       std::vector<PairPlan>& plans) {
    if (npairs < 0 || (npairs > 0 && !pairs)) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad pair list");
    plans.clear();
    plans.reserve(npairs);
    for (int i = 0; i < npairs; ++i) {
        OS_TRY(check_view(m, pairs[2 * i]));
        OS_TRY(check_view(m, pairs[2 * i + 1]));
        plans.push_back(plan_pair(m, pairs[2 * i], pairs[2 * i + 1], limit, lowres));
    }
    return OSFM_OK;
}

int64_t osfm_match_pairs_result_size(osfm_matcher* m, const int32_t* pairs, int npairs) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (require_committed(m) != OSFM_OK) return OSFM_ERR_STATE;
    std::vector<PairPlan> plans;
    int r = build_plans(m, pairs, npairs, 0, false, plans);
    if (r != OSFM_OK) return r;
    int64_t total = 0;
    for (PairPlan const& p : plans) total += static_cast<int64_t>(p.len12) + p.len21;
    return total;
    OSFM_TRY_END(m, -6)
}

static int match_pairs_dense(osfm_matcher* m, std::vector<PairPlan>& plans, OutputMode mode, int only_kind,
                             int32_t* matches, int64_t* offsets, int32_t* n_consistent) {
    CU_TRY(m, cudaSetDevice(m->device));
    m->phases.reset();
    CU_TRY(m, cudaEventRecord(m->ev[2], m->stream));
    int64_t host_base = 0;
    int r = for_each_batch(plans, [&](size_t first, size_t last, int64_t dense) -> int {
        std::vector<PairPlan> sub(plans.begin() + first, plans.begin() + last);
        OS_TRY(run_batch(m, sub, dense, mode, only_kind));
        if (dense > 0 && matches)
            CU_TRY(m, cudaMemcpyAsync(matches + host_base, m->d_dense.p, sizeof(int32_t) * dense,
                                      cudaMemcpyDeviceToHost, m->stream));
        if (n_consistent)
            CU_TRY(m, cudaMemcpyAsync(n_consistent + first, m->d_counts.p, sizeof(int32_t) * (last - first),
                                      cudaMemcpyDeviceToHost, m->stream));
        CU_TRY(m, cudaStreamSynchronize(m->stream));
        OS_TRY(collect_scan_time(m));
        if (offsets)
            for (size_t i = first; i < last; ++i) {
                offsets[2 * i] = host_base + plans[i].out12;
                offsets[2 * i + 1] = host_base + plans[i].out21;
            }
        host_base += dense;
        return OSFM_OK;
    });
    if (r != OSFM_OK) return r;
    if (offsets) offsets[2 * plans.size()] = host_base;
    CU_TRY(m, cudaEventRecord(m->ev[3], m->stream));
    CU_TRY(m, cudaEventSynchronize(m->ev[3]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, m->ev[2], m->ev[3]);
    m->stats.last_total_ms = ms;
    publish_phase_times(m);
    m->stats.last_comparisons = comparisons_of(plans);
    OS_TRY(finish_staging(m));
    return read_counters(m);
}

static int dense_dispatch(osfm_matcher* m, std::vector<PairPlan>& plans, OutputMode mode, int only_kind,
                          int32_t* matches, int64_t* offsets, int32_t* n_consistent);

int osfm_match_pairs(osfm_matcher* m, const int32_t* pairs, int npairs, int32_t* matches, int64_t* offsets,
                     int32_t* n_consistent) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    OS_TRY(require_committed(m));
    std::vector<PairPlan> plans;
    OS_TRY(build_plans(m, pairs, npairs, 0, false, plans));
    return dense_dispatch(m, plans, kFiltered, -1, matches, offsets, n_consistent);
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// The window of pairs that follows (view_1, view_2) in the reference's enumeration
// (bundler_matching.cc:92-93: i -> view_1 = (int)(0.5 + sqrt(0.25 + 2 i)), view_2 = i - view_1 (view_1 - 1) / 2,
// i.e. view_1 ascending, view_2 = 0 .. view_1 - 1), at most `limit` pairs.
static void lookahead_window(int num_views, int v1, int v2, int limit, std::vector<int32_t>& pairs) {
    pairs.clear();
    for (int a = v1; a < num_views && static_cast<int>(pairs.size()) < 2 * limit; ++a)
        for (int b = (a == v1 ? v2 : 0); b < a && static_cast<int>(pairs.size()) < 2 * limit; ++b) {
            pairs.push_back(a);
            pairs.push_back(b);
        }
}

static int64_t flat_pair_index(int v1, int v2) { return static_cast<int64_t>(v1) * (v1 - 1) / 2 + v2; }

int osfm_match_set_lookahead(osfm_matcher* m, int max_pairs) {
    OSFM_TRY_BEGIN
    if (!m || max_pairs < 0) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    m->lookahead = max_pairs;
    m->cache_full.clear();
    m->cache_lowres.clear();
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// Serves osfm_match_pair from the look-ahead cache, filling it first if (view_1, view_2) is not in it.
static int compact_to_host(osfm_matcher* m, const int32_t* pairs, int npairs, int32_t* match_ij, int64_t capacity_ij,
                           int64_t* list_offset, bool sift_only, int min_count, int32_t* counts_out);

static int pair_from_lookahead(osfm_matcher* m, int v1, int v2, int32_t* matches_1_2, int* len_1_2,
                               int32_t* matches_2_1, int* len_2_1, int* n_consistent) {
    // The cache holds the window's correspondence lists ((i, j) pairs in the combined SIFT + SURF
    // index space, an eighth of the dense vectors' bytes to bring back), computed by the same
    // sharded call as osfm_match_pairs_compact; a pair's dense Matching::Result is rebuilt from its
    // list when it is asked for: after the mutual filter the two vectors hold exactly those pairs.
    osfm_matcher::PairCache& c = m->cache_full;
    int64_t const idx = flat_pair_index(v1, v2);
    if (c.first < 0 || idx < c.first || idx >= c.first + c.count) {
        constexpr int64_t kMaxCacheEntries = 1ll << 27;     // 1 GiB of pinned host memory at most
        std::vector<int32_t> pairs;
        lookahead_window(m->num_views, v1, v2, m->lookahead, pairs);
        std::vector<PairPlan> plans;
        OS_TRY(build_plans(m, pairs.data(), static_cast<int>(pairs.size() / 2), 0, false, plans));
        int64_t total = 0;
        size_t keep = 0;
        for (; keep < plans.size(); ++keep) {
            PairPlan const& p = plans[keep];
            int64_t const d = std::min(p.n1[0], p.n2[0]) + std::min(p.n1[1], p.n2[1]);   // a list cannot be longer
            if (keep > 0 && total + d > kMaxCacheEntries) break;
            total += d;
        }
        plans.resize(keep);
        c.clear();
        if (2 * static_cast<size_t>(total) + 2 > c.host_cap) {
            CU_TRY(m, cudaSetDevice(m->device));
            if (c.host) cudaFreeHost(c.host);
            c.host = nullptr; c.host_cap = 0;
            size_t const want = 2 * static_cast<size_t>(total) + static_cast<size_t>(total) / 4 + 1024;
            CU_TRY(m, cudaHostAlloc(reinterpret_cast<void**>(&c.host), want * sizeof(int32_t), cudaHostAllocDefault));
            c.host_cap = want;
        }
        c.offsets.assign(plans.size() + 1, 0);
        c.counts.assign(plans.size(), 0);
        c.lens.resize(2 * plans.size());
        for (size_t p = 0; p < plans.size(); ++p) { c.lens[2 * p] = plans[p].len12; c.lens[2 * p + 1] = plans[p].len21; }
        OS_TRY(compact_to_host(m, pairs.data(), static_cast<int>(plans.size()), c.host, static_cast<int64_t>(c.host_cap / 2),
                               c.offsets.data(), false, 0, c.counts.data()));
        c.first = idx;
        c.count = static_cast<int>(plans.size());
    }
    size_t const k = static_cast<size_t>(idx - c.first);
    int const l12 = c.lens[2 * k], l21 = c.lens[2 * k + 1];
    if (matches_1_2) std::fill(matches_1_2, matches_1_2 + l12, -1);
    if (matches_2_1) std::fill(matches_2_1, matches_2_1 + l21, -1);
    for (int64_t e = c.offsets[k]; e < c.offsets[k + 1]; ++e) {
        int32_t const i = c.host[2 * e], jj = c.host[2 * e + 1];
        if (matches_1_2) matches_1_2[i] = jj;
        if (matches_2_1) matches_2_1[jj] = i;
    }
    if (len_1_2) *len_1_2 = l12;
    if (len_2_1) *len_2_1 = l21;
    if (n_consistent) *n_consistent = c.counts[k];
    return OSFM_OK;
}

int osfm_match_pair(osfm_matcher* m, int view_1_id, int view_2_id, int32_t* matches_1_2, int* len_1_2,
                    int32_t* matches_2_1, int* len_2_1, int* n_consistent) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    OS_TRY(require_committed(m));
    if (m->lookahead > 1 && view_1_id > view_2_id && view_2_id >= 0 && view_1_id < m->num_views)
        return pair_from_lookahead(m, view_1_id, view_2_id, matches_1_2, len_1_2, matches_2_1, len_2_1, n_consistent);
    int32_t pr[2] = {view_1_id, view_2_id};
    std::vector<PairPlan> plans;
    OS_TRY(build_plans(m, pr, 1, 0, false, plans));
    std::vector<int32_t> buf(static_cast<size_t>(plans[0].len12) + plans[0].len21 + 1);
    int64_t offs[3];
    int32_t cnt = 0;
    OS_TRY(match_pairs_dense(m, plans, kFiltered, -1, buf.data(), offs, &cnt));
    if (matches_1_2 && plans[0].len12 > 0) memcpy(matches_1_2, buf.data() + offs[0], sizeof(int32_t) * plans[0].len12);
    if (matches_2_1 && plans[0].len21 > 0) memcpy(matches_2_1, buf.data() + offs[1], sizeof(int32_t) * plans[0].len21);
    if (len_1_2) *len_1_2 = plans[0].len12;
    if (len_2_1) *len_2_1 = plans[0].len21;
    if (n_consistent) *n_consistent = cnt;
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_pair_twoway(osfm_matcher* m, int kind, int view_1_id, int view_2_id, int32_t* matches_1_2,
                           int32_t* matches_2_1) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    OS_TRY(require_committed(m));
    if (kind != OSFM_KIND_SIFT_U8 && kind != OSFM_KIND_SURF_S8) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "unknown kind %d", kind);
    OS_TRY(check_view(m, view_1_id));
    OS_TRY(check_view(m, view_2_id));
    // twoway_match itself has no "view_1 must be non-empty" rule: both vectors always
    // have the full set sizes (matching.h:121-122).
    PairPlan p;
    p.v1 = view_1_id; p.v2 = view_2_id;
    for (int kd = 0; kd < 2; ++kd) { p.n1[kd] = 0; p.n2[kd] = 0; }
    p.n1[kind] = m->kind[kind].n[view_1_id];
    p.n2[kind] = m->kind[kind].n[view_2_id];
    p.len12 = p.n1[kind]; p.len21 = p.n2[kind];
    p.out12 = p.out21 = 0;
    std::vector<PairPlan> plans(1, p);
    std::vector<int32_t> buf(static_cast<size_t>(p.len12) + p.len21 + 1);
    int64_t offs[3];
    OS_TRY(match_pairs_dense(m, plans, kTwoway, kind, buf.data(), offs, nullptr));
    if (matches_1_2 && p.len12 > 0) memcpy(matches_1_2, buf.data() + offs[0], sizeof(int32_t) * p.len12);
    if (matches_2_1 && p.len21 > 0) memcpy(matches_2_1, buf.data() + offs[1], sizeof(int32_t) * p.len21);
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_pair_lowres(osfm_matcher* m, int view_1_id, int view_2_id, size_t num_features, int* n_consistent) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    OS_TRY(require_committed(m));
    if (num_features == 0 || num_features > INT32_MAX) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad num_features");
    if (m->lookahead > 1 && view_1_id > view_2_id && view_2_id >= 0 && view_1_id < m->num_views) {
        // the low-res gate is cheap (500 x 500 per pair): a window of 16 x the look-ahead
        osfm_matcher::PairCache& c = m->cache_lowres;
        int64_t const idx = flat_pair_index(view_1_id, view_2_id);
        if (c.first < 0 || c.num_features != static_cast<int>(num_features) || idx < c.first || idx >= c.first + c.count) {
            std::vector<int32_t> pairs;
            int const limit = static_cast<int>(std::min<int64_t>(16ll * m->lookahead, 1 << 20));
            lookahead_window(m->num_views, view_1_id, view_2_id, limit, pairs);
            std::vector<PairPlan> plans;
            OS_TRY(build_plans(m, pairs.data(), static_cast<int>(pairs.size() / 2), static_cast<int>(num_features), true, plans));
            c.clear();
            c.counts.assign(plans.size(), 0);
            OS_TRY(dense_dispatch(m, plans, kFiltered, -1, nullptr, nullptr, c.counts.data()));
            c.first = idx;
            c.count = static_cast<int>(plans.size());
            c.num_features = static_cast<int>(num_features);
        }
        if (n_consistent) *n_consistent = c.counts[static_cast<size_t>(idx - c.first)];
        return OSFM_OK;
    }
    int32_t pr[2] = {view_1_id, view_2_id};
    std::vector<PairPlan> plans;
    OS_TRY(build_plans(m, pr, 1, static_cast<int>(num_features), true, plans));
    int32_t cnt = 0;
    // count_consistent_matches of the unfiltered two-way result equals the number of
    // survivors of the mutual filter (matching.cc:39-47 vs :19-36).
    OS_TRY(match_pairs_dense(m, plans, kFiltered, -1, nullptr, nullptr, &cnt));
    if (n_consistent) *n_consistent = cnt;
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// ---- float path ---------------------------------------------------------------------------

// The tensor-core filter of the float path (float_tc_kernels.cuh) for one pair: d1 / d2 are the
// zero-padded n x 128 float sets on the device.  Afterwards (in stream order) m->d_oneway holds the
// result of every row the filter decided, m->d_flist / m->d_fmeta the rows it left, per direction.
static int float_filter(osfm_matcher* m, const float* d1, int n1, const float* d2, int n2, float sq_lowe, float sq_dist) {
    int const n1p = (n1 + kFtN - 1) / kFtN * kFtN, n2p = (n2 + kFtN - 1) / kFtN * kFtN;
    size_t const f1 = static_cast<size_t>(n1p) * kFDim, f2 = static_cast<size_t>(n2p) * kFDim;
    CU_TRY(m, m->d_fsplit.reserve(2 * (f1 + f2)));
    CU_TRY(m, m->d_fnorm.reserve(static_cast<size_t>(n1) + n2));
    CU_TRY(m, m->d_ftop.reserve(static_cast<size_t>(n1) + n2));
    CU_TRY(m, m->d_flist.reserve(static_cast<size_t>(n1) + n2));
    if (!m->d_fmeta) CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&m->d_fmeta), 4 * sizeof(int)));
    CU_TRY(m, cudaMemsetAsync(m->d_fmeta, 0, 4 * sizeof(int), m->stream));
    float* const hi1 = m->d_fsplit.p;
    float* const lo1 = hi1 + f1;
    float* const hi2 = lo1 + f1;
    float* const lo2 = hi2 + f2;
    float_split_kernel<<<(n1p * 32 + 255) / 256, 256, 0, m->stream>>>(d1, n1, n1p, hi1, lo1, m->d_fnorm.p, m->d_fmeta + 2);
    float_split_kernel<<<(n2p * 32 + 255) / 256, 256, 0, m->stream>>>(d2, n2, n2p, hi2, lo2, m->d_fnorm.p + n1, m->d_fmeta + 3);
    CU_TRY(m, cudaGetLastError());
    CUtensorMap t_hi1, t_lo1, t_hi2, t_lo2;
    OS_TRY(encode_tmap_f32(m, &t_hi1, hi1, n1p));
    OS_TRY(encode_tmap_f32(m, &t_lo1, lo1, n1p));
    OS_TRY(encode_tmap_f32(m, &t_hi2, hi2, n2p));
    OS_TRY(encode_tmap_f32(m, &t_lo2, lo2, n2p));
    CU_TRY(m, cudaFuncSetAttribute(float_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtSmemBytes));
    int const items = (n1 + kFtM - 1) / kFtM + (n2 + kFtM - 1) / kFtM;
    float_filter_kernel<<<std::min(m->num_sms, items), kFtThreads, kFtSmemBytes, m->stream>>>(
        t_hi1, t_lo1, t_hi2, t_lo2, n1, n2, m->d_ftop.p);
    CU_TRY(m, cudaGetLastError());
    float_decide_kernel<<<(n1 + n2 + 255) / 256, 256, 0, m->stream>>>(
        m->d_ftop.p, m->d_fnorm.p, m->d_fnorm.p + n1, m->d_fmeta + 2, n1, n2, sq_lowe, sq_dist,
        m->d_oneway.p, m->d_flist.p, m->d_fmeta);
    CU_TRY(m, cudaGetLastError());
    m->stats.kernel_launches += 4;
    return OSFM_OK;
}

int osfm_match_twoway_f32(osfm_matcher* m, const float* set_1, int n1, const float* set_2, int n2, int dim,
                          float lowe_ratio_threshold, float distance_threshold,
                          int32_t* matches_1_2, int32_t* matches_2_1) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->stream) return fail(m, OSFM_ERR_STATE, "handle was not created successfully");
    if (n1 < 0 || n2 < 0 || dim <= 0 || dim > kFDim || (dim & 3) != 0)
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "float path needs n >= 0 and dim a multiple of 4 in (0, %d]", kFDim);
    if ((n1 > 0 && (!set_1 || !matches_1_2)) || (n2 > 0 && (!set_2 || !matches_2_1)))
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "null pointer");
    // oneway_match: either set empty -> every entry -1 (matching.h:121-124)
    if (n1 == 0 || n2 == 0) {
        for (int i = 0; i < n1; ++i) matches_1_2[i] = -1;
        for (int i = 0; i < n2; ++i) matches_2_1[i] = -1;
        return OSFM_OK;
    }
    CU_TRY(m, cudaSetDevice(m->device));
    size_t const f1 = static_cast<size_t>(n1) * kFDim, f2 = static_cast<size_t>(n2) * kFDim;
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    CU_TRY(m, m->d_ftmp.reserve(f1 + f2));
    CU_TRY(m, m->d_oneway.reserve(static_cast<size_t>(n1) + n2));
    float* const d1 = m->d_ftmp.p;
    float* const d2 = m->d_ftmp.p + f1;
    if (dim < kFDim) CU_TRY(m, cudaMemsetAsync(m->d_ftmp.p, 0, (f1 + f2) * sizeof(float), m->stream));
    CU_TRY(m, cudaMemcpy2DAsync(d1, kFDim * sizeof(float), set_1, dim * sizeof(float), dim * sizeof(float), n1,
                                cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpy2DAsync(d2, kFDim * sizeof(float), set_2, dim * sizeof(float), dim * sizeof(float), n2,
                                cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaFuncSetAttribute(float_oneway_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFloatSmemBytes));
    float const sq_lowe = lowe_ratio_threshold * lowe_ratio_threshold;    // MATH_POW2 in float
    float const sq_dist = distance_threshold * distance_threshold;
    // Large pairs: the tensor-core filter decides nearly every row; the exact kernel then sees only
    // the rows on its list.  Small pairs: the exact kernel alone.
    bool const filter_first = m->float_mode == 2 ||
        (m->float_mode == 0 && static_cast<int64_t>(n1) * n2 >= (int64_t(1) << 20));
    const int32_t* list1 = nullptr;
    const int32_t* list2 = nullptr;
    const int* cnt1 = nullptr;
    const int* cnt2 = nullptr;
    if (filter_first) {
        OS_TRY(float_filter(m, d1, n1, d2, n2, sq_lowe, sq_dist));
        list1 = m->d_flist.p;
        list2 = m->d_flist.p + n1;
        cnt1 = m->d_fmeta;
        cnt2 = m->d_fmeta + 1;
    }
    if (filter_first) {
        // the listed rows: candidate tiles sliced over blockIdx.y, merged by float_finish_kernel
        int const s1n = std::min(64, std::max(1, ((n2 + kFN - 1) / kFN) / 4));
        int const s2n = std::min(64, std::max(1, ((n1 + kFN - 1) / kFN) / 4));
        CU_TRY(m, m->d_fparts.reserve(static_cast<size_t>(n1) * s1n + static_cast<size_t>(n2) * s2n));
        FloatRowState* const p1 = m->d_fparts.p;
        FloatRowState* const p2 = p1 + static_cast<size_t>(n1) * s1n;
        float_oneway_kernel<<<dim3((n1 + kFM - 1) / kFM, s1n), kFloatThreads, kFloatSmemBytes, m->stream>>>(
            d1, n1, d2, n2, sq_lowe, sq_dist, m->d_oneway.p, list1, cnt1, p1);
        float_oneway_kernel<<<dim3((n2 + kFM - 1) / kFM, s2n), kFloatThreads, kFloatSmemBytes, m->stream>>>(
            d2, n2, d1, n1, sq_lowe, sq_dist, m->d_oneway.p + n1, list2, cnt2, p2);
        float_finish_kernel<<<(n1 + 255) / 256, 256, 0, m->stream>>>(p1, s1n, list1, cnt1, n1, sq_lowe, sq_dist, m->d_oneway.p);
        float_finish_kernel<<<(n2 + 255) / 256, 256, 0, m->stream>>>(p2, s2n, list2, cnt2, n2, sq_lowe, sq_dist, m->d_oneway.p + n1);
        m->stats.kernel_launches += 2;
    } else {
        float_oneway_kernel<<<(n1 + kFM - 1) / kFM, kFloatThreads, kFloatSmemBytes, m->stream>>>(
            d1, n1, d2, n2, sq_lowe, sq_dist, m->d_oneway.p, nullptr, nullptr, nullptr);
        float_oneway_kernel<<<(n2 + kFM - 1) / kFM, kFloatThreads, kFloatSmemBytes, m->stream>>>(
            d2, n2, d1, n1, sq_lowe, sq_dist, m->d_oneway.p + n1, nullptr, nullptr, nullptr);
    }
    CU_TRY(m, cudaGetLastError());
    m->stats.kernel_launches += 2;
    CU_TRY(m, cudaMemcpyAsync(matches_1_2, m->d_oneway.p, sizeof(int32_t) * n1, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaMemcpyAsync(matches_2_1, m->d_oneway.p + n1, sizeof(int32_t) * n2, cudaMemcpyDeviceToHost, m->stream));
    int left[2] = {0, 0};
    if (filter_first) CU_TRY(m, cudaMemcpyAsync(left, m->d_fmeta, sizeof left, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    if (filter_first) {
        m->stats.float_filter_rows += static_cast<int64_t>(n1) + n2;
        m->stats.float_exact_rows += static_cast<int64_t>(left[0]) + left[1];
    }
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_debug_float_filter(osfm_matcher* m, const float* set_1, int n1, const float* set_2, int n2, int dim,
                                  float* s1, float* s2, int32_t* j1) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->stream) return fail(m, OSFM_ERR_STATE, "handle was not created successfully");
    if (n1 <= 0 || n2 <= 0 || dim <= 0 || dim > kFDim || (dim & 3) != 0 || !set_1 || !set_2 || !s1 || !s2 || !j1)
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "float filter dump needs two non-empty sets, dim a multiple of 4 in (0, %d]", kFDim);
    CU_TRY(m, cudaSetDevice(m->device));
    size_t const f1 = static_cast<size_t>(n1) * kFDim, f2 = static_cast<size_t>(n2) * kFDim;
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    CU_TRY(m, m->d_ftmp.reserve(f1 + f2));
    CU_TRY(m, m->d_oneway.reserve(static_cast<size_t>(n1) + n2));
    float* const d1 = m->d_ftmp.p;
    float* const d2 = m->d_ftmp.p + f1;
    if (dim < kFDim) CU_TRY(m, cudaMemsetAsync(m->d_ftmp.p, 0, (f1 + f2) * sizeof(float), m->stream));
    CU_TRY(m, cudaMemcpy2DAsync(d1, kFDim * sizeof(float), set_1, dim * sizeof(float), dim * sizeof(float), n1,
                                cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpy2DAsync(d2, kFDim * sizeof(float), set_2, dim * sizeof(float), dim * sizeof(float), n2,
                                cudaMemcpyHostToDevice, m->stream));
    OS_TRY(float_filter(m, d1, n1, d2, n2, 0.64f, INFINITY));
    std::vector<FloatTopRow> top(static_cast<size_t>(n1) + n2);
    CU_TRY(m, cudaMemcpyAsync(top.data(), m->d_ftop.p, sizeof(FloatTopRow) * top.size(), cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    for (size_t i = 0; i < top.size(); ++i) { s1[i] = top[i].s1; s2[i] = top[i].s2; j1[i] = top[i].j1; }
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// ---- batched, device-resident, compacted ------------------------------------------------

// capacity_ij < 0: use (and grow) the handle's own list buffer m->d_list; *d_used receives it.
// sift_only: the SIFT part of every pair (device-resident multi-GPU path); otherwise SIFT and
// SURF in the combined index space of pairwise_match.  Pairs with fewer than min_count
// consistent matches get an empty list.  counts_out (may be null) receives every pair's count.
static int compact_core(osfm_matcher* m, const int32_t* pairs, int npairs, int32_t* d_match_ij,
                        int64_t capacity_ij, int64_t* list_offset, int32_t** d_used,
                        bool sift_only = true, int min_count = 0, int32_t* counts_out = nullptr) {
    OS_TRY(require_committed(m));
    if (!list_offset) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "list_offset is null");
    std::vector<PairPlan> plans;
    OS_TRY(build_plans(m, pairs, npairs, 0, false, plans));
    if (capacity_ij < 0) {
        // a pair has at most min(n1, n2) mutual matches
        int64_t cap = 0;
        for (PairPlan const& p : plans)
            cap += std::min(p.n1[0], p.n2[0]) + (sift_only ? 0 : std::min(p.n1[1], p.n2[1]));
        CU_TRY(m, cudaSetDevice(m->device));
        CU_TRY(m, m->d_list.reserve(static_cast<size_t>(std::max<int64_t>(cap, 1))));
        d_match_ij = reinterpret_cast<int32_t*>(m->d_list.p);
        capacity_ij = cap;
    }
    if (d_used) *d_used = d_match_ij;
    if (sift_only)
        for (PairPlan& p : plans) {
            p.n1[1] = p.n2[1] = 0;
            p.len12 = p.n1[0]; p.len21 = p.n2[0];
        }
    CU_TRY(m, cudaSetDevice(m->device));
    m->phases.reset();
    CU_TRY(m, cudaEventRecord(m->ev[2], m->stream));
    int64_t list_base = 0;
    bool overflow = false;
    std::vector<int32_t> counts;
    std::vector<int64_t> loff;
    std::vector<PairPart> parts;
    int r = for_each_batch(plans, [&](size_t first, size_t last, int64_t dense) -> int {
        std::vector<PairPlan> sub(plans.begin() + first, plans.begin() + last);
        OS_TRY(run_batch(m, sub, dense, kFiltered, sift_only ? 0 : -1));
        size_t const np = last - first;
        counts.resize(np);
        CU_TRY(m, cudaMemcpyAsync(counts.data(), m->d_counts.p, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, m->stream));
        CU_TRY(m, cudaStreamSynchronize(m->stream));
        OS_TRY(collect_scan_time(m));
        loff.resize(np);
        parts.clear();
        for (size_t i = 0; i < np; ++i) {
            if (counts_out) counts_out[first + i] = counts[i];
            if (counts[i] < min_count) counts[i] = 0;
            list_offset[first + i] = list_base;
            loff[i] = list_base;
            list_base += counts[i];
            if (counts[i] > 0) {
                PairPart pt;
                memset(&pt, 0, sizeof pt);
                pt.out12 = sub[i].out12;
                pt.n1 = sub[i].len12;
                pt.pair = static_cast<int32_t>(i);
                parts.push_back(pt);
            }
        }
        if (list_base > capacity_ij || !d_match_ij) { overflow = true; return OSFM_OK; }
        if (parts.empty()) return OSFM_OK;
        CU_TRY(m, m->d_listoff.reserve(np));
        CU_TRY(m, m->d_parts.reserve(parts.size()));
        CU_TRY(m, cudaMemcpyAsync(m->d_listoff.p, loff.data(), sizeof(int64_t) * np, cudaMemcpyHostToDevice, m->stream));
        CU_TRY(m, cudaMemcpyAsync(m->d_parts.p, parts.data(), sizeof(PairPart) * parts.size(), cudaMemcpyHostToDevice, m->stream));
        CU_TRY(m, m->phases.mark(m->stream, kPhCompact));
        compact_kernel<<<static_cast<unsigned>(parts.size()), 1024, 0, m->stream>>>(
            m->d_parts.p, m->d_dense.p, m->d_listoff.p, reinterpret_cast<int2*>(d_match_ij));
        CU_TRY(m, cudaGetLastError());
        CU_TRY(m, m->phases.mark(m->stream, -1));
        m->stats.kernel_launches++;
        // no host round trip here: the next batch is ordered behind this kernel by the stream,
        // and the pageable sources above were staged before cudaMemcpyAsync returned
        return OSFM_OK;
    });
    if (r != OSFM_OK) return r;
    list_offset[plans.size()] = list_base;
    CU_TRY(m, cudaEventRecord(m->ev[3], m->stream));
    CU_TRY(m, cudaEventSynchronize(m->ev[3]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, m->ev[2], m->ev[3]);
    m->stats.last_total_ms = ms;
    publish_phase_times(m);
    m->stats.last_comparisons = comparisons_of(plans);
    OS_TRY(finish_staging(m));
    OS_TRY(read_counters(m));
    if (overflow) return fail(m, OSFM_ERR_OUT_OF_MEMORY, "match list needs %lld entries, capacity %lld",
                              (long long)list_base, (long long)capacity_ij);
    return OSFM_OK;
}

}  // extern "C" (templates below)

// ---- multi-device: every batched call is cut into one contiguous range of pairs per device ------
//
// Image pairs are independent (bundler_matching.cc:74-132 has no cross-pair state), so there is
// no data-path collective: the pool is replicated once at commit (replicate_to_peers), every device
// matches its range of the caller's pair list on its own handle and stream from its own host thread,
// and the results land in the caller's buffers in pair order.

struct Shard { osfm_matcher* h; size_t first, last; int rc; };

// Contiguous ranges of (nearly) equal cost sum(n1 * n2), one per device.
static std::vector<Shard> make_shards(osfm_matcher* m, const std::vector<PairPlan>& plans) {
    size_t const nd = 1 + m->peers.size(), n = plans.size();
    std::vector<double> cost(n);
    double total = 0.0;
    for (size_t i = 0; i < n; ++i) {
        cost[i] = 1.0;   // a pair is never free (fixed per-pair work), so empty pairs spread as well
        for (int kd = 0; kd < 2; ++kd) cost[i] += static_cast<double>(plans[i].n1[kd]) * plans[i].n2[kd];
        total += cost[i];
    }
    std::vector<Shard> out;
    size_t at = 0;
    double acc = 0.0;
    for (size_t d = 0; d < nd; ++d) {
        double const goal = total * static_cast<double>(d + 1) / static_cast<double>(nd);
        size_t const first = at;
        while (at < n && (d + 1 == nd || acc + 0.5 * cost[at] <= goal)) acc += cost[at++];
        out.push_back({d == 0 ? m : m->peers[d - 1], first, at, OSFM_OK});
    }
    return out;
}

// Runs fn(shard) for every non-empty shard: shard 0 (the primary, whose lock the caller holds) on
// this thread, the others on a thread each under their own handle's lock.
template <typename Fn>
static int run_sharded(osfm_matcher* m, std::vector<Shard>& shards, Fn fn) {
    std::vector<std::thread> threads;
    try {
        for (size_t d = 1; d < shards.size(); ++d) {
            if (shards[d].first == shards[d].last) continue;
            threads.emplace_back([&shards, &fn, d] {
                Shard& sh = shards[d];
                std::lock_guard<std::recursive_mutex> lock(sh.h->mu);
                try {                       // an exception must not leave a thread
                    sh.rc = fn(sh);
                } catch (const std::bad_alloc&) {
                    sh.rc = fail(sh.h, OSFM_ERR_OUT_OF_MEMORY, "out of host memory");
                } catch (...) {
                    sh.rc = fail(sh.h, OSFM_ERR_INTERNAL, "unexpected exception");
                }
            });
        }
    } catch (...) {
        for (std::thread& t : threads) t.join();
        return fail(m, OSFM_ERR_INTERNAL, "cannot start a host thread per device");
    }
    if (shards[0].first < shards[0].last) shards[0].rc = fn(shards[0]);
    for (std::thread& t : threads) t.join();
    cudaSetDevice(m->device);
    for (Shard& sh : shards)
        if (sh.rc != OSFM_OK) {
            if (sh.h != m) m->err = "device " + std::to_string(sh.h->device) + ": " + sh.h->err;
            return sh.rc;
        }
    return OSFM_OK;
}

// Sums the per-device statistics of the last sharded call into the primary's.
static void merge_peer_stats(osfm_matcher* m, const std::vector<Shard>& shards) {
    double total_ms = 0.0, scan_ms = 0.0, phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t cmp = 0;
    for (Shard const& sh : shards) {
        if (sh.first == sh.last) continue;
        osfm_match_stats const& st = sh.h->stats;
        total_ms = std::max(total_ms, st.last_total_ms);       // the devices work side by side
        scan_ms = std::max(scan_ms, st.last_scan_ms);
        for (int i = 0; i < 8; ++i) phase[i] = std::max(phase[i], st.last_phase_ms[i]);
        cmp += st.last_comparisons;
    }
    m->stats.last_total_ms = total_ms;
    m->stats.last_scan_ms = scan_ms;
    for (int i = 0; i < 8; ++i) m->stats.last_phase_ms[i] = phase[i];
    m->stats.last_comparisons = cmp;
}

// Dense results of a plan list into HOST buffers, on one device or sharded over all of them.
static int dense_dispatch(osfm_matcher* m, std::vector<PairPlan>& plans, OutputMode mode, int only_kind,
                          int32_t* matches, int64_t* offsets, int32_t* n_consistent) {
    if (m->peers.empty() || plans.size() < 2)
        return match_pairs_dense(m, plans, mode, only_kind, matches, offsets, n_consistent);
    std::vector<Shard> shards = make_shards(m, plans);
    std::vector<int64_t> base(plans.size() + 1, 0);
    for (size_t i = 0; i < plans.size(); ++i) base[i + 1] = base[i] + plans[i].len12 + plans[i].len21;
    int const rc = run_sharded(m, shards, [&](Shard& sh) -> int {
        std::vector<PairPlan> sub(plans.begin() + sh.first, plans.begin() + sh.last);
        std::vector<int64_t> off(2 * sub.size() + 1, 0);
        int const r = match_pairs_dense(sh.h, sub, mode, only_kind, matches ? matches + base[sh.first] : nullptr,
                                        offsets ? off.data() : nullptr, n_consistent ? n_consistent + sh.first : nullptr);
        if (r == OSFM_OK && offsets)
            for (size_t i = 0; i < 2 * sub.size(); ++i) offsets[2 * sh.first + i] = base[sh.first] + off[i];
        return r;
    });
    if (rc != OSFM_OK) return rc;
    if (offsets) offsets[2 * plans.size()] = base[plans.size()];
    merge_peer_stats(m, shards);
    return OSFM_OK;
}

// Compacted correspondence lists of a pair list into a HOST buffer, in pair order, on one device or
// sharded over all of them.  sift_only / min_count / counts_out as in compact_core.
static int compact_to_host(osfm_matcher* m, const int32_t* pairs, int npairs, int32_t* match_ij, int64_t capacity_ij,
                           int64_t* list_offset, bool sift_only, int min_count, int32_t* counts_out) {
    if (!list_offset) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "list_offset is null");
    if (m->peers.empty() || npairs < 2) {
        int32_t* d = nullptr;
        OS_TRY(compact_core(m, pairs, npairs, nullptr, -1, list_offset, &d, sift_only, min_count, counts_out));
        int64_t const total = list_offset[npairs];
        if (total > capacity_ij || (total > 0 && !match_ij))
            return fail(m, OSFM_ERR_OUT_OF_MEMORY, "match list needs %lld entries, capacity %lld",
                        (long long)total, (long long)capacity_ij);
        if (total > 0) {
            CU_TRY(m, cudaMemcpyAsync(match_ij, d, sizeof(int32_t) * 2 * total, cudaMemcpyDeviceToHost, m->stream));
            CU_TRY(m, cudaStreamSynchronize(m->stream));
        }
        return OSFM_OK;
    }
    OS_TRY(require_committed(m));
    std::vector<PairPlan> plans;
    OS_TRY(build_plans(m, pairs, npairs, 0, false, plans));
    std::vector<Shard> shards = make_shards(m, plans);
    std::vector<int32_t*> d_lists(shards.size(), nullptr);
    std::vector<std::vector<int64_t>> loff(shards.size());
    // 1. every device matches its range; its lists stay in its own list buffer
    OS_TRY(run_sharded(m, shards, [&](Shard& sh) -> int {
        size_t const d = static_cast<size_t>(&sh - shards.data());
        size_t const cnt = sh.last - sh.first;
        loff[d].assign(cnt + 1, 0);
        return compact_core(sh.h, pairs + 2 * sh.first, static_cast<int>(cnt), nullptr, -1, loff[d].data(), &d_lists[d],
                            sift_only, min_count, counts_out ? counts_out + sh.first : nullptr);
    }));
    // 2. where each device's lists go in the caller's buffer
    std::vector<int64_t> base(shards.size() + 1, 0);
    for (size_t d = 0; d < shards.size(); ++d) {
        int64_t const total = loff[d].empty() ? 0 : loff[d].back();
        base[d + 1] = base[d] + total;
        for (size_t i = 0; i + 1 < loff[d].size(); ++i) list_offset[shards[d].first + i] = base[d] + loff[d][i];
    }
    list_offset[npairs] = base[shards.size()];
    merge_peer_stats(m, shards);
    if (base[shards.size()] > capacity_ij || (base[shards.size()] > 0 && !match_ij))
        return fail(m, OSFM_ERR_OUT_OF_MEMORY, "match list needs %lld entries, capacity %lld",
                    (long long)base[shards.size()], (long long)capacity_ij);
    // 3. every device copies its lists to their place, side by side
    return run_sharded(m, shards, [&](Shard& sh) -> int {
        size_t const d = static_cast<size_t>(&sh - shards.data());
        int64_t const total = base[d + 1] - base[d];
        if (total == 0) return OSFM_OK;
        CU_TRY(sh.h, cudaSetDevice(sh.h->device));
        CU_TRY(sh.h, cudaMemcpyAsync(match_ij + 2 * base[d], d_lists[d], sizeof(int32_t) * 2 * total, cudaMemcpyDeviceToHost,
                                     sh.h->stream));
        CU_TRY(sh.h, cudaStreamSynchronize(sh.h->stream));
        return OSFM_OK;
    });
}

extern "C" {

int osfm_match_pairs_compact_device(osfm_matcher* m, const int32_t* pairs, int npairs, int32_t* d_match_ij,
                                    int64_t capacity_ij, int64_t* list_offset) {
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (capacity_ij < 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "negative capacity");
    return compact_core(m, pairs, npairs, d_match_ij, capacity_ij, list_offset, nullptr);
}

int osfm_match_pairs_compact(osfm_matcher* m, const int32_t* pairs, int npairs, int32_t* match_ij,
                             int64_t capacity_ij, int64_t* list_offset) {
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (capacity_ij < 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "negative capacity");
    return compact_to_host(m, pairs, npairs, match_ij, capacity_ij, list_offset, true, 0, nullptr);
}

// ---- two-view gates (bundler::Matching::two_view_matching up to RANSAC) --------------------------

void osfm_match_two_view_default_options(osfm_two_view_options* o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->use_lowres_matching = 0;       // bundler_matching.h:66
    o->num_lowres_features = 500;     // :68
    o->min_lowres_matches = 5;        // :70
    o->min_feature_matches = 24;      // :62
    o->match_num_previous_frames = 0; // :72
}

int osfm_match_two_view_candidates(osfm_matcher* m, const osfm_two_view_options* opts, const int32_t* pairs,
                                   int npairs, int32_t* match_ij, int64_t capacity_ij, int64_t* list_offset,
                                   int32_t* status, int32_t* count) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    OS_TRY(require_committed(m));
    if (!opts || !list_offset || !status || !count || npairs < 0 || (npairs > 0 && !pairs) || capacity_ij < 0)
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad argument");
    // 1. the rules that need no matching (bundler_matching.cc:92-100)
    std::vector<int32_t> todo;      // indices into pairs[]
    auto feats = [&](int v) { return static_cast<int64_t>(m->kind[0].n[v]) + m->kind[1].n[v]; };
    for (int i = 0; i < npairs; ++i) {
        int const v1 = pairs[2 * i], v2 = pairs[2 * i + 1];
        OS_TRY(check_view(m, v1));
        OS_TRY(check_view(m, v2));
        status[i] = OSFM_TWO_VIEW_SKIPPED;
        count[i] = 0;
        if (opts->match_num_previous_frames != 0 && v2 + opts->match_num_previous_frames < v1) continue;
        if (feats(v1) == 0 || feats(v2) == 0) continue;
        todo.push_back(i);
    }
    // 2. the low-resolution gate (bundler_matching.cc:146-158), all eligible pairs in one batch
    std::vector<int32_t> full;
    if (opts->use_lowres_matching) {
        if (opts->num_lowres_features <= 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad num_lowres_features");
        std::vector<int32_t> lr_pairs, lr_index;
        for (int32_t i : todo) {
            if (feats(pairs[2 * i]) * feats(pairs[2 * i + 1]) > 1000000) {
                lr_pairs.push_back(pairs[2 * i]);
                lr_pairs.push_back(pairs[2 * i + 1]);
                lr_index.push_back(i);
            } else {
                full.push_back(i);
            }
        }
        if (!lr_index.empty()) {
            std::vector<PairPlan> plans;
            OS_TRY(build_plans(m, lr_pairs.data(), static_cast<int>(lr_index.size()), opts->num_lowres_features, true, plans));
            std::vector<int32_t> lr_counts(lr_index.size(), 0);
            // count_consistent_matches of the unfiltered result = survivors of the mutual filter
            OS_TRY(dense_dispatch(m, plans, kFiltered, -1, nullptr, nullptr, lr_counts.data()));
            for (size_t k = 0; k < lr_index.size(); ++k) {
                int32_t const i = lr_index[k];
                if (lr_counts[k] < opts->min_lowres_matches) {
                    status[i] = OSFM_TWO_VIEW_LOWRES_REJECTED;
                    count[i] = lr_counts[k];
                } else {
                    full.push_back(i);
                }
            }
            std::sort(full.begin(), full.end());
        }
    } else {
        full = todo;
    }
    // 3. full matching of the remaining pairs, the match-count threshold (:161-169) and the
    //    correspondence lists (:171-192)
    for (int i = 0; i <= npairs; ++i) list_offset[i] = 0;
    if (full.empty()) return OSFM_OK;
    std::vector<int32_t> fp;
    for (int32_t i : full) { fp.push_back(pairs[2 * i]); fp.push_back(pairs[2 * i + 1]); }
    int const nf = static_cast<int>(full.size());
    std::vector<int64_t> loff(nf + 1, 0);
    std::vector<int32_t> cnt(nf, 0);
    int const thr = std::max(8, opts->min_feature_matches);
    {
        int const rc = compact_to_host(m, fp.data(), nf, match_ij, capacity_ij, loff.data(), false, thr, cnt.data());
        if (rc != OSFM_OK) {
            if (rc == OSFM_ERR_OUT_OF_MEMORY) list_offset[npairs] = loff[nf];
            return rc;
        }
    }
    // scatter the per-pair results back to the caller's pair order (lists stay in `full` order,
    // which is ascending pair index: offsets are monotone)
    int k = 0;
    int64_t at = 0;
    for (int i = 0; i < npairs; ++i) {
        list_offset[i] = at;
        if (k < nf && full[k] == i) {
            count[i] = cnt[k];
            status[i] = cnt[k] < thr ? OSFM_TWO_VIEW_TOO_FEW_MATCHES : OSFM_TWO_VIEW_OK;
            at += loff[k + 1] - loff[k];
            ++k;
        }
    }
    list_offset[npairs] = at;
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// ---- tracks (bundler::Tracks::compute) -------------------------------------------------------

int osfm_tracks_compute(osfm_matcher* m, int num_views, const int32_t* features_per_view, const int32_t* pair_views,
                        const int64_t* list_offset, const int32_t* match_ij, int npairs,
                        int32_t* track_of_feature, int32_t* num_tracks, int32_t* num_conflicting) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->stream) return fail(m, OSFM_ERR_STATE, "handle was not created successfully");
    if (num_views < 0 || npairs < 0 || (num_views > 0 && !features_per_view) ||
        (npairs > 0 && (!pair_views || !list_offset)) || !num_tracks)
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad argument");
    std::vector<int64_t> base(num_views + 1, 0);
    for (int v = 0; v < num_views; ++v) {
        if (features_per_view[v] < 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "negative feature count");
        base[v + 1] = base[v] + features_per_view[v];
    }
    int64_t const n64 = base[num_views];
    if (n64 > INT32_MAX) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "more than 2^31 features");
    int const n = static_cast<int>(n64);
    int64_t const nedges = npairs > 0 ? list_offset[npairs] : 0;
    for (int p = 0; p < npairs; ++p) {
        OS_TRY((pair_views[2 * p] >= 0 && pair_views[2 * p] < num_views && pair_views[2 * p + 1] >= 0 &&
                pair_views[2 * p + 1] < num_views && list_offset[p] <= list_offset[p + 1])
                   ? OSFM_OK : fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad pair %d", p));
    }
    *num_tracks = 0;
    if (num_conflicting) *num_conflicting = 0;
    if (n == 0) return OSFM_OK;
    if (!track_of_feature || (nedges > 0 && !match_ij)) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "null buffer");
    CU_TRY(m, cudaSetDevice(m->device));

    // device scratch: grown on demand, kept with the handle
    uint64_t cap = 1024;
    while (cap < 2ull * static_cast<uint64_t>(n)) cap <<= 1;
    int const nblocks = (n + kScanBlock - 1) / kScanBlock;
    size_t const ni = sizeof(int) * static_cast<size_t>(n);
    // one int array holds parent | root | size | conflict | flag | id | out | block sums | 4 counters
    size_t const ints = 7 * static_cast<size_t>(n) + static_cast<size_t>(nblocks) + 1 + 4;
    CU_TRY(m, m->tr_ints.reserve(ints));
    CU_TRY(m, m->tr_table.reserve(cap));
    CU_TRY(m, m->tr_meta.reserve(static_cast<size_t>(num_views) + 1 + static_cast<size_t>(npairs) + 1));
    CU_TRY(m, m->tr_meta32.reserve(static_cast<size_t>(std::max(num_views, 1)) + 2 * static_cast<size_t>(npairs)));
    int* const d_parent = m->tr_ints.p;
    int* const d_root = d_parent + n;
    int* const d_size = d_root + n;
    int* const d_conflict = d_size + n;
    int* const d_flag = d_conflict + n;
    int* const d_id = d_flag + n;
    int32_t* const d_out = d_id + n;
    int* const d_bsum = d_out + n;
    int* const d_small = d_bsum + nblocks + 1;
    int64_t* const d_base = m->tr_meta.p;
    int64_t* const d_off = d_base + num_views + 1;
    int32_t* const d_vn = m->tr_meta32.p;
    int32_t* const d_pv = d_vn + std::max(num_views, 1);
    unsigned long long* const d_table = m->tr_table.p;
    CU_TRY(m, cudaMemsetAsync(d_size, 0, 2 * ni, m->stream));                       // size, conflict
    CU_TRY(m, cudaMemsetAsync(d_small, 0, sizeof(int) * 4, m->stream));
    CU_TRY(m, cudaMemsetAsync(d_table, 0xff, sizeof(unsigned long long) * cap, m->stream));
    CU_TRY(m, cudaMemcpyAsync(d_base, base.data(), sizeof(int64_t) * (num_views + 1), cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpyAsync(d_vn, features_per_view, sizeof(int32_t) * num_views, cudaMemcpyHostToDevice, m->stream));
    int const g = (n + 255) / 256;
    tracks_init_kernel<<<g, 256, 0, m->stream>>>(d_parent, n);
    if (nedges > 0) {
        CU_TRY(m, m->d_list.reserve(static_cast<size_t>(nedges)));
        int2* const d_ij = m->d_list.p;
        CU_TRY(m, cudaMemcpyAsync(d_pv, pair_views, sizeof(int32_t) * 2 * npairs, cudaMemcpyHostToDevice, m->stream));
        CU_TRY(m, cudaMemcpyAsync(d_off, list_offset, sizeof(int64_t) * (npairs + 1), cudaMemcpyHostToDevice, m->stream));
        CU_TRY(m, cudaMemcpyAsync(d_ij, match_ij, sizeof(int2) * nedges, cudaMemcpyHostToDevice, m->stream));
        int64_t const ge = (nedges + 255) / 256;
        tracks_union_kernel<<<static_cast<unsigned>(ge), 256, 0, m->stream>>>(d_parent, d_pv, d_off, npairs, d_ij, nedges,
                                                                             d_base, d_vn, d_small + 0);
    }
    tracks_root_kernel<<<g, 256, 0, m->stream>>>(d_parent, n, d_root, d_size);
    tracks_conflict_kernel<<<g, 256, 0, m->stream>>>(d_root, n, d_size, d_base, num_views, d_table, cap - 1, d_conflict);
    tracks_flag_kernel<<<g, 256, 0, m->stream>>>(d_root, d_size, d_conflict, n, d_flag, d_small + 1);
    scan_partial_kernel<<<nblocks, 256, 0, m->stream>>>(d_flag, n, d_bsum);
    scan_sums_kernel<<<1, 1024, 0, m->stream>>>(d_bsum, nblocks, d_small + 2);
    scan_apply_kernel<<<nblocks, 256, 0, m->stream>>>(d_flag, n, d_bsum, d_id);
    tracks_assign_kernel<<<g, 256, 0, m->stream>>>(d_root, d_id, n, d_out);
    CU_TRY(m, cudaGetLastError());
    m->stats.kernel_launches += nedges > 0 ? 9 : 8;
    int small[4];
    CU_TRY(m, cudaMemcpyAsync(track_of_feature, d_out, ni, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaMemcpyAsync(small, d_small, sizeof small, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    if (small[0] != 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "%d matches refer to features outside their view", small[0]);
    *num_tracks = small[2];
    if (num_conflicting) *num_conflicting = small[1];
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// ---- RANSAC for the fundamental matrix ------------------------------------------------------------

// The reference draws from std::rand().  Under glibc that is random() behind a lock (about
// 20 ns a call, and a pair needs 8000+ calls); RandSequence reads the very same generator
// without the lock: setstate() parks the process-wide generator on a scratch state and hands
// out its real state, setstate_r() attaches a private random_data to that state, random_r()
// steps it (the same additive-feedback recurrence, the same values), and the destructor stores
// the advanced position back and makes it the process-wide state again.  Whoever calls rand()
// afterwards continues exactly where the reference's own draws would have left the sequence.
// Not for use while another thread calls rand() (neither is the reference).
class RandSequence {
public:
#if defined(__GLIBC__)
    RandSequence() {
        memset(&park_rd_, 0, sizeof park_rd_);
        memset(&rd_, 0, sizeof rd_);
        // two valid scratch states: one for the process-wide generator to rest on, one for rd_
        // to be initialised with (setstate_r writes through the state it leaves)
        if (initstate_r(1u, park_, sizeof park_, &park_rd_) != 0 || initstate_r(1u, spare_, sizeof spare_, &rd_) != 0) return;
        real_ = setstate(park_);
        if (!real_) return;
        if (setstate_r(real_, &rd_) != 0) { setstate(real_); real_ = nullptr; return; }
    }
    ~RandSequence() {
        if (!real_) return;
        setstate_r(spare_, &rd_);      // writes the position reached into the real state
        setstate(real_);
    }
    int next() {
        if (!real_) return std::rand();
        int32_t v;
        random_r(&rd_, &v);
        return static_cast<int>(v);
    }
private:
    alignas(8) char park_[128];
    alignas(8) char spare_[128];
    struct random_data park_rd_, rd_;
    char* real_ = nullptr;
#else
    RandSequence() = default;
    int next() { return std::rand(); }
#endif
    RandSequence(const RandSequence&) = delete;
    RandSequence& operator=(const RandSequence&) = delete;
};

// RansacFundamental::estimate_8_point, ransac_fundamental.cc:70-76: rand() % count into an
// ordered set until it holds eight; the set is then read in ascending order.
static inline void order2(int32_t& a, int32_t& b) {
    int32_t const lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo; b = hi;
}

// RandSequence takes over the process-wide generator while it draws: two handles (or two threads
// on one) must not do that at the same time.  A thread that calls rand() itself meanwhile is as
// undefined as it is against the reference's own std::rand() loop.
static std::mutex g_rand_mutex;

static void draw_samples_for_pairs(int npairs, const int64_t* list_offset, int max_iterations, int32_t* out) {
    std::lock_guard<std::mutex> rand_lock(g_rand_mutex);
    RandSequence seq;
    for (int p = 0; p < npairs; ++p) {
        // rand() % count without a division per draw (Lemire's fastmod: exact for 32-bit operands)
        uint32_t const count = static_cast<uint32_t>(list_offset[p + 1] - list_offset[p]);
        uint64_t const magic = ~0ull / count + 1;
        for (int it = 0; it < max_iterations; ++it, out += 8) {
            // the std::set of the reference: distinct values, read in ascending order.  Kept
            // unsorted while drawing (a duplicate is rare, so that branch predicts), sorted once
            // by a 19-exchange network.
            int32_t s[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
            int have = 0;
            while (have < 8) {
                uint64_t const low = magic * static_cast<uint32_t>(seq.next());
                int32_t const v = static_cast<int32_t>((static_cast<unsigned __int128>(low) * count) >> 64);
                int dup = 0;
                for (int k = 0; k < 8; ++k) dup |= (s[k] == v);
                if (dup) continue;
                s[have++] = v;
            }
            order2(s[0], s[1]); order2(s[2], s[3]); order2(s[4], s[5]); order2(s[6], s[7]);
            order2(s[0], s[2]); order2(s[1], s[3]); order2(s[4], s[6]); order2(s[5], s[7]);
            order2(s[1], s[2]); order2(s[5], s[6]); order2(s[0], s[4]); order2(s[3], s[7]);
            order2(s[1], s[5]); order2(s[2], s[6]);
            order2(s[1], s[4]); order2(s[3], s[6]);
            order2(s[2], s[4]); order2(s[3], s[5]);
            order2(s[3], s[4]);
            for (int k = 0; k < 8; ++k) out[k] = s[k];
        }
    }
}

int osfm_ransac_draw_samples(int npairs, const int64_t* list_offset, int max_iterations, int32_t* samples) {
    OSFM_TRY_BEGIN
    if (npairs < 0 || max_iterations < 0 || (npairs > 0 && (!list_offset || !samples))) return OSFM_ERR_INVALID_ARGUMENT;
    for (int p = 0; p < npairs; ++p) {
        int64_t const count = list_offset[p + 1] - list_offset[p];
        if (count < 8 || count > INT32_MAX) return OSFM_ERR_INVALID_ARGUMENT;    // the reference throws below 8
    }
    draw_samples_for_pairs(npairs, list_offset, max_iterations, samples);
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

int osfm_ransac_fundamental(osfm_matcher* m, int num_views, const int32_t* features_per_view, const float* positions,
                            const int32_t* pair_views, const int64_t* list_offset, const int32_t* match_ij,
                            int npairs, const int32_t* samples, int max_iterations, double threshold,
                            int32_t* inlier_ij, int64_t* inlier_offset, double* fundamental) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    if (!m->stream) return fail(m, OSFM_ERR_STATE, "handle was not created successfully");
    if (num_views < 0 || npairs < 0 || max_iterations < 0 || (num_views > 0 && !features_per_view) || !inlier_offset ||
        (npairs > 0 && (!pair_views || !list_offset || !positions || !match_ij || !inlier_ij)))
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad argument");
    inlier_offset[0] = 0;
    if (npairs == 0) return OSFM_OK;
    std::vector<int64_t> base(num_views + 1, 0);
    for (int v = 0; v < num_views; ++v) {
        if (features_per_view[v] < 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "negative feature count");
        base[v + 1] = base[v] + features_per_view[v];
    }
    if (list_offset[0] != 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "list_offset[0] must be 0");
    for (int p = 0; p < npairs; ++p) {
        int64_t const count = list_offset[p + 1] - list_offset[p];
        if (pair_views[2 * p] < 0 || pair_views[2 * p] >= num_views || pair_views[2 * p + 1] < 0 ||
            pair_views[2 * p + 1] >= num_views)
            return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad pair %d", p);
        if (count < 8 || count > INT32_MAX)
            return fail(m, OSFM_ERR_INVALID_ARGUMENT, "pair %d: at least 8 matches required", p);
    }
    int64_t const nmatches = list_offset[npairs];
    int const per_pair = 8 * max_iterations;
    // Pairs go through in chunks of about 8 MB of samples: enough fits (260 k at 1000
    // iterations) for the work fetching of the iteration kernel to even out its lanes, and the
    // per-fit scratch (samples, matrices, inlier counts) is sized for one chunk, not for the job
    const char* chunk_env = getenv("OSFM_RANSAC_CHUNK_PAIRS");      // test knob: forces small chunks
    int const chunk = chunk_env ? std::max(1, std::min(npairs, atoi(chunk_env)))
                                : std::max(1, std::min(npairs, (1 << 21) / std::max(per_pair, 1)));
    int64_t const chunk_fits = static_cast<int64_t>(chunk) * max_iterations;
    int64_t const scratch_fits = std::max<int64_t>(chunk_fits, 1);
    CU_TRY(m, cudaSetDevice(m->device));
    CU_TRY(m, m->rs_xy.reserve(static_cast<size_t>(nmatches)));
    CU_TRY(m, m->rs_out.reserve(static_cast<size_t>(2 * nmatches)));              // input lists | inlier lists
    CU_TRY(m, m->rs_pos.reserve(static_cast<size_t>(std::max<int64_t>(base[num_views], 1))));
    CU_TRY(m, m->rs_samples.reserve(static_cast<size_t>(scratch_fits * 8)));
    CU_TRY(m, m->rs_F.reserve(static_cast<size_t>(scratch_fits * 9) + 9 * static_cast<size_t>(npairs)));
    CU_TRY(m, m->rs_cnt.reserve(static_cast<size_t>(scratch_fits) + static_cast<size_t>(npairs) + 4));
    CU_TRY(m, m->tr_meta.reserve(static_cast<size_t>(num_views) + 1 + static_cast<size_t>(npairs) + 1));
    CU_TRY(m, m->tr_meta32.reserve(static_cast<size_t>(std::max(num_views, 1)) + 2 * static_cast<size_t>(npairs)));
    int64_t* const d_base = m->tr_meta.p;
    int64_t* const d_off = d_base + num_views + 1;
    int32_t* const d_vn = m->tr_meta32.p;
    int32_t* const d_pv = d_vn + std::max(num_views, 1);
    int2* const d_ij = m->rs_out.p;
    int2* const d_inl = d_ij + nmatches;
    double* const d_bestF = m->rs_F.p + scratch_fits * 9;
    int* const d_count = m->rs_cnt.p + scratch_fits;
    int* const d_bad = d_count + npairs;
    cudaStream_t const st = m->stream;
    CU_TRY(m, cudaMemsetAsync(d_bad, 0, sizeof(int) * 2, st));
    CU_TRY(m, cudaMemcpyAsync(d_base, base.data(), sizeof(int64_t) * (num_views + 1), cudaMemcpyHostToDevice, st));
    CU_TRY(m, cudaMemcpyAsync(d_vn, features_per_view, sizeof(int32_t) * num_views, cudaMemcpyHostToDevice, st));
    CU_TRY(m, cudaMemcpyAsync(d_pv, pair_views, sizeof(int32_t) * 2 * npairs, cudaMemcpyHostToDevice, st));
    CU_TRY(m, cudaMemcpyAsync(d_off, list_offset, sizeof(int64_t) * (npairs + 1), cudaMemcpyHostToDevice, st));
    CU_TRY(m, cudaMemcpyAsync(d_ij, match_ij, sizeof(int2) * nmatches, cudaMemcpyHostToDevice, st));
    if (base[num_views] > 0)
        CU_TRY(m, cudaMemcpyAsync(m->rs_pos.p, positions, sizeof(float2) * base[num_views], cudaMemcpyHostToDevice, st));
    ransac_gather_kernel<<<static_cast<unsigned>((nmatches + 255) / 256), 256, 0, st>>>(
        d_pv, d_off, npairs, d_ij, nmatches, d_base, d_vn, m->rs_pos.p, m->rs_xy.p, d_bad + 0);
    double const thr2 = threshold * threshold;       // ransac_fundamental.cc:97
    // When the samples are drawn here, the draws of one chunk (host) run while the device works
    // on the chunk before.
    CU_TRY(m, m->rs_stage1.reserve(static_cast<size_t>(std::max<int64_t>(chunk_fits, 1)) * (kBidiagDoubles + 81) + 1));
    double* const d_bd = m->rs_stage1.p;
    double* const d_vv = d_bd + std::max<int64_t>(chunk_fits, 1) * kBidiagDoubles;
    unsigned long long* const d_next = reinterpret_cast<unsigned long long*>(d_vv + std::max<int64_t>(chunk_fits, 1) * 81);
    const char* gk_thr_env = getenv("OSFM_GK_THREADS");       // tuning knob: 32, 64 or 128
    int const gk_threads = gk_thr_env ? std::max(32, std::min(kGkThreads, atoi(gk_thr_env) / 32 * 32)) : kGkThreads;
    size_t const gk_smem = sizeof(double) * 81 * gk_threads;
    CU_TRY(m, cudaFuncSetAttribute(ransac_gk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(gk_smem)));
    const char* gk_env = getenv("OSFM_GK_BEGIN_BATCH");      // tuning knob (tools/ransac_probe.py)
    int const gk_batch = gk_env ? std::max(1, std::min(32, atoi(gk_env))) : kGkBeginBatch;
    int gk_per_sm = 1;
    CU_TRY(m, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gk_per_sm, ransac_gk_kernel, gk_threads, gk_smem));
    int64_t const gk_ctas = static_cast<int64_t>(std::max(gk_per_sm, 1)) * m->num_sms;
    if (!samples && per_pair > 0) {
        size_t const want = static_cast<size_t>(chunk) * per_pair;
        if (want > m->rs_stage_ints) {
            for (int k = 0; k < 2; ++k) {
                if (m->rs_stage[k]) cudaFreeHost(m->rs_stage[k]);
                m->rs_stage[k] = nullptr;
                CU_TRY(m, cudaHostAlloc(reinterpret_cast<void**>(&m->rs_stage[k]), want * sizeof(int32_t), cudaHostAllocDefault));
                if (!m->rs_stage_free[k]) CU_TRY(m, cudaEventCreateWithFlags(&m->rs_stage_free[k], cudaEventDisableTiming));
            }
            m->rs_stage_ints = want;
        }
    }
    int launches = 1;
    // When the samples are drawn here the first chunks are small (1/8, 1/4, 1/2 of a chunk): the
    // device has nothing to do while the first chunk is drawn, so that one should be short.
    int np = 0;
    for (int p0 = 0, c = 0; p0 < npairs; p0 += np, ++c) {
        int const ramp = (!samples && !chunk_env && c < 3) ? std::max(1, chunk >> (3 - c)) : chunk;
        np = std::min(ramp, npairs - p0);
        int64_t const fits = static_cast<int64_t>(np) * max_iterations;
        int64_t const f0 = static_cast<int64_t>(p0) * max_iterations;
        if (fits > 0) {
            const int32_t* src = samples ? samples + f0 * 8 : nullptr;
            if (!samples) {
                int const buf = c & 1;
                if (c >= 2) CU_TRY(m, cudaEventSynchronize(m->rs_stage_free[buf]));       // its last copy has left
                draw_samples_for_pairs(np, list_offset + p0, max_iterations, m->rs_stage[buf]);
                src = m->rs_stage[buf];
            }
            CU_TRY(m, cudaMemcpyAsync(m->rs_samples.p, src, sizeof(int32_t) * 8 * fits, cudaMemcpyHostToDevice, st));
            if (!samples) CU_TRY(m, cudaEventRecord(m->rs_stage_free[c & 1], st));
            ransac_bidiag_kernel<<<static_cast<unsigned>((fits + 127) / 128), 128, 0, st>>>(
                d_off + p0, np, max_iterations, m->rs_samples.p, m->rs_xy.p, d_bd, d_vv, d_bad + 1);
            CU_TRY(m, cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), st));
            ransac_gk_kernel<<<static_cast<unsigned>(std::min<int64_t>(gk_ctas, (fits + gk_threads - 1) / gk_threads)),
                               gk_threads, gk_smem, st>>>(fits, d_bd, d_vv, d_next, m->rs_F.p, gk_batch);
            ransac_rank2_kernel<<<static_cast<unsigned>((fits + 127) / 128), 128, 0, st>>>(fits, m->rs_F.p);
            ransac_count_kernel<<<static_cast<unsigned>((fits * 32 + 255) / 256), 256, 0, st>>>(
                d_off + p0, np, max_iterations, m->rs_xy.p, m->rs_F.p, thr2, m->rs_cnt.p);
            launches += 4;
        }
        ransac_select_kernel<<<np, 256, 0, st>>>(d_off + p0, max_iterations, m->rs_xy.p, d_ij, m->rs_F.p,
                                                 m->rs_cnt.p, thr2, d_inl, d_count + p0, d_bestF + 9 * p0);
        launches += 1;
    }
    CU_TRY(m, cudaGetLastError());
    m->stats.kernel_launches += launches;
    std::vector<int> count(static_cast<size_t>(npairs) + 2);
    std::vector<int2> inl(static_cast<size_t>(nmatches));
    CU_TRY(m, cudaMemcpyAsync(count.data(), d_count, sizeof(int) * (npairs + 2), cudaMemcpyDeviceToHost, st));
    CU_TRY(m, cudaMemcpyAsync(inl.data(), d_inl, sizeof(int2) * nmatches, cudaMemcpyDeviceToHost, st));
    if (fundamental)
        CU_TRY(m, cudaMemcpyAsync(fundamental, d_bestF, sizeof(double) * 9 * npairs, cudaMemcpyDeviceToHost, st));
    CU_TRY(m, cudaStreamSynchronize(st));
    if (count[npairs] != 0)
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "%d matches refer to features outside their view", count[npairs]);
    if (count[npairs + 1] != 0)
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "%d samples are not eight ascending indices into their pair's list",
                    count[npairs + 1]);
    for (int p = 0; p < npairs; ++p) {
        if (count[p] > 0)
            memcpy(inlier_ij + 2 * inlier_offset[p], inl.data() + list_offset[p], sizeof(int2) * count[p]);
        inlier_offset[p + 1] = inlier_offset[p] + count[p];
    }
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// ---- the whole two-view stage: gates, RANSAC, inlier threshold --------------------------------------

void osfm_match_ransac_default_options(osfm_ransac_options* o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->max_iterations = 1000;          // RansacFundamental::Options, ransac_fundamental.h:88-93
    o->threshold = 0.0015;
    o->min_matching_inliers = 12;      // bundler::Matching::Options, bundler_matching.h:66
}

int osfm_match_two_view(osfm_matcher* m, const osfm_two_view_options* opts, const osfm_ransac_options* ransac,
                        const float* positions, const int32_t* pairs, int npairs, int32_t* match_ij,
                        int64_t capacity_ij, int64_t* list_offset, int32_t* status, int32_t* count) {
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    OSFM_TRY_BEGIN
    // the views must not change between the stages: one lock for the whole call
    std::lock_guard<std::recursive_mutex> whole_call(m->mu);
    if (!ransac || ransac->max_iterations < 0) {
        std::lock_guard<std::recursive_mutex> lock(m->mu);
        return fail(m, OSFM_ERR_INVALID_ARGUMENT, "bad RANSAC options");
    }
    // 1. bundler_matching.cc:92-192: pair rules, low-res gate, full match, match-count threshold
    OS_TRY(osfm_match_two_view_candidates(m, opts, pairs, npairs, match_ij, capacity_ij, list_offset, status, count));
    std::vector<int32_t> ok, pv;
    std::vector<int64_t> off(1, 0);
    for (int i = 0; i < npairs; ++i) {
        if (status[i] != OSFM_TWO_VIEW_OK) continue;
        ok.push_back(i);
        pv.push_back(pairs[2 * i]);
        pv.push_back(pairs[2 * i + 1]);
        off.push_back(list_offset[i + 1]);      // rejected pairs have empty lists: the lists of the others are back to back
    }
    int const nok = static_cast<int>(ok.size());
    if (nok == 0) return OSFM_OK;
    std::vector<int32_t> fpv;
    {
        std::lock_guard<std::recursive_mutex> lock(m->mu);
        if (!positions) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "positions required");
        fpv.resize(m->kind[0].n.size());
        for (size_t v = 0; v < fpv.size(); ++v) fpv[v] = m->kind[0].n[v] + m->kind[1].n[v];
    }
    // 2. :194-201: RANSAC; the samples are drawn inside, in the order the reference reaches the pairs
    std::vector<int32_t> inl(static_cast<size_t>(2 * off[nok]));
    std::vector<int64_t> inl_off(static_cast<size_t>(nok) + 1, 0);
    OS_TRY(osfm_ransac_fundamental(m, static_cast<int>(fpv.size()), fpv.data(), positions, pv.data(), off.data(), match_ij,
                                   nok, nullptr, ransac->max_iterations, ransac->threshold, inl.data(),
                                   inl_off.data(), nullptr));
    // 3. :203-220: the inlier threshold; the inliers replace the candidate lists
    int const thr = std::max(8, ransac->min_matching_inliers);
    int64_t at = 0;
    int k = 0;
    for (int i = 0; i < npairs; ++i) {
        list_offset[i] = at;
        if (status[i] != OSFM_TWO_VIEW_OK) continue;
        int64_t const n = inl_off[k + 1] - inl_off[k];
        count[i] = static_cast<int32_t>(n);
        if (n < thr) {
            status[i] = OSFM_TWO_VIEW_TOO_FEW_INLIERS;
        } else {
            memcpy(match_ij + 2 * at, inl.data() + 2 * inl_off[k], sizeof(int32_t) * 2 * n);
            at += n;
        }
        ++k;
    }
    list_offset[npairs] = at;
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

// ---- on-disk format (MVE prebundle) ------------------------------------------------------------

int osfm_io_save_prebundle(const char* path, int num_views, const int32_t* features_per_view,
                           const float* positions, const uint8_t* colors, int npairs,
                           const int32_t* pair_views, const int64_t* list_offset, const int32_t* match_ij) {
    OSFM_TRY_BEGIN
    if (!path || num_views < 0 || npairs < 0 || (num_views > 0 && !features_per_view) ||
        (npairs > 0 && (!pair_views || !list_offset)))
        return OSFM_ERR_INVALID_ARGUMENT;
    for (int p = 0; p < npairs; ++p) {
        int64_t const k = list_offset[p + 1] - list_offset[p];
        if (k < 0 || k > INT32_MAX || (k > 0 && !match_ij)) return OSFM_ERR_INVALID_ARGUMENT;
    }
    return save_prebundle(path, num_views, features_per_view, positions, colors, npairs, pair_views, list_offset,
                          match_ij) == 0 ? OSFM_OK : OSFM_ERR_IO;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

struct osfm_prebundle {
    PrebundleData d;
};

int osfm_io_load_prebundle(const char* path, osfm_prebundle** out, int* num_views, int64_t* num_positions,
                           int64_t* num_colors, int* npairs, int64_t* num_matches) {
    OSFM_TRY_BEGIN
    if (!path || !out) return OSFM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    osfm_prebundle* h = new (std::nothrow) osfm_prebundle();
    if (!h) return OSFM_ERR_OUT_OF_MEMORY;
    int const r = load_prebundle(path, &h->d);
    if (r != 0) { delete h; return r == 1 ? OSFM_ERR_IO : OSFM_ERR_INVALID_ARGUMENT; }
    if (num_views) *num_views = static_cast<int>(h->d.n_positions.size());
    if (num_positions) *num_positions = static_cast<int64_t>(h->d.positions.size() / 2);
    if (num_colors) *num_colors = static_cast<int64_t>(h->d.colors.size() / 3);
    if (npairs) *npairs = static_cast<int>(h->d.pair_views.size() / 2);
    if (num_matches) *num_matches = h->d.list_offset.back();
    *out = h;
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

int osfm_io_prebundle_get(const osfm_prebundle* h, int32_t* n_positions, int32_t* n_colors, float* positions,
                          uint8_t* colors, int32_t* pair_views, int64_t* list_offset, int32_t* match_ij) {
    OSFM_TRY_BEGIN
    if (!h) return OSFM_ERR_INVALID_ARGUMENT;
    PrebundleData const& d = h->d;
    auto cp = [](void* dst, const void* src, size_t bytes) { if (dst && bytes) memcpy(dst, src, bytes); };
    cp(n_positions, d.n_positions.data(), sizeof(int32_t) * d.n_positions.size());
    cp(n_colors, d.n_colors.data(), sizeof(int32_t) * d.n_colors.size());
    cp(positions, d.positions.data(), sizeof(float) * d.positions.size());
    cp(colors, d.colors.data(), d.colors.size());
    cp(pair_views, d.pair_views.data(), sizeof(int32_t) * d.pair_views.size());
    cp(list_offset, d.list_offset.data(), sizeof(int64_t) * d.list_offset.size());
    cp(match_ij, d.match_ij.data(), sizeof(int32_t) * d.match_ij.size());
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

void osfm_io_prebundle_free(osfm_prebundle* h) { delete h; }

// ---- on-disk formats (tracks.txt, AAA_BBB.txt) -------------------------------------------------

static int track_table_from_ids(int num_views, const int32_t* features_per_view, const int32_t* track_of_feature,
                                int num_tracks, const float* positions, double image_width, const uint8_t* colors,
                                TrackTable* t) {
    if (num_views < 0 || num_tracks < 0 || (num_views > 0 && !features_per_view)) return OSFM_ERR_INVALID_ARGUMENT;
    int64_t total = 0;
    for (int v = 0; v < num_views; ++v) {
        if (features_per_view[v] < 0) return OSFM_ERR_INVALID_ARGUMENT;
        total += features_per_view[v];
    }
    if (total > 0 && !track_of_feature) return OSFM_ERR_INVALID_ARGUMENT;
    return build_track_table(num_views, features_per_view, track_of_feature, num_tracks, positions, image_width,
                             colors, t) == 0 ? OSFM_OK : OSFM_ERR_INVALID_ARGUMENT;
}

int osfm_io_save_tracks(const char* path, int num_views, const int32_t* features_per_view,
                        const int32_t* track_of_feature, int num_tracks, const float* positions,
                        double image_width, const uint8_t* colors) {
    OSFM_TRY_BEGIN
    if (!path) return OSFM_ERR_INVALID_ARGUMENT;
    TrackTable t;
    int const rc = track_table_from_ids(num_views, features_per_view, track_of_feature, num_tracks, positions,
                                        image_width, colors, &t);
    if (rc != OSFM_OK) return rc;
    return save_tracks(path, t) == 0 ? OSFM_OK : OSFM_ERR_IO;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

int osfm_io_save_pairwise_tracks(const char* folder, int num_views, const int32_t* features_per_view,
                                 const int32_t* track_of_feature, int num_tracks, const float* positions,
                                 double image_width, int* files_written) {
    OSFM_TRY_BEGIN
    if (!folder) return OSFM_ERR_INVALID_ARGUMENT;
    TrackTable t;
    int const rc = track_table_from_ids(num_views, features_per_view, track_of_feature, num_tracks, positions,
                                        image_width, nullptr, &t);
    if (rc != OSFM_OK) return rc;
    std::vector<int32_t> ids(static_cast<size_t>(num_views));
    for (int v = 0; v < num_views; ++v) ids[v] = v;
    int const n = save_pairwise_tracks(folder, t, num_views, ids.data());
    if (n < 0) return OSFM_ERR_IO;
    if (files_written) *files_written = n;
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

struct osfm_track_table {
    TrackTable t;
};

int osfm_io_load_tracks(const char* path, osfm_track_table** out, int64_t* num_tracks, int64_t* num_features) {
    OSFM_TRY_BEGIN
    if (!path || !out) return OSFM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    osfm_track_table* h = new (std::nothrow) osfm_track_table();
    if (!h) return OSFM_ERR_OUT_OF_MEMORY;
    int const r = load_tracks(path, &h->t);
    if (r != 0) { delete h; return r == 1 ? OSFM_ERR_IO : OSFM_ERR_INVALID_ARGUMENT; }
    if (num_tracks) *num_tracks = static_cast<int64_t>(h->t.offset.size()) - 1;
    if (num_features) *num_features = static_cast<int64_t>(h->t.features.size());
    *out = h;
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

int osfm_io_track_table_get(const osfm_track_table* h, int64_t* track_offset, uint32_t* ids, float* xy,
                            uint32_t* rgb) {
    OSFM_TRY_BEGIN
    if (!h) return OSFM_ERR_INVALID_ARGUMENT;
    TrackTable const& t = h->t;
    if (track_offset) memcpy(track_offset, t.offset.data(), sizeof(int64_t) * t.offset.size());
    for (size_t i = 0; i < t.features.size(); ++i) {
        TrackFeature const& o = t.features[i];
        if (ids) { ids[3 * i] = o.view; ids[3 * i + 1] = o.local_id; ids[3 * i + 2] = o.global_id; }
        if (xy) { xy[2 * i] = o.x; xy[2 * i + 1] = o.y; }
        if (rgb) { rgb[3 * i] = o.r; rgb[3 * i + 1] = o.g; rgb[3 * i + 2] = o.b; }
    }
    return OSFM_OK;
    OSFM_TRY_END(nullptr, OSFM_ERR_INTERNAL)
}

void osfm_io_track_table_free(osfm_track_table* h) { delete h; }

// ---- introspection ----------------------------------------------------------------------

int osfm_match_get_stats(const osfm_matcher* m, osfm_match_stats* out) {
    if (!m || !out) return OSFM_ERR_INVALID_ARGUMENT;
    *out = m->stats;
    for (osfm_matcher const* p : m->peers) {        // a multi-device matcher reports the sum over its devices
        out->kernel_launches += p->stats.kernel_launches;
        out->scan_items += p->stats.scan_items;
        out->candidate_rows += p->stats.candidate_rows;
        out->slow_rows += p->stats.slow_rows;
        out->exact_rows += p->stats.exact_rows;
        out->exact_wide_rows += p->stats.exact_wide_rows;
        out->float_filter_rows += p->stats.float_filter_rows;
        out->float_exact_rows += p->stats.float_exact_rows;
        out->claimed_rows += p->stats.claimed_rows;
        out->reverse_candidate_rows += p->stats.reverse_candidate_rows;
        out->reverse_restricted_pairs += p->stats.reverse_restricted_pairs;
        out->self_check_failures += p->stats.self_check_failures;
    }
    return OSFM_OK;
}

int osfm_match_debug_set_both_directions(osfm_matcher* m, int on) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    m->both_directions = on == 1;
    m->reverse_mode = on == 2 ? 2 : 0;
    for (osfm_matcher* p : m->peers) { p->both_directions = m->both_directions; p->reverse_mode = m->reverse_mode; }
    return OSFM_OK;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_debug_set_float_path(osfm_matcher* m, int mode) {
    if (!m || mode < 0 || mode > 2) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    m->float_mode = mode;
    return OSFM_OK;
}

int osfm_match_debug_set_exact_path(osfm_matcher* m, int mode) {
    if (!m || mode < 0 || mode > 1) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    m->exact_mode = mode;
    for (osfm_matcher* p : m->peers) p->exact_mode = mode;
    return OSFM_OK;
}

int osfm_match_debug_set_scan_mode(osfm_matcher* m, int mode) {
    if (!m || mode < 0 || mode > 2) return OSFM_ERR_INVALID_ARGUMENT;   /* 3-5: dump / trace entry points */
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    m->scan_mode = mode;
    return OSFM_OK;
}

static int debug_dump(osfm_matcher* m, int kind, int view_q, int view_c, int32_t* out, int64_t out_ints, int mode) {
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    OS_TRY(require_committed(m));
    CU_TRY(m, cudaSetDevice(m->device));
    OS_TRY(ensure_views(m, m->num_views - 1));
    if (kind != 0 && kind != 1) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "unknown kind %d", kind);
    OS_TRY(check_view(m, view_q));
    OS_TRY(check_view(m, view_c));
    CU_TRY(m, cudaSetDevice(m->device));
    int const nq = m->kind[kind].n[view_q], nc = m->kind[kind].n[view_c];
    if (nq <= 0 || nc <= 0) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "empty view");
    int64_t const ld = static_cast<int64_t>(kBlockN) * ((nc + kBlockN - 1) / kBlockN);
    int64_t const ints = static_cast<int64_t>(nq) * (mode == 4 ? ld / 2 : ld);
    if (!out || out_ints < ints) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "dump buffer too small");
    int32_t* d = nullptr;
    CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&d), sizeof(int32_t) * ints));
    std::vector<JobSpec> specs(1, JobSpec{view_q, nq, view_c, nc, false, -1});
    std::vector<int64_t> out_row;
    int r = run_jobs(m, kind, specs, out_row, d, ld, mode);
    if (r == OSFM_OK) {
        cudaError_t e = cudaStreamSynchronize(m->stream);
        if (e == cudaSuccess) e = cudaMemcpy(out, d, sizeof(int32_t) * ints, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) r = cuda_fail(m, e, "dump copy");
    }
    cudaFree(d);
    return r;
}

int osfm_match_debug_dump_similarity(osfm_matcher* m, int kind, int view_q, int view_c, int32_t* out, int64_t out_ints) {
    return debug_dump(m, kind, view_q, view_c, out, out_ints, 3);
}

int osfm_match_debug_trace(osfm_matcher* m, const int32_t* pairs, int npairs, int64_t* out, int64_t out_words) {
    OSFM_TRY_BEGIN
    if (!m) return OSFM_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::recursive_mutex> lock(m->mu);
    OS_TRY(require_committed(m));
    CU_TRY(m, cudaSetDevice(m->device));
    OS_TRY(ensure_views(m, m->num_views - 1));
    int64_t const words = static_cast<int64_t>(kScanThreads / 32) * kTraceEvents * 4;   // 20 warps
    if (!out || out_words < words) return fail(m, OSFM_ERR_INVALID_ARGUMENT, "trace buffer too small (%lld words)", (long long)words);
    CU_TRY(m, cudaSetDevice(m->device));
    std::vector<JobSpec> specs;
    for (int i = 0; i < npairs; ++i) {
        int const a = pairs[2 * i], b = pairs[2 * i + 1];
        OS_TRY(check_view(m, a));
        OS_TRY(check_view(m, b));
        specs.push_back({a, m->kind[0].n[a], b, m->kind[0].n[b], false, -1});
        specs.push_back({b, m->kind[0].n[b], a, m->kind[0].n[a], false, -1});
    }
    int64_t* d = nullptr;
    CU_TRY(m, cudaMalloc(reinterpret_cast<void**>(&d), sizeof(int64_t) * words));
    CU_TRY(m, cudaMemset(d, 0, sizeof(int64_t) * words));
    std::vector<int64_t> out_row;
    int r = run_jobs(m, 0, specs, out_row, reinterpret_cast<int32_t*>(d), 0, 5);
    if (r == OSFM_OK) {
        cudaError_t e = cudaStreamSynchronize(m->stream);
        if (e == cudaSuccess) e = cudaMemcpy(out, d, sizeof(int64_t) * words, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) r = cuda_fail(m, e, "trace copy");
    }
    cudaFree(d);
    return r;
    OSFM_TRY_END(m, OSFM_ERR_INTERNAL)
}

int osfm_match_debug_dump_packed(osfm_matcher* m, int kind, int view_q, int view_c, uint32_t* out, int64_t out_words) {
    return debug_dump(m, kind, view_q, view_c, reinterpret_cast<int32_t*>(out), out_words, 4);
}

}  // extern "C"
