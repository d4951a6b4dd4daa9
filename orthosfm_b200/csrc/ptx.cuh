// ptx.cuh -- thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences).  No CUTLASS dependency.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace osfm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------- mbarrier

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Makes mbarrier.init visible to the async proxy (TMA / tcgen05.commit).
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Watchdog: a wait that has not completed after ~4 s records which barrier it
// was (code) and traps, so that a protocol bug shows up as a launch failure with
// a diagnostic instead of hanging the GPU until an external timeout.
struct HangReport {
    unsigned int flag;
    unsigned int code;
    unsigned int block;
    unsigned int thread;
    unsigned int parity;
    unsigned int aux;
};
// Points at mapped pinned host memory (set once per process by the host code), so the
// report survives the trap that kills the context.
__device__ HangReport* g_hang_report = nullptr;

__device__ __noinline__ void report_hang(uint32_t code, uint32_t parity, uint32_t aux) {
    HangReport* const r = g_hang_report;
    if (r != nullptr && atomicCAS(&r->flag, 0u, 1u) == 0u) {
        r->code = code;
        r->block = blockIdx.x;
        r->thread = threadIdx.x;
        r->parity = parity;
        r->aux = aux;
        __threadfence_system();
    }
    __trap();
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t code = 0,
                                          uint32_t aux = 0) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        // a failed try has already slept in hardware; look at the clock only now and then
        if ((++spins & 0xffffu) == 0u) {   // rarely: reading the global timer is slow
            uint64_t const now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) report_hang(code, parity, aux);
        }
    }
}

// ---------------------------------------------------------------- TMA

__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// 2-D tiled load global -> shared, completion signalled on an mbarrier (bytes).
// c0 = coordinate in the innermost (contiguous) dimension, c1 = row.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar,
                                            int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM

// Executed by one full warp.  Writes the TMEM base address to *smem_result.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, 8-bit integer operands, int32 accumulate.
// One thread issues on behalf of the CTA.
__device__ __forceinline__ void mma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Arrives (count 1) on an mbarrier once all previously issued tcgen05.mma of this
// thread have completed.  Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane
// (taddr.lane + i), registers v[0..31] = columns taddr.col + 0..31.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, int32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// Same shape with .pack::16b: the low 16 bits of two adjacent columns share one register,
// so the 32 registers cover 64 consecutive columns: v[k] = column taddr.col + 2k (bits 0-15)
// and taddr.col + 2k + 1 (bits 16-31).  (Layout checked on hardware by
// tests/test_gpu_parity.py::test_packed_tmem_layout.)
__device__ __forceinline__ void tmem_ld_32x32b_x32_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// x16 flavour: 16 registers = 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32b_x16_pack16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for a K-major operand tile laid out by TMA with
// SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), base
// 1024-byte aligned.  Advancing along K inside the 128-byte swizzle span is done
// by adding (bytes >> 4) to the start-address field.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);  // start address, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                       // LBO (unused for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // SBO = 1024 B, bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                       // layout type SWIZZLE_128B
    return d;
}

// Same, for 64-byte rows (SURF, K = 64 bytes) laid out with SWIZZLE_64B:
// 8-row groups are 512 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;  // SWIZZLE_64B
    return d;
}

// tcgen05 instruction descriptor for kind::i8 (dense, K-major A and B, int32 D).
// a_signed / b_signed: 0 = unsigned 8-bit, 1 = signed 8-bit.
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n, int a_signed, int b_signed) {
    return (2u << 4)                              // D format: S32
           | (static_cast<uint32_t>(a_signed) << 7)   // A format
           | (static_cast<uint32_t>(b_signed) << 10)  // B format
           | (static_cast<uint32_t>(n >> 3) << 17)    // N >> 3
           | (static_cast<uint32_t>(m >> 4) << 24);   // M >> 4
}

// One lane of the (converged) warp.  Code that issues TMA / tcgen05 instructions runs
// warp-uniformly up to this point, so that their operands live in uniform registers; inside a
// divergent "if (lane == 0)" the compiler has to assume per-lane values and wraps every such
// instruction in a register-to-uniform-register "waterfall" loop of a dozen vector instructions.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// Marks a value that is the same in every lane as such for the compiler.
__device__ __forceinline__ int warp_uniform(int x) { return __shfl_sync(0xffffffffu, x, 0); }
__device__ __forceinline__ uint32_t warp_uniform(uint32_t x) { return __shfl_sync(0xffffffffu, x, 0); }

__device__ __forceinline__ void named_barrier_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace osfm
