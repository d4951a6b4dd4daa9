// float_kernels.cuh -- the float descriptor path: Matching::twoway_match<float>.
//
// Reference (paths relative to /root/reference):
//   float_inner_prod, SSE3 branch            src/mve/sfm/nearest_neighbor.cc:141-176
//   NearestNeighbor<float>::find             src/mve/sfm/nearest_neighbor.cc:272-289
//   Matching::oneway_match<float>            src/mve/sfm/matching.h:114-146
//
// This path is compiled out of ExhaustiveMatching (DISCRETIZE_DESCRIPTORS 1,
// exhaustive_matching.h:21) and only reachable through the static template, so it is built
// for exactness, not for the tensor cores: every inner product is formed in the reference's
// own summation order -- four partial sums over elements k = j (mod 4), each updated with a
// separately rounded multiply and add (no FMA contraction), then (s0 + s1) + (s2 + s3) as
// the two _mm_hadd_ps produce -- so the similarities, and with them the match sets, are
// bit-identical to the reference's SSE3 build rather than merely within its tie tolerance.
//
// One CTA owns 64 query rows and walks over the candidates in tiles of 64; a thread owns a
// 4 x 4 block of similarities (64 partial sums in registers).  Row states (best, index of the
// last best, second best) are order-independent for floats -- there is no 16-bit truncation
// here -- so they are merged across threads at the end.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "common.cuh"

namespace osfm {

constexpr int kFM = 64;            // query rows per CTA
constexpr int kFN = 64;            // candidate rows per tile
constexpr int kFDim = 128;         // padded descriptor length (zeros do not change any sum)
constexpr int kFPitch = kFDim + 4; // shared-memory row pitch in floats
constexpr int kFloatThreads = 256;
constexpr int kFloatSmemBytes = (kFM + kFN) * kFPitch * 4 + kFM * 16 * 16;

struct FloatRowState {
    float b1, b2;
    int i1, pad;
};

// A row's final state -> its entry of the match vector.
__device__ __forceinline__ int float_row_result(float b1, float b2, int i1, float sq_lowe, float sq_dist)
{
    // std::max(0.0f, 2.0f - 2.0f * ip) (nearest_neighbor.cc:287-288), mul and sub separately rounded
    float d1 = __fsub_rn(2.0f, __fmul_rn(2.0f, b1));
    float d2 = __fsub_rn(2.0f, __fmul_rn(2.0f, b2));
    d1 = 0.0f < d1 ? d1 : 0.0f;
    d2 = 0.0f < d2 ? d2 : 0.0f;
    bool ok = !(d1 > sq_dist);                           // matching.h:138
    if (ok && __fdiv_rn(d1, d2) > sq_lowe) ok = false;   // :140-143, NaN accepts
    return ok ? i1 : -1;
}

// set_q: n_q x 128 floats, set_c: n_c x 128 floats (both zero-padded beyond the descriptor
// length).  out[i] = index of the match of query i in set_c, or -1.
// row_list (may be null): only the query rows row_list[0 .. *list_count) are evaluated -- the rows
// the tensor-core filter could not decide (float_tc_kernels.cuh); the grid covers n_q rows and the
// CTAs beyond the list leave at once.  A short list would leave the device to a handful of CTAs
// walking all candidates, so in that mode blockIdx.y slices the candidate tiles: every CTA writes
// its rows' states for its slice to parts[list position * gridDim.y + slice], and
// float_finish_kernel merges the slices (row states merge in any order, see above) and applies the
// tests.
__global__ void __launch_bounds__(kFloatThreads)
float_oneway_kernel(const float* __restrict__ set_q, int n_q, const float* __restrict__ set_c, int n_c,
                    float sq_lowe, float sq_dist, int32_t* __restrict__ out,
                    const int32_t* __restrict__ row_list, const int* __restrict__ list_count,
                    FloatRowState* __restrict__ parts)
{
    extern __shared__ float4 fsmem4[];
    float* const As = reinterpret_cast<float*>(fsmem4);
    float* const Bs = As + kFM * kFPitch;
    FloatRowState* const merge = reinterpret_cast<FloatRowState*>(Bs + kFN * kFPitch);

    int const tx = threadIdx.x & 15;   // column block: candidates tx*4 .. tx*4+3 of a tile
    int const ty = threadIdx.x >> 4;   // row block: queries ty*4 .. ty*4+3
    int const row0 = blockIdx.x * kFM;
    int const n_rows = row_list != nullptr ? min(*list_count, n_q) : n_q;
    if (row0 >= n_rows) return;
    auto query_row = [&](int r) { return row_list != nullptr ? row_list[row0 + r] : row0 + r; };

    // query tile (rows past the end are zero)
    for (int e = threadIdx.x; e < kFM * (kFDim / 4); e += kFloatThreads) {
        int const r = e / (kFDim / 4), c4 = e % (kFDim / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(set_q + static_cast<int64_t>(query_row(r)) * kFDim) + c4);
        *reinterpret_cast<float4*>(As + r * kFPitch + c4 * 4) = v;
    }

    float b1[4], b2[4];
    int i1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { b1[r] = 0.0f; b2[r] = 0.0f; i1[r] = 0; }   // nearest_neighbor.cc:276-279

    int const tiles = (n_c + kFN - 1) / kFN;
    int const tile_lo = parts != nullptr ? static_cast<int>(static_cast<int64_t>(tiles) * blockIdx.y / gridDim.y) : 0;
    int const tile_hi = parts != nullptr ? static_cast<int>(static_cast<int64_t>(tiles) * (blockIdx.y + 1) / gridDim.y) : tiles;
    for (int col0 = tile_lo * kFN; col0 < tile_hi * kFN; col0 += kFN) {
        __syncthreads();   // previous tile fully consumed (and the query tile is visible)
        for (int e = threadIdx.x; e < kFN * (kFDim / 4); e += kFloatThreads) {
            int const r = e / (kFDim / 4), c4 = e % (kFDim / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col0 + r < n_c) v = __ldg(reinterpret_cast<const float4*>(set_c + static_cast<int64_t>(col0 + r) * kFDim) + c4);
            *reinterpret_cast<float4*>(Bs + r * kFPitch + c4 * 4) = v;
        }
        __syncthreads();

        float s[4][4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[r][c][j] = 0.0f;

#pragma unroll 4
        for (int k4 = 0; k4 < kFDim / 4; ++k4) {
            float4 a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(As + (ty * 4 + r) * kFPitch + k4 * 4);
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(Bs + (tx * 4 + c) * kFPitch + k4 * 4);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // sum = _mm_add_ps(sum, _mm_mul_ps(q, e)): separately rounded (:164)
                    s[r][c][0] = __fadd_rn(s[r][c][0], __fmul_rn(a[r].x, b[c].x));
                    s[r][c][1] = __fadd_rn(s[r][c][1], __fmul_rn(a[r].y, b[c].y));
                    s[r][c][2] = __fadd_rn(s[r][c][2], __fmul_rn(a[r].z, b[c].z));
                    s[r][c][3] = __fadd_rn(s[r][c][3], __fmul_rn(a[r].w, b[c].w));
                }
        }

#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int const col = col0 + tx * 4 + c;
                // two _mm_hadd_ps (:165-166): (s0 + s1) + (s2 + s3)
                float const ip = __fadd_rn(__fadd_rn(s[r][c][0], s[r][c][1]), __fadd_rn(s[r][c][2], s[r][c][3]));
                if (col < n_c && ip >= b2[r]) {          // :170-184
                    if (ip >= b1[r]) { b2[r] = b1[r]; b1[r] = ip; i1[r] = col; }
                    else             { b2[r] = ip; }
                }
            }
    }

    // merge the 16 column blocks of every row: best = largest (value, index); second = the
    // largest of all second bests and of the bests that lost
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        FloatRowState st;
        st.b1 = b1[r]; st.b2 = b2[r]; st.i1 = i1[r]; st.pad = 0;
        merge[(ty * 4 + r) * 16 + tx] = st;
    }
    __syncthreads();
    if (threadIdx.x < kFM) {
        int const r = threadIdx.x;
        FloatRowState best = merge[r * 16];
        float second = best.b2;
        for (int t = 1; t < 16; ++t) {
            FloatRowState const o = merge[r * 16 + t];
            second = fmaxf(second, o.b2);
            if (o.b1 > best.b1 || (o.b1 == best.b1 && o.i1 > best.i1)) {
                second = fmaxf(second, best.b1);
                best.b1 = o.b1; best.i1 = o.i1;
            } else {
                second = fmaxf(second, o.b1);
            }
        }
        if (row0 + r < n_rows) {
            if (parts != nullptr) {
                best.b2 = second;
                parts[static_cast<int64_t>(row0 + r) * gridDim.y + blockIdx.y] = best;
            } else {
                out[query_row(r)] = float_row_result(best.b1, second, best.i1, sq_lowe, sq_dist);
            }
        }
    }
}

// The candidate slices of the listed rows (float_oneway_kernel with parts): one thread per row.
__global__ void __launch_bounds__(256) float_finish_kernel(const FloatRowState* __restrict__ parts, int slices,
                                                           const int32_t* __restrict__ row_list,
                                                           const int* __restrict__ list_count, int n_q,
                                                           float sq_lowe, float sq_dist, int32_t* __restrict__ out)
{
    int const k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= min(*list_count, n_q)) return;
    FloatRowState best = parts[static_cast<int64_t>(k) * slices];
    float second = best.b2;
    for (int t = 1; t < slices; ++t) {
        FloatRowState const o = parts[static_cast<int64_t>(k) * slices + t];
        second = fmaxf(second, o.b2);
        if (o.b1 > best.b1 || (o.b1 == best.b1 && o.i1 > best.i1)) {
            second = fmaxf(second, best.b1);
            best.b1 = o.b1; best.i1 = o.i1;
        } else {
            second = fmaxf(second, o.b1);
        }
    }
    out[row_list[k]] = float_row_result(best.b1, second, best.i1, sq_lowe, sq_dist);
}

}  // namespace osfm
