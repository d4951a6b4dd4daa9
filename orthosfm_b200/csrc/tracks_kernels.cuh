// tracks_kernels.cuh -- feature tracks from pairwise match lists.
//
// What it replaces (reference, paths relative to /root/reference):
//   sfm::bundler::Tracks::compute and remove_invalid_tracks
//   (src/mve/sfm/bundler_tracks.cc:47-203; unify_tracks :23-43).
//
// The reference walks the match lists sequentially, propagates track ids and merges two
// tracks into the larger one when a match connects them.  Whatever the order, the tracks it
// ends with are the connected components of the graph whose nodes are (view, feature) and
// whose edges are the matches; it then drops every component that holds two features of one
// view.  Only the numbering of the tracks and the order of the features inside a track depend
// on the order of the walk.  Here:
//   1. lock-free union-find over the edges (the smaller node id becomes the root),
//   2. every node learns its root; components are sized,
//   3. a component with two nodes of one view is found through an open-addressing hash set of
//      (root, view) keys -- a second insertion of the same key is the conflict,
//   4. the surviving components are numbered in ascending order of their smallest node
//      (a prefix sum over the root flags), and every feature gets its track id or -1:
//      exactly Viewport::track_ids, up to the numbering.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace osfm {

__device__ __forceinline__ int uf_find(int* __restrict__ parent, int x) {
    // path halving; parents only ever decrease, so concurrent unions cannot create a cycle
    int p = parent[x];
    while (p != x) {
        int const g = parent[p];
        if (g != p) parent[x] = g;
        x = p;
        p = g;
    }
    return x;
}

__global__ void __launch_bounds__(256) tracks_init_kernel(int* __restrict__ parent, int n) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = i;
}

// One thread per match.  pair_of_edge is found by binary search in the list offsets.
__global__ void __launch_bounds__(256) tracks_union_kernel(int* __restrict__ parent,
                                                           const int32_t* __restrict__ pair_views,
                                                           const int64_t* __restrict__ list_offset, int npairs,
                                                           const int2* __restrict__ ij, int64_t nedges,
                                                           const int64_t* __restrict__ view_base,
                                                           const int32_t* __restrict__ view_n,
                                                           int* __restrict__ bad_edges)
{
    int64_t const e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= nedges) return;
    int lo = 0, hi = npairs;
    while (hi - lo > 1) {
        int const mid = (lo + hi) >> 1;
        if (list_offset[mid] <= e) lo = mid; else hi = mid;
    }
    int const v1 = pair_views[2 * lo], v2 = pair_views[2 * lo + 1];
    int2 const m = ij[e];
    if (m.x < 0 || m.x >= view_n[v1] || m.y < 0 || m.y >= view_n[v2]) {
        atomicAdd(bad_edges, 1);
        return;
    }
    int a = static_cast<int>(view_base[v1] + m.x);
    int b = static_cast<int>(view_base[v2] + m.y);
    // union: hook the larger root under the smaller one
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) break;
        if (a < b) { int const t = a; a = b; b = t; }      // a > b
        int const old = atomicCAS(parent + a, a, b);
        if (old == a) break;
        a = old;                                           // somebody hooked a meanwhile: retry
    }
}

// root[i] for every node (read-only walk: path compression here would let one thread overwrite
// the final root another thread has just stored with a mere ancestor), component sizes.
__global__ void __launch_bounds__(256) tracks_root_kernel(const int* __restrict__ parent, int n,
                                                          int* __restrict__ root, int* __restrict__ size) {
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = i;
    while (true) {
        int const p = parent[r];
        if (p == r) break;
        r = p;
    }
    root[i] = r;
    atomicAdd(size + r, 1);
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// (root, view) keys of every node in a component of two or more nodes go into a hash set; a
// key that is already there marks the component as conflicting.  table: cap entries
// (power of two), all ~0ull.
__global__ void __launch_bounds__(256) tracks_conflict_kernel(const int* __restrict__ root, int n,
                                                              const int* __restrict__ size,
                                                              const int64_t* __restrict__ view_base, int nviews,
                                                              unsigned long long* __restrict__ table, uint64_t cap_mask,
                                                              int* __restrict__ conflict)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int const r = root[i];
    if (size[r] < 2) return;
    // the view of node i: last view whose base is <= i
    int lo = 0, hi = nviews;
    while (hi - lo > 1) {
        int const mid = (lo + hi) >> 1;
        if (view_base[mid] <= i) lo = mid; else hi = mid;
    }
    unsigned long long const key = (static_cast<unsigned long long>(static_cast<uint32_t>(r)) << 32) | static_cast<uint32_t>(lo);
    uint64_t h = mix64(key) & cap_mask;
    while (true) {
        unsigned long long const old = atomicCAS(table + h, ~0ull, key);
        if (old == ~0ull) return;                 // inserted
        if (old == key) { conflict[r] = 1; return; }
        h = (h + 1) & cap_mask;
    }
}

// flag[i] = 1 if node i is the root of a surviving track.
__global__ void __launch_bounds__(256) tracks_flag_kernel(const int* __restrict__ root, const int* __restrict__ size,
                                                          const int* __restrict__ conflict, int n,
                                                          int* __restrict__ flag, int* __restrict__ num_conflicting)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool const is_root = root[i] == i && size[i] >= 2;
    flag[i] = (is_root && !conflict[i]) ? 1 : 0;
    if (is_root && conflict[i]) atomicAdd(num_conflicting, 1);
}

constexpr int kScanBlock = 1024;     // elements per CTA in the prefix sum (256 threads x 4)

// Exclusive prefix sum in three launches: per-CTA sums, scan of the sums (single CTA), add.
__global__ void __launch_bounds__(256) scan_partial_kernel(const int* __restrict__ in, int n, int* __restrict__ block_sum) {
    __shared__ int s[8];
    int const base = blockIdx.x * kScanBlock;
    int v = 0;
    for (int k = 0; k < 4; ++k) {
        int const i = base + k * 256 + threadIdx.x;
        if (i < n) v += in[i];
    }
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s[w];
        block_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int* __restrict__ block_sum, int nblocks, int* __restrict__ total) {
    // single CTA, sequential over chunks of 1024 (nblocks = n / 1024: 32 K for 32 M nodes)
    __shared__ int s[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int const i = base + threadIdx.x;
        int const v = i < nblocks ? block_sum[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int const add = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
            __syncthreads();
            s[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < nblocks) block_sum[i] = carry + s[threadIdx.x] - v;     // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// track id of every root (exclusive prefix of the flags), then of every feature.
__global__ void __launch_bounds__(256) scan_apply_kernel(const int* __restrict__ flag, int n,
                                                         const int* __restrict__ block_sum, int* __restrict__ id_of_root)
{
    __shared__ int s[256];
    int const base = blockIdx.x * kScanBlock + threadIdx.x * 4;
    int f[4], t = 0;
    for (int k = 0; k < 4; ++k) { f[k] = (base + k < n) ? flag[base + k] : 0; t += f[k]; }
    s[threadIdx.x] = t;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        int const add = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += add;
        __syncthreads();
    }
    int run = block_sum[blockIdx.x] + s[threadIdx.x] - t;
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) id_of_root[base + k] = f[k] ? run : -1;
        run += f[k];
    }
}

__global__ void __launch_bounds__(256) tracks_assign_kernel(const int* __restrict__ root, const int* __restrict__ id_of_root,
                                                            int n, int32_t* __restrict__ track_of_feature)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) track_of_feature[i] = id_of_root[root[i]];
}

}  // namespace osfm
