// mis_sort.cuh -- hand-written device primitives for the neighbour-structure build:
// exclusive scan (uint32 counts -> uint32/uint64 offsets) and a stable LSD radix sort
// of (uint32 key, uint32 value) pairs, 8 bits per pass.
//
// Replaces the radix sort + cell_starts/cell_ends construction inside
// wp.HashGrid.build (sim.py:126-127; Warp's native hash grid, not in the reference tree).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mis {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 2048 elements per block

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* >= 32 entries */) {
    // warp scan
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < (blockDim.x >> 5)) ? smem[lane] : T(0);
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        smem[lane] = wi - w;                       // exclusive warp offsets
        if (lane == 31) smem[32] = wi;             // block total
    }
    __syncthreads();
    T res = smem[warp] + incl - v;
    if (total) *total = smem[32];
    __syncthreads();
    return res;
}

// pass 1: per-tile totals
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint32_t* __restrict__ in, long long n, T* __restrict__ tile_sums) {
    __shared__ T sm[40];
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    T s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += (T)in[base + k];
    T tot;
    block_exclusive_scan<T>(s, &tot, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// pass 2: one block scans the tile totals in place (exclusive), writes grand total to *total
template <typename T>
__global__ void __launch_bounds__(1024) scan_tile_offsets(T* __restrict__ tile_sums, int ntiles, T* __restrict__ total) {
    __shared__ T sm[40];
    T carry = 0;
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + threadIdx.x;
        T v = i < ntiles ? tile_sums[i] : T(0);
        T tot;
        T ex = block_exclusive_scan<T>(v, &tot, sm);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// pass 3: per-tile exclusive scan + tile base.  out may alias in when T is uint32.
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const uint32_t* __restrict__ in, long long n, const T* __restrict__ tile_offsets, T* __restrict__ out) {
    __shared__ T sm[40];
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    T s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; s += (T)v[k]; }
    T ex = block_exclusive_scan<T>(s, (T*)nullptr, sm) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += (T)v[k]; }
}

// Exclusive scan of n uint32 counts into n offsets of type T; out[n] = total.
// tile_tmp: >= ceil(n / 2048) entries of T.  Returns the number of kernels launched.
template <typename T>
inline int exclusive_scan(const uint32_t* in, T* out, long long n, T* tile_tmp, cudaStream_t st) {
    if (n <= 0) { cudaMemsetAsync(out, 0, sizeof(T), st); return 0; }
    int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    scan_tile_sums<T><<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile_tmp);
    scan_tile_offsets<T><<<1, 1024, 0, st>>>(tile_tmp, ntiles, out + n);
    scan_apply<T><<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile_tmp, out);
    return 3;
}

// ---------------------------------------------------------------- radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                      // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 2048 keys per block
constexpr int RS_RADIX = 256;

// digit histogram of one tile; hist is digit-major: hist[digit * nblocks + block]
__global__ void __launch_bounds__(RS_THREADS) rs_histogram(const uint32_t* __restrict__ keys, int n, int shift,
                                                           uint32_t* __restrict__ hist, int nblocks) {
    __shared__ uint32_t h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        int idx = base + k * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// stable scatter of one tile.  Element order inside a tile is (warp, round, lane), which
// is ascending index, so equal digits keep their input order.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                         uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                         int n, int shift, const uint32_t* __restrict__ hist_scanned, int nblocks) {
    __shared__ uint32_t wcount[RS_WARPS][RS_RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * RS_RADIX; k += RS_THREADS) (&wcount[0][0])[k] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE + warp * (RS_ITEMS * 32);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int idx = base + r * 32 + lane;
        bool valid = idx < n;
        key[r] = valid ? keys_in[idx] : 0xffffffffu;
        val[r] = valid ? (vals_in ? vals_in[idx] : (uint32_t)idx) : 0u;
        uint32_t digit = (key[r] >> shift) & 255u;
        unsigned vmask = __ballot_sync(0xffffffffu, valid);
        unsigned peers = __match_any_sync(0xffffffffu, digit) & vmask;
        int leader = peers ? (__ffs(peers) - 1) : lane;
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = wcount[warp][digit];
            wcount[warp][digit] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[r] = prev + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // thread t owns digit t: exclusive scan over the warps of this tile + global base
        uint32_t run = hist_scanned[threadIdx.x * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = wcount[w][threadIdx.x];
            wcount[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int idx = base + r * 32 + lane;
        if (idx < n) {
            uint32_t digit = (key[r] >> shift) & 255u;
            uint32_t pos = wcount[warp][digit] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

struct RadixSortTemp {
    uint32_t* keys_alt = nullptr;
    uint32_t* vals_alt = nullptr;
    uint32_t* hist = nullptr;          // 256 * nblocks
    uint32_t* hist_scanned = nullptr;  // 256 * nblocks + 1
    uint32_t* tile_tmp = nullptr;      // scan temp: ceil(256 * nblocks / 2048)
    int nblocks = 0;                   // ceil(n / RS_TILE)
};

// Sorts keys (ascending, stable) carrying values; the first pass takes the input index as
// the value.  `bits` key bits are significant.  An even number of passes is run so the
// result lands back in (keys, vals).  Returns the number of kernels launched.
inline int radix_sort_pairs(uint32_t* keys, uint32_t* vals, int n, int bits, RadixSortTemp& t, cudaStream_t st) {
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    if (passes & 1) passes++;
    int launches = 0;
    int nb = (n + RS_TILE - 1) / RS_TILE;                 // t is sized for the largest sort; a short one (the tile launch order) runs short grids and scans
    if (nb < 1) nb = 1;
    if (nb > t.nblocks) nb = t.nblocks;
    for (int p = 0; p < passes; p++) {
        const bool even = (p & 1) == 0;
        const uint32_t* src_k = even ? keys : t.keys_alt;
        const uint32_t* src_v = (p == 0) ? nullptr : (even ? vals : t.vals_alt);
        uint32_t* dst_k = even ? t.keys_alt : keys;
        uint32_t* dst_v = even ? t.vals_alt : vals;
        const int shift = 8 * p;
        rs_histogram<<<nb, RS_THREADS, 0, st>>>(src_k, n, shift, t.hist, nb);
        launches += 1 + exclusive_scan<uint32_t>(t.hist, t.hist_scanned, (long long)RS_RADIX * nb, t.tile_tmp, st);
        rs_scatter<<<nb, RS_THREADS, 0, st>>>(src_k, src_v, dst_k, dst_v, n, shift, t.hist_scanned, nb);
        launches++;
    }
    return launches;
}

}  // namespace mis
