// mis_tilebuild.cuh -- the neighbour lists built from shared-memory cell tiles (replaces the per-thread 27-cell walks of
// mis_neighbors.cuh / mis_cluster.cuh whenever a 27-cell neighbourhood fits a tile of <= 4096 particles).
//
// wp.HashGrid's query (sim.py:161,178,203,224) visits every particle of the 27 cells around x0_i; the kernels keep those with
// sqrt(|x0_i - x0_j|^2)/h < 2.  Build once, exactly (same fp32 predicate as dist2_exact, mis_neighbors.cuh), in three steps:
//   k_tile_walk_bits : one CTA per cell, the 27 cells' x0 records staged in shared memory by TMA bulk copies; one warp per own
//                      particle tests all ~1 700 tile records 32 at a time and keeps the outcome as a BITMASK over the tile
//                      (<= 512 B per particle).  The distance tests -- the whole cost of the build -- run once: the neighbour
//                      count is the popcount, and the union list of a cluster of two particles of the same cell is the OR of
//                      their masks (counted here from shared memory).
//   (exclusive scans of the counts: CSR offsets)
//   k_bits_expand    : one warp per particle (or per in-cell cluster) turns mask bits into list entries in ascending tile order,
//                      which IS the 27-cell walk order (cells z-y-x, ascending slot): the exact lists are bit-identical to the
//                      per-thread walk's.  Written at once as slot ids (uint32: set-up kernels, export, cluster lists) and as
//                      transposed uint16 tile offsets (mis_tile.cuh).
// Clusters whose two members sit in different cells (the pair straddles a cell boundary of the slot order) take their union list
// from the members' exact lists (k_cluster_merge).
#pragma once
#include "mis_tile.cuh"

namespace mis {

constexpr int TB_THREADS = 256;
constexpr int TB_MAX_WORDS = 128;              // tile <= 4096 particles
constexpr int TB_OWN = 4;                      // own particles a warp tests against one tile record (register tile)

// masks in shared memory: groups of TB_OWN particles, inside a group word-major, so that the TB_OWN ballots of one tile word are
// ONE 16-byte store
__host__ __device__ inline int tb_smem_bytes(int W, int max_own) {
    return TILE_HDR + W * 32 * 16 + (max_own + TB_OWN - 1) / TB_OWN * TB_OWN * W * 4;
}
__device__ __forceinline__ int tb_mask_idx(int p, int w, int W) { return ((p / TB_OWN) * W + w) * TB_OWN + (p % TB_OWN); }

__global__ void __launch_bounds__(TB_THREADS) k_tile_walk_bits(const uint32_t* __restrict__ order, const int* __restrict__ tab, const float4* __restrict__ x0m,
                                                               int n, float d2_limit, int W, int max_own, uint32_t* __restrict__ bits,
                                                               uint32_t* __restrict__ nbr_count, int* __restrict__ max_k,
                                                               uint32_t* __restrict__ cl_count /* clusters of 2, or null */) {
    static_assert(TB_OWN == 4, "the mask store is one uint4");
    extern __shared__ __align__(128) unsigned char smem[];
    const float4* const planes[1] = {x0m};
    tile_issue<1, 1>(smem, tab + (size_t)(order ? order[blockIdx.x] : blockIdx.x) * TT_STRIDE, planes);
    const int* row = reinterpret_cast<const int*>(smem + 64);
    const int own_start = row[TT_OWN_START], own_count = row[TT_OWN_COUNT], own_pref = row[TT_PREF + 13], total = row[TT_PREF + 27];
    float4* tile = reinterpret_cast<float4*>(smem + TILE_HDR);
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem + TILE_HDR + (size_t)W * 32 * 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wc = (total + 31) >> 5;                                   // words this cell's tile fills; the rest of the W words are zero
    // records behind the tile's end, up to the word boundary: far away (d2 = inf fails the test), so the loop carries no bound check
    for (int t = total + (int)threadIdx.x; t < Wc * 32; t += TB_THREADS) tile[t] = make_float4(3.0e30f, 3.0e30f, 3.0e30f, 0.f);
    tile_wait(smem);
    int kmax = 0;
    for (int p0 = warp * TB_OWN; p0 < own_count; p0 += (TB_THREADS / 32) * TB_OWN) {
        // one LDS.128 per lane and tile word serves TB_OWN own particles (the loop was bound by shared-memory wavefronts with one)
        float4 pi[TB_OWN];
#pragma unroll
        for (int u = 0; u < TB_OWN; u++) pi[u] = tile[own_pref + min(p0 + u, own_count - 1)];
        uint32_t* mrow = masks + (size_t)(p0 / TB_OWN) * W * TB_OWN;
        for (int w = 0; w < Wc; w++) {
            const float4 q = tile[w * 32 + lane];
            uint32_t m[TB_OWN];
#pragma unroll
            for (int u = 0; u < TB_OWN; u++) m[u] = __ballot_sync(0xffffffffu, dist2_exact(pi[u], q) < d2_limit);
            if (lane == 0) *reinterpret_cast<uint4*>(mrow + w * TB_OWN) = make_uint4(m[0], m[1], m[2], m[3]);
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < TB_OWN; u++) {
            const int p = p0 + u;
            if (p >= own_count) break;
            const int self = own_pref + p;                               // the particle passes its own test (d2 = 0): cleared here
            int count = 0;
            for (int w = lane; w < W; w += 32) {
                uint32_t word = 0u;
                if (w < Wc) {
                    word = mrow[w * TB_OWN + u];
                    if (w == (self >> 5)) { word &= ~(1u << (self & 31)); mrow[w * TB_OWN + u] = word; }
                }
                bits[(size_t)(own_start + p) * W + w] = word;
                count += __popc(word);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
            if (lane == 0) nbr_count[own_start + p] = (uint32_t)count;
            kmax = max(kmax, count);
        }
    }
    if (lane == 0 && kmax > 0) atomicMax(max_k, kmax);
    if (!cl_count) return;
    __syncthreads();
    // clusters (2 cc, 2 cc + 1) with both members in this cell: union = OR of the two masks (each mask excludes its own particle
    // and holds the mate iff the mate is a neighbour, exactly the rule of k_cluster_walk)
    const int c0 = (own_start + 1) >> 1, c1 = (own_start + own_count) >> 1;           // cc in [c0, c1): 2 cc >= own_start, 2 cc + 1 < end
    for (int cc = c0 + warp; cc < c1; cc += TB_THREADS / 32) {
        const int pa = 2 * cc - own_start;
        int cnt = 0;
        for (int w = lane; w < Wc; w += 32) cnt += __popc(masks[tb_mask_idx(pa, w, W)] | masks[tb_mask_idx(pa + 1, w, W)]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) cl_count[cc] = (uint32_t)cnt;
    }
}

// ---------------------------------------------------------------- expansion, one CTA per cell
// Mask bits -> list entries in ascending tile order.  The tile-index -> slot map is the same for every particle of a cell, so it
// is built once per CTA in shared memory.  One warp per own particle (MODE 0) or per in-cell cluster of two (MODE 1, OR of the two
// masks): lane l holds words l, l + 32, ... of the mask; a warp scan of the popcounts gives every word its first entry, each lane
// then peels the bits of ITS words into a per-warp staging buffer, and the finished list leaves with coalesced 128-byte stores
// (the one-bit-per-lane expansion it replaces issued ~20 instructions and two scattered stores for 4.6 entries per word).
// MODE 0 writes nbr (slot ids) and the transposed uint16 tile lists (pads = the particle itself), MODE 1 writes cl.
constexpr int TX_THREADS = 256;
constexpr int TX_MAX_K = 512;                  // staging capacity per warp: denser scenes take k_bits_expand
__host__ __device__ inline int tx_smem_bytes(int W, int KP, int mode) { return TILE_HDR + W * 32 * 4 + (TX_THREADS / 32) * KP * (mode == 0 ? 6 : 8); }
__device__ __forceinline__ int tile_list_pos(int e) { return (e & ~(TILE_BLOCK - 1)) | ((e % TILE_G) * 8) | ((e / TILE_G) & 7); }

template <int MODE>
__global__ void __launch_bounds__(TX_THREADS) k_tile_expand(const uint32_t* __restrict__ bits, int W, int KP /* multiple of TILE_BLOCK, >= max k */,
                                                            const int* __restrict__ tab, const unsigned long long* __restrict__ out_start,
                                                            uint32_t* __restrict__ out_slots, const unsigned long long* __restrict__ blk_start,
                                                            uint32_t* __restrict__ t_off, unsigned short* __restrict__ lists16) {
    static_assert(TILE_G == 8 && TILE_BLOCK == 64, "tile_list_pos");
    extern __shared__ __align__(128) unsigned char smem[];
    int* row = reinterpret_cast<int*>(smem + 64);
    if (threadIdx.x < TT_STRIDE) row[threadIdx.x] = tab[(size_t)blockIdx.x * TT_STRIDE + threadIdx.x];
    __syncthreads();
    const int own_start = row[TT_OWN_START], own_count = row[TT_OWN_COUNT], own_pref = row[TT_PREF + 13], total = row[TT_PREF + 27];
    uint32_t* map = reinterpret_cast<uint32_t*>(smem + TILE_HDR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < 27; k += TX_THREADS / 32) {
        const int t0 = row[TT_PREF + k], t1 = row[TT_PREF + k + 1], s0 = row[k];
        for (int t = t0 + lane; t < t1; t += 32) map[t] = (uint32_t)(s0 + (t - t0));
    }
    __syncthreads();
    const int Wc = (total + 31) >> 5;
    const int cap = MODE == 0 ? KP : 2 * KP;
    uint32_t* st_slots = reinterpret_cast<uint32_t*>(smem + TILE_HDR + (size_t)W * 32 * 4 + (size_t)warp * KP * (MODE == 0 ? 6 : 8));
    unsigned short* st16 = reinterpret_cast<unsigned short*>(st_slots + KP);             // MODE 0
    const int first = MODE == 0 ? 0 : (own_start + 1) >> 1;
    const int last = MODE == 0 ? own_count : (own_start + own_count) >> 1;
    for (int it = first + warp; it < last; it += TX_THREADS / 32) {
        const int sa = MODE == 0 ? own_start + it : 2 * it;
        uint32_t word[TB_MAX_WORDS / 32];
        int start[TB_MAX_WORDS / 32];
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < TB_MAX_WORDS / 32; q++) {
            const int w = q * 32 + lane;
            word[q] = 0u; start[q] = 0;
            if (q * 32 < Wc) {                                            // uniform
                if (w < Wc) {
                    word[q] = bits[(size_t)sa * W + w];
                    if (MODE == 1) word[q] |= bits[(size_t)(sa + 1) * W + w];
                }
                const int c = __popc(word[q]);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                start[q] = cnt + incl - c;
                cnt += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (cnt > cap) continue;                                          // cannot happen: the host sizes KP from the largest list
#pragma unroll
        for (int q = 0; q < TB_MAX_WORDS / 32; q++) {
            uint32_t m = word[q];
            int e = start[q];
            const int t0 = (q * 32 + lane) * 32;
            while (m) {
                const int t = t0 + __ffs((int)m) - 1;
                m &= m - 1u;
                st_slots[e] = map[t];
                if (MODE == 0) st16[tile_list_pos(e)] = (unsigned short)(t * 16);
                e++;
            }
        }
        __syncwarp();
        if (MODE == 0) {
            const int nb = (cnt + TILE_BLOCK - 1) / TILE_BLOCK;
            const unsigned short self16 = (unsigned short)((own_pref + it) * 16);
            for (int e = cnt + lane; e < nb * TILE_BLOCK; e += 32) st16[tile_list_pos(e)] = self16;      // x0_ij = 0: no contribution
            __syncwarp();
            const unsigned long long b0 = blk_start[sa];
            if (lane == 0) t_off[sa] = (uint32_t)b0;
            uint4* dst = reinterpret_cast<uint4*>(lists16 + b0 * TILE_BLOCK);
            const uint4* src = reinterpret_cast<const uint4*>(st16);
            for (int i = lane; i < nb * (TILE_BLOCK / 8); i += 32) dst[i] = src[i];
        }
        const unsigned long long ob = out_start[MODE == 0 ? sa : it];
        for (int e = lane; e < cnt; e += 32) out_slots[ob + e] = st_slots[e];
        __syncwarp();
    }
}

// MODE 0: one warp per particle s -> nbr (slot ids) and the transposed uint16 tile list; MODE 1: one warp per cluster of two
// particles of one cell -> cl (slot ids of the union).  Entries in ascending tile index.
template <int MODE>
__global__ void __launch_bounds__(256) k_bits_expand(const uint32_t* __restrict__ bits, int W, const int* __restrict__ cell_lin_sorted,
                                                     const int* __restrict__ cell_start, const unsigned long long* __restrict__ pos,
                                                     const int* __restrict__ tab, int n,
                                                     const unsigned long long* __restrict__ out_start, uint32_t* __restrict__ out_slots,
                                                     const uint32_t* __restrict__ nbr_count, const unsigned long long* __restrict__ blk_start,
                                                     uint32_t* __restrict__ t_off, unsigned short* __restrict__ lists16) {
    __shared__ int s_tab[8][56];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int sa, sb = -1;
    if (MODE == 0) {
        if (item >= n) return;
        sa = item;
    } else {
        sa = 2 * item; sb = sa + 1;
        if (sb >= n) return;                                           // a lone last particle: k_cluster_merge
        if (cell_lin_sorted[sa] != cell_lin_sorted[sb]) return;       // straddling cluster: k_cluster_merge
    }
    const int lin = cell_lin_sorted[sa];
    const int* row = tab + (size_t)pos[cell_start[lin]] * TT_STRIDE;
    if (lane < 28) s_tab[warp][lane] = row[TT_PREF + lane];
    if (lane < 27) s_tab[warp][28 + lane] = row[lane];
    __syncwarp();
    const int* pref = s_tab[warp];
    const int* start = s_tab[warp] + 28;
    const unsigned long long base = out_start[MODE == 0 ? sa : item];
    unsigned short* l16 = nullptr;
    int self_off = 0;
    if (MODE == 0) {
        const unsigned long long b0 = blk_start[sa];
        if (lane == 0) t_off[sa] = (uint32_t)b0;
        l16 = lists16 + b0 * TILE_BLOCK;
        self_off = pref[13] + (sa - start[13]);
    }
    // lane l keeps words l, l + 32, ... of the mask; the warp then expands ONE word at a time, lane b owning bit b: the entry index
    // is the running count plus the rank of the bit inside its word, so the list comes out in ascending tile order whatever the
    // bit pattern, every lane does the same amount of work, and the neighbour cell of a tile index is found once per word (one
    // ballot over the 28 prefix values the lanes hold) and walked forward per bit
    uint32_t mine[TB_MAX_WORDS / 32];
#pragma unroll
    for (int q = 0; q < TB_MAX_WORDS / 32; q++) {
        const int w = q * 32 + lane;
        uint32_t word = 0u;
        if (w < W) {
            word = bits[(size_t)sa * W + w];
            if (MODE == 1) word |= bits[(size_t)sb * W + w];
        }
        mine[q] = word;
    }
    const int my_pref = lane < 28 ? pref[lane] : 0x7fffffff;
    const uint32_t lt = (1u << lane) - 1u;
    int done = 0;
#pragma unroll
    for (int q = 0; q < TB_MAX_WORDS / 32; q++) {
        for (int wl = 0; wl < 32 && q * 32 + wl < W; wl++) {
            const uint32_t word = __shfl_sync(0xffffffffu, mine[q], wl);
            if (word == 0u) continue;
            const int t0 = (q * 32 + wl) * 32;
            int kk = __popc(__ballot_sync(0xffffffffu, my_pref <= t0) & 0x0fffffffu) - 1;       // largest kk with pref[kk] <= t0 (pref[0] = 0)
            if ((word >> lane) & 1u) {
                const int t = t0 + lane;
                while (kk < 26 && pref[kk + 1] <= t) kk++;
                const int e = done + __popc(word & lt);
                out_slots[base + e] = (uint32_t)(start[kk] + (t - pref[kk]));
                if (MODE == 0) {
                    const int blk = e / TILE_BLOCK, r = e % TILE_BLOCK;
                    l16[(size_t)blk * TILE_BLOCK + (r % TILE_G) * 8 + r / TILE_G] = (unsigned short)(t * 16);
                }
            }
            done += __popc(word);
        }
    }
    if (MODE == 0) {
        const int cnt = (int)nbr_count[sa];
        const int nb = (cnt + TILE_BLOCK - 1) / TILE_BLOCK;
        for (int e = cnt + lane; e < nb * TILE_BLOCK; e += 32) {       // pads behind the last entry: the particle itself (x0_ij = 0)
            const int blk = e / TILE_BLOCK, r = e % TILE_BLOCK;
            l16[(size_t)blk * TILE_BLOCK + (r % TILE_G) * 8 + r / TILE_G] = (unsigned short)(self_off * 16);
        }
    }
}

// sort key of the longest-first launch order of the tile kernels: pairs of the cell
__global__ void __launch_bounds__(256) k_tile_work(const int* __restrict__ tab, int n_active, const unsigned long long* __restrict__ nbr_start,
                                                   uint32_t* __restrict__ work_key, uint32_t* __restrict__ work_val) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_active) return;
    const int* row = tab + (size_t)k * TT_STRIDE;
    const unsigned long long pairs = nbr_start[row[TT_OWN_START] + row[TT_OWN_COUNT]] - nbr_start[row[TT_OWN_START]];
    work_key[k] = 0x000fffffu - (uint32_t)(pairs < 0x000fffffull ? pairs : 0x000fffffull);      // ascending sort = descending work
    work_val[k] = (uint32_t)k;
}

}  // namespace mis
