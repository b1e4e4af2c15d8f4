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
// masks).  The expansion is ENTRY-parallel: lane l produces entries l, l + 32, ... whatever the bit pattern -- the masks are
// blobs (a tile word = 32 neighbouring records is mostly inside or mostly outside a support sphere), and a lane that peels the
// bits of its own words ran at 11 of 32 lanes (ncu, 1 400 instructions per particle).  The unit is a BYTE of the mask: a warp
// scan of the word popcounts (plus three popcounts inside the word) gives every byte its first entry; each non-empty byte
// leaves its index as a marker at that entry, a running max-scan over the entries turns the markers into "byte of entry e", and
// the entry's rank inside the byte picks the bit from a 256 x 8 table.  Stores go straight to global memory: slot ids
// coalesced, the transposed uint16 blocks as 2-byte stores that fill a 128-byte block with two consecutive instructions
// (merged in L2).
// MODE 0 writes nbr (slot ids) and the transposed uint16 tile lists (pads = the particle itself), MODE 1 writes cl.
constexpr int TX_THREADS = 256;
constexpr int TX_MAX_K = 1024;                 // marker capacity per warp (MODE 1: 2x): denser scenes take k_bits_expand
constexpr int TX_LUT = 2048;                   // bytes: position of the r-th set bit of byte v at [8 v + r]
// per warp: W words, 4 W uint16 byte offsets (= 2 W ints), cap markers
__host__ __device__ inline int tx_warp_ints(int W, int KP, int mode) { return 3 * ((W + 1) & ~1) + (mode == 0 ? KP : 2 * KP); }   // W rounded to even: 8-byte rows
__host__ __device__ inline int tx_smem_bytes(int W, int KP, int mode) {
    return TILE_HDR + TX_LUT + W * 32 * 4 + (TX_THREADS / 32) * tx_warp_ints(W, KP, mode) * 4;
}
__device__ __forceinline__ int tile_list_pos(int e) { return (e & ~(TILE_BLOCK - 1)) | ((e % TILE_G) * 8) | ((e / TILE_G) & 7); }

template <int MODE>
__global__ void __launch_bounds__(TX_THREADS) k_tile_expand(const uint32_t* __restrict__ bits, int W, int KP /* >= the longest exact list */,
                                                            const int* __restrict__ tab, const unsigned long long* __restrict__ out_start,
                                                            uint32_t* __restrict__ out_slots, const unsigned long long* __restrict__ blk_start,
                                                            uint32_t* __restrict__ t_off, unsigned short* __restrict__ lists16) {
    static_assert(TILE_G == 8 && TILE_BLOCK == 64, "tile_list_pos");
    static_assert(TX_THREADS == 256, "one thread per table row");
    extern __shared__ __align__(128) unsigned char smem[];
    int* row = reinterpret_cast<int*>(smem + 64);
    if (threadIdx.x < TT_STRIDE) row[threadIdx.x] = tab[(size_t)blockIdx.x * TT_STRIDE + threadIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cap = MODE == 0 ? KP : 2 * KP;
    unsigned char* lut = smem + TILE_HDR;
    uint32_t* map = reinterpret_cast<uint32_t*>(smem + TILE_HDR + TX_LUT);
    uint32_t* s_wd = reinterpret_cast<uint32_t*>(smem + TILE_HDR + TX_LUT + (size_t)W * 32 * 4) + (size_t)warp * tx_warp_ints(W, KP, MODE);   // word w
    const int Wp = (W + 1) & ~1;
    unsigned short* s_pre = reinterpret_cast<unsigned short*>(s_wd + Wp);       // first entry of byte u (4 per word)
    int* s_mk = reinterpret_cast<int*>(s_wd + 3 * Wp);                          // marker (u + 1) at the first entry of a non-empty byte u
    for (int e = lane; e < cap; e += 32) s_mk[e] = 0;                    // every marker is cleared again by the entry that reads it
    {
        const uint32_t v = threadIdx.x;
        unsigned long long pk = 0ull;
        int r = 0;
#pragma unroll
        for (int b = 0; b < 8; b++)
            if ((v >> b) & 1u) { pk |= (unsigned long long)b << (8 * r); r++; }
        reinterpret_cast<unsigned long long*>(lut)[v] = pk;
    }
    __syncthreads();
    const int own_start = row[TT_OWN_START], own_count = row[TT_OWN_COUNT], own_pref = row[TT_PREF + 13], total = row[TT_PREF + 27];
    for (int k = warp; k < 27; k += TX_THREADS / 32) {
        const int t0 = row[TT_PREF + k], t1 = row[TT_PREF + k + 1], s0 = row[k];
        for (int t = t0 + lane; t < t1; t += 32) map[t] = (uint32_t)(s0 + (t - t0));
    }
    __syncthreads();
    const int Wc = (total + 31) >> 5;
    const int first = MODE == 0 ? 0 : (own_start + 1) >> 1;
    const int last = MODE == 0 ? own_count : (own_start + own_count) >> 1;
    for (int it = first + warp; it < last; it += TX_THREADS / 32) {
        const int sa = MODE == 0 ? own_start + it : 2 * it;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < TB_MAX_WORDS / 32; q++) {
            if (q * 32 < Wc) {                                            // uniform
                const int w = q * 32 + lane;
                uint32_t v = 0u;
                if (w < Wc) {
                    v = bits[(size_t)sa * W + w];
                    if (MODE == 1) v |= bits[(size_t)(sa + 1) * W + w];
                }
                const int c = __popc(v);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
                const int p0 = cnt + incl - c;
                cnt += __shfl_sync(0xffffffffu, incl, 31);
                if (w < Wc && cnt <= cap) {
                    const int p1 = p0 + __popc(v & 0xffu), p2 = p1 + __popc(v & 0xff00u), p3 = p2 + __popc(v & 0xff0000u);
                    s_wd[w] = v;
                    *reinterpret_cast<uint2*>(s_pre + 4 * w) = make_uint2((uint32_t)p0 | ((uint32_t)p1 << 16), (uint32_t)p2 | ((uint32_t)p3 << 16));
                    if (v & 0xffu) s_mk[p0] = 4 * w + 1;
                    if (v & 0xff00u) s_mk[p1] = 4 * w + 2;
                    if (v & 0xff0000u) s_mk[p2] = 4 * w + 3;
                    if (v & 0xff000000u) s_mk[p3] = 4 * w + 4;
                }
            }
        }
        if (cnt > cap) {                                                  // cannot happen (the host sizes KP from the longest list); keep the markers clean
            __syncwarp();
            for (int e = lane; e < cap; e += 32) s_mk[e] = 0;
            __syncwarp();
            continue;
        }
        __syncwarp();
        const unsigned long long ob = out_start[MODE == 0 ? sa : it];
        unsigned short* l16 = nullptr;
        unsigned short self16 = 0;
        int e_end = cnt;
        if (MODE == 0) {
            const unsigned long long b0 = blk_start[sa];
            if (lane == 0) t_off[sa] = (uint32_t)b0;
            l16 = lists16 + b0 * TILE_BLOCK;
            self16 = (unsigned short)((own_pref + it) * 16);              // pads behind the last entry: the particle itself (x0_ij = 0)
            e_end = (cnt + TILE_BLOCK - 1) / TILE_BLOCK * TILE_BLOCK;
        }
        int carry = 0;
        for (int e0 = 0; e0 < e_end; e0 += 32) {
            const int e = e0 + lane;
            int m = 0;
            if (e < cnt) { m = s_mk[e]; s_mk[e] = 0; }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, m, o); if (lane >= o) m = max(m, x); }
            m = max(m, carry);                                            // m - 1 = the byte entry e belongs to
            carry = __shfl_sync(0xffffffffu, m, 31);
            if (e < cnt) {
                const int u = m - 1;
                const uint32_t byte = (s_wd[u >> 2] >> ((u & 3) * 8)) & 0xffu;
                const int t = u * 8 + lut[byte * 8 + (e - s_pre[u])];
                out_slots[ob + e] = map[t];
                if (MODE == 0) l16[tile_list_pos(e)] = (unsigned short)(t * 16);
            } else if (MODE == 0) {
                l16[tile_list_pos(e)] = self16;
            }
        }
        __syncwarp();                                                     // s_wd / s_pre / s_mk are rewritten for the next item
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
    const unsigned long long units = pairs >> 4;                                                // 16 bits (two radix passes) in units of 16 pairs
    work_key[k] = 0x0000ffffu - (uint32_t)(units < 0x0000ffffull ? units : 0x0000ffffull);      // ascending sort = descending work
    work_val[k] = (uint32_t)k;
}

}  // namespace mis
