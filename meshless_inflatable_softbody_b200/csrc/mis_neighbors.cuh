// mis_neighbors.cuh -- cell binning of the reference positions x0 and the static
// neighbour lists.
//
// Replaces wp.HashGrid.build(init_position, 2h) (sim.py:123-127) and every
// wp.hash_grid_query(grid, x0_i, 2h) walk (sim.py:161,178,203,224).  The reference is
// Total-Lagrangian: all queries are centred on x0, so the neighbour set
//     N(i) = { j != i : sqrt(|x0_i - x0_j|^2) / h < 2 }       (support of W, sim.py:133-151)
// never changes.  It is built once here (exactly, with IEEE fp32 operations in the
// reference's order so that the lists are bit-identical to the oracle's) and the step
// kernels stream it instead of re-walking 27 cells and re-testing ~6.4x too many
// candidates per pass.
#pragma once
#include "mis_math.cuh"
#include <limits.h>

namespace mis {

// integer cell coordinate exactly as Warp's hash grid: int(p * cell_width_inv), truncation.
__device__ __forceinline__ int cell_coord(float p, float inv_cw) { return __float2int_rz(__fmul_rn(p, inv_cw)); }

__device__ __forceinline__ uint32_t part1by2(uint32_t x) {
    x &= 0x000003ffu;
    x = (x ^ (x << 16)) & 0xff0000ffu;
    x = (x ^ (x << 8)) & 0x0300f00fu;
    x = (x ^ (x << 4)) & 0x030c30c3u;
    x = (x ^ (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return part1by2(x) | (part1by2(y) << 1) | (part1by2(z) << 2);
}

// per particle (caller order): Warp's integer cell coords and linear cell index (exported, bit-exact), and the coordinates the
// internal binning uses.  Warp truncates toward zero (int(p / cw)), which makes every cell that touches a coordinate plane twice
// as wide (x in (-cw, cw) is ONE cell): up to 8x the particles in a cell at the origin.  Any binning with cells >= the support
// radius gives the same neighbour sets, so the internal structure bins with floor(): uniform cells, balanced tiles.
__global__ void __launch_bounds__(256) k_cell_coords(const float* __restrict__ x0, int n, float inv_cw,
                                                     int gx, int gy, int gz,
                                                     int* __restrict__ coords, int* __restrict__ bcoords, int* __restrict__ cell_index,
                                                     uint32_t* __restrict__ subkey,
                                                     int* __restrict__ bounds /* min xyz, max xyz of the binning coords */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cx = INT_MAX, cy = INT_MAX, cz = INT_MAX, mx = INT_MIN, my = INT_MIN, mz = INT_MIN;
    if (i < n) {
        const float sx_ = __fmul_rn(x0[3 * i + 0], inv_cw), sy_ = __fmul_rn(x0[3 * i + 1], inv_cw), sz_ = __fmul_rn(x0[3 * i + 2], inv_cw);
        const int tx = __float2int_rz(sx_), ty = __float2int_rz(sy_), tz = __float2int_rz(sz_);      // Warp: truncation
        coords[3 * i + 0] = tx; coords[3 * i + 1] = ty; coords[3 * i + 2] = tz;
        cx = mx = __float2int_rd(sx_); cy = my = __float2int_rd(sy_); cz = mz = __float2int_rd(sz_);  // binning: floor
        bcoords[3 * i + 0] = cx; bcoords[3 * i + 1] = cy; bcoords[3 * i + 2] = cz;
        // position inside the cell in eighths: orders the particles of a cell along a fine Morton curve (locality of the
        // neighbour runs; no effect on results)
        float fx = sx_ - (float)cx, fy = sy_ - (float)cy, fz = sz_ - (float)cz;
        int sx = min(max((int)(fx * 8.f), 0), 7), sy = min(max((int)(fy * 8.f), 0), 7), sz = min(max((int)(fz * 8.f), 0), 7);
        subkey[i] = morton3((uint32_t)sx, (uint32_t)sy, (uint32_t)sz);
        // hash_grid_index: +2^20 origin, clamp at 0, mod dim, x fastest
        const int origin = 1 << 20;
        int hx = max(tx + origin, 0) % gx, hy = max(ty + origin, 0) % gy, hz = max(tz + origin, 0) % gz;
        cell_index[i] = hz * (gx * gy) + hy * gx + hx;
    }
    // warp-reduce the bounds, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx = min(cx, __shfl_xor_sync(0xffffffffu, cx, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cy = min(cy, __shfl_xor_sync(0xffffffffu, cy, o)); my = max(my, __shfl_xor_sync(0xffffffffu, my, o));
        cz = min(cz, __shfl_xor_sync(0xffffffffu, cz, o)); mz = max(mz, __shfl_xor_sync(0xffffffffu, mz, o));
    }
    // an atomic only where it would move the bound: after the first few warps almost none does (31 k warps x 6 atomics on six
    // addresses serialised in L2 and were 90 % of this kernel's time at n = 1e6)
    if ((threadIdx.x & 31) == 0) {
        const volatile int* vb = bounds;
        if (cx < vb[0]) atomicMin(bounds + 0, cx);
        if (cy < vb[1]) atomicMin(bounds + 1, cy);
        if (cz < vb[2]) atomicMin(bounds + 2, cz);
        if (mx > vb[3]) atomicMax(bounds + 3, mx);
        if (my > vb[4]) atomicMax(bounds + 4, my);
        if (mz > vb[5]) atomicMax(bounds + 5, mz);
    }
}

// Morton code with its own bit count per axis: level b of an axis is present only while that axis still has bits (x below y below
// z inside a level, as in morton3).  With equal counts this IS morton3; with unequal ones it is morton3 with the always-zero
// bits of the short axes squeezed out -- the same order, fewer key bits: a scene fits 32 bits whenever its dense cell table
// (<= 2^28 cells) exists, however elongated it is.
__device__ __forceinline__ uint32_t morton3_axes(uint32_t x, uint32_t y, uint32_t z, int3 nbits) {
    uint32_t key = 0u;
    int pos = 0;
    const int mb = max(nbits.x, max(nbits.y, nbits.z));
    for (int b = 0; b < mb; b++) {
        if (b < nbits.x) { key |= ((x >> b) & 1u) << pos; pos++; }
        if (b < nbits.y) { key |= ((y >> b) & 1u) << pos; pos++; }
        if (b < nbits.z) { key |= ((z >> b) & 1u) << pos; pos++; }
    }
    return key;
}
// key = Morton(cell - cmin) << sub_bits | top sub_bits bits of the 9-bit in-cell Morton code
__global__ void __launch_bounds__(256) k_cell_keys(const int* __restrict__ coords, const uint32_t* __restrict__ subkey, int n, int3 cmin,
                                                   int3 nbits, int sub_bits, uint32_t* __restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t cx = (uint32_t)(coords[3 * i] - cmin.x), cy = (uint32_t)(coords[3 * i + 1] - cmin.y), cz = (uint32_t)(coords[3 * i + 2] - cmin.z);
    const uint32_t cell = (nbits.x == nbits.y && nbits.y == nbits.z) ? morton3(cx, cy, cz) : morton3_axes(cx, cy, cz, nbits);
    keys[i] = (cell << sub_bits) | (sub_bits ? subkey[i] >> (9 - sub_bits) : 0u);
}

// after the sort: dense cell table, inverse permutation, cell-sorted x0
__global__ void __launch_bounds__(256) k_cell_table(const uint32_t* __restrict__ keys_sorted, const uint32_t* __restrict__ perm,
                                                    const int* __restrict__ coords, const float* __restrict__ x0, int n,
                                                    int3 cmin, int3 cdim, int sub_bits,
                                                    int* __restrict__ cell_start, int* __restrict__ cell_end,
                                                    int* __restrict__ cell_lin_sorted, int* __restrict__ inv_perm,
                                                    float4* __restrict__ x0m) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t id = perm[s];
    uint32_t key = keys_sorted[s] >> sub_bits;
    int cx = coords[3 * id] - cmin.x, cy = coords[3 * id + 1] - cmin.y, cz = coords[3 * id + 2] - cmin.z;
    int lin = (cz * cdim.y + cy) * cdim.x + cx;
    cell_lin_sorted[s] = lin;
    inv_perm[id] = s;
    if (s == 0 || (keys_sorted[s - 1] >> sub_bits) != key) cell_start[lin] = s;
    if (s == n - 1 || (keys_sorted[s + 1] >> sub_bits) != key) cell_end[lin] = s + 1;
    float* dst = reinterpret_cast<float*>(x0m + s);       // keep .w (mass)
    dst[0] = x0[3 * id]; dst[1] = x0[3 * id + 1]; dst[2] = x0[3 * id + 2];
}

// ---------------------------------------------------------------- in-cell pairing
// The step kernels gather ONE union neighbour list per cluster of consecutive slots, so every slot pair (2k, 2k+1) should be a
// pair of close particles: the union of two support spheres (radius 4 lattice spacings) at distance d holds 1.19x (d = 1 spacing)
// to 1.5x (d = 2.4, what a plain in-cell Morton order gives) the entries of one, and every extra entry is a wasted pair evaluation.
// One warp per cell re-orders the cell's slot range by a greedy nearest-neighbour matching: seed = first unmatched particle along
// the Morton curve, mate = its nearest unmatched particle (ties: lower index).  Pairs are aligned to the GLOBAL slot parity (a cell
// that starts on an odd slot gives its first particle to the pair that straddles the boundary).  Deterministic: a rebuild yields
// the same order.  Cells with more than PAIR_MAX particles keep the Morton order.
constexpr int PAIR_MAX = 512;
__global__ void __launch_bounds__(128) k_pair_cells(const float4* __restrict__ x0m, const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                                                    int ncells, const uint32_t* __restrict__ perm_in, uint32_t* __restrict__ perm_out) {
    __shared__ float4 pos[4][PAIR_MAX];
    __shared__ unsigned short order[4][PAIR_MAX];
    __shared__ uint32_t used[4][PAIR_MAX / 32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * 4 + w;
    if (cell >= ncells) return;
    const int b = cell_start[cell], e = cell_end[cell];
    const int m = e - b;
    if (m <= 0) return;
    if (m > PAIR_MAX) {
        for (int i = lane; i < m; i += 32) perm_out[b + i] = perm_in[b + i];
        return;
    }
    for (int i = lane; i < m; i += 32) pos[w][i] = x0m[b + i];
    if (lane < PAIR_MAX / 32) used[w][lane] = 0u;
    __syncwarp();
    int out = 0;
    auto take = [&](int i) {                       // all lanes call with the same i
        if (lane == 0) { used[w][i >> 5] |= 1u << (i & 31); order[w][out] = (unsigned short)i; }
        out++;
        __syncwarp();
    };
    auto first_unused = [&]() {
        int f = -1;
        for (int wd = 0; wd < (m + 31) / 32 && f < 0; wd++) {
            uint32_t freebits = ~used[w][wd];
            if (wd == (m - 1) / 32 && (m & 31)) freebits &= (1u << (m & 31)) - 1u;
            if (freebits) f = wd * 32 + __ffs(freebits) - 1;
        }
        return f;
    };
    if (b & 1) take(0);                            // odd global slot: half of the pair that straddles the cell boundary
    while (out < m) {
        const int seed = first_unused();
        take(seed);
        if (out >= m) break;
        const float4 ps = pos[w][seed];
        float best = 3.0e38f;
        int bi = 0x7fffffff;
        for (int i = lane; i < m; i += 32) {
            if ((used[w][i >> 5] >> (i & 31)) & 1u) continue;
            const float4 q = pos[w][i];
            const float dx = q.x - ps.x, dy = q.y - ps.y, dz = q.z - ps.z;
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 < best) { best = d2; bi = i; }                  // ascending i: the lowest index wins a tie inside a lane
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        take(bi);
    }
    for (int i = lane; i < m; i += 32) perm_out[b + i] = perm_in[b + order[w][i]];
}
// sorted positions and the inverse permutation for the final slot order (mass in .w stays where it is: the order is reproducible)
__global__ void __launch_bounds__(256) k_apply_order(const uint32_t* __restrict__ perm, const float* __restrict__ x0, int n,
                                                     int* __restrict__ inv_perm, float4* __restrict__ x0m) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t id = perm[s];
    inv_perm[id] = s;
    float* dst = reinterpret_cast<float*>(x0m + s);
    dst[0] = x0[3 * id]; dst[1] = x0[3 * id + 1]; dst[2] = x0[3 * id + 2];
}

// Exact membership test.  d2 is formed as ((dx*dx + dy*dy) + dz*dz) with no FMA; the
// reference predicate sqrt(d2)/h < 2 is monotone in d2, so it equals d2 < d2_limit with
// d2_limit = the smallest float for which the predicate is false (found on the host with
// the same correctly-rounded sqrt and divide).
__device__ __forceinline__ float dist2_exact(float4 a, float4 b) {
    float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// One thread per cell-sorted particle walks its 27 cells.  fill == 0: count; fill == 1: write.
__global__ void __launch_bounds__(128) k_nbr_walk(const float4* __restrict__ x0m, const int* __restrict__ cell_lin_sorted,
                                                  const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                                                  int3 cdim, int n, float d2_limit, int fill,
                                                  const unsigned long long* __restrict__ nbr_start,
                                                  uint32_t* __restrict__ nbr, uint32_t* __restrict__ nbr_count, int* __restrict__ max_k) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int lin = cell_lin_sorted[s];
    int cx = lin % cdim.x, cy = (lin / cdim.x) % cdim.y, cz = lin / (cdim.x * cdim.y);
    float4 p = x0m[s];
    uint32_t count = 0;
    uint32_t* out = fill ? nbr + nbr_start[s] : nullptr;
    for (int dz = -1; dz <= 1; dz++) {
        int z = cz + dz;
        if (z < 0 || z >= cdim.z) continue;
        for (int dy = -1; dy <= 1; dy++) {
            int y = cy + dy;
            if (y < 0 || y >= cdim.y) continue;
            for (int dx = -1; dx <= 1; dx++) {
                int x = cx + dx;
                if (x < 0 || x >= cdim.x) continue;
                int c = (z * cdim.y + y) * cdim.x + x;
                int b = cell_start[c], e = cell_end[c];
                // four candidates per trip: the loads do not depend on the tests, so they overlap (the walk is latency-bound)
                for (int t = b; t < e; t += 4) {
                    const int t1 = min(t + 1, e - 1), t2 = min(t + 2, e - 1), t3 = min(t + 3, e - 1);
                    const float4 q0 = x0m[t], q1 = x0m[t1], q2 = x0m[t2], q3 = x0m[t3];
                    const bool k0 = t != s && dist2_exact(p, q0) < d2_limit;
                    const bool k1 = t + 1 < e && t1 != s && dist2_exact(p, q1) < d2_limit;
                    const bool k2 = t + 2 < e && t2 != s && dist2_exact(p, q2) < d2_limit;
                    const bool k3 = t + 3 < e && t3 != s && dist2_exact(p, q3) < d2_limit;
                    if (fill) {
                        if (k0) out[count] = (uint32_t)t;
                        if (k1) out[count + k0] = (uint32_t)t1;
                        if (k2) out[count + k0 + k1] = (uint32_t)t2;
                        if (k3) out[count + k0 + k1 + k2] = (uint32_t)t3;
                    }
                    count += (uint32_t)k0 + (uint32_t)k1 + (uint32_t)k2 + (uint32_t)k3;
                }
            }
        }
    }
    if (!fill) {
        nbr_count[s] = count;
        atomicMax(max_k, (int)count);
    }
}

// Union neighbour list of every cluster of C consecutive slots from the members' exact lists (one warp per cluster):
//   union = list(member 0) ++ [ j in list(member 1) : j is not a neighbour of member 0 ] ++ ...
// "not a neighbour" is one distance test against the earlier members (same exact predicate), so no search is needed and the work is
// k tests per member instead of 1 728 candidates.  The set equals { j : j is an exact neighbour of at least one member } (a member
// is listed when it is a neighbour of another member; its own contribution is zero in the step kernels).  fill == 0 counts.
template <int C>
__global__ void __launch_bounds__(128) k_cluster_merge(const float4* __restrict__ x0m, const unsigned long long* __restrict__ nbr_start,
                                                       const uint32_t* __restrict__ nbr, int n, float d2_limit, int fill,
                                                       const unsigned long long* __restrict__ cl_start, uint32_t* __restrict__ cl,
                                                       uint32_t* __restrict__ cl_count,
                                                       const int* __restrict__ cell_lin_sorted = nullptr /* given: only clusters that straddle a cell boundary */,
                                                       const int* __restrict__ list = nullptr /* given: warp w works on cluster list[w] */, int list_n = 0) {
    const int nc = (n + C - 1) / C;
    int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (list) {
        if (c >= list_n) return;
        c = list[c];
    }
    if (c >= nc) return;
    const int i0 = c * C;
    const int m = min(C, n - i0);
    if (cell_lin_sorted && m == C) {          // clusters inside one cell get their union from the tile bitmasks (mis_tilebuild.cuh)
        bool same = true;
#pragma unroll
        for (int q = 1; q < C; q++) same = same && cell_lin_sorted[i0 + q] == cell_lin_sorted[i0];
        if (same) return;
    }
    float4 p[C];
#pragma unroll
    for (int q = 0; q < C; q++) p[q] = x0m[min(i0 + q, n - 1)];
    uint32_t count = 0;
    uint32_t* out = fill ? cl + cl_start[c] : nullptr;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int q = 0; q < C; q++) {
        if (q >= m) break;
        const unsigned long long b = nbr_start[i0 + q];
        const int cnt = (int)(nbr_start[i0 + q + 1] - b);
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int k = k0 + lane;
            bool keep = k < cnt;
            uint32_t j = 0;
            if (keep) {
                j = nbr[b + k];
                if (q > 0) {
                    const float4 pj = x0m[j];
#pragma unroll
                    for (int r = 0; r < q; r++)
                        if ((int)j != i0 + r && dist2_exact(p[r], pj) < d2_limit) keep = false;     // already listed by member r
                }
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            if (fill && keep) out[count + __popc(mask & lt)] = j;
            count += __popc(mask);
        }
    }
    if (!fill && lane == 0) cl_count[c] = count;
}

// clusters of two whose members sit in different cells, and a lone last particle: the work list of k_cluster_merge in the tiled
// build (3 % of the clusters; a warp per cluster just to find that out cost 2 x 80 us at n = 1e6)
__global__ void __launch_bounds__(256) k_straddle_list(const int* __restrict__ cell_lin_sorted, int n, int* __restrict__ list, int* __restrict__ count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int nc = (n + 1) / 2;
    bool st = false;
    if (c < nc) { const int a = 2 * c, b = a + 1; st = b >= n || cell_lin_sorted[a] != cell_lin_sorted[b]; }
    const uint32_t m = __ballot_sync(0xffffffffu, st);
    if (!m) return;
    const int leader = __ffs((int)m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (st) list[base + __popc(m & ((1u << lane) - 1u))] = c;
}

// CSR export in caller ids: row i (caller id) <- row inv_perm[i] (sorted), entries mapped by perm
__global__ void __launch_bounds__(256) k_export_counts(const uint32_t* __restrict__ nbr_count, const int* __restrict__ inv_perm, int n,
                                                       uint32_t* __restrict__ counts_orig) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts_orig[i] = nbr_count[inv_perm[i]];
}
__global__ void __launch_bounds__(256) k_export_lists(const unsigned long long* __restrict__ nbr_start, const uint32_t* __restrict__ nbr,
                                                      const int* __restrict__ inv_perm, const uint32_t* __restrict__ perm, int n,
                                                      const long long* __restrict__ offsets_orig, int* __restrict__ out) {
    // one warp per caller row
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    int s = inv_perm[row];
    unsigned long long b = nbr_start[s], e = nbr_start[s + 1];
    long long o = offsets_orig[row];
    for (unsigned long long k = b + lane; k < e; k += 32) out[o + (long long)(k - b)] = (int)perm[nbr[k]];
}

}  // namespace mis
