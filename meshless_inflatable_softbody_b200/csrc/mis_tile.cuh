// mis_tile.cuh -- the per-step gather kernels with the 27-cell neighbourhood of a cell staged in shared memory.
//
// Same arithmetic as mis_cluster.cuh (compute_A_pq / compute_nabla_u / compute_elastic_forces, sim.py:170-235), other
// data movement.  One CTA owns one cell of the hash grid (cell width 2h = the support radius, sim.py:127): every
// neighbour of every particle of that cell lives in the 27 surrounding cells, and because the particles are
// cell-sorted each of those cells is ONE contiguous slot range of every state array.  The CTA therefore copies 27
// ranges per array into shared memory with 1-D TMA bulk copies (cp.async.bulk, completion counted on an mbarrier) and
// all its gathers become shared-memory loads:
//
//   k_deform_t : tile = x0m + xcur          (32 B / particle);  G lanes per particle stride over its EXACT neighbour
//                list and accumulate A_i, B_i; the per-particle tail (polar rotation, F, S) is a separate, fully
//                populated launch (k_deform_fin), so the gather loop carries no rotation code or registers.
//   k_force_t  : tile = x0m + R, S, V       (80 B / particle);  same loop shape; tail = the fused integrate epilogue.
//
// Lists.  Exact per-particle lists (no union with a cluster mate: no wasted pair evaluations), stored as uint16 BYTE
// offsets into the tile (tile index x 16), in blocks of 8 G entries transposed so that lane g of a group loads one
// 16-byte word holding entries g, g + G, ... of the block: in pass e of a block the G lanes of a group read entries
// G e .. G e + G - 1, i.e. mostly consecutive tile records (conflict-free LDS.128), and the index stream is one coalesced
// 16-byte load per lane per block.  2 B per pair instead of 4 B per union entry.
//
// Limits: a tile holds at most CAP particles (template constant, so every plane offset is an immediate); scenes whose
// densest 27-cell neighbourhood exceeds the largest instantiation fall back to the mis_cluster.cuh kernels.
#pragma once
#include "mis_cluster.cuh"

namespace mis {

constexpr int TT_STRIDE = 64;         // ints per active cell: [0..26] start slot of neighbour cell k, [27..54] tile prefix, [55] cell id, [56] own start, [57] own count
constexpr int TT_PREF = 27, TT_CELL = 55, TT_OWN_START = 56, TT_OWN_COUNT = 57;
constexpr int TILE_G = 8;             // lanes per particle
constexpr int TILE_BLOCK = 8 * TILE_G;    // list entries per block (one 16-byte word per lane)

struct TileView {
    const uint32_t* order;            // CTA b works on active cell order[b]: cells sorted by descending pair count (longest first)
    const int* tab;                   // TT_STRIDE ints per active cell
    int n_active;
    const uint4* lists;               // transposed uint16 blocks, TILE_G words per block
    const uint32_t* t_off;            // per slot: first block of its list
    const uint32_t* t_cnt;            // per slot: exact neighbour count
    float4* AB;                       // 5 planes of n float4: A (9), B (9), 2 unused
};

// ---------------------------------------------------------------- PTX helpers (names local to this header)
__device__ __forceinline__ uint32_t t_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void t_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"     // suspend-time hint: a waiting warp sleeps instead of
            "selp.u32 %0, 1, 0, p;\n\t"                                           // spinning through the issue slots of the SM's other CTAs
            "}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    } while (!done);
}
__device__ __forceinline__ void t_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4 + %5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint4 ldg_idx(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---------------------------------------------------------------- build
// flag[s] = 1 where slot s is the first slot of its cell (the cells in slot order are the active cells)
__global__ void __launch_bounds__(256) k_tile_flag(const int* __restrict__ cell_lin_sorted, const int* __restrict__ cell_start, int n, uint32_t* __restrict__ flag) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) flag[s] = cell_start[cell_lin_sorted[s]] == s ? 1u : 0u;
}
// the table row of every non-empty cell, one warp per cell.  by_cell: item = a cell of the dense table (compact scenes: fewer
// cells than particles); else item = a slot, and the first slot of each cell does the work (sparse scenes whose dense table is
// mostly empty).  max_out[0] = largest tile, [1] = largest own count
__global__ void __launch_bounds__(256) k_tile_tab(const uint32_t* __restrict__ flag, const unsigned long long* __restrict__ pos,
                                                  const int* __restrict__ cell_lin_sorted, const int* __restrict__ cell_start,
                                                  const int* __restrict__ cell_end, int3 cdim, int items, int by_cell,
                                                  int* __restrict__ tab, int* __restrict__ max_out) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= items) return;
    int lin, s;
    if (by_cell) {
        lin = item; s = cell_start[lin];
        if (cell_end[lin] <= s) return;
    } else {
        s = item;
        if (!flag[s]) return;
        lin = cell_lin_sorted[s];
    }
    const int k = (int)pos[s];
    const int cx = lin % cdim.x, cy = (lin / cdim.x) % cdim.y, cz = lin / (cdim.x * cdim.y);
    int start = 0, cnt = 0;
    if (lane < 27) {
        const int x = cx + lane % 3 - 1, y = cy + (lane / 3) % 3 - 1, z = cz + lane / 9 - 1;
        if (x >= 0 && x < cdim.x && y >= 0 && y < cdim.y && z >= 0 && z < cdim.z) {
            const int c = (z * cdim.y + y) * cdim.x + x;
            start = cell_start[c]; cnt = cell_end[c] - start;
            if (cnt < 0) cnt = 0;
        }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int* row = tab + (size_t)k * TT_STRIDE;
    if (lane < 27) { row[lane] = start; row[TT_PREF + lane] = incl - cnt; }
    const int total = __shfl_sync(0xffffffffu, incl, 26);
    if (lane == 27) row[TT_PREF + 27] = total;
    if (lane == 28) row[TT_CELL] = lin;
    if (lane == 13) { row[TT_OWN_START] = start; row[TT_OWN_COUNT] = cnt; atomicMax(max_out + 1, cnt); }
    if (lane == 0) atomicMax(max_out, total);
}
// blocks per particle
__global__ void __launch_bounds__(256) k_tile_count(const uint32_t* __restrict__ nbr_count, int n, uint32_t* __restrict__ nblocks) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) nblocks[s] = (nbr_count[s] + TILE_BLOCK - 1) / TILE_BLOCK;
}
// one warp per particle: exact list (slot ids, 27-cell walk order) -> transposed uint16 tile byte offsets; pads = the particle itself
// (zero contribution: x0_ij = 0)
__global__ void __launch_bounds__(256) k_tile_lists(const unsigned long long* __restrict__ nbr_start, const uint32_t* __restrict__ nbr,
                                                    const uint32_t* __restrict__ nbr_count, const int* __restrict__ cell_lin_sorted,
                                                    const int* __restrict__ cell_start, const unsigned long long* __restrict__ pos,
                                                    const int* __restrict__ tab, int3 cdim, int n,
                                                    const unsigned long long* __restrict__ blk_start, uint32_t* __restrict__ t_off,
                                                    unsigned short* __restrict__ lists) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n) return;
    const int lin = cell_lin_sorted[s];
    const int cx = lin % cdim.x, cy = (lin / cdim.x) % cdim.y, cz = lin / (cdim.x * cdim.y);
    const int* row = tab + (size_t)pos[cell_start[lin]] * TT_STRIDE;
    const unsigned long long b = nbr_start[s];
    const int cnt = (int)nbr_count[s];
    const unsigned long long blk0 = blk_start[s];
    if (lane == 0) t_off[s] = (uint32_t)blk0;
    const int nb = (cnt + TILE_BLOCK - 1) / TILE_BLOCK;
    const int self_off = row[TT_PREF + 13] + (s - row[13]);
    unsigned short* out = lists + blk0 * TILE_BLOCK;
    for (int e = lane; e < nb * TILE_BLOCK; e += 32) {
        int off = self_off;
        if (e < cnt) {
            const int t = (int)nbr[b + e];
            const int lt = cell_lin_sorted[t];
            const int tx = lt % cdim.x, ty = (lt / cdim.x) % cdim.y, tz = lt / (cdim.x * cdim.y);
            const int kk = (tz - cz + 1) * 9 + (ty - cy + 1) * 3 + (tx - cx + 1);
            off = row[TT_PREF + kk] + (t - row[kk]);
        }
        // entry e of the list = entry (e % TILE_BLOCK) of block e / TILE_BLOCK -> lane g = r % G, position r / G inside that lane's word
        const int blk = e / TILE_BLOCK, r = e % TILE_BLOCK;
        out[(size_t)blk * TILE_BLOCK + (r % TILE_G) * 8 + r / TILE_G] = (unsigned short)(off * 16);
    }
}

// ---------------------------------------------------------------- shared-memory tile load
// smem layout: [0,8) mbarrier | [64, 64 + 4 TT_STRIDE) table row | [TILE_HDR, ...) NP planes of CAP float4
constexpr int TILE_HDR = 64 + 4 * TT_STRIDE + 64;      // 384: keeps the planes 128-byte aligned
template <int CAP, int NP>
__device__ __forceinline__ void tile_issue(unsigned char* smem, const int* __restrict__ tab_row, const float4* const (&plane)[NP]) {
    int* row = reinterpret_cast<int*>(smem + 64);
    const uint32_t bar = t_smem_u32(smem);
    if (threadIdx.x < TT_STRIDE) row[threadIdx.x] = tab_row[threadIdx.x];
    if (threadIdx.x == 0) {
        t_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0) t_mbar_expect_tx(bar, (uint32_t)row[TT_PREF + 27] * 16u * NP);
        __syncwarp();
        if (lane < 27) {
            const int cnt = row[TT_PREF + lane + 1] - row[TT_PREF + lane];
            if (cnt > 0) {
                const uint32_t dst = t_smem_u32(smem + TILE_HDR) + (uint32_t)row[TT_PREF + lane] * 16u;
#pragma unroll
                for (int q = 0; q < NP; q++) t_bulk_g2s(dst + q * CAP * 16, plane[q] + row[lane], (uint32_t)cnt * 16u, bar);
            }
        }
    }
}
// One warp polls the mbarrier; the others park on the CTA barrier (no issue slots: a spinning CTA took 7.6 % of the deformation
// kernel's executed instructions away from the SM's other CTAs).  The bulk copies' writes are visible to the polling warp through
// the mbarrier's acquire and to everyone else through the barrier that follows it.
__device__ __forceinline__ void tile_wait(unsigned char* smem) {
    if (threadIdx.x < 32) t_mbar_wait(t_smem_u32(smem), 0);
    __syncthreads();
}
template <int CAP, int NP>
__device__ __forceinline__ void tile_load(unsigned char* smem, const int* __restrict__ tab_row, const float4* const (&plane)[NP]) {
    tile_issue<CAP, NP>(smem, tab_row, plane);
    tile_wait(smem);
}

#ifndef MIS_TILE_THREADS_D
#define MIS_TILE_THREADS_D 256
#endif
#ifndef MIS_TILE_THREADS_F
#define MIS_TILE_THREADS_F 512
#endif
constexpr int TILE_THREADS_D = MIS_TILE_THREADS_D;
constexpr int TILE_THREADS_F = MIS_TILE_THREADS_F;

template <int CAP> constexpr int tile_smem_deform() { return TILE_HDR + 2 * CAP * 16; }
template <int CAP> constexpr int tile_smem_force() { return TILE_HDR + 5 * CAP * 16; }

// ---------------------------------------------------------------- k_deform_t
// A_i = sum_j (W_ij m_j) dx_ij d0_ij^T, B_i = sum_j dx_ij (V_j nabla_W_ij)^T  (see k_deform_c), exact lists, shared-memory gathers
template <int CAP, int MINB>
__global__ void __launch_bounds__(TILE_THREADS_D, MINB) k_deform_t(View s, Consts c, TileView t) {
    extern __shared__ __align__(128) unsigned char smem[];
    const float4* const planes[2] = {s.x0m, s.xcur};
    tile_issue<CAP, 2>(smem, t.tab + (size_t)t.order[blockIdx.x] * TT_STRIDE, planes);
    const int* row = reinterpret_cast<const int*>(smem + 64);
    const int own_start = row[TT_OWN_START], own_count = row[TT_OWN_COUNT], own_pref = row[TT_PREF + 13];
    const uint32_t tile = t_smem_u32(smem + TILE_HDR);
    constexpr int XOFF = CAP * 16;
    constexpr int GROUPS = TILE_THREADS_D / TILE_G;
    const int gl = threadIdx.x % TILE_G;
    const int n = s.n;
    // The list metadata and the first index block of a particle are two dependent global loads (~2 us from DRAM): they are
    // issued one particle ahead -- for the first particle while the tile is still in flight, for the next one during the
    // last block of the current one -- so a group never starts a particle with a cold load chain.
    int p = threadIdx.x / TILE_G;
    int cnt = 0;
    const uint4* lst = t.lists;
    uint4 w4 = make_uint4(0, 0, 0, 0);
    if (p < own_count) {
        cnt = (int)t.t_cnt[own_start + p];
        lst = t.lists + (size_t)t.t_off[own_start + p] * TILE_G + gl;
        if (cnt > 0) w4 = ldg_idx(lst);
    }
    tile_wait(smem);
    while (p < own_count) {
        const int i = own_start + p;
        const int pn = p + GROUPS;
        int cnt_n = 0;
        uint32_t off_n = 0;
        if (pn < own_count) { cnt_n = (int)t.t_cnt[own_start + pn]; off_n = t.t_off[own_start + pn]; }
        bool skip = false;
        if (s.push) { const int2 pc = s.push[i]; skip = pc.x == PUSH_GHOST && pc.y >= 2; }   // outer-layer ghost: nobody reads its R, S
        const uint32_t self = tile + (uint32_t)(own_pref + p) * 16u;
        const float4 p0i = lds128<0>(self), pxi = lds128<XOFF>(self);
        const int nb = skip ? 0 : (cnt + TILE_BLOCK - 1) / TILE_BLOCK;
        float A[9], B[9];
#pragma unroll
        for (int k = 0; k < 9; k++) { A[k] = 0.f; B[k] = 0.f; }
        auto eval = [&](uint32_t off) {
            const uint32_t a = tile + off;
            const float4 q0 = lds128<0>(a), qx = lds128<XOFF>(a);
            const float d0x = q0.x - p0i.x, d0y = q0.y - p0i.y, d0z = q0.z - p0i.z;
            float w, beta;
            kernel_W_and_coef(d0x * d0x + d0y * d0y + d0z * d0z, c, w, beta);
            w *= q0.w;                                         // W_ij m_j
            const float dx = qx.x - pxi.x, dy = qx.y - pxi.y, dz = qx.z - pxi.z;
            const float tx = w * d0x, ty = w * d0y, tz = w * d0z;
            A[0] += dx * tx; A[1] += dx * ty; A[2] += dx * tz;
            A[3] += dy * tx; A[4] += dy * ty; A[5] += dy * tz;
            A[6] += dz * tx; A[7] += dz * ty; A[8] += dz * tz;
            const float nbv = -beta * qx.w;                    // nabla_W(x0_i - x0_j) = (-beta) d0, times V_j
            const float gx = nbv * d0x, gy = nbv * d0y, gz = nbv * d0z;
            B[0] += dx * gx; B[1] += dx * gy; B[2] += dx * gz;
            B[3] += dy * gx; B[4] += dy * gy; B[5] += dy * gz;
            B[6] += dz * gx; B[7] += dz * gy; B[8] += dz * gz;
        };
        const uint4* lst_n = t.lists + (size_t)off_n * TILE_G + gl;
        uint4 w4n = make_uint4(0, 0, 0, 0);
        if (nb == 0 && cnt_n > 0) w4n = ldg_idx(lst_n);
        for (int blk = 0; blk < nb; blk++) {
            const uint4 cur = w4;
            const uint32_t wd[4] = {cur.x, cur.y, cur.z, cur.w};
            if (blk + 1 < nb) {
                w4 = ldg_idx(lst + (size_t)(blk + 1) * TILE_G);
#pragma unroll
                for (int e = 0; e < 8; e++) eval((wd[e >> 1] >> (16 * (e & 1))) & 0xffffu);
            } else {
                if (cnt_n > 0) w4n = ldg_idx(lst_n);           // first block of this group's next particle
                const int rem = cnt - blk * TILE_BLOCK;        // entries of the last block; entry of (lane gl, pass e) is e G + gl
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if (e * TILE_G + gl < rem) eval((wd[e >> 1] >> (16 * (e & 1))) & 0xffffu);
            }
        }
        if (!skip) {
#pragma unroll
            for (int k = 0; k < 9; k++) { A[k] = group_sum<TILE_G>(A[k]); B[k] = group_sum<TILE_G>(B[k]); }
            if (gl == 0) {
                t.AB[i] = make_float4(A[0], A[1], A[2], A[3]);
                t.AB[n + i] = make_float4(A[4], A[5], A[6], A[7]);
                t.AB[2 * (size_t)n + i] = make_float4(A[8], B[0], B[1], B[2]);
                t.AB[3 * (size_t)n + i] = make_float4(B[3], B[4], B[5], B[6]);
                t.AB[4 * (size_t)n + i] = make_float4(B[7], B[8], 0.f, 0.f);
            }
        }
        p = pn; cnt = cnt_n; lst = lst_n; w4 = w4n;
    }
}

// per-particle tail of the deformation pass: R = polar(A), N = R^T B - K, F = I + N^T, S (sim.py:185-216); one thread per slot
__global__ void __launch_bounds__(128) k_deform_fin(View s, Consts c, const float4* __restrict__ AB) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = s.n;
    if (i >= n) return;
    if (s.push) { const int2 pc = s.push[i]; if (pc.x == PUSH_GHOST && pc.y >= 2) return; }
    const float4 a0 = AB[i], a1 = AB[n + i], a2 = AB[2 * (size_t)n + i], a3 = AB[3 * (size_t)n + i], a4 = AB[4 * (size_t)n + i];
    const float Am[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x};
    const float Bm[9] = {a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w, a4.x, a4.y};
    float R[9], Nm[9];
    if (c.identity_rot) {
#pragma unroll
        for (int q = 0; q < 9; q++) R[q] = (q % 4 == 0) ? 1.f : 0.f;
    } else {
        polar_rotation(Am, R);
    }
    const float4 k0 = s.Ks[i], k1 = s.Ks[n + i], k2 = s.Ks[2 * (size_t)n + i];
    const float K[9] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w, k2.x};
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++)
            Nm[3 * r + q] = R[0 * 3 + r] * Bm[0 * 3 + q] + R[1 * 3 + r] * Bm[1 * 3 + q] + R[2 * 3 + r] * Bm[2 * 3 + q] - K[3 * r + q];
    const float F[9] = {1.f + Nm[0], Nm[3], Nm[6], Nm[1], 1.f + Nm[4], Nm[7], Nm[2], Nm[5], 1.f + Nm[8]};
    const float4 ml = s.matl[i];
    float S[6];
    stress_svk(F, ml.x, ml.y, ml.z, c, S);
    const float vol = s.xcur[i].w;
    s.RS[i] = make_float4(R[0], R[1], R[2], R[3]);
    s.RS[n + i] = make_float4(R[4], R[5], R[6], R[7]);
    s.RS[2 * (size_t)n + i] = make_float4(R[8], S[0], S[1], S[2]);
    s.RS[3 * (size_t)n + i] = make_float4(S[3], S[4], S[5], vol);
    s.Fd[i] = make_float4(F[0], F[1], F[2], F[3]);
    s.Fd[n + i] = make_float4(F[4], F[5], F[6], F[7]);
    s.Fd[2 * (size_t)n + i] = make_float4(F[8], 0.f, 0.f, 0.f);
    if (s.Apq) {
#pragma unroll
        for (int q = 0; q < 9; q++) s.Apq[9 * (size_t)i + q] = Am[q];
    }
}

// ---------------------------------------------------------------- k_force_t
// force_i = 0.5 V_i [ sum_j V_j R_j F_i S_j nabla_W_ij + R_i F_i S_i G_i ]  (see k_force_c; F_i multiplies S_j: sim.py:233)
template <int CAP>
__global__ void __launch_bounds__(TILE_THREADS_F, 1) k_force_t(View s, Consts c, TileView t, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = s.n;
    const float4* const planes[5] = {s.x0m, s.RS, s.RS + n, s.RS + 2 * (size_t)n, s.RS + 3 * (size_t)n};
    tile_load<CAP, 5>(smem, t.tab + (size_t)t.order[blockIdx.x] * TT_STRIDE, planes);
    const int* row = reinterpret_cast<const int*>(smem + 64);
    const int own_start = row[TT_OWN_START], own_count = row[TT_OWN_COUNT], own_pref = row[TT_PREF + 13];
    const uint32_t tile = t_smem_u32(smem + TILE_HDR);
    constexpr int P = CAP * 16;
    const int gl = threadIdx.x % TILE_G;
    for (int p = threadIdx.x / TILE_G; p < own_count; p += TILE_THREADS_F / TILE_G) {
        const int i = own_start + p;
        if (s.push && s.push[i].x == PUSH_GHOST) continue;     // ghosts are integrated by their owner
        const uint32_t self = tile + (uint32_t)(own_pref + p) * 16u;
        const float4 p0i = lds128<0>(self);
        const float4 f0 = s.Fd[i], f1 = s.Fd[n + i], f2 = s.Fd[2 * (size_t)n + i];
        const float F[9] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x};
        const int cnt = (int)t.t_cnt[i];
        const uint4* __restrict__ lst = t.lists + (size_t)t.t_off[i] * TILE_G + gl;
        const int nb = (cnt + TILE_BLOCK - 1) / TILE_BLOCK;
        float ax = 0.f, ay = 0.f, az = 0.f;
        auto eval = [&](uint32_t off) {
            const uint32_t a = tile + off;
            const float4 q0 = lds128<0>(a), r0 = lds128<P>(a), r1 = lds128<2 * P>(a), r2 = lds128<3 * P>(a), r3 = lds128<4 * P>(a);
            const float d0x = q0.x - p0i.x, d0y = q0.y - p0i.y, d0z = q0.z - p0i.z;
            const float nbv = -kernel_gradW_coef(d0x * d0x + d0y * d0y + d0z * d0z, c) * r3.w;      // V_j folded in
            const float nx = nbv * d0x, ny = nbv * d0y, nz = nbv * d0z;
            const float tx = r2.y * nx + r2.z * ny + r2.w * nz;                                     // t = S_j n
            const float ty = r2.z * nx + r3.x * ny + r3.y * nz;
            const float tz = r2.w * nx + r3.y * ny + r3.z * nz;
            const float ux = F[0] * tx + F[1] * ty + F[2] * tz;                                     // u = F_i t
            const float uy = F[3] * tx + F[4] * ty + F[5] * tz;
            const float uz = F[6] * tx + F[7] * ty + F[8] * tz;
            ax += r0.x * ux + r0.y * uy + r0.z * uz;                                                // a += R_j u
            ay += r0.w * ux + r1.x * uy + r1.y * uz;
            az += r1.z * ux + r1.w * uy + r2.x * uz;
        };
        uint4 w4 = nb > 0 ? ldg_idx(lst) : make_uint4(0, 0, 0, 0);
        for (int blk = 0; blk < nb; blk++) {
            const uint4 cur = w4;
            if (blk + 1 < nb) w4 = ldg_idx(lst + (size_t)(blk + 1) * TILE_G);
            const uint32_t wd[4] = {cur.x, cur.y, cur.z, cur.w};
            if (blk + 1 < nb) {
#pragma unroll
                for (int e = 0; e < 8; e++) eval((wd[e >> 1] >> (16 * (e & 1))) & 0xffffu);
            } else {
                const int rem = cnt - blk * TILE_BLOCK;
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if (e * TILE_G + gl < rem) eval((wd[e >> 1] >> (16 * (e & 1))) & 0xffffu);
            }
        }
        ax = group_sum<TILE_G>(ax); ay = group_sum<TILE_G>(ay); az = group_sum<TILE_G>(az);
        if (gl == 0) {
            const float4 r0 = lds128<P>(self), r1 = lds128<2 * P>(self), r2 = lds128<3 * P>(self), r3 = lds128<4 * P>(self);
            const float4 gs = s.Ks[2 * (size_t)n + i];
            const float gx = gs.y, gy = gs.z, gz = gs.w;
            const float tx = r2.y * gx + r2.z * gy + r2.w * gz;
            const float ty = r2.z * gx + r3.x * gy + r3.y * gz;
            const float tz = r2.w * gx + r3.y * gy + r3.z * gz;
            const float ux = F[0] * tx + F[1] * ty + F[2] * tz;
            const float uy = F[3] * tx + F[4] * ty + F[5] * tz;
            const float uz = F[6] * tx + F[7] * ty + F[8] * tz;
            const float hv = 0.5f * r3.w;
            float3 fel;
            fel.x = hv * (ax + r0.x * ux + r0.y * uy + r0.z * uz);
            fel.y = hv * (ay + r0.w * ux + r1.x * uy + r1.y * uz);
            fel.z = hv * (az + r1.z * ux + r1.w * uy + r2.x * uz);
            const float4 px = s.xcur[i];
            if (!(px.w < 3.0e38f)) fel = make_float3(0.f, 0.f, 0.f);   // isolated particle (rho = 0, V = m/0): the reference loop never runs
            integrate_epilogue(s, c, i, fel, mode, p0i, px);
        }
    }
}

}  // namespace mis
