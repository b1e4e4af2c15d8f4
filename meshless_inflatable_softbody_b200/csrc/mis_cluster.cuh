// mis_cluster.cuh -- the per-step gather kernels, register-tiled over particle clusters.
//
// One simulation step (loop body sim.py:352-358) is two launches:
//   k_deform_c : compute_A_pq (sim.py:170-183) -> compute_R_i (185-191) ->
//                compute_nabla_u (193-209) -> compute_sigma (212-216): writes R_i, S_i, F_i.
//   k_force_c  : compute_elastic_forces (sim.py:218-235) with R_j, S_j gathered per neighbour
//                instead of re-derived per candidate, fused with part_2 of this step
//                (sim.py:253-258) and part_1 of the next one (sim.py:247-251).
//
// Work decomposition.  Particles are in cell-sorted order with a fine Morton curve inside each
// cell, so C consecutive slots ("a cluster") are spatial neighbours and share most of their
// neighbourhood.  Each cluster has ONE static neighbour list: the union of its members' exact
// lists (built once by k_cluster_walk; the reference is Total-Lagrangian, sim.py:161,178,203,224).
// G lanes cooperate on a cluster and stride over the union list; every gathered neighbour record
// (the L1-bound part of the reference's data flow) is used for all C members from registers.
// A union entry that is not a true neighbour of member p contributes exactly zero to p because the
// cubic-spline kernel and its gradient vanish for q >= 2 (sim.py:139-141,149-151); the member itself
// contributes zero because x0_ij = 0 (outer products with a zero vector).  3x3 / 3-vector partial
// sums are combined with xor-shuffles inside the G-lane group; lane p then finishes member p.
//
// Layout (cell-sorted slot s), float4 everywhere so every access is one 16-byte vector:
//   x0m[s]   = (x0.x, x0.y, x0.z, mass)           static
//   xv[b][s] = (x.x,  x.y,  x.z,  volume)         ping-pong; volume static
//   vel[s], f1[s] (force_1 of the current frame), fel[s], fext[s], freem[s], matl[s] = (mu, lam, ratio, rho)
//   RS: 4 planes of n float4 (plane p at RS + p*n): lanes reading consecutive neighbours touch
//       consecutive 16-byte words of one plane:
//       RS0 = (R00 R01 R02 R10)  RS1 = (R11 R12 R20 R21)  RS2 = (R22 Sxx Sxy Sxz)  RS3 = (Syy Syz Szz V)
//   Fd: 3 planes  Fd0 = (F00 F01 F02 F10)  Fd1 = (F11 F12 F20 F21)  Fd2 = (F22 - - -)
//   Ks: 3 planes, static sums over the exact neighbour list (k_static_K):
//       K_i = sum_j (x0_j - x0_i) (V_j nabla_W_ij)^T  packed like Fd, and Ks2.yzw = sum_j V_j nabla_W_ij
#pragma once
#include "mis_math.cuh"
#include "mis_neighbors.cuh"

namespace mis {

struct View {
    int n;
    const float4* x0m;
    const float4* xcur;
    float4* xnext;
    float4* vel;
    float4* f1;
    float4* fel;
    const float4* fext;
    const float4* freem;
    const float4* matl;
    float4* RS;
    float4* Fd;
    const float4* Ks;                 // static K_i and G_i (3 planes)
    float* Apq;                       // optional (keep_fields), 9 floats per slot
    const unsigned long long* cl_start;   // union list offsets per cluster
    const uint32_t* cl;                   // union lists (slot ids)
    const float4* fcon;                   // optional obstacle-contact force at the current positions (DeepSDF extension)
    // fused halo push (slab-partitioned scenes): per slot two destination codes (peer << 28 | slot on that peer), -1 = none;
    // .x == -2 marks a ghost slot, which only its owner's push may write.  peer_x[p] = the position buffer of the NEXT frame
    // on peer p (peer-mapped memory, NVLink stores).
    const int2* push;
    float4* peer_x[4];
};

constexpr int PUSH_GHOST = -2;
constexpr int PUSH_SLOT_BITS = 28;

// Redundant work of a slab rank: a ghost's force is never used (its owner integrates it), and an outer-layer ghost's R, S neither
// (only owned particles gather them).  role: 1 = any ghost, 2 = outer-layer ghost.  True if every member of the cluster is >= role.
template <int C>
__device__ __forceinline__ bool cluster_is_ghost(const View& s, int cc, int role) {
    if (!s.push) return false;
    bool all = true;
#pragma unroll
    for (int p = 0; p < C; p++) {
        const int i = cc * C + p;
        if (i < s.n) { const int2 pc = s.push[i]; all = all && pc.x == PUSH_GHOST && pc.y >= role; }
    }
    return all;
}

// new position of slot i: local store + stores into the ghost slots of the peers that mirror the particle
__device__ __forceinline__ void store_next(const View& s, int i, float4 v) {
    if (s.push) {
        const int2 pc = s.push[i];
        if (pc.x == PUSH_GHOST) return;
        s.xnext[i] = v;
        if (pc.x >= 0) s.peer_x[pc.x >> PUSH_SLOT_BITS][pc.x & ((1 << PUSH_SLOT_BITS) - 1)] = v;
        if (pc.y >= 0) s.peer_x[pc.y >> PUSH_SLOT_BITS][pc.y & ((1 << PUSH_SLOT_BITS) - 1)] = v;
    } else {
        s.xnext[i] = v;
    }
}

// One exchange = every rank publishes its epoch to its peers and waits for theirs.  Runs after the kernel that pushed
// (stream order): the pushes are complete when this kernel starts; fence + release make them visible system-wide before
// the flag.  A peer's pushes are visible here once its flag is (acquire).  Single block, thread p handles peer p.
struct HaloSync {
    int n_peers;
    int wait;                             // 0: publish only (the host orders the ranks)
    unsigned* epoch;                      // [0] exchanges done by this rank
    unsigned* my_flags;                   // [p] last epoch published by peer p (written remotely)
    unsigned* peer_flag[4];               // my entry in peer p's flag array (peer-mapped)
    int* err;
    unsigned long long timeout_ns;
};
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void __launch_bounds__(32) k_halo_sync(HaloSync h) {
    __shared__ unsigned e_s;
    if (threadIdx.x == 0) {
        const unsigned e = h.epoch[0] + 1u;
        h.epoch[0] = e;
        e_s = e;
    }
    __syncthreads();
    const unsigned e = e_s;
    if ((int)threadIdx.x < h.n_peers) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(h.peer_flag[threadIdx.x]), "r"(e) : "memory");
        const unsigned long long t0 = global_ns();
        const unsigned* f = h.my_flags + threadIdx.x;
        for (; h.wait;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - e) >= 0) break;
            if (*(volatile int*)h.err) break;                                   // an earlier wait timed out: do not stall every later step
            if (global_ns() - t0 > h.timeout_ns) { atomicExch(h.err, 1); break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

// per-slot push table from (caller id, peer, remote slot) triples and the ghost list
__global__ void __launch_bounds__(256) k_push_fill(int2* __restrict__ push, const int* __restrict__ inv_perm, int n_push, const int* __restrict__ ids,
                                                   const int* __restrict__ peer, const int* __restrict__ slot, int n_ghost, const int* __restrict__ ghost_ids,
                                                   const int* __restrict__ ghost_layer) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_push) {
        const int code = (peer[k] << PUSH_SLOT_BITS) | slot[k];
        int* e = (int*)(push + inv_perm[ids[k]]);
        if (atomicCAS(e, -1, code) != -1) atomicCAS(e + 1, -1, code);      // second mirror of the same particle
    }
    if (k < n_ghost) push[inv_perm[ghost_ids[k]]] = make_int2(PUSH_GHOST, ghost_layer ? ghost_layer[k] : 1);
}
__global__ void __launch_bounds__(256) k_slots_of(const int* __restrict__ inv_perm, const int* __restrict__ ids, int count, int* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) out[k] = inv_perm[ids[k]];
}
__global__ void __launch_bounds__(256) k_set_volumes(float4* __restrict__ xv0, float4* __restrict__ xv1, const int* __restrict__ inv_perm,
                                                     const int* __restrict__ ids, int count, const float* __restrict__ vol) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int sl = inv_perm[ids[k]];
    xv0[sl].w = vol[k]; xv1[sl].w = vol[k];
}

enum ForceMode { MODE_PRIME = 0, MODE_STEP = 1, MODE_EULER = 2, MODE_EVAL = 3 };

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifndef MIS_STEP_THREADS
#define MIS_STEP_THREADS 128
#endif
#ifndef MIS_STEP_MIN_BLOCKS
#define MIS_STEP_MIN_BLOCKS 1
#endif
constexpr int STEP_THREADS = MIS_STEP_THREADS;
#ifndef MIS_IDX_AHEAD
#define MIS_IDX_AHEAD 2
#endif
constexpr int IDX_AHEAD = MIS_IDX_AHEAD;
constexpr int LIST_PAD = 1024;                  // zero entries behind the last union list: >= (IDX_AHEAD + 3) * G for every G
// union-list entries are read once per launch.  Keeping them out of L1 (MIS_IDX_NOALLOC) was measured SLOWER on B200
// (deform 155 vs 145 us at n = 1e5): the default is a plain cached load.
__device__ __forceinline__ uint32_t ld_idx(const uint32_t* p) {
#ifdef MIS_IDX_NOALLOC
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}     // union-list indices are fetched this many entries (of one lane) ahead; even, >= 2

// ---------------------------------------------------------------- union lists
// One thread per cluster walks the 27 cells around each distinct member cell (cells already
// covered by an earlier member's walk are skipped) and keeps every slot that is an exact
// neighbour (same predicate as k_nbr_walk) of at least one member.  fill == 0 counts.
template <int C>
__global__ void __launch_bounds__(128) k_cluster_walk(const float4* __restrict__ x0m, const int* __restrict__ cell_lin_sorted,
                                                      const int* __restrict__ cell_start, const int* __restrict__ cell_end,
                                                      int3 cdim, int n, float d2_limit, int fill,
                                                      const unsigned long long* __restrict__ cl_start,
                                                      uint32_t* __restrict__ cl, uint32_t* __restrict__ cl_count) {
    const int nc = (n + C - 1) / C;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int i0 = c * C;
    const int m = min(C, n - i0);
    float4 p[C];
    int mx[C], my[C], mz[C];
#pragma unroll
    for (int q = 0; q < C; q++) {
        const int i = min(i0 + q, n - 1);
        p[q] = x0m[i];
        const int lin = cell_lin_sorted[i];
        mx[q] = lin % cdim.x; my[q] = (lin / cdim.x) % cdim.y; mz[q] = lin / (cdim.x * cdim.y);
    }
    uint32_t count = 0;
    uint32_t* out = fill ? cl + cl_start[c] : nullptr;
#pragma unroll
    for (int q = 0; q < C; q++) {
        if (q >= m) break;
        bool dup = false;
#pragma unroll
        for (int r = 0; r < q; r++) dup = dup || (mx[r] == mx[q] && my[r] == my[q] && mz[r] == mz[q]);
        if (dup) continue;
        for (int dz = -1; dz <= 1; dz++) {
            const int z = mz[q] + dz;
            if (z < 0 || z >= cdim.z) continue;
            for (int dy = -1; dy <= 1; dy++) {
                const int y = my[q] + dy;
                if (y < 0 || y >= cdim.y) continue;
                for (int dx = -1; dx <= 1; dx++) {
                    const int x = mx[q] + dx;
                    if (x < 0 || x >= cdim.x) continue;
                    bool seen = false;
#pragma unroll
                    for (int r = 0; r < q; r++)
                        seen = seen || (abs(x - mx[r]) <= 1 && abs(y - my[r]) <= 1 && abs(z - mz[r]) <= 1);
                    if (seen) continue;
                    const int cell = (z * cdim.y + y) * cdim.x + x;
                    const int b = cell_start[cell], e = cell_end[cell];
                    for (int t = b; t < e; t++) {
                        const float4 pt = x0m[t];
                        bool any = false;
#pragma unroll
                        for (int r = 0; r < C; r++)
                            any = any || (r < m && t != i0 + r && dist2_exact(p[r], pt) < d2_limit);
                        if (any) {
                            if (fill) out[count] = (uint32_t)t;
                            count++;
                        }
                    }
                }
            }
        }
    }
    if (!fill) cl_count[c] = count;
}

// ---------------------------------------------------------------- k_volume
// compute_v_i, sim.py:154-167: rho_i = sum_{j != i} m_j W(x0_i - x0_j), V_i = m_i / rho_i.
// Set-up kernel over the exact per-particle lists.
template <int G>
__global__ void __launch_bounds__(STEP_THREADS) k_volume(const float4* __restrict__ x0m, const unsigned long long* __restrict__ nbr_start,
                                                         const uint32_t* __restrict__ nbr, int n, Consts c, int self_density,
                                                         float4* __restrict__ xv0, float4* __restrict__ xv1, float4* __restrict__ matl) {
    const int gid = (blockIdx.x * STEP_THREADS + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const int i = min(gid, n - 1);
    const float4 pi = x0m[i];
    const unsigned long long b = nbr_start[i];
    const int cnt = (int)(nbr_start[i + 1] - b);
    const uint32_t* __restrict__ lst = nbr + b;
    float rho = 0.f;
    for (int k = gl; k < cnt; k += G) {
        const float4 pj = x0m[lst[k]];
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        rho += pj.w * kernel_W(dx * dx + dy * dy + dz * dz, c);
    }
    rho = group_sum<G>(rho);
    if (self_density) rho += pi.w * c.sigma;          // W(0) = sigma (sim_taichi.py:97)
    if (gl == 0 && gid < n) {
        float vol = pi.w / rho;
        xv0[i].w = vol; xv1[i].w = vol;
        matl[i].w = rho;
    }
}

// Static sums of compute_nabla_u / compute_elastic_forces.  With g_ij = V_j nabla_W(x0_i - x0_j):
//   N_i = sum_j (R_i^T dx_ij - d0_ij) g_ij^T = R_i^T (sum_j dx_ij g_ij^T) - K_i,   K_i = sum_j d0_ij g_ij^T
//   the self term of the force (sim.py:232) needs G_i = sum_j g_ij.
// Both depend on x0 and V only: computed here once per set_mass.
template <int G>
__global__ void __launch_bounds__(STEP_THREADS) k_static_K(const float4* __restrict__ x0m, const float4* __restrict__ xv,
                                                           const unsigned long long* __restrict__ nbr_start,
                                                           const uint32_t* __restrict__ nbr, int n, Consts c, float4* __restrict__ Ks) {
    const int gid = (blockIdx.x * STEP_THREADS + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const int i = min(gid, n - 1);
    const float4 p0i = x0m[i];
    const unsigned long long b = nbr_start[i];
    const int cnt = (int)(nbr_start[i + 1] - b);
    const uint32_t* __restrict__ lst = nbr + b;
    float K[9], g3[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 9; k++) K[k] = 0.f;
    for (int k = gl; k < cnt; k += G) {
        const uint32_t j = lst[k];
        const float4 p0 = x0m[j];
        const float d0x = p0.x - p0i.x, d0y = p0.y - p0i.y, d0z = p0.z - p0i.z;
        const float nb = -kernel_gradW_coef(d0x * d0x + d0y * d0y + d0z * d0z, c) * xv[j].w;
        const float gx = nb * d0x, gy = nb * d0y, gz = nb * d0z;
        K[0] += d0x * gx; K[1] += d0x * gy; K[2] += d0x * gz;
        K[3] += d0y * gx; K[4] += d0y * gy; K[5] += d0y * gz;
        K[6] += d0z * gx; K[7] += d0z * gy; K[8] += d0z * gz;
        g3[0] += gx; g3[1] += gy; g3[2] += gz;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) K[k] = group_sum<G>(K[k]);
#pragma unroll
    for (int k = 0; k < 3; k++) g3[k] = group_sum<G>(g3[k]);
    if (gl == 0 && gid < n) {
        Ks[i] = make_float4(K[0], K[1], K[2], K[3]);
        Ks[n + i] = make_float4(K[4], K[5], K[6], K[7]);
        Ks[2 * (size_t)n + i] = make_float4(K[8], g3[0], g3[1], g3[2]);
    }
}

// ---------------------------------------------------------------- k_deform_c
// One pass over the union list accumulates, for each member i of the cluster,
//   A_i = sum_j (W_ij m_j) dx_ij d0_ij^T      (compute_A_pq, sim.py:170-183)
//   B_i = sum_j dx_ij (V_j nabla_W_ij)^T       (the dynamic half of compute_nabla_u)
// with dx_ij = x_j - x_i, d0_ij = x0_j - x0_i; both weights are scalars times d0_ij.
// Then R_i = polar(A_i), N_i = R_i^T B_i - K_i, F_i = I + N_i^T, S_i = compute_sigma(F_i).
// FAITHFUL2 = true keeps the reference's two-loop evaluation order (u = R^T dx - d0 formed per
// pair, sim.py:207-208); results agree to the fp32 reorder floor.
struct DeformJ { float4 p0, px; };

template <int C, int G, bool FAITHFUL2>
__global__ void __launch_bounds__(STEP_THREADS, MIS_STEP_MIN_BLOCKS) k_deform_c(View s, Consts c) {
    const int n = s.n;
    const int nc = (n + C - 1) / C;
    const int cid = (blockIdx.x * STEP_THREADS + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const int cc = min(cid, nc - 1);
    float4 p0i[C], pxi[C];
#pragma unroll
    for (int p = 0; p < C; p++) {
        const int i = min(cc * C + p, n - 1);
        p0i[p] = s.x0m[i]; pxi[p] = s.xcur[i];
    }
    const unsigned long long b = s.cl_start[cc];
    const bool skip = cluster_is_ghost<C>(s, cc, 2);           // outer-layer ghosts: nobody reads their R, S
    const int cnt = skip ? 0 : (int)(s.cl_start[cc + 1] - b);
    const uint32_t* __restrict__ lst = s.cl + b;
    const float4* __restrict__ x0m = s.x0m;
    const float4* __restrict__ xcur = s.xcur;

    float A[C][9], B[C][9];
#pragma unroll
    for (int p = 0; p < C; p++)
#pragma unroll
        for (int k = 0; k < 9; k++) { A[p][k] = 0.f; B[p][k] = 0.f; }

    // Software pipeline: indices are fetched IDX_AHEAD entries ahead (2, 4 and 8 measured the same on B200: the index stream is not
    // what the loop waits for); the gathered record of entry k + G is in flight while entry k is consumed.  Unrolled by two (dA: even entries of this lane,
    // dB: odd ones) so the in-flight record is never moved between registers.  Index loads are guarded here (measured: 134 vs 140 us unguarded; the force kernel is the other way round, 110.5 vs 112.6 us, and reads past the end of a list: the next cluster's list or the LIST_PAD zero entries behind the last one).
    auto eval = [&](const DeformJ& d) {
#pragma unroll
        for (int p = 0; p < C; p++) {
            const float d0x = d.p0.x - p0i[p].x, d0y = d.p0.y - p0i[p].y, d0z = d.p0.z - p0i[p].z;
            float w, beta;
            kernel_W_and_coef(d0x * d0x + d0y * d0y + d0z * d0z, c, w, beta);
            w *= d.p0.w;                                      // W_ij m_j
            const float dx = d.px.x - pxi[p].x, dy = d.px.y - pxi[p].y, dz = d.px.z - pxi[p].z;
            const float tx = w * d0x, ty = w * d0y, tz = w * d0z;
            A[p][0] += dx * tx; A[p][1] += dx * ty; A[p][2] += dx * tz;
            A[p][3] += dy * tx; A[p][4] += dy * ty; A[p][5] += dy * tz;
            A[p][6] += dz * tx; A[p][7] += dz * ty; A[p][8] += dz * tz;
            if (!FAITHFUL2) {
                // nabla_W(x0_i - x0_j) = beta (x0_i - x0_j) = (-beta) d0
                const float nb = -beta * d.px.w;              // * V_j
                const float gx = nb * d0x, gy = nb * d0y, gz = nb * d0z;
                B[p][0] += dx * gx; B[p][1] += dx * gy; B[p][2] += dx * gz;
                B[p][3] += dy * gx; B[p][4] += dy * gy; B[p][5] += dy * gz;
                B[p][6] += dz * gx; B[p][7] += dz * gy; B[p][8] += dz * gz;
            }
        }
    };
    {
        int k = gl;
        uint32_t q[IDX_AHEAD];                                  // q[t] = index of entry k + (t + 1) G
#pragma unroll
        for (int t = 0; t < IDX_AHEAD; t++) q[t] = (k + (t + 1) * G < cnt) ? ld_idx(lst + k + (t + 1) * G) : 0u;
        DeformJ dA, dB;
        {
            const uint32_t j0 = (k < cnt) ? ld_idx(lst + k) : 0u;
            dA.p0 = x0m[j0]; dA.px = xcur[j0];
        }
        while (k < cnt) {
            dB.p0 = x0m[q[0]]; dB.px = xcur[q[0]];            // entry k + G
            const uint32_t n0 = (k + (IDX_AHEAD + 1) * G < cnt) ? ld_idx(lst + k + (IDX_AHEAD + 1) * G) : 0u;
            eval(dA);
            if (k + G >= cnt) break;
            dA.p0 = x0m[q[1]]; dA.px = xcur[q[1]];            // entry k + 2G
            const uint32_t n1 = (k + (IDX_AHEAD + 2) * G < cnt) ? ld_idx(lst + k + (IDX_AHEAD + 2) * G) : 0u;
            eval(dB);
#pragma unroll
            for (int t = 0; t + 2 < IDX_AHEAD; t++) q[t] = q[t + 2];
            q[IDX_AHEAD - 2] = n0; q[IDX_AHEAD - 1] = n1;
            k += 2 * G;
        }
    }
#pragma unroll
    for (int p = 0; p < C; p++)
#pragma unroll
        for (int q = 0; q < 9; q++) A[p][q] = group_sum<G>(A[p][q]);

    float Am[9], Nm[9], R[9];
    if (FAITHFUL2) {
        // every lane needs every member's rotation for the second loop
        float Rall[C][9], N[C][9];
#pragma unroll
        for (int p = 0; p < C; p++) {
            if (c.identity_rot) {
#pragma unroll
                for (int q = 0; q < 9; q++) Rall[p][q] = (q % 4 == 0) ? 1.f : 0.f;
            } else {
                polar_rotation(A[p], Rall[p]);
            }
#pragma unroll
            for (int q = 0; q < 9; q++) N[p][q] = 0.f;
        }
        for (int kk = gl; kk < cnt; kk += G) {
            const uint32_t j = lst[kk];
            const float4 p0 = x0m[j];
            const float4 px = xcur[j];
#pragma unroll
            for (int p = 0; p < C; p++) {
                const float d0x = p0.x - p0i[p].x, d0y = p0.y - p0i[p].y, d0z = p0.z - p0i[p].z;
                const float nb = -kernel_gradW_coef(d0x * d0x + d0y * d0y + d0z * d0z, c) * px.w;
                const float dx = px.x - pxi[p].x, dy = px.y - pxi[p].y, dz = px.z - pxi[p].z;
                const float* Rp = Rall[p];
                const float ux = Rp[0] * dx + Rp[3] * dy + Rp[6] * dz - d0x;
                const float uy = Rp[1] * dx + Rp[4] * dy + Rp[7] * dz - d0y;
                const float uz = Rp[2] * dx + Rp[5] * dy + Rp[8] * dz - d0z;
                const float gx = nb * d0x, gy = nb * d0y, gz = nb * d0z;
                N[p][0] += ux * gx; N[p][1] += ux * gy; N[p][2] += ux * gz;
                N[p][3] += uy * gx; N[p][4] += uy * gy; N[p][5] += uy * gz;
                N[p][6] += uz * gx; N[p][7] += uz * gy; N[p][8] += uz * gz;
            }
        }
#pragma unroll
        for (int p = 0; p < C; p++)
#pragma unroll
            for (int q = 0; q < 9; q++) N[p][q] = group_sum<G>(N[p][q]);
        // lane p keeps member p
#pragma unroll
        for (int q = 0; q < 9; q++) { Am[q] = A[0][q]; Nm[q] = N[0][q]; R[q] = Rall[0][q]; }
#pragma unroll
        for (int p = 1; p < C; p++)
            if (gl == p) {
#pragma unroll
                for (int q = 0; q < 9; q++) { Am[q] = A[p][q]; Nm[q] = N[p][q]; R[q] = Rall[p][q]; }
            }
    } else {
#pragma unroll
        for (int p = 0; p < C; p++)
#pragma unroll
            for (int q = 0; q < 9; q++) B[p][q] = group_sum<G>(B[p][q]);
        float Bm[9];
#pragma unroll
        for (int q = 0; q < 9; q++) { Am[q] = A[0][q]; Bm[q] = B[0][q]; }
#pragma unroll
        for (int p = 1; p < C; p++)
            if (gl == p) {
#pragma unroll
                for (int q = 0; q < 9; q++) { Am[q] = A[p][q]; Bm[q] = B[p][q]; }
            }
        const int i = cc * C + gl;
        if (gl < C && cid < nc && i < n) {
            if (c.identity_rot) {
#pragma unroll
                for (int q = 0; q < 9; q++) R[q] = (q % 4 == 0) ? 1.f : 0.f;
            } else {
                polar_rotation(Am, R);
            }
            const float4 k0 = s.Ks[i], k1 = s.Ks[n + i], k2 = s.Ks[2 * (size_t)n + i];
            const float K[9] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w, k2.x};
            // N = R^T B - K
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int q = 0; q < 3; q++)
                    Nm[3 * r + q] = R[0 * 3 + r] * Bm[0 * 3 + q] + R[1 * 3 + r] * Bm[1 * 3 + q] + R[2 * 3 + r] * Bm[2 * 3 + q] - K[3 * r + q];
        }
    }

    const int i = cc * C + gl;
    if (gl < C && cid < nc && i < n && !skip) {
        // def_grad = I + N^T
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
}

// ---------------------------------------------------------------- integrate (literal order, no FMA)
__device__ __forceinline__ float3 total_force(float3 fext, float3 fel, float3 v, float y, const Consts& c, float3 fcon) {
    // ((external + elastic) - damping * v) + penalty(x)      sim.py:250,256-257
    // penalty = ground plane (sim.py:238-244) + optional obstacle contact (zero without a DeepSDF obstacle)
    float pen = ground_penalty_y(y, c);
    float3 f;
    f.x = __fadd_rn(__fsub_rn(__fadd_rn(fext.x, fel.x), __fmul_rn(c.damping, v.x)), fcon.x);
    f.y = __fadd_rn(__fsub_rn(__fadd_rn(fext.y, fel.y), __fmul_rn(c.damping, v.y)), __fadd_rn(pen, fcon.y));
    f.z = __fadd_rn(__fsub_rn(__fadd_rn(fext.z, fel.z), __fmul_rn(c.damping, v.z)), fcon.z);
    return f;
}
// x + cw_mul(dt * v + 0.5 * dt * dt * force / m, free)        sim.py:251
__device__ __forceinline__ float part1_axis(float x, float v, float f, float m, float fr, const Consts& c) {
    return __fadd_rn(x, __fmul_rn(__fadd_rn(__fmul_rn(c.dt, v), __fdiv_rn(__fmul_rn(c.half_dt2, f), m)), fr));
}
// v + cw_mul(dt * (f1 + f2) / (2 m), free)                    sim.py:258
__device__ __forceinline__ float part2_axis(float v, float f1, float f2, float m, float fr, const Consts& c) {
    return __fadd_rn(v, __fmul_rn(__fdiv_rn(__fmul_rn(c.dt, __fadd_rn(f1, f2)), __fmul_rn(2.f, m)), fr));
}

__device__ __forceinline__ void integrate_epilogue(const View& s, const Consts& c, int i, float3 fel, int mode,
                                                   float4 p0i, float4 pxi) {
    if (mode == MODE_EVAL) { s.fel[i] = make_float4(fel.x, fel.y, fel.z, 0.f); return; }
    const float m = p0i.w;
    const float3 fext = xyz(s.fext[i]);
    const float3 fr = xyz(s.freem[i]);
    float3 v = xyz(s.vel[i]);
    float3 x = xyz(pxi);
    const float3 fcon = s.fcon ? xyz(s.fcon[i]) : make_float3(0.f, 0.f, 0.f);
    if (mode == MODE_EULER) {
        // sim_taichi.py:161-172: force = ext + el + (-damping v); v' = v + dt f / m * free; x' = x + dt v' * free
        float3 f;
        f.x = __fadd_rn(__fadd_rn(__fadd_rn(fext.x, fel.x), __fmul_rn(-c.damping, v.x)), fcon.x);
        f.y = __fadd_rn(__fadd_rn(__fadd_rn(fext.y, fel.y), __fmul_rn(-c.damping, v.y)), fcon.y);
        f.z = __fadd_rn(__fadd_rn(__fadd_rn(fext.z, fel.z), __fmul_rn(-c.damping, v.z)), fcon.z);
        float3 vn, xn;
        vn.x = __fadd_rn(v.x, __fmul_rn(__fdiv_rn(__fmul_rn(c.dt, f.x), m), fr.x));
        vn.y = __fadd_rn(v.y, __fmul_rn(__fdiv_rn(__fmul_rn(c.dt, f.y), m), fr.y));
        vn.z = __fadd_rn(v.z, __fmul_rn(__fdiv_rn(__fmul_rn(c.dt, f.z), m), fr.z));
        xn.x = __fadd_rn(x.x, __fmul_rn(__fmul_rn(c.dt, vn.x), fr.x));
        xn.y = __fadd_rn(x.y, __fmul_rn(__fmul_rn(c.dt, vn.y), fr.y));
        xn.z = __fadd_rn(x.z, __fmul_rn(__fmul_rn(c.dt, vn.z), fr.z));
        s.vel[i] = make_float4(vn.x, vn.y, vn.z, 0.f);
        s.fel[i] = make_float4(fel.x, fel.y, fel.z, 0.f);
        store_next(s, i, make_float4(xn.x, xn.y, xn.z, pxi.w));
        return;
    }
    if (mode == MODE_STEP) {
        // part_2 of this step: force_1 was stored by the previous part_1 (same inputs, same value)
        const float3 F1 = xyz(s.f1[i]);
        const float3 F2 = total_force(fext, fel, v, x.y, c, fcon);
        v.x = part2_axis(v.x, F1.x, F2.x, m, fr.x, c);
        v.y = part2_axis(v.y, F1.y, F2.y, m, fr.y, c);
        v.z = part2_axis(v.z, F1.z, F2.z, m, fr.z, c);
        s.vel[i] = make_float4(v.x, v.y, v.z, 0.f);
    }
    // part_1 of the next step from (x, v, fel) of the now-current frame
    const float3 F1n = total_force(fext, fel, v, x.y, c, fcon);
    float3 xn;
    xn.x = part1_axis(x.x, v.x, F1n.x, m, fr.x, c);
    xn.y = part1_axis(x.y, v.y, F1n.y, m, fr.y, c);
    xn.z = part1_axis(x.z, v.z, F1n.z, m, fr.z, c);
    s.f1[i] = make_float4(F1n.x, F1n.y, F1n.z, 0.f);
    s.fel[i] = make_float4(fel.x, fel.y, fel.z, 0.f);
    store_next(s, i, make_float4(xn.x, xn.y, xn.z, pxi.w));
}

// ---------------------------------------------------------------- k_force_c
// force_i = sum_j 0.5 (R_j f_ij - R_i f_ji),  f_ij = V_j F_i S_j (V_i nabla_W_ij),
//                                             f_ji = -V_i F_i S_i (V_j nabla_W_ij)
//         = 0.5 V_i [ sum_j V_j R_j F_i S_j nabla_W_ij  +  R_i F_i S_i G_i ],  G_i = sum_j V_j nabla_W_ij (static)
// (F_i, not F_j, multiplies S_j: sim.py:233.)  SYM = true is sim_taichi.py:147-158, where f_ij uses
// F_j: force_i = 0.5 V_i [ sum_j V_j R_j F_j S_j nabla_W_ij + R_i F_i S_i G_i ] (antisymmetric pair term).
struct ForceJ { float4 p0, r0, r1, r2, r3, f0, f1, f2; };   // f0..f2 (F_j) are only loaded by the symmetric (sim_taichi.py) pair force

template <int C, int G, bool SYM>
__device__ __forceinline__ void force_cluster(const View& s, const Consts& c, int mode, int cc, int gl, bool valid) {
    const int n = s.n;
    const float4* __restrict__ x0m = s.x0m;
    const float4* __restrict__ RS0 = s.RS;
    const float4* __restrict__ RS1 = s.RS + n;
    const float4* __restrict__ RS2 = s.RS + 2 * (size_t)n;
    const float4* __restrict__ RS3 = s.RS + 3 * (size_t)n;
    const float4* __restrict__ Fd0 = s.Fd;
    const float4* __restrict__ Fd1 = s.Fd + n;
    const float4* __restrict__ Fd2 = s.Fd + 2 * (size_t)n;
    float3 p0i[C];
    float F[SYM ? 1 : C][9];
#pragma unroll
    for (int p = 0; p < C; p++) {
        const int i = min(cc * C + p, n - 1);
        p0i[p] = xyz(x0m[i]);
        if (!SYM) {
            const float4 f0 = Fd0[i], f1 = Fd1[i], f2 = Fd2[i];
            F[p][0] = f0.x; F[p][1] = f0.y; F[p][2] = f0.z; F[p][3] = f0.w;
            F[p][4] = f1.x; F[p][5] = f1.y; F[p][6] = f1.z; F[p][7] = f1.w; F[p][8] = f2.x;
        }
    }
    const unsigned long long b = s.cl_start[cc];
    const bool skip = cluster_is_ghost<C>(s, cc, 1);           // ghosts are integrated by their owner
    const int cnt = skip ? 0 : (int)(s.cl_start[cc + 1] - b);
    const uint32_t* __restrict__ lst = s.cl + b;

    float a[C][3];
#pragma unroll
    for (int p = 0; p < C; p++) { a[p][0] = 0.f; a[p][1] = 0.f; a[p][2] = 0.f; }

    auto load = [&](ForceJ& d, uint32_t j) {
        d.p0 = x0m[j]; d.r0 = RS0[j]; d.r1 = RS1[j]; d.r2 = RS2[j]; d.r3 = RS3[j];
        if (SYM) { d.f0 = Fd0[j]; d.f1 = Fd1[j]; d.f2 = Fd2[j]; }
    };
    auto eval = [&](const ForceJ& d) {
        float Fj[9];
        if (SYM) {
            Fj[0] = d.f0.x; Fj[1] = d.f0.y; Fj[2] = d.f0.z; Fj[3] = d.f0.w; Fj[4] = d.f1.x; Fj[5] = d.f1.y; Fj[6] = d.f1.z; Fj[7] = d.f1.w; Fj[8] = d.f2.x;
        }
#pragma unroll
        for (int p = 0; p < C; p++) {
            const float d0x = d.p0.x - p0i[p].x, d0y = d.p0.y - p0i[p].y, d0z = d.p0.z - p0i[p].z;
            const float nb = -kernel_gradW_coef(d0x * d0x + d0y * d0y + d0z * d0z, c) * d.r3.w;   // V_j folded in
            const float nx = nb * d0x, ny = nb * d0y, nz = nb * d0z;                                // V_j nabla_W_ij
            // t = S_j n     (S = r2.y r2.z r2.w / r3.x r3.y / r3.z)
            const float tx = d.r2.y * nx + d.r2.z * ny + d.r2.w * nz;
            const float ty = d.r2.z * nx + d.r3.x * ny + d.r3.y * nz;
            const float tz = d.r2.w * nx + d.r3.y * ny + d.r3.z * nz;
            // u = F t
            const float* Fp = SYM ? Fj : F[SYM ? 0 : p];
            const float ux = Fp[0] * tx + Fp[1] * ty + Fp[2] * tz;
            const float uy = Fp[3] * tx + Fp[4] * ty + Fp[5] * tz;
            const float uz = Fp[6] * tx + Fp[7] * ty + Fp[8] * tz;
            // a += R_j u
            a[p][0] += d.r0.x * ux + d.r0.y * uy + d.r0.z * uz;
            a[p][1] += d.r0.w * ux + d.r1.x * uy + d.r1.y * uz;
            a[p][2] += d.r1.z * ux + d.r1.w * uy + d.r2.x * uz;
        }
    };
    {
        // same pipeline as k_deform_c: indices IDX_AHEAD entries ahead, one gathered record in flight, unrolled by two
        int k = gl;
        uint32_t q[IDX_AHEAD];
#pragma unroll
        for (int t = 0; t < IDX_AHEAD; t++) q[t] = ld_idx(lst + k + (t + 1) * G);
        ForceJ dA, dB;
        load(dA, ld_idx(lst + k));
        while (k < cnt) {
            load(dB, q[0]);
            const uint32_t n0 = ld_idx(lst + k + (IDX_AHEAD + 1) * G);
            eval(dA);
            if (k + G >= cnt) break;
            load(dA, q[1]);
            const uint32_t n1 = ld_idx(lst + k + (IDX_AHEAD + 2) * G);
            eval(dB);
#pragma unroll
            for (int t = 0; t + 2 < IDX_AHEAD; t++) q[t] = q[t + 2];
            q[IDX_AHEAD - 2] = n0; q[IDX_AHEAD - 1] = n1;
            k += 2 * G;
        }
    }
#pragma unroll
    for (int p = 0; p < C; p++) {
        a[p][0] = group_sum<G>(a[p][0]); a[p][1] = group_sum<G>(a[p][1]); a[p][2] = group_sum<G>(a[p][2]);
    }
    float ax = a[0][0], ay = a[0][1], az = a[0][2];
#pragma unroll
    for (int p = 1; p < C; p++)
        if (gl == p) { ax = a[p][0]; ay = a[p][1]; az = a[p][2]; }

    const int i = cc * C + gl;
    if (gl < C && valid && i < n && !skip) {
        const float4 r0 = RS0[i], r1 = RS1[i], r2 = RS2[i], r3 = RS3[i];
        const float4 f0 = Fd0[i], f1 = Fd1[i], f2 = Fd2[i];
        const float4 gs = s.Ks[2 * (size_t)n + i];
        const float gx = gs.y, gy = gs.z, gz = gs.w;
        const float tx = r2.y * gx + r2.z * gy + r2.w * gz;
        const float ty = r2.z * gx + r3.x * gy + r3.y * gz;
        const float tz = r2.w * gx + r3.y * gy + r3.z * gz;
        const float ux = f0.x * tx + f0.y * ty + f0.z * tz;
        const float uy = f0.w * tx + f1.x * ty + f1.y * tz;
        const float uz = f1.z * tx + f1.w * ty + f2.x * tz;
        const float hv = 0.5f * r3.w;
        float3 fel;
        fel.x = hv * (ax + r0.x * ux + r0.y * uy + r0.z * uz);
        fel.y = hv * (ay + r0.w * ux + r1.x * uy + r1.y * uz);
        fel.z = hv * (az + r1.z * ux + r1.w * uy + r2.x * uz);
        const float4 p0 = x0m[i];
        const float4 px = s.xcur[i];
        if (!(px.w < 3.0e38f)) fel = make_float3(0.f, 0.f, 0.f);   // isolated particle (rho = 0, V = m/0): the reference loop never runs
        integrate_epilogue(s, c, i, fel, mode, p0, px);
    }
}

template <int C, int G, bool SYM>
__global__ void __launch_bounds__(STEP_THREADS, MIS_STEP_MIN_BLOCKS) k_force_c(View s, Consts c, int mode) {
    const int nc = (s.n + C - 1) / C;
    const int cid = (blockIdx.x * STEP_THREADS + threadIdx.x) / G;
    force_cluster<C, G, SYM>(s, c, mode, min(cid, nc - 1), threadIdx.x % G, cid < nc);
}

// The same gather with the clusters handed out PER SM.  A block of k_force_c takes 16 consecutive clusters, but the blocks that
// are resident on one SM at a time are 148 blocks apart in the slot order: four unrelated neighbourhoods (~140 KB of records each)
// share one L1 and a quarter of the gathers go to L2 (ncu: L1 hit rate 75 %, half of the stall samples on the first use of a
// gathered record).  Here every SM owns a contiguous, equally loaded range of clusters (first[]); each warp of a persistent grid
// reads %smid and takes the next 32 / G clusters of ITS SM's range (one atomic per warp), so the 16 warps of an SM work on ~64
// adjacent clusters -- two cells -- whose neighbourhoods overlap.  A warp whose range is empty steals from the following SMs, so
// the result does not depend on where the hardware put the blocks (or on an SM being busy with the contact chain).
struct SmQueue {
    const int* first;     // nsm + 1 cluster boundaries (equal union-list work per range)
    int* ctr;             // nsm counters, zeroed before the launch
    int nsm;
};
template <int C, int G, bool SYM>
__global__ void __launch_bounds__(STEP_THREADS, MIS_STEP_MIN_BLOCKS) k_force_p(View s, Consts c, int mode, SmQueue q) {
    constexpr int PER_WARP = 32 / G;
    const int lane = threadIdx.x & 31, gl = threadIdx.x % G, gw = lane / G;
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    auto drain = [&](int t) {                                  // take clusters from queue t until it is empty
        const int lo = q.first[t], hi = q.first[t + 1];
        for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(q.ctr + t, PER_WARP);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lo + base >= hi) break;
            const int cc = lo + base + gw;
            force_cluster<C, G, SYM>(s, c, mode, min(cc, hi - 1), gl, cc < hi);
        }
    };
    drain((int)(sm % (unsigned)q.nsm));
    // steal: the 32 lanes look at 32 other queues at once (a plain load each), the warp drains the nearest one that still has work
    for (int k0 = 1; k0 < q.nsm; k0 += 32) {
        const int k = k0 + lane;
        const int t = (int)((sm + (unsigned)k) % (unsigned)q.nsm);
        bool has = false;
        if (k < q.nsm) has = *(volatile int*)(q.ctr + t) < q.first[t + 1] - q.first[t];
        unsigned m = __ballot_sync(0xffffffffu, has);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1u;
            drain(__shfl_sync(0xffffffffu, t, src));
        }
    }
}
// first[t] = the first cluster whose union list starts at or behind t / nsm of all union entries
__global__ void __launch_bounds__(256) k_sm_ranges(const unsigned long long* __restrict__ cl_start, int nc, int nsm, int* __restrict__ first) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nsm) return;
    if (t == nsm) { first[t] = nc; return; }
    const unsigned long long target = cl_start[nc] / (unsigned long long)nsm * (unsigned long long)t;
    int lo = 0, hi = nc;                                      // smallest cc with cl_start[cc] >= target
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cl_start[mid] >= target) hi = mid; else lo = mid + 1; }
    first[t] = lo;
}

// force_1 + part_1 again from the stored elastic force (external force or Dirichlet mask changed between steps)
__global__ void __launch_bounds__(256) k_reintegrate(View s, Consts c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    integrate_epilogue(s, c, i, xyz(s.fel[i]), MODE_PRIME, s.x0m[i], s.xcur[i]);
}

// part_2 / part_1 (or the Euler update) from the STORED elastic force: the epilogue of the force kernel as its own launch.  Scenes
// with an obstacle run the force gather in MODE_EVAL so that it does not have to wait for the contact chain (which only this
// kernel's total force needs): same arithmetic, same operands, same results as the fused epilogue.
__global__ void __launch_bounds__(256) k_integrate(View s, Consts c, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    if (s.push && s.push[i].x == PUSH_GHOST) return;               // ghosts are integrated by their owner
    integrate_epilogue(s, c, i, xyz(s.fel[i]), mode, s.x0m[i], s.xcur[i]);
}

// ---------------------------------------------------------------- small per-particle kernels
// gather caller-order arrays into cell-sorted slots
__global__ void __launch_bounds__(256) k_gather_vec3(const float* __restrict__ src, const uint32_t* __restrict__ perm, int n, float4* __restrict__ dst, int keep_w) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t id = perm[s];
    float4 o = make_float4(src[3 * id], src[3 * id + 1], src[3 * id + 2], 0.f);
    if (keep_w) o.w = dst[s].w;
    dst[s] = o;
}
__global__ void __launch_bounds__(256) k_gather_w(const float* __restrict__ src, const uint32_t* __restrict__ perm, int n, float4* __restrict__ dst) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) dst[s].w = src[perm[s]];
}
// mu, lam from E, nu (sim.py:288-300), literal order
__global__ void __launch_bounds__(256) k_material(const float* __restrict__ E, const float* __restrict__ nu, const uint32_t* __restrict__ perm, int n, float4* __restrict__ matl) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t id = perm[s];
    float e = E[id], v = nu[id];
    float mu = __fdiv_rn(e, __fmul_rn(2.f, __fadd_rn(1.f, v)));
    float lam = __fdiv_rn(__fmul_rn(e, v), __fmul_rn(__fadd_rn(1.f, v), __fsub_rn(1.f, __fmul_rn(2.f, v))));
    matl[s].x = mu; matl[s].y = lam;
}
// ratio = 0.5 tanh(k x) + 0.5 (sim.py:107-110)
__global__ void __launch_bounds__(256) k_design(const float* __restrict__ x, const uint32_t* __restrict__ perm, int n, float tanh_k, float4* __restrict__ matl) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) matl[s].z = 0.5f * tanhf(tanh_k * x[perm[s]]) + 0.5f;
}
// startup (sim.py:261-266): x = x0, v = v0
__global__ void __launch_bounds__(256) k_startup(const float4* __restrict__ x0m, int n, float3 v0, float4* __restrict__ xcur, float4* __restrict__ vel) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float4 p = x0m[s];
    xcur[s] = make_float4(p.x, p.y, p.z, xcur[s].w);
    vel[s] = make_float4(v0.x, v0.y, v0.z, 0.f);
}
// export cell-sorted float4 -> caller-order vec3
__global__ void __launch_bounds__(256) k_export_vec3(const float4* __restrict__ src, const int* __restrict__ inv_perm, int n, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = src[inv_perm[i]];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}
// halo plumbing: subset of particles by caller id <-> packed vec3 buffer
__global__ void __launch_bounds__(256) k_subset_gather(const float4* __restrict__ src, const int* __restrict__ inv_perm, const int* __restrict__ ids, int count, float* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const float4 v = src[inv_perm[ids[k]]];
    out[3 * (size_t)k] = v.x; out[3 * (size_t)k + 1] = v.y; out[3 * (size_t)k + 2] = v.z;
}
__global__ void __launch_bounds__(256) k_subset_scatter(float4* __restrict__ dst, const int* __restrict__ inv_perm, const int* __restrict__ ids, int count, const float* __restrict__ in) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    float4* d = dst + inv_perm[ids[k]];
    d->x = in[3 * (size_t)k]; d->y = in[3 * (size_t)k + 1]; d->z = in[3 * (size_t)k + 2];   // .w (volume) untouched
}
// compute_loss, sim.py:269-273: l += |x_i - xt_i|^2 + time_step |v_i - vt_i|^2 per particle (fp32 terms as in the reference; the
// reference sums them with fp32 atomics in arbitrary order -- here: fp64 block partials, then a fixed-order final sum: deterministic).
__global__ void __launch_bounds__(256) k_loss_partial(const float4* __restrict__ x, const float4* __restrict__ v, const int* __restrict__ inv_perm,
                                                      const float* __restrict__ tx, const float* __restrict__ tv, int n, float dt,
                                                      double* __restrict__ partial) {
    __shared__ double sm[8];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int sl = inv_perm[i];
        const float4 px = x[sl], pv = v[sl];
        const float dx = px.x - tx[3 * (size_t)i], dy = px.y - tx[3 * (size_t)i + 1], dz = px.z - tx[3 * (size_t)i + 2];
        const float ex = pv.x - tv[3 * (size_t)i], ey = pv.y - tv[3 * (size_t)i + 1], ez = pv.z - tv[3 * (size_t)i + 2];
        const float a = dx * dx + dy * dy + dz * dz;                   // wp.length_sq
        const float b = (ex * ex + ey * ey + ez * ez) * dt;
        acc += (double)a + (double)b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sm[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void k_loss_final(const double* __restrict__ partial, int nblocks, double* __restrict__ loss) {
    double t = 0.0;
    for (int b = 0; b < nblocks; b++) t += partial[b];
    loss[0] += t;
}

// export fields: which = 0 R, 1 S (full symmetric 3x3), 2 F, 3 A, 4 rho, 5 vol
__global__ void __launch_bounds__(256) k_export_field(View s, const int* __restrict__ inv_perm, int which, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    int p = inv_perm[i];
    if (which == 0) {
        float4 r0 = s.RS[p], r1 = s.RS[s.n + p], r2 = s.RS[2 * (size_t)s.n + p];
        float R[9] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
        for (int k = 0; k < 9; k++) dst[9 * (size_t)i + k] = R[k];
    } else if (which == 1) {
        float4 r2 = s.RS[2 * (size_t)s.n + p], r3 = s.RS[3 * (size_t)s.n + p];
        float S[9] = {r2.y, r2.z, r2.w, r2.z, r3.x, r3.y, r2.w, r3.y, r3.z};
        for (int k = 0; k < 9; k++) dst[9 * (size_t)i + k] = S[k];
    } else if (which == 2) {
        float4 f0 = s.Fd[p], f1 = s.Fd[s.n + p], f2 = s.Fd[2 * (size_t)s.n + p];
        float F[9] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x};
        for (int k = 0; k < 9; k++) dst[9 * (size_t)i + k] = F[k];
    } else if (which == 3) {
        for (int k = 0; k < 9; k++) dst[9 * (size_t)i + k] = s.Apq[9 * (size_t)p + k];
    } else if (which == 4) {
        dst[i] = s.matl[p].w;
    } else {
        dst[i] = s.xcur[p].w;
    }
}

}  // namespace mis
