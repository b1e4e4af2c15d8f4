// mis_step.cuh -- the per-step kernels.
//
// One simulation step (loop body sim.py:352-358) is two launches:
//   k_deform  : compute_A_pq (sim.py:170-183) -> compute_R_i (185-191) ->
//               compute_nabla_u (193-209) -> compute_sigma (212-216), per particle:
//               writes the rotation R_i, the stress S_i and def_grad F_i.
//   k_force   : compute_elastic_forces (sim.py:218-235) with R_j, S_j read per neighbour
//               instead of re-derived per candidate, fused with part_2 of this step
//               (sim.py:253-258) and part_1 of the next one (sim.py:247-251).
//
// Work decomposition: G lanes (8, 16 or 32) cooperate on one particle and stride over
// its static neighbour list; 3x3 / 3-vector partial sums are combined with xor-shuffles
// inside the G-lane group.  Particles are in cell-sorted (Morton) order, so the lanes of
// a group read runs of consecutive float4 records and the groups of a block share their
// neighbourhoods in L1.
//
// Layout (cell-sorted slot s):
//   x0m[s]   = (x0.x, x0.y, x0.z, mass)           static
//   xv[b][s] = (x.x,  x.y,  x.z,  volume)         ping-pong; volume static
//   vel[s]   = (v, -)     f1[s] = (force_1 of the current frame, -)
//   fel[s]   = (elastic force of the current frame, -)
//   fext[s], freem[s] = external force, Dirichlet mask      matl[s] = (mu, lam, ratio, rho)
//   RS: 4 float4 planes of n entries (plane p at RS + p*n), so that the lanes of a group
//       reading consecutive neighbours touch consecutive 16-byte words of ONE plane:
//       RS0 = (R00 R01 R02 R10)  RS1 = (R11 R12 R20 R21)  RS2 = (R22 Sxx Sxy Sxz)  RS3 = (Syy Syz Szz V)
//   Fd: 3 planes  Fd0 = (F00 F01 F02 F10)  Fd1 = (F11 F12 F20 F21)  Fd2 = (F22 - - -)
//   Ks: 3 planes, static K_i = sum_j (x0_j - x0_i) (V_j nabla_W_ij)^T (same packing as Fd)
#pragma once
#include "mis_math.cuh"

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
    const float4* Ks;                 // static K_i (3 planes)
    float* Apq;                       // optional (keep_fields), 9 floats per slot
    const unsigned long long* nbr_start;
    const uint32_t* nbr;
};

enum ForceMode { MODE_PRIME = 0, MODE_STEP = 1, MODE_EULER = 2, MODE_EVAL = 3 };

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int STEP_THREADS = 256;

// ---------------------------------------------------------------- k_volume
// compute_v_i, sim.py:154-167: rho_i = sum_{j != i} m_j W(x0_i - x0_j), V_i = m_i / rho_i.
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

// Static part of compute_nabla_u (sim.py:193-209).  With g_ij = V_j nabla_W(x0_i - x0_j):
//   N_i = sum_j (R_i^T dx_ij - d0_ij) g_ij^T = R_i^T (sum_j dx_ij g_ij^T) - K_i,
//   K_i = sum_j d0_ij g_ij^T  depends on x0, V only: computed here once per set_mass.
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
    float K[9];
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
    }
#pragma unroll
    for (int k = 0; k < 9; k++) K[k] = group_sum<G>(K[k]);
    if (gl == 0 && gid < n) {
        Ks[i] = make_float4(K[0], K[1], K[2], K[3]);
        Ks[n + i] = make_float4(K[4], K[5], K[6], K[7]);
        Ks[2 * (size_t)n + i] = make_float4(K[8], 0.f, 0.f, 0.f);
    }
}

// ---------------------------------------------------------------- k_deform
// One pass over the neighbour list accumulates both moment sums
//   A_i = sum_j (W_ij m_j) dx_ij d0_ij^T      (compute_A_pq, sim.py:170-183)
//   B_i = sum_j dx_ij (V_j nabla_W_ij)^T       (the dynamic half of compute_nabla_u)
// with dx_ij = x_j - x_i, d0_ij = x0_j - x0_i; both weights are scalars times d0_ij.
// Then R_i = polar(A_i), N_i = R_i^T B_i - K_i, F_i = I + N_i^T, S_i = compute_sigma(F_i).
// FAITHFUL2 = true keeps the reference's two-loop evaluation order (u = R^T dx - d0 formed per
// pair, sim.py:207-208) for accuracy studies; results agree to the fp32 reorder floor.
template <int G, bool FAITHFUL2>
__global__ void __launch_bounds__(STEP_THREADS) k_deform(View s, Consts c) {
    const int gid = (blockIdx.x * STEP_THREADS + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const int n = s.n;
    const int i = min(gid, n - 1);
    const float4 p0i = s.x0m[i];
    const float4 pxi = s.xcur[i];
    const unsigned long long b = s.nbr_start[i];
    const int cnt = (int)(s.nbr_start[i + 1] - b);
    const uint32_t* __restrict__ lst = s.nbr + b;

    float A[9], B[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { A[k] = 0.f; B[k] = 0.f; }
    for (int k = gl; k < cnt; k += G) {
        const uint32_t j = lst[k];
        const float4 p0 = s.x0m[j];
        const float4 px = s.xcur[j];
        const float d0x = p0.x - p0i.x, d0y = p0.y - p0i.y, d0z = p0.z - p0i.z;
        float w, beta;
        kernel_W_and_coef(d0x * d0x + d0y * d0y + d0z * d0z, c, w, beta);
        w *= p0.w;                                        // W_ij m_j
        const float dx = px.x - pxi.x, dy = px.y - pxi.y, dz = px.z - pxi.z;
        const float tx = w * d0x, ty = w * d0y, tz = w * d0z;
        A[0] += dx * tx; A[1] += dx * ty; A[2] += dx * tz;
        A[3] += dy * tx; A[4] += dy * ty; A[5] += dy * tz;
        A[6] += dz * tx; A[7] += dz * ty; A[8] += dz * tz;
        if (!FAITHFUL2) {
            // nabla_W(x0_i - x0_j) = beta (x0_i - x0_j) = (-beta) d0
            const float nb = -beta * px.w;                // * V_j
            const float gx = nb * d0x, gy = nb * d0y, gz = nb * d0z;
            B[0] += dx * gx; B[1] += dx * gy; B[2] += dx * gz;
            B[3] += dy * gx; B[4] += dy * gy; B[5] += dy * gz;
            B[6] += dz * gx; B[7] += dz * gy; B[8] += dz * gz;
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) A[k] = group_sum<G>(A[k]);

    float R[9];
    if (c.identity_rot) {
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.f : 0.f;
    } else {
        polar_rotation(A, R);
    }

    float N[9];
    if (FAITHFUL2) {
#pragma unroll
        for (int k = 0; k < 9; k++) N[k] = 0.f;
        for (int k = gl; k < cnt; k += G) {
            const uint32_t j = lst[k];
            const float4 p0 = s.x0m[j];
            const float4 px = s.xcur[j];
            const float d0x = p0.x - p0i.x, d0y = p0.y - p0i.y, d0z = p0.z - p0i.z;
            const float nb = -kernel_gradW_coef(d0x * d0x + d0y * d0y + d0z * d0z, c) * px.w;
            const float dx = px.x - pxi.x, dy = px.y - pxi.y, dz = px.z - pxi.z;
            const float ux = R[0] * dx + R[3] * dy + R[6] * dz - d0x;
            const float uy = R[1] * dx + R[4] * dy + R[7] * dz - d0y;
            const float uz = R[2] * dx + R[5] * dy + R[8] * dz - d0z;
            const float gx = nb * d0x, gy = nb * d0y, gz = nb * d0z;
            N[0] += ux * gx; N[1] += ux * gy; N[2] += ux * gz;
            N[3] += uy * gx; N[4] += uy * gy; N[5] += uy * gz;
            N[6] += uz * gx; N[7] += uz * gy; N[8] += uz * gz;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) N[k] = group_sum<G>(N[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) B[k] = group_sum<G>(B[k]);
        const float4 k0 = s.Ks[i], k1 = s.Ks[n + i], k2 = s.Ks[2 * (size_t)n + i];
        const float K[9] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w, k2.x};
        // N = R^T B - K
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int q = 0; q < 3; q++)
                N[3 * r + q] = R[0 * 3 + r] * B[0 * 3 + q] + R[1 * 3 + r] * B[1 * 3 + q] + R[2 * 3 + r] * B[2 * 3 + q] - K[3 * r + q];
    }

    if (gl == 0 && gid < n) {
        // def_grad = I + N^T
        float F[9] = {1.f + N[0], N[3], N[6], N[1], 1.f + N[4], N[7], N[2], N[5], 1.f + N[8]};
        const float4 ml = s.matl[i];
        float S[6];
        stress_svk(F, ml.x, ml.y, ml.z, c, S);
        s.RS[i] = make_float4(R[0], R[1], R[2], R[3]);
        s.RS[n + i] = make_float4(R[4], R[5], R[6], R[7]);
        s.RS[2 * (size_t)n + i] = make_float4(R[8], S[0], S[1], S[2]);
        s.RS[3 * (size_t)n + i] = make_float4(S[3], S[4], S[5], pxi.w);
        s.Fd[i] = make_float4(F[0], F[1], F[2], F[3]);
        s.Fd[n + i] = make_float4(F[4], F[5], F[6], F[7]);
        s.Fd[2 * (size_t)n + i] = make_float4(F[8], 0.f, 0.f, 0.f);
        if (s.Apq) {
#pragma unroll
            for (int k = 0; k < 9; k++) s.Apq[9 * (size_t)i + k] = A[k];
        }
    }
}

// ---------------------------------------------------------------- integrate (literal order, no FMA)
__device__ __forceinline__ float3 total_force(float3 fext, float3 fel, float3 v, float y, const Consts& c) {
    // ((external + elastic) - damping * v) + penalty(x)      sim.py:250,256-257
    float pen = ground_penalty_y(y, c);
    float3 f;
    f.x = __fadd_rn(__fsub_rn(__fadd_rn(fext.x, fel.x), __fmul_rn(c.damping, v.x)), 0.f);
    f.y = __fadd_rn(__fsub_rn(__fadd_rn(fext.y, fel.y), __fmul_rn(c.damping, v.y)), pen);
    f.z = __fadd_rn(__fsub_rn(__fadd_rn(fext.z, fel.z), __fmul_rn(c.damping, v.z)), 0.f);
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
    if (mode == MODE_EULER) {
        // sim_taichi.py:161-172: force = ext + el + (-damping v); v' = v + dt f / m * free; x' = x + dt v' * free
        float3 f;
        f.x = __fadd_rn(__fadd_rn(fext.x, fel.x), __fmul_rn(-c.damping, v.x));
        f.y = __fadd_rn(__fadd_rn(fext.y, fel.y), __fmul_rn(-c.damping, v.y));
        f.z = __fadd_rn(__fadd_rn(fext.z, fel.z), __fmul_rn(-c.damping, v.z));
        float3 vn, xn;
        vn.x = __fadd_rn(v.x, __fmul_rn(__fdiv_rn(__fmul_rn(c.dt, f.x), m), fr.x));
        vn.y = __fadd_rn(v.y, __fmul_rn(__fdiv_rn(__fmul_rn(c.dt, f.y), m), fr.y));
        vn.z = __fadd_rn(v.z, __fmul_rn(__fdiv_rn(__fmul_rn(c.dt, f.z), m), fr.z));
        xn.x = __fadd_rn(x.x, __fmul_rn(__fmul_rn(c.dt, vn.x), fr.x));
        xn.y = __fadd_rn(x.y, __fmul_rn(__fmul_rn(c.dt, vn.y), fr.y));
        xn.z = __fadd_rn(x.z, __fmul_rn(__fmul_rn(c.dt, vn.z), fr.z));
        s.vel[i] = make_float4(vn.x, vn.y, vn.z, 0.f);
        s.fel[i] = make_float4(fel.x, fel.y, fel.z, 0.f);
        s.xnext[i] = make_float4(xn.x, xn.y, xn.z, pxi.w);
        return;
    }
    if (mode == MODE_STEP) {
        // part_2 of this step: force_1 was stored by the previous part_1 (same inputs, same value)
        const float3 F1 = xyz(s.f1[i]);
        const float3 F2 = total_force(fext, fel, v, x.y, c);
        v.x = part2_axis(v.x, F1.x, F2.x, m, fr.x, c);
        v.y = part2_axis(v.y, F1.y, F2.y, m, fr.y, c);
        v.z = part2_axis(v.z, F1.z, F2.z, m, fr.z, c);
        s.vel[i] = make_float4(v.x, v.y, v.z, 0.f);
    }
    // part_1 of the next step from (x, v, fel) of the now-current frame
    const float3 F1n = total_force(fext, fel, v, x.y, c);
    float3 xn;
    xn.x = part1_axis(x.x, v.x, F1n.x, m, fr.x, c);
    xn.y = part1_axis(x.y, v.y, F1n.y, m, fr.y, c);
    xn.z = part1_axis(x.z, v.z, F1n.z, m, fr.z, c);
    s.f1[i] = make_float4(F1n.x, F1n.y, F1n.z, 0.f);
    s.fel[i] = make_float4(fel.x, fel.y, fel.z, 0.f);
    s.xnext[i] = make_float4(xn.x, xn.y, xn.z, pxi.w);
}

// ---------------------------------------------------------------- k_force
// force_i = sum_j 0.5 (R_j f_ij - R_i f_ji),  f_ij = V_j F_i S_j (V_i nabla_W_ij),
//                                             f_ji = -V_i F_i S_i (V_j nabla_W_ij)
//         = 0.5 V_i [ sum_j V_j R_j F_i S_j nabla_W_ij  +  R_i F_i S_i sum_j V_j nabla_W_ij ]
// (F_i, not F_j, multiplies S_j: sim.py:233.)
template <int G>
__global__ void __launch_bounds__(STEP_THREADS) k_force(View s, Consts c, int mode) {
    const int gid = (blockIdx.x * STEP_THREADS + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const int i = min(gid, s.n - 1);
    const float4 p0i = s.x0m[i];
    const int n = s.n;
    const float4* __restrict__ RS0 = s.RS;
    const float4* __restrict__ RS1 = s.RS + n;
    const float4* __restrict__ RS2 = s.RS + 2 * (size_t)n;
    const float4* __restrict__ RS3 = s.RS + 3 * (size_t)n;
    const float4 f0 = s.Fd[i], f1 = s.Fd[n + i], f2 = s.Fd[2 * (size_t)n + i];
    const float F[9] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x};
    const unsigned long long b = s.nbr_start[i];
    const int cnt = (int)(s.nbr_start[i + 1] - b);
    const uint32_t* __restrict__ lst = s.nbr + b;

    float ax = 0.f, ay = 0.f, az = 0.f;     // sum_j V_j R_j F_i S_j nw
    float gx = 0.f, gy = 0.f, gz = 0.f;     // sum_j V_j nw
    for (int k = gl; k < cnt; k += G) {
        const uint32_t j = lst[k];
        const float4 p0 = s.x0m[j];
        const float4 r0 = RS0[j], r1 = RS1[j], r2 = RS2[j], r3 = RS3[j];
        const float d0x = p0.x - p0i.x, d0y = p0.y - p0i.y, d0z = p0.z - p0i.z;
        const float nb = -kernel_gradW_coef(d0x * d0x + d0y * d0y + d0z * d0z, c) * r3.w;   // V_j folded in
        const float nx = nb * d0x, ny = nb * d0y, nz = nb * d0z;                            // V_j nabla_W_ij
        // t = S_j n     (S = r2.y r2.z r2.w / r3.x r3.y / r3.z)
        const float tx = r2.y * nx + r2.z * ny + r2.w * nz;
        const float ty = r2.z * nx + r3.x * ny + r3.y * nz;
        const float tz = r2.w * nx + r3.y * ny + r3.z * nz;
        // u = F_i t
        const float ux = F[0] * tx + F[1] * ty + F[2] * tz;
        const float uy = F[3] * tx + F[4] * ty + F[5] * tz;
        const float uz = F[6] * tx + F[7] * ty + F[8] * tz;
        // a += R_j u
        ax += r0.x * ux + r0.y * uy + r0.z * uz;
        ay += r0.w * ux + r1.x * uy + r1.y * uz;
        az += r1.z * ux + r1.w * uy + r2.x * uz;
        gx += nx; gy += ny; gz += nz;
    }
    ax = group_sum<G>(ax); ay = group_sum<G>(ay); az = group_sum<G>(az);
    gx = group_sum<G>(gx); gy = group_sum<G>(gy); gz = group_sum<G>(gz);

    if (gl == 0 && gid < s.n) {
        const float4 r0 = RS0[i], r1 = RS1[i], r2 = RS2[i], r3 = RS3[i];
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
        if (cnt == 0) fel = make_float3(0.f, 0.f, 0.f);   // isolated particle: the reference loop never runs (V = m/0)
        integrate_epilogue(s, c, i, fel, mode, p0i, s.xcur[i]);
    }
}

// ---------------------------------------------------------------- k_force_sym
// sim_taichi.py:147-158: f_ij uses F_j, so the pair term is antisymmetric:
//   force_i = 0.5 V_i sum_j V_j (P_j + P_i) nabla_W_ij,  P = R F S  (first Piola in the rotated frame)
// P_j is rebuilt per neighbour from (R_j, S_j) and F_j.
template <int G>
__global__ void __launch_bounds__(STEP_THREADS) k_force_sym(View s, Consts c, int mode) {
    const int gid = (blockIdx.x * STEP_THREADS + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const int i = min(gid, s.n - 1);
    const float4 p0i = s.x0m[i];
    const int n = s.n;
    const float4* __restrict__ RS0 = s.RS;
    const float4* __restrict__ RS1 = s.RS + n;
    const float4* __restrict__ RS2 = s.RS + 2 * (size_t)n;
    const float4* __restrict__ RS3 = s.RS + 3 * (size_t)n;
    const unsigned long long b = s.nbr_start[i];
    const int cnt = (int)(s.nbr_start[i + 1] - b);
    const uint32_t* __restrict__ lst = s.nbr + b;
    float ax = 0.f, ay = 0.f, az = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    for (int k = gl; k < cnt; k += G) {
        const uint32_t j = lst[k];
        const float4 p0 = s.x0m[j];
        const float4 r0 = RS0[j], r1 = RS1[j], r2 = RS2[j], r3 = RS3[j];
        const float4 f0 = s.Fd[j], f1 = s.Fd[n + j], f2 = s.Fd[2 * (size_t)n + j];
        const float d0x = p0.x - p0i.x, d0y = p0.y - p0i.y, d0z = p0.z - p0i.z;
        const float nb = -kernel_gradW_coef(d0x * d0x + d0y * d0y + d0z * d0z, c) * r3.w;
        const float nx = nb * d0x, ny = nb * d0y, nz = nb * d0z;
        const float tx = r2.y * nx + r2.z * ny + r2.w * nz;
        const float ty = r2.z * nx + r3.x * ny + r3.y * nz;
        const float tz = r2.w * nx + r3.y * ny + r3.z * nz;
        const float ux = f0.x * tx + f0.y * ty + f0.z * tz;
        const float uy = f0.w * tx + f1.x * ty + f1.y * tz;
        const float uz = f1.z * tx + f1.w * ty + f2.x * tz;
        ax += r0.x * ux + r0.y * uy + r0.z * uz;
        ay += r0.w * ux + r1.x * uy + r1.y * uz;
        az += r1.z * ux + r1.w * uy + r2.x * uz;
        gx += nx; gy += ny; gz += nz;
    }
    ax = group_sum<G>(ax); ay = group_sum<G>(ay); az = group_sum<G>(az);
    gx = group_sum<G>(gx); gy = group_sum<G>(gy); gz = group_sum<G>(gz);
    if (gl == 0 && gid < s.n) {
        const float4 r0 = RS0[i], r1 = RS1[i], r2 = RS2[i], r3 = RS3[i];
        const float4 f0 = s.Fd[i], f1 = s.Fd[n + i], f2 = s.Fd[2 * (size_t)n + i];
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
        if (cnt == 0) fel = make_float3(0.f, 0.f, 0.f);   // isolated particle: the reference loop never runs (V = m/0)
        integrate_epilogue(s, c, i, fel, mode, p0i, s.xcur[i]);
    }
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
