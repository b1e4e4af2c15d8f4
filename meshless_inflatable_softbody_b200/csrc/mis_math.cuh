// mis_math.cuh -- per-pair and per-particle device arithmetic of the meshless step.
//
// Cubic-spline kernel and gradient (sim.py:133-151), rotation extraction
// (sim.py:185-191), St.Venant-Kirchhoff stress with the stiffness design factor
// (sim.py:212-216), ground penalty (sim.py:238-244).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mis {

// Scene constants, passed by value to every kernel (sim.py:25-26,65,68-69,215).
struct Consts {
    float h, inv_h;
    float sigma;        // 1 / (pi h^3)
    float grad_c1;      // sigma / h^2
    float grad_c2;      // 0.75 * sigma / h      (1<=q<2 branch: -c2 (2-q)^2 / r)
    float sigma4;       // sigma / 4
    float grad_q1;      // 3 sigma / h           beta = (grad_q1 t1^2 - grad_q2 t2^2) / r
    float grad_q2;      // 0.75 sigma / h
    float dt, half_dt2, damping;
    float k_col, col_range;
    float stiff_a, stiff_b;
    int   identity_rot, euler, no_contact, symmetric_pair;
};

struct Mat3 { float m[9]; };   // row-major

__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 xyz(float4 a) { return make_float3(a.x, a.y, a.z); }

// W(|xij|) for a listed neighbour (q < 2 by construction of the list), sim.py:133-141.
// r2 = |xij|^2.  rsqrt-based: value parity is to fp32 round-off, membership of the
// neighbour set is decided exactly elsewhere (mis_neighbors.cuh).
// Branch-free form used by all three helpers: with t2 = max(2 - q, 0), t1 = max(1 - q, 0) the piecewise cubic of sim.py:137-141
// is the cubic B-spline identity
//     W = sigma (t2^3 / 4 - t1^3),        dW/dq = sigma (3 t1^2 - 3/4 t2^2),        nabla_W = (dW/dq) / (r h) * xij
// (q < 1: 1/4 (2-q)^3 - (1-q)^3 = 1 - 1.5 q^2 + 0.75 q^3; 1 <= q < 2: the first term alone; q >= 2: zero), so no lane diverges on
// the branch of the kernel and the compiler needs neither selects nor a reconvergence point inside the pair loop.
__device__ __forceinline__ float kernel_W(float r2, const Consts& c) {
    float rinv = rsqrtf(fmaxf(r2, 1e-30f));
    float q = r2 * rinv * c.inv_h;
    float t2 = fmaxf(2.f - q, 0.f), t1 = fmaxf(1.f - q, 0.f);
    return c.sigma * (0.25f * t2 * t2 * t2 - t1 * t1 * t1);
}

// nabla_W(xij) = beta(|xij|) * xij, sim.py:143-151.  Returns beta.
//   q < 1    : sigma (-3 + 2.25 q) / h^2
//   1<=q<2   : sigma/4 * (-3) (2-q)^2 / (q h^2) = -0.75 sigma (2-q)^2 / (r h)
// 1/sqrt(x) for x >= 1e-30 (always a normal number here): the bare MUFU.RSQ.  rsqrtf() wraps the same instruction in a denormal
// rescue (FSETP + two predicated FMULs) that can never trigger behind the fmaxf guard; same result, three instructions fewer per pair.
__device__ __forceinline__ float rsqrt_normal(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float kernel_gradW_coef(float r2, const Consts& c) {
    float rinv = rsqrt_normal(fmaxf(r2, 1e-30f));
    float q = r2 * rinv * c.inv_h;
    float t2 = fmaxf(2.f - q, 0.f), t1 = fmaxf(1.f - q, 0.f);
    float e = c.grad_q1 * (t1 * t1);
    e = fmaf(-c.grad_q2, t2 * t2, e);                 // sigma / h (3 t1^2 - 3/4 t2^2)
    return e * rinv;
}

// Both at once (pass A needs W, pass F needs beta; the fused kernel wants both).
__device__ __forceinline__ void kernel_W_and_coef(float r2, const Consts& c, float& w, float& beta) {
    float rinv = rsqrt_normal(fmaxf(r2, 1e-30f));
    float q = r2 * rinv * c.inv_h;
    float t2 = fmaxf(2.f - q, 0.f), t1 = fmaxf(1.f - q, 0.f);
    float t2s = t2 * t2, t1s = t1 * t1;
    w = (c.sigma4 * t2s) * t2 - (c.sigma * t1s) * t1;
    float e = c.grad_q1 * t1s;
    e = fmaf(-c.grad_q2, t2s, e);
    beta = e * rinv;
}

// ---------------------------------------------------------------- rotation
// R = U V^T of the SVD of A with det U = det V = +1 (sim.py:185-191, wp.svd3):
// cyclic Jacobi on A^T A for V, Gram-Schmidt of the two dominant columns of A V
// plus a cross product for U.  Same sequence as the oracle's polar_rotation.
__device__ __forceinline__ void jacobi_rot(float S[3][3], float V[3][3], int p, int q) {
    float apq = S[p][q];
    if (fabsf(apq) <= 1e-30f) return;
    float theta = (S[q][q] - S[p][p]) / (2.f * apq);
    float t = 1.f / (fabsf(theta) + sqrtf(theta * theta + 1.f));
    if (theta < 0.f) t = -t;
    float cs = 1.f / sqrtf(t * t + 1.f);
    float sn = t * cs;
    S[p][p] -= t * apq;
    S[q][q] += t * apq;
    S[p][q] = 0.f; S[q][p] = 0.f;
    int r = 3 - p - q;
    float srp = S[r][p], srq = S[r][q];
    S[r][p] = cs * srp - sn * srq; S[p][r] = S[r][p];
    S[r][q] = sn * srp + cs * srq; S[q][r] = S[r][q];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float vkp = V[k][p], vkq = V[k][q];
        V[k][p] = cs * vkp - sn * vkq;
        V[k][q] = sn * vkp + cs * vkq;
    }
}

__device__ __noinline__ void polar_rotation(const float A[9], float R[9]) {
    float S[3][3], V[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            S[i][j] = A[0 * 3 + i] * A[0 * 3 + j] + A[1 * 3 + i] * A[1 * 3 + j] + A[2 * 3 + i] * A[2 * 3 + j];
#pragma unroll 1
    for (int sweep = 0; sweep < 6; sweep++) {
        jacobi_rot(S, V, 0, 1);
        jacobi_rot(S, V, 0, 2);
        jacobi_rot(S, V, 1, 2);
    }
    float B[3][3], nrm[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            B[i][j] = A[i * 3 + 0] * V[0][j] + A[i * 3 + 1] * V[1][j] + A[i * 3 + 2] * V[2][j];
#pragma unroll
    for (int j = 0; j < 3; j++) nrm[j] = B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j];
#define MIS_CSWAP(a, b)                                                      \
    if (nrm[a] < nrm[b]) {                                                   \
        float tn = nrm[a]; nrm[a] = nrm[b]; nrm[b] = tn;                     \
        _Pragma("unroll") for (int k = 0; k < 3; k++) {                      \
            float tb = B[k][a]; B[k][a] = B[k][b]; B[k][b] = -tb;            \
            float tv = V[k][a]; V[k][a] = V[k][b]; V[k][b] = -tv;            \
        }                                                                    \
    }
    MIS_CSWAP(0, 1) MIS_CSWAP(0, 2) MIS_CSWAP(1, 2)
#undef MIS_CSWAP
    float n0 = sqrtf(nrm[0]);
    if (!(n0 > 1e-30f)) {
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.f : 0.f;
        return;
    }
    float u0[3] = {B[0][0] / n0, B[1][0] / n0, B[2][0] / n0};
    float d = u0[0] * B[0][1] + u0[1] * B[1][1] + u0[2] * B[2][1];
    float w1[3] = {B[0][1] - d * u0[0], B[1][1] - d * u0[1], B[2][1] - d * u0[2]};
    float n1 = sqrtf(w1[0] * w1[0] + w1[1] * w1[1] + w1[2] * w1[2]);
    float u1[3];
    if (n1 > 1e-30f) {
        u1[0] = w1[0] / n1; u1[1] = w1[1] / n1; u1[2] = w1[2] / n1;
    } else {
        float a0 = fabsf(u0[0]), a1 = fabsf(u0[1]), a2 = fabsf(u0[2]);
        int k = (a0 <= a1 && a0 <= a2) ? 0 : (a1 <= a2 ? 1 : 2);
        float e[3] = {k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f};
        float dd = k == 0 ? u0[0] : (k == 1 ? u0[1] : u0[2]);
        float ww[3] = {e[0] - dd * u0[0], e[1] - dd * u0[1], e[2] - dd * u0[2]};
        float nn = sqrtf(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
        u1[0] = ww[0] / nn; u1[1] = ww[1] / nn; u1[2] = ww[2] / nn;
    }
    float u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            R[i * 3 + j] = u0[i] * V[j][0] + u1[i] * V[j][1] + u2[i] * V[j][2];
}

// S = (2 mu E + lam tr(E) I) * (a - b ratio), E = (F^T F - I)/2, sim.py:212-216.
// Symmetric: returns (xx, xy, xz, yy, yz, zz).
__device__ __forceinline__ void stress_svk(const float F[9], float mu, float lam, float ratio, const Consts& c, float S[6]) {
    float exx = 0.5f * (F[0] * F[0] + F[3] * F[3] + F[6] * F[6] - 1.f);
    float eyy = 0.5f * (F[1] * F[1] + F[4] * F[4] + F[7] * F[7] - 1.f);
    float ezz = 0.5f * (F[2] * F[2] + F[5] * F[5] + F[8] * F[8] - 1.f);
    float exy = 0.5f * (F[0] * F[1] + F[3] * F[4] + F[6] * F[7]);
    float exz = 0.5f * (F[0] * F[2] + F[3] * F[5] + F[6] * F[8]);
    float eyz = 0.5f * (F[1] * F[2] + F[4] * F[5] + F[7] * F[8]);
    float k = c.stiff_a - ratio * c.stiff_b;
    float ltr = lam * (exx + eyy + ezz);
    float m2 = 2.f * mu;
    S[0] = (m2 * exx + ltr) * k;
    S[1] = (m2 * exy) * k;
    S[2] = (m2 * exz) * k;
    S[3] = (m2 * eyy + ltr) * k;
    S[4] = (m2 * eyz) * k;
    S[5] = (m2 * ezz + ltr) * k;
}

// sim.py:238-244: quadratic ground penalty, y component only.
__device__ __forceinline__ float ground_penalty_y(float y, const Consts& c) {
    if (c.no_contact || !(y < c.col_range)) return 0.f;
    float delta = c.col_range - y;
    return delta * delta * c.k_col;
}

}  // namespace mis
