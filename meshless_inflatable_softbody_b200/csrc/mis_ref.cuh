// mis_ref.cuh -- the same step written once more, templated on the working precision and in the reference's own evaluation
// order, plus its REVERSE pass.
//
// Two users:
//   * real = double: the Taichi prototype's precision (options.py:3 real = ti.f64; sim_taichi.py:93-182) -- MisParams.fp64.
//     The hot fp32 path (mis_tile.cuh / mis_cluster.cuh) is tuned for float4 records and fp32 accumulation and is not
//     instantiated in double; these kernels are the plain data flow of sim.py:154-258 over the exact neighbour lists, one group
//     of 8 lanes per particle.
//   * real = float or double: the adjoint of the rollout (sim.py:346-372: wp.Tape() around the loop, tape.backward(l),
//     x.grad) -- mis_rollout_grad.  The reference keeps all 3 001 frames of every array for its tape; here the forward
//     trajectory is checkpointed every K frames and a segment is recomputed before it is swept backwards.
//
// Adjoint of the elastic force (lam_i = dL/d fel_i), with a_ij = W_ij m_j, g_ij = V_j nabla_W(x0_i - x0_j), dx_ij = x_j - x_i,
// d0_ij = x0_j - x0_i, G_i = sum_j g_ij (forward: A_i = sum a_ij dx_ij d0_ij^T, R_i = polar(A_i), B_i = sum dx_ij g_ij^T,
// N_i = R_i^T B_i - K_i, F_i = I + N_i^T, E_i = (F_i^T F_i - I)/2, S_i = k_i (2 mu E_i + lam tr(E_i) I), k_i = a - b ratio_i,
// fel_i = V_i/2 [ sum_j R_j F_i S_j g_ij + R_i F_i S_i G_i ]   (sim.py:218-235; sim_taichi.py:147-158 uses F_j S_j)):
//   pair (i, j):  Rb_j += V_i/2 lam_i (F_i S_j g_ij)^T,  Fb_i += V_i/2 (R_j^T lam_i)(S_j g_ij)^T,  Sb_j += V_i/2 (F_i^T R_j^T lam_i) g_ij^T
//   self term  :  Rb_i += V_i/2 lam_i (F_i S_i G_i)^T,  Fb_i += V_i/2 (R_i^T lam_i)(S_i G_i)^T,  Sb_i += V_i/2 (F_i^T R_i^T lam_i) G_i^T
//   per particle: Ss = sym(Sb); Eb = k (2 mu Ss + lam tr(Ss) I); kb = <Ss, 2 mu E + lam tr(E) I>; ratio_b = -b kb;
//                 Fb += F Eb; Nb = Fb^T; Rb += B Nb^T; Bb = R Nb; Ab = polar_backward(A, R, Rb)
//   positions   : dxb_ij = a_ij Ab_i d0_ij + Bb_i g_ij;  xb_j += dxb_ij;  xb_i -= dxb_ij
// Every scatter (index j) is turned into a gather over the symmetric neighbour relation (j in N(i) <=> i in N(j)).
// polar_backward: with A = R M (M symmetric), Y = R^T Rb, y = axial(Y - Y^T)/2, u = (tr(M) I - M)^-1 y: Ab = 2 R [u]x ... (see below).
#pragma once
#include "mis_math.cuh"

namespace mis {
namespace ref {

template <typename T> struct RP {               // scene constants in the working precision
    T h, dt, damping, k_col, col_range, stiff_a, stiff_b, tanh_k;
    int identity_rot, euler, no_contact, symmetric_pair, self_density;
};

template <typename T> struct RS_ {              // state, cell-sorted slot order (the exact lists index slots)
    int n;
    const unsigned long long* nbr_start;
    const uint32_t* nbr;
    T *x0, *m, *vol, *rho, *mu, *lam, *ratio, *fext, *freem;       // static (n or 3n)
    T *x, *xn, *v, *fel, *feln;                                     // dynamic (3n)
    T *A, *R, *F, *S, *B;                                           // per-particle matrices of the current evaluation (9n)
};

constexpr int RG = 8;                            // lanes per particle
constexpr int RTHREADS = 128;

__device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float t_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double t_abs(double x) { return fabs(x); }
__device__ __forceinline__ float t_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double t_tanh(double x) { return tanh(x); }
template <typename T> __device__ __forceinline__ T t_pi() { return (T)3.14159265358979323846; }

template <typename T> __device__ __forceinline__ T gsum(T v) {
#pragma unroll
    for (int o = RG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sim.py:133-141 / utils.py:25-33
template <typename T> __device__ __forceinline__ T Wk(const T r[3], T h) {
    const T q = t_sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / h;
    T ret = (T)0;
    if (q < (T)1) ret = (T)1 / (t_pi<T>() * h * h * h) * ((T)1 - (T)1.5 * q * q + (T)0.75 * q * q * q);
    else if (q < (T)2) ret = (T)1 / ((T)4 * t_pi<T>() * h * h * h) * ((T)2 - q) * ((T)2 - q) * ((T)2 - q);
    return ret;
}
// sim.py:143-151 / utils.py:35-43
template <typename T> __device__ __forceinline__ void nablaWk(const T r[3], T h, T out[3]) {
    const T q = t_sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / h;
    T c = (T)0;
    if (q < (T)1) c = (T)1 / (t_pi<T>() * h * h * h) * ((T)-3 / (h * h) + (T)0.75 * (T)3 * q / (h * h));
    else if (q < (T)2) c = (T)1 / ((T)4 * t_pi<T>() * h * h * h) * (T)-3 * ((T)2 - q) * ((T)2 - q) / (q * h * h);
    out[0] = c * r[0]; out[1] = c * r[1]; out[2] = c * r[2];
}

// ---------------------------------------------------------------- 3x3 helpers (row-major)
template <typename T> __device__ __forceinline__ void mm(const T a[9], const T b[9], T c[9]) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
template <typename T> __device__ __forceinline__ void mtm(const T a[9], const T b[9], T c[9]) {      // a^T b
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
template <typename T> __device__ __forceinline__ void mmt(const T a[9], const T b[9], T c[9]) {      // a b^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
template <typename T> __device__ __forceinline__ void mv(const T a[9], const T v[3], T o[3]) {
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = a[3 * i] * v[0] + a[3 * i + 1] * v[1] + a[3 * i + 2] * v[2];
}
template <typename T> __device__ __forceinline__ void mtv(const T a[9], const T v[3], T o[3]) {      // a^T v
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = a[i] * v[0] + a[3 + i] * v[1] + a[6 + i] * v[2];
}
template <typename T> __device__ __forceinline__ void ld9(const T* p, int i, T o[9]) {
#pragma unroll
    for (int k = 0; k < 9; k++) o[k] = p[9 * (size_t)i + k];
}
template <typename T> __device__ __forceinline__ void st9(T* p, int i, const T o[9]) {
#pragma unroll
    for (int k = 0; k < 9; k++) p[9 * (size_t)i + k] = o[k];
}
template <typename T> __device__ __forceinline__ void ld3(const T* p, int i, T o[3]) {
    o[0] = p[3 * (size_t)i]; o[1] = p[3 * (size_t)i + 1]; o[2] = p[3 * (size_t)i + 2];
}

// rotation of the polar decomposition, same sequence as mis_math.cuh::polar_rotation (sim.py:185-191)
template <typename T> __device__ __forceinline__ void jrot(T S[3][3], T V[3][3], int p, int q) {
    const T apq = S[p][q];
    if (t_abs(apq) <= (T)(sizeof(T) == 8 ? 1e-300 : 1e-30)) return;
    const T theta = (S[q][q] - S[p][p]) / ((T)2 * apq);
    T t = (T)1 / (t_abs(theta) + t_sqrt(theta * theta + (T)1));
    if (theta < (T)0) t = -t;
    const T cs = (T)1 / t_sqrt(t * t + (T)1), sn = t * cs;
    S[p][p] -= t * apq; S[q][q] += t * apq; S[p][q] = (T)0; S[q][p] = (T)0;
    const int r = 3 - p - q;
    const T srp = S[r][p], srq = S[r][q];
    S[r][p] = cs * srp - sn * srq; S[p][r] = S[r][p];
    S[r][q] = sn * srp + cs * srq; S[q][r] = S[r][q];
#pragma unroll
    for (int k = 0; k < 3; k++) { const T vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
}
template <typename T> __device__ __noinline__ void polar_t(const T A[9], T R[9]) {
    T S[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) S[i][j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
    const int sweeps = sizeof(T) == 8 ? 12 : 6;
#pragma unroll 1
    for (int s = 0; s < sweeps; s++) { jrot(S, V, 0, 1); jrot(S, V, 0, 2); jrot(S, V, 1, 2); }
    T Bm[3][3], nrm[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Bm[i][j] = A[3 * i] * V[0][j] + A[3 * i + 1] * V[1][j] + A[3 * i + 2] * V[2][j];
    for (int j = 0; j < 3; j++) nrm[j] = Bm[0][j] * Bm[0][j] + Bm[1][j] * Bm[1][j] + Bm[2][j] * Bm[2][j];
    auto cswap = [&](int a, int b) {
        if (nrm[a] < nrm[b]) {
            T tn = nrm[a]; nrm[a] = nrm[b]; nrm[b] = tn;
            for (int k = 0; k < 3; k++) {
                T tb = Bm[k][a]; Bm[k][a] = Bm[k][b]; Bm[k][b] = -tb;
                T tv = V[k][a]; V[k][a] = V[k][b]; V[k][b] = -tv;
            }
        }
    };
    cswap(0, 1); cswap(0, 2); cswap(1, 2);
    const T n0 = t_sqrt(nrm[0]);
    if (!(n0 > (T)1e-30)) { for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? (T)1 : (T)0; return; }
    T u0[3] = {Bm[0][0] / n0, Bm[1][0] / n0, Bm[2][0] / n0};
    const T d = u0[0] * Bm[0][1] + u0[1] * Bm[1][1] + u0[2] * Bm[2][1];
    T w1[3] = {Bm[0][1] - d * u0[0], Bm[1][1] - d * u0[1], Bm[2][1] - d * u0[2]};
    const T n1 = t_sqrt(w1[0] * w1[0] + w1[1] * w1[1] + w1[2] * w1[2]);
    T u1[3];
    if (n1 > (T)1e-30) { u1[0] = w1[0] / n1; u1[1] = w1[1] / n1; u1[2] = w1[2] / n1; }
    else {
        const T a0 = t_abs(u0[0]), a1 = t_abs(u0[1]), a2 = t_abs(u0[2]);
        const int k = (a0 <= a1 && a0 <= a2) ? 0 : (a1 <= a2 ? 1 : 2);
        T e[3] = {k == 0 ? (T)1 : (T)0, k == 1 ? (T)1 : (T)0, k == 2 ? (T)1 : (T)0};
        const T dd = u0[k];
        T ww[3] = {e[0] - dd * u0[0], e[1] - dd * u0[1], e[2] - dd * u0[2]};
        const T nn = t_sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
        u1[0] = ww[0] / nn; u1[1] = ww[1] / nn; u1[2] = ww[2] / nn;
    }
    const T u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[3 * i + j] = u0[i] * V[j][0] + u1[i] * V[j][1] + u2[i] * V[j][2];
}

// Adjoint of R = polar(A).  A = R M with M = R^T A symmetric.  A perturbation dA rotates R by dR = R [w]x with
// (tr(M) I - M) w = axial(R^T dA - dA^T R)  (the skew part of R^T dA = [w]x M + dM).  With Y = R^T Rb and y = axial of its
// skew part (y_k = (Y_ji - Y_ij)/2 cyclic), <Rb, dR> = 2 y.w = 2 u.axial(R^T dA - dA^T R), u = (tr(M) I - M)^-1 y (symmetric
// matrix), and u.axial(Z - Z^T) = <[u]x, Z> for any Z: Ab = 2 R [u]x.
template <typename T> __device__ __forceinline__ void polar_backward(const T A[9], const T R[9], const T Rb[9], T Ab[9]) {
    T M[9], Y[9];
    mtm(R, A, M);
    mtm(R, Rb, Y);
    const T y[3] = {(T)0.5 * (Y[7] - Y[5]), (T)0.5 * (Y[2] - Y[6]), (T)0.5 * (Y[3] - Y[1])};
    const T tr = M[0] + M[4] + M[8];
    // H = tr I - sym(M); solve H u = y by the adjugate (H is SPD when A is non-singular with det > 0)
    const T h00 = tr - M[0], h11 = tr - M[4], h22 = tr - M[8];
    const T h01 = -(T)0.5 * (M[1] + M[3]), h02 = -(T)0.5 * (M[2] + M[6]), h12 = -(T)0.5 * (M[5] + M[7]);
    const T c00 = h11 * h22 - h12 * h12, c01 = h02 * h12 - h01 * h22, c02 = h01 * h12 - h02 * h11;
    const T c11 = h00 * h22 - h02 * h02, c12 = h01 * h02 - h00 * h12, c22 = h00 * h11 - h01 * h01;
    const T det = h00 * c00 + h01 * c01 + h02 * c02;
    T u[3] = {(T)0, (T)0, (T)0};
    if (t_abs(det) > (T)1e-30) {
        const T id = (T)1 / det;
        u[0] = (c00 * y[0] + c01 * y[1] + c02 * y[2]) * id;
        u[1] = (c01 * y[0] + c11 * y[1] + c12 * y[2]) * id;
        u[2] = (c02 * y[0] + c12 * y[1] + c22 * y[2]) * id;
    }
    const T U[9] = {(T)0, -u[2], u[1], u[2], (T)0, -u[0], -u[1], u[0], (T)0};
    T RU[9];
    mm(R, U, RU);
#pragma unroll
    for (int k = 0; k < 9; k++) Ab[k] = (T)2 * RU[k];
}

// S = (2 mu E + lam tr(E) I) (a - b ratio), E = (F^T F - I)/2     sim.py:212-216
template <typename T> __device__ __forceinline__ void sigma_t(const T F[9], T mu, T lam, T ratio, const RP<T>& c, T S[9]) {
    T E[9];
    mtm(F, F, E);
#pragma unroll
    for (int k = 0; k < 9; k++) E[k] = (T)0.5 * (E[k] - ((k % 4 == 0) ? (T)1 : (T)0));
    const T tr = E[0] + E[4] + E[8];
    const T kf = c.stiff_a - ratio * c.stiff_b;
#pragma unroll
    for (int k = 0; k < 9; k++) S[k] = ((T)2 * mu * E[k] + ((k % 4 == 0) ? lam * tr : (T)0)) * kf;
}

// ---------------------------------------------------------------- forward kernels
#define MIS_REF_GROUP()                                                           \
    const int gid = (blockIdx.x * RTHREADS + threadIdx.x) / RG;                   \
    const int gl = threadIdx.x % RG;                                              \
    const int i = min(gid, s.n - 1);                                              \
    const unsigned long long lb = s.nbr_start[i];                                 \
    const int cnt = (int)(s.nbr_start[i + 1] - lb);                               \
    const uint32_t* __restrict__ lst = s.nbr + lb;

// compute_v_i (sim.py:154-167; sim_taichi.py:93-100 includes j == i)
template <typename T> __global__ void __launch_bounds__(RTHREADS) kr_volume(RS_<T> s, RP<T> c) {
    MIS_REF_GROUP();
    T xi[3]; ld3(s.x0, i, xi);
    T rho = (T)0;
    for (int k = gl; k < cnt; k += RG) {
        const int j = (int)lst[k];
        T xj[3]; ld3(s.x0, j, xj);
        const T r[3] = {xi[0] - xj[0], xi[1] - xj[1], xi[2] - xj[2]};
        rho += s.m[j] * Wk(r, c.h);
    }
    rho = gsum(rho);
    if (c.self_density) { const T z[3] = {(T)0, (T)0, (T)0}; rho += s.m[i] * Wk(z, c.h); }
    if (gl == 0 && gid < s.n) { s.rho[i] = rho; s.vol[i] = s.m[i] / rho; }
}

// compute_A_pq (sim.py:170-183)
template <typename T> __global__ void __launch_bounds__(RTHREADS) kr_Apq(RS_<T> s, RP<T> c, const T* __restrict__ x) {
    MIS_REF_GROUP();
    T x0i[3], xi[3]; ld3(s.x0, i, x0i); ld3(x, i, xi);
    T A[9];
#pragma unroll
    for (int k = 0; k < 9; k++) A[k] = (T)0;
    for (int k = gl; k < cnt; k += RG) {
        const int j = (int)lst[k];
        T x0j[3], xj[3]; ld3(s.x0, j, x0j); ld3(x, j, xj);
        const T r[3] = {x0i[0] - x0j[0], x0i[1] - x0j[1], x0i[2] - x0j[2]};
        const T w = Wk(r, c.h) * s.m[j];
        const T dx[3] = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) A[3 * a + b] += w * dx[a] * (-r[b]);
    }
#pragma unroll
    for (int k = 0; k < 9; k++) A[k] = gsum(A[k]);
    if (gl == 0 && gid < s.n) st9(s.A, i, A);
}

// compute_nabla_u + compute_sigma (sim.py:193-216): R, B, F, S of the evaluation at x
template <typename T> __global__ void __launch_bounds__(RTHREADS) kr_nabla_u(RS_<T> s, RP<T> c, const T* __restrict__ x) {
    MIS_REF_GROUP();
    T x0i[3], xi[3]; ld3(s.x0, i, x0i); ld3(x, i, xi);
    T A[9], R[9];
    ld9(s.A, i, A);
    if (c.identity_rot) {
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? (T)1 : (T)0;
    } else {
        polar_t(A, R);
    }
    T N[9], B[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { N[k] = (T)0; B[k] = (T)0; }
    for (int k = gl; k < cnt; k += RG) {
        const int j = (int)lst[k];
        T x0j[3], xj[3]; ld3(s.x0, j, x0j); ld3(x, j, xj);
        const T r[3] = {x0i[0] - x0j[0], x0i[1] - x0j[1], x0i[2] - x0j[2]};
        T nw[3]; nablaWk(r, c.h, nw);
        const T vj = s.vol[j];
        const T dx[3] = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
        T u[3]; mtv(R, dx, u);
        u[0] -= -r[0]; u[1] -= -r[1]; u[2] -= -r[2];                     // - (x0_j - x0_i)
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) { N[3 * a + b] += vj * u[a] * nw[b]; B[3 * a + b] += vj * dx[a] * nw[b]; }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) { N[k] = gsum(N[k]); B[k] = gsum(B[k]); }
    if (gl == 0 && gid < s.n) {
        const T F[9] = {(T)1 + N[0], N[3], N[6], N[1], (T)1 + N[4], N[7], N[2], N[5], (T)1 + N[8]};
        T S[9];
        sigma_t(F, s.mu[i], s.lam[i], s.ratio[i], c, S);
        st9(s.R, i, R); st9(s.B, i, B); st9(s.F, i, F); st9(s.S, i, S);
    }
}

// compute_elastic_forces (sim.py:218-235; symmetric_pair: sim_taichi.py:147-158)
template <typename T> __global__ void __launch_bounds__(RTHREADS) kr_force(RS_<T> s, RP<T> c, T* __restrict__ fel) {
    MIS_REF_GROUP();
    T x0i[3]; ld3(s.x0, i, x0i);
    T Ri[9], Fi[9], Si[9];
    ld9(s.R, i, Ri); ld9(s.F, i, Fi); ld9(s.S, i, Si);
    const T vi = s.vol[i];
    T f[3] = {(T)0, (T)0, (T)0};
    for (int k = gl; k < cnt; k += RG) {
        const int j = (int)lst[k];
        T x0j[3]; ld3(s.x0, j, x0j);
        const T r[3] = {x0i[0] - x0j[0], x0i[1] - x0j[1], x0i[2] - x0j[2]};
        T nw[3]; nablaWk(r, c.h, nw);
        const T vj = s.vol[j];
        T Rj[9], Sj[9], Fj[9];
        ld9(s.R, j, Rj); ld9(s.S, j, Sj);
        if (c.symmetric_pair) ld9(s.F, j, Fj);
        const T gj[3] = {vj * nw[0], vj * nw[1], vj * nw[2]};            // V_j n_w
        const T gi[3] = {vi * nw[0], vi * nw[1], vi * nw[2]};            // V_i n_w
        T t[3], u[3], fji[3], fij[3], a[3], b[3];
        mv(Si, gj, t); mv(Fi, t, u);                                     // f_ji = -V_i F_i S_i (V_j n_w)
        fji[0] = -vi * u[0]; fji[1] = -vi * u[1]; fji[2] = -vi * u[2];
        mv(Sj, gi, t); mv(c.symmetric_pair ? Fj : Fi, t, u);             // f_ij = V_j F S_j (V_i n_w)
        fij[0] = vj * u[0]; fij[1] = vj * u[1]; fij[2] = vj * u[2];
        mv(Rj, fij, a); mv(Ri, fji, b);
        f[0] += (T)0.5 * (a[0] - b[0]); f[1] += (T)0.5 * (a[1] - b[1]); f[2] += (T)0.5 * (a[2] - b[2]);
    }
    f[0] = gsum(f[0]); f[1] = gsum(f[1]); f[2] = gsum(f[2]);
    if (gl == 0 && gid < s.n) {
        if (!(vi < (T)3.0e38)) { f[0] = f[1] = f[2] = (T)0; }             // isolated particle: the reference loop never runs
        fel[3 * (size_t)i] = f[0]; fel[3 * (size_t)i + 1] = f[1]; fel[3 * (size_t)i + 2] = f[2];
    }
}

template <typename T> __device__ __forceinline__ T penalty_y(T y, const RP<T>& c) {        // sim.py:238-244
    if (c.no_contact || !(y < c.col_range)) return (T)0;
    const T d = c.col_range - y;
    return d * d * c.k_col;
}
template <typename T> __device__ __forceinline__ T penalty_dy(T y, const RP<T>& c) {       // d penalty / d y
    if (c.no_contact || !(y < c.col_range)) return (T)0;
    return (T)-2 * (c.col_range - y) * c.k_col;
}

// part_1 (sim.py:247-251)
template <typename T> __global__ void __launch_bounds__(256) kr_part1(RS_<T> s, RP<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const T m = s.m[i];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const size_t k = 3 * (size_t)i + a;
        const T force = s.fext[k] + s.fel[k] - c.damping * s.v[k] + (a == 1 ? penalty_y(s.x[3 * (size_t)i + 1], c) : (T)0);
        s.xn[k] = s.x[k] + (c.dt * s.v[k] + (T)0.5 * c.dt * c.dt * force / m) * s.freem[k];
    }
}
// part_2 (sim.py:253-258)
template <typename T> __global__ void __launch_bounds__(256) kr_part2(RS_<T> s, RP<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const T m = s.m[i];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const size_t k = 3 * (size_t)i + a;
        const T f1 = s.fext[k] + s.fel[k] - c.damping * s.v[k] + (a == 1 ? penalty_y(s.x[3 * (size_t)i + 1], c) : (T)0);
        const T f2 = s.fext[k] + s.feln[k] - c.damping * s.v[k] + (a == 1 ? penalty_y(s.xn[3 * (size_t)i + 1], c) : (T)0);
        s.v[k] = s.v[k] + (c.dt * (f1 + f2) / ((T)2 * m)) * s.freem[k];
    }
}
// advance (sim_taichi.py:161-172): force = ext + el + (-damping v); v' = v + dt f / m * free; x' = x + dt v' * free
template <typename T> __global__ void __launch_bounds__(256) kr_euler(RS_<T> s, RP<T> c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const T m = s.m[i];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const size_t k = 3 * (size_t)i + a;
        const T force = s.fext[k] + s.fel[k] + (-c.damping * s.v[k]);
        const T vn = s.v[k] + c.dt * force / m * s.freem[k];
        s.v[k] = vn;
        s.xn[k] = s.x[k] + c.dt * vn * s.freem[k];
    }
}

// ---------------------------------------------------------------- reverse kernels
// adjoint of the force gather: lam = dL/d fel (3n) -> Rb, Fb, Sb (9n each) at the evaluation stored in s.R, s.F, s.S
template <typename T> __global__ void __launch_bounds__(RTHREADS) kr_adj_force(RS_<T> s, RP<T> c, const T* __restrict__ lam,
                                                                               T* __restrict__ Rb, T* __restrict__ Fb, T* __restrict__ Sb) {
    MIS_REF_GROUP();
    T x0p[3]; ld3(s.x0, i, x0p);
    T Rp[9], Fp[9], Sp[9], lp[3];
    ld9(s.R, i, Rp); ld9(s.F, i, Fp); ld9(s.S, i, Sp); ld3(lam, i, lp);
    const T vp = s.vol[i];
    T rb[9], fb[9], sb[9], G[3] = {(T)0, (T)0, (T)0};
#pragma unroll
    for (int k = 0; k < 9; k++) { rb[k] = (T)0; fb[k] = (T)0; sb[k] = (T)0; }
    for (int k = gl; k < cnt; k += RG) {
        const int q = (int)lst[k];
        T x0q[3]; ld3(s.x0, q, x0q);
        const T r[3] = {x0p[0] - x0q[0], x0p[1] - x0q[1], x0p[2] - x0q[2]};
        T nw[3]; nablaWk(r, c.h, nw);                                  // nabla_W(x0_p - x0_q)
        const T vq = s.vol[q];
        T Rq[9], Fq[9], Sq[9], lq[3];
        ld9(s.R, q, Rq); ld9(s.F, q, Fq); ld9(s.S, q, Sq); ld3(lam, q, lq);
        const T gpq[3] = {vq * nw[0], vq * nw[1], vq * nw[2]};         // g_pq = V_q nabla_W(x0_p - x0_q)
        const T gqp[3] = {-vp * nw[0], -vp * nw[1], -vp * nw[2]};      // g_qp = V_p nabla_W(x0_q - x0_p)
        G[0] += gpq[0]; G[1] += gpq[1]; G[2] += gpq[2];
        T t[3], u[3], w[3];
        if (!c.symmetric_pair) {
            // p as i: Fb_p += V_p/2 (R_q^T lam_p)(S_q g_pq)^T
            mtv(Rq, lp, u); mv(Sq, gpq, t);
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) fb[3 * a + b] += (T)0.5 * vp * u[a] * t[b];
            // p as j: Rb_p += V_q/2 lam_q (F_q S_p g_qp)^T; Sb_p += V_q/2 (F_q^T R_p^T lam_q) g_qp^T
            mv(Sp, gqp, t); mv(Fq, t, u);
            mtv(Rp, lq, t); mtv(Fq, t, w);
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) { rb[3 * a + b] += (T)0.5 * vq * lq[a] * u[b]; sb[3 * a + b] += (T)0.5 * vq * w[a] * gqp[b]; }
        } else {
            // T_qp = V_q/2 lam_q^T R_p F_p S_p g_qp: everything lands on p
            mv(Sp, gqp, t); mv(Fp, t, u);                               // F_p S_p g_qp
            T rl[3]; mtv(Rp, lq, rl);                                   // R_p^T lam_q
            mtv(Fp, rl, w);                                             // F_p^T R_p^T lam_q
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    rb[3 * a + b] += (T)0.5 * vq * lq[a] * u[b];
                    fb[3 * a + b] += (T)0.5 * vq * rl[a] * t[b];
                    sb[3 * a + b] += (T)0.5 * vq * w[a] * gqp[b];
                }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) { rb[k] = gsum(rb[k]); fb[k] = gsum(fb[k]); sb[k] = gsum(sb[k]); }
    G[0] = gsum(G[0]); G[1] = gsum(G[1]); G[2] = gsum(G[2]);
    if (gl == 0 && gid < s.n) {
        // self term V_p/2 lam_p^T R_p F_p S_p G_p
        T t[3], u[3], rl[3], w[3];
        mv(Sp, G, t); mv(Fp, t, u); mtv(Rp, lp, rl); mtv(Fp, rl, w);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                rb[3 * a + b] += (T)0.5 * vp * lp[a] * u[b];
                fb[3 * a + b] += (T)0.5 * vp * rl[a] * t[b];
                sb[3 * a + b] += (T)0.5 * vp * w[a] * G[b];
            }
        if (!(vp < (T)3.0e38)) {
#pragma unroll
            for (int k = 0; k < 9; k++) { rb[k] = (T)0; fb[k] = (T)0; sb[k] = (T)0; }
        }
        st9(Rb, i, rb); st9(Fb, i, fb); st9(Sb, i, sb);
    }
}

// per particle: (Rb, Fb, Sb) -> (Ab, Bb) and the stiffness-design gradient; Rb/Fb are overwritten with Ab/Bb
template <typename T> __global__ void __launch_bounds__(128) kr_adj_particle(RS_<T> s, RP<T> c, T* __restrict__ Rb, T* __restrict__ Fb,
                                                                             const T* __restrict__ Sb, T* __restrict__ ratio_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    T A[9], R[9], F[9], B[9], rb[9], fb[9], sb[9];
    ld9(s.A, i, A); ld9(s.R, i, R); ld9(s.F, i, F); ld9(s.B, i, B); ld9(Rb, i, rb); ld9(Fb, i, fb); ld9(Sb, i, sb);
    T Ss[9];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) Ss[3 * a + b] = (T)0.5 * (sb[3 * a + b] + sb[3 * b + a]);
    T E[9];
    mtm(F, F, E);
#pragma unroll
    for (int k = 0; k < 9; k++) E[k] = (T)0.5 * (E[k] - ((k % 4 == 0) ? (T)1 : (T)0));
    const T mu = s.mu[i], la = s.lam[i];
    const T kf = c.stiff_a - s.ratio[i] * c.stiff_b;
    const T trE = E[0] + E[4] + E[8], trS = Ss[0] + Ss[4] + Ss[8];
    T kb = (T)0, Eb[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        kb += Ss[k] * ((T)2 * mu * E[k] + ((k % 4 == 0) ? la * trE : (T)0));
        Eb[k] = kf * ((T)2 * mu * Ss[k] + ((k % 4 == 0) ? la * trS : (T)0));
    }
    ratio_b[i] += -c.stiff_b * kb;
    T FE[9];
    mm(F, Eb, FE);
#pragma unroll
    for (int k = 0; k < 9; k++) fb[k] += FE[k];
    T Nb[9];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) Nb[3 * a + b] = fb[3 * b + a];                  // F = I + N^T
    T BNt[9], Bb[9], Ab[9];
    mmt(B, Nb, BNt);                                                                // N = R^T B - K: Rb += B Nb^T
#pragma unroll
    for (int k = 0; k < 9; k++) rb[k] += BNt[k];
    mm(R, Nb, Bb);                                                                  // Bb = R Nb
    if (c.identity_rot) {
#pragma unroll
        for (int k = 0; k < 9; k++) Ab[k] = (T)0;
    } else {
        polar_backward(A, R, rb, Ab);
    }
    st9(Rb, i, Ab); st9(Fb, i, Bb);
}

// positions: xb_p += sum_q [ a_qp Ab_q d0_qp + Bb_q g_qp ] - sum_q [ a_pq Ab_p d0_pq + Bb_p g_pq ]
template <typename T> __global__ void __launch_bounds__(RTHREADS) kr_adj_pos(RS_<T> s, RP<T> c, const T* __restrict__ Ab, const T* __restrict__ Bb,
                                                                             T* __restrict__ xb) {
    MIS_REF_GROUP();
    T x0p[3]; ld3(s.x0, i, x0p);
    T Ap[9], Bp[9];
    ld9(Ab, i, Ap); ld9(Bb, i, Bp);
    const T vp = s.vol[i], mp = s.m[i];
    T acc[3] = {(T)0, (T)0, (T)0};
    for (int k = gl; k < cnt; k += RG) {
        const int q = (int)lst[k];
        T x0q[3]; ld3(s.x0, q, x0q);
        const T r[3] = {x0p[0] - x0q[0], x0p[1] - x0q[1], x0p[2] - x0q[2]};       // x0_p - x0_q = -d0_pq = d0_qp
        T nw[3]; nablaWk(r, c.h, nw);
        const T w = Wk(r, c.h);
        T Aq[9], Bq[9];
        ld9(Ab, q, Aq); ld9(Bb, q, Bq);
        const T vq = s.vol[q], mq = s.m[q];
        // pair (q, p): a_qp = W m_p, d0_qp = x0_p - x0_q = r, g_qp = V_p nabla_W(x0_q - x0_p) = -V_p nw
        T t1[3], t2[3];
        const T gqp[3] = {-vp * nw[0], -vp * nw[1], -vp * nw[2]};
        mv(Aq, r, t1); mv(Bq, gqp, t2);
        // pair (p, q): a_pq = W m_q, d0_pq = -r, g_pq = V_q nw
        T t3[3], t4[3];
        const T mr[3] = {-r[0], -r[1], -r[2]};
        const T gpq[3] = {vq * nw[0], vq * nw[1], vq * nw[2]};
        mv(Ap, mr, t3); mv(Bp, gpq, t4);
#pragma unroll
        for (int a = 0; a < 3; a++) acc[a] += (w * mp * t1[a] + t2[a]) - (w * mq * t3[a] + t4[a]);
    }
    acc[0] = gsum(acc[0]); acc[1] = gsum(acc[1]); acc[2] = gsum(acc[2]);
    if (gl == 0 && gid < s.n) { xb[3 * (size_t)i] += acc[0]; xb[3 * (size_t)i + 1] += acc[1]; xb[3 * (size_t)i + 2] += acc[2]; }
}

// Reverse of one velocity-Verlet step f -> f + 1 (sim.py:247-258).  In: xb, vb = adjoints of (x_{f+1}, v_{f+1}) WITHOUT the
// elastic term of frame f + 1; felb = adjoint of fel_{f+1} collected so far (from step f + 1's force_1).  Two phases around the
// elastic adjoint J(x_{f+1})^T felb:
//   phase A: q = free dt/(2m) vb;  felb += q;  xb.y += penalty'(x_{f+1}) q.y;  (q kept in qbuf)
//   [caller: xb += J^T felb, ratio gradient]
//   phase B: r = free dt^2/(2m) xb; g = q + r (adjoint of force_1; q alone is the adjoint of force_2, and both contain -damping v_f);
//            xb' = xb (+ penalty'(x_f) g.y on y); vb' = vb + free dt xb - damping (g + q); felb' = g
template <typename T> __global__ void __launch_bounds__(256) kr_adj_stepA(RS_<T> s, RP<T> c, const T* __restrict__ x_next, T* __restrict__ xb,
                                                                          const T* __restrict__ vb, T* __restrict__ felb, T* __restrict__ qbuf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const T m = s.m[i];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const size_t k = 3 * (size_t)i + a;
        const T q = s.freem[k] * c.dt / ((T)2 * m) * vb[k];
        qbuf[k] = q;
        felb[k] += q;
        if (a == 1) xb[k] += penalty_dy(x_next[k], c) * q;
    }
}
template <typename T> __global__ void __launch_bounds__(256) kr_adj_stepB(RS_<T> s, RP<T> c, const T* __restrict__ x_prev, T* __restrict__ xb,
                                                                          T* __restrict__ vb, T* __restrict__ felb, const T* __restrict__ qbuf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const T m = s.m[i];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const size_t k = 3 * (size_t)i + a;
        const T xbn = xb[k];
        const T r = s.freem[k] * (T)0.5 * c.dt * c.dt / m * xbn;
        const T g = qbuf[k] + r;
        vb[k] = vb[k] + s.freem[k] * c.dt * xbn - c.damping * (g + qbuf[k]);
        xb[k] = xbn + (a == 1 ? penalty_dy(x_prev[k], c) * g : (T)0);
        felb[k] = g;
    }
}
// compute_loss (sim.py:269-273) and its gradient at one target frame: L += |x - xt|^2 + dt |v - vt|^2
template <typename T> __global__ void __launch_bounds__(256) kr_loss(RS_<T> s, RP<T> c, const T* __restrict__ x, const T* __restrict__ v,
                                                                     const float* __restrict__ tx, const float* __restrict__ tv, const uint32_t* __restrict__ perm,
                                                                     T* __restrict__ xb, T* __restrict__ vb, double* __restrict__ loss) {
    __shared__ double sm[8];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += gridDim.x * blockDim.x) {
        const uint32_t id = perm[i];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const T dx = x[3 * (size_t)i + a] - (T)tx[3 * (size_t)id + a];
            const T dv = v[3 * (size_t)i + a] - (T)tv[3 * (size_t)id + a];
            acc += (double)(dx * dx) + (double)(dv * dv * c.dt);
            if (xb) { xb[3 * (size_t)i + a] += (T)2 * dx; vb[3 * (size_t)i + a] += (T)2 * c.dt * dv; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sm[w];
        if (loss) atomicAdd(loss, t);
    }
}

// statics in the working precision from the raw fp32 inputs (cell-sorted): mu, lam (sim.py:288-300), ratio (sim.py:107-110)
template <typename T> __global__ void __launch_bounds__(256) kr_material(int n, const T* __restrict__ E, const T* __restrict__ nu, const T* __restrict__ design,
                                                                         T tanh_k, T* __restrict__ mu, T* __restrict__ lam, T* __restrict__ ratio) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T e = E[i], v = nu[i];
    mu[i] = e / ((T)2 * ((T)1 + v));
    lam[i] = e * v / (((T)1 + v) * ((T)1 - (T)2 * v));
    ratio[i] = (T)0.5 * t_tanh(tanh_k * design[i]) + (T)0.5;
}
// d ratio / d design folded into the accumulated ratio gradient, written in caller order
template <typename T> __global__ void __launch_bounds__(256) kr_design_grad(int n, const T* __restrict__ ratio_b, const T* __restrict__ design, T tanh_k,
                                                                            const uint32_t* __restrict__ perm, float* __restrict__ out32, double* __restrict__ out64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T th = t_tanh(tanh_k * design[i]);
    const T g = ratio_b[i] * (T)0.5 * tanh_k * ((T)1 - th * th);
    if (out32) out32[perm[i]] = (float)g;
    if (out64) out64[perm[i]] = (double)g;
}
// caller order (fp32 or fp64, n x dim) <-> slot order (T)
template <typename T, typename U> __global__ void __launch_bounds__(256) kr_gather(const U* __restrict__ src, const uint32_t* __restrict__ perm, int n, int dim, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = perm[i];
    for (int a = 0; a < dim; a++) dst[(size_t)dim * i + a] = (T)src[(size_t)dim * id + a];
}
template <typename T, typename U> __global__ void __launch_bounds__(256) kr_scatter(const T* __restrict__ src, const uint32_t* __restrict__ perm, int n, int dim, U* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = perm[i];
    for (int a = 0; a < dim; a++) dst[(size_t)dim * id + a] = (U)src[(size_t)dim * i + a];
}
// components [off, off + dim) of a float4 array (slot order) -> T array
template <typename T> __global__ void __launch_bounds__(256) kr_pull(const float4* __restrict__ src, int n, int off, int dim, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    const float c[4] = {v.x, v.y, v.z, v.w};
    for (int a = 0; a < dim; a++) dst[(size_t)dim * i + a] = (T)c[off + a];
}
template <typename T, typename U> __global__ void __launch_bounds__(256) kr_cast(const U* __restrict__ src, long long count, T* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = (T)src[i];
}
template <typename T> __global__ void __launch_bounds__(256) kr_fill3(int n, T a, T b, T cc, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[3 * (size_t)i] = a; dst[3 * (size_t)i + 1] = b; dst[3 * (size_t)i + 2] = cc;
}

#undef MIS_REF_GROUP
}  // namespace ref
}  // namespace mis
