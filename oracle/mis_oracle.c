/*
 * mis_oracle.c -- CPU restatement of the reference's per-step meshless particle
 * update (sim.py), used ONLY as test infrastructure.
 *
 *   THIS IS THE ORACLE.  It is the checker, never the product: only tests/,
 *   __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 *   may build, load or call it.  The CUDA product path never links it.
 *
 *   PARITY UNPINNED (by the reference): the reference tree holds no tests, no
 *   golden vectors and cannot run here (NVIDIA Warp / Taichi absent, assets at
 *   non-existent absolute paths, device="cuda" hard-coded).  The pins this
 *   oracle does have are (1) the known-answer properties of the reference's
 *   own mathematics (tests/test_oracle_known_answers.py), (2) an independent
 *   fp64 numpy restatement (oracle/np_oracle.py) and (3) brute-force
 *   neighbour search.  Two pieces of arithmetic live in an un-vendored,
 *   un-pinned third-party dependency (NVIDIA Warp, `warp-lang`, no version in
 *   the reference tree): wp.HashGrid / hash_grid_query and wp.svd3.  Their
 *   published behaviour is restated below (hash grid: truncating cell coords,
 *   +2^20 offset, mod dim, x-fastest 27-cell walk; svd3: Jacobi on A^T A for V,
 *   orthogonal-triangular factorisation for U, both proper rotations).
 *
 * Every function cites the sim.py lines it follows.  All arithmetic is fp32,
 * in the literal operation order of the Python source (operator precedence,
 * left-to-right association), compiled with -ffp-contract=off so no FMA is
 * formed.
 *
 * Modes of the step:
 *   ORC_FAITHFUL  per-candidate svd3 + compute_sigma inside the force loop,
 *                 every candidate of the 27-cell walk visited (sim.py:218-235
 *                 as written).  The timed "reference CPU path".
 *   ORC_CACHED    R_j and S_j evaluated once per particle and candidates with
 *                 q >= 2 skipped.  Bit-identical results to ORC_FAITHFUL (both
 *                 are pure functions of per-particle data; skipped candidates
 *                 contribute exact zeros) at ~1/50 of the cost; used for the
 *                 longer parity runs.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Working precision.  Default: float (real = wp.float32, sim.py:22).  -DORC_DOUBLE builds the same source in double
 * (libmis_oracle_f64.so): the Taichi prototype's precision (real = ti.f64, options.py:3) and an "exact arithmetic"
 * reference for the fp32 build.  Every literal below is exactly representable in both.                            */
#ifdef ORC_DOUBLE
typedef double real;
#define R_SQRT sqrt
#define R_FABS fabs
#define R_TANH tanh
#else
typedef float real;
#define R_SQRT sqrtf
#define R_FABS fabsf
#define R_TANH tanhf
#endif

typedef struct { real x, y, z; } v3;
typedef struct { real m[3][3]; } m33;

typedef struct {
    real h;            /* sim.py:25  */
    real damping;      /* sim.py:26  */
    real dt;           /* sim.py:65  */
    real k_col;        /* sim.py:68  */
    real col_range;    /* sim.py:69  */
    int   grid_x, grid_y, grid_z; /* sim.py:123-125 */
    /* variant switches (sim_taichi.py deltas); 0 = sim.py */
    int   symmetric_pair;   /* 1: f_ij uses F_j (sim_taichi.py:157)       */
    int   identity_rot;     /* 1: R = I (sim_taichi.py:129)               */
    int   self_density;     /* 1: rho includes j == i (sim_taichi.py:97)  */
    int   euler;            /* 1: symplectic Euler (sim_taichi.py:167-172)*/
    int   no_contact;       /* 1: no ground penalty (sim_taichi.py)       */
    real stiff_a, stiff_b; /* stiffness factor = a - b*ratio (200,199 | 1,1) */
} OrcParams;

typedef struct {
    int n;
    OrcParams p;
    /* static */
    v3 *x0; real *mass, *rho, *vol, *E, *nu, *mu, *lam, *design, *ratio;
    v3 *fext, *free_;
    /* dynamic (current frame and next frame) */
    v3 *x, *v, *xn, *vn, *fel, *feln;
    m33 *A, *F, *Rc, *Sc;
    /* hash grid (sim.py:123-127) */
    real cell_width, cell_width_inv;
    int *point_cell, *point_ids, *cell_start, *cell_end;
    int ncells;
    int threads;
    int order;   /* 0: reference walk order, 1: reversed (noise-floor probe) */
} Orc;

/* ------------------------------------------------------------------ vec/mat
 * Warp's vec/mat operators, in Warp's evaluation order (mul(mat,vec) and
 * mul(mat,mat) accumulate k = 0,1,2; dot accumulates x,y,z).               */
static inline v3 v3_make(real x, real y, real z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_scale(real s, v3 a) { return v3_make(s * a.x, s * a.y, s * a.z); }
static inline v3 v3_div(v3 a, real s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline v3 v3_cwmul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline real v3_length(v3 a) { return R_SQRT(a.x * a.x + a.y * a.y + a.z * a.z); }

static inline m33 m33_zero(void) { m33 r; memset(&r, 0, sizeof r); return r; }
static inline m33 m33_identity(void) { m33 r = m33_zero(); r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.f; return r; }
static inline m33 m33_outer(v3 a, v3 b) {
    m33 r; const real av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = av[i] * bv[j];
    return r;
}
static inline m33 m33_scale(real s, m33 a) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a.m[i][j] = s * a.m[i][j]; return a; }
static inline m33 m33_add(m33 a, m33 b) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a.m[i][j] = a.m[i][j] + b.m[i][j]; return a; }
static inline m33 m33_sub(m33 a, m33 b) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a.m[i][j] = a.m[i][j] - b.m[i][j]; return a; }
static inline m33 m33_transpose(m33 a) { m33 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
static inline m33 m33_mul(m33 a, m33 b) {
    m33 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        real s = a.m[i][0] * b.m[0][j];
        s = s + a.m[i][1] * b.m[1][j];
        s = s + a.m[i][2] * b.m[2][j];
        r.m[i][j] = s;
    }
    return r;
}
static inline v3 m33_mulv(m33 a, v3 b) {
    real r[3];
    for (int i = 0; i < 3; i++) {
        real s = a.m[i][0] * b.x;
        s = s + a.m[i][1] * b.y;
        s = s + a.m[i][2] * b.z;
        r[i] = s;
    }
    return v3_make(r[0], r[1], r[2]);
}
static inline real m33_trace(m33 a) { return a.m[0][0] + a.m[1][1] + a.m[2][2]; }

/* ------------------------------------------------------------------ kernels */
#define ORC_PI ((real)3.14159265358979323846)   /* real(wp.pi) */

/* sim.py:133-141 */
static inline real W_kernel(v3 xij, real h) {
    real q = v3_length(xij) / h;
    real ret = 0.f;
    if (q < 1.f) {
        ret = 1.f / (ORC_PI * h * h * h) * (1.f - 1.5f * q * q + 0.75f * q * q * q);
    } else if (q >= 1.f && q < 2.f) {
        ret = 1.f / (4.f * ORC_PI * h * h * h) * (2.f - q) * (2.f - q) * (2.f - q);
    }
    return ret;
}

/* sim.py:143-151 */
static inline v3 nabla_W_kernel(v3 xij, real h) {
    real q = v3_length(xij) / h;
    v3 ret = v3_make(0.f, 0.f, 0.f);
    if (q < 1.f) {
        v3 a = v3_div(v3_div(v3_scale(-3.f, xij), h), h);
        v3 b = v3_div(v3_div(v3_scale(0.75f * 3.f * q, xij), h), h);
        ret = v3_scale(1.f / (ORC_PI * h * h * h), v3_add(a, b));
    } else if (q >= 1.f && q < 2.f) {
        real c = 1.f / (4.f * ORC_PI * h * h * h) * -3.f * (2.f - q) * (2.f - q);
        ret = v3_div(v3_scale(c, xij), q * h * h);
    }
    return ret;
}

/* ------------------------------------------------------------------ svd3
 * sim.py:185-191 compute_R_i = U V^T of wp.svd3(A).  Warp's svd3 (McAdams et
 * al. 2011: Jacobi eigen-analysis of A^T A gives V, an orthogonal-triangular
 * factorisation of A V gives U, both kept proper rotations, singular values
 * sorted descending, any reflection carried by the sign of sigma_3) is not in
 * the reference tree.  Restated with exact (not approximate) Givens angles and
 * a fixed sweep count so the result is the rotation of the polar
 * decomposition to fp32 round-off.  The CUDA path implements the same
 * sequence; tests compare the two within tolerance, not bitwise.            */
#define ORC_JACOBI_SWEEPS 6
static inline void jacobi_rot(real S[3][3], real V[3][3], int p, int q) {
    real apq = S[p][q];
    if (R_FABS(apq) <= 1e-30f) return;
    real theta = (S[q][q] - S[p][p]) / (2.f * apq);
    real t = 1.f / (R_FABS(theta) + R_SQRT(theta * theta + 1.f));
    if (theta < 0.f) t = -t;
    real c = 1.f / R_SQRT(t * t + 1.f);
    real s = t * c;
    /* S <- J^T S J with J = [[c, s], [-s, c]] on (p,q) */
    real spp = S[p][p], sqq = S[q][q];
    S[p][p] = spp - t * apq;
    S[q][q] = sqq + t * apq;
    S[p][q] = 0.f; S[q][p] = 0.f;
    int r = 3 - p - q;
    real srp = S[r][p], srq = S[r][q];
    S[r][p] = c * srp - s * srq; S[p][r] = S[r][p];
    S[r][q] = s * srp + c * srq; S[q][r] = S[r][q];
    for (int k = 0; k < 3; k++) {
        real vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq;
        V[k][q] = s * vkp + c * vkq;
    }
}

static m33 polar_rotation(m33 Am) {
    real S[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        real s = Am.m[0][i] * Am.m[0][j];
        s = s + Am.m[1][i] * Am.m[1][j];
        s = s + Am.m[2][i] * Am.m[2][j];
        S[i][j] = s;
    }
    for (int sweep = 0; sweep < ORC_JACOBI_SWEEPS; sweep++) {
        jacobi_rot(S, V, 0, 1);
        jacobi_rot(S, V, 0, 2);
        jacobi_rot(S, V, 1, 2);
    }
    /* B = A V, column norms^2 */
    real B[3][3], nrm[3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        real s = Am.m[i][0] * V[0][j];
        s = s + Am.m[i][1] * V[1][j];
        s = s + Am.m[i][2] * V[2][j];
        B[i][j] = s;
    }
    for (int j = 0; j < 3; j++) nrm[j] = B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j];
    /* sort columns by descending norm; a swap negates one column so det V stays +1 */
#define ORC_CSWAP(a, b)                                                           \
    if (nrm[a] < nrm[b]) {                                                        \
        real tn = nrm[a]; nrm[a] = nrm[b]; nrm[b] = tn;                          \
        for (int k = 0; k < 3; k++) {                                             \
            real tb = B[k][a]; B[k][a] = B[k][b]; B[k][b] = -tb;                 \
            real tv = V[k][a]; V[k][a] = V[k][b]; V[k][b] = -tv;                 \
        }                                                                         \
    }
    ORC_CSWAP(0, 1) ORC_CSWAP(0, 2) ORC_CSWAP(1, 2)
#undef ORC_CSWAP
    /* U: Gram-Schmidt on the two dominant columns, third = cross (det U = +1) */
    real U[3][3];
    real n0 = R_SQRT(nrm[0]);
    if (!(n0 > 1e-30f)) return m33_identity();
    real u0[3] = {B[0][0] / n0, B[1][0] / n0, B[2][0] / n0};
    real d = u0[0] * B[0][1] + u0[1] * B[1][1] + u0[2] * B[2][1];
    real w1[3] = {B[0][1] - d * u0[0], B[1][1] - d * u0[1], B[2][1] - d * u0[2]};
    real n1 = R_SQRT(w1[0] * w1[0] + w1[1] * w1[1] + w1[2] * w1[2]);
    real u1[3];
    if (n1 > 1e-30f) {
        u1[0] = w1[0] / n1; u1[1] = w1[1] / n1; u1[2] = w1[2] / n1;
    } else {
        /* rank-1 A: any unit vector orthogonal to u0 */
        int k = (R_FABS(u0[0]) <= R_FABS(u0[1]) && R_FABS(u0[0]) <= R_FABS(u0[2])) ? 0
              : (R_FABS(u0[1]) <= R_FABS(u0[2]) ? 1 : 2);
        real e[3] = {0, 0, 0}; e[k] = 1.f;
        real dd = u0[k];
        real ww[3] = {e[0] - dd * u0[0], e[1] - dd * u0[1], e[2] - dd * u0[2]};
        real nn = R_SQRT(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
        u1[0] = ww[0] / nn; u1[1] = ww[1] / nn; u1[2] = ww[2] / nn;
    }
    real u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    for (int k = 0; k < 3; k++) { U[k][0] = u0[k]; U[k][1] = u1[k]; U[k][2] = u2[k]; }
    m33 R;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        real s = U[i][0] * V[j][0];
        s = s + U[i][1] * V[j][1];
        s = s + U[i][2] * V[j][2];
        R.m[i][j] = s;
    }
    return R;
}

/* sim.py:185-191 (sim_taichi.py:129 overwrites R with identity) */
static inline m33 compute_R_i(const Orc *o, m33 A) {
    if (o->p.identity_rot) return m33_identity();
    return polar_rotation(A);
}

/* sim.py:212-216 */
static inline m33 compute_sigma(const Orc *o, m33 F, real mu, real lam, real ratio) {
    m33 E = m33_scale(0.5f, m33_sub(m33_mul(m33_transpose(F), F), m33_identity()));
    m33 s = m33_add(m33_scale(2.f * mu, E), m33_scale(lam * m33_trace(E), m33_identity()));
    return m33_scale(o->p.stiff_a - ratio * o->p.stiff_b, s);   /* mat * scalar: same products */
}

/* sim.py:238-244 */
static inline v3 collision_penalty(const Orc *o, v3 pos) {
    v3 pen = v3_make(0.f, 0.f, 0.f);
    if (!o->p.no_contact && pos.y < o->p.col_range) {
        real delta = o->p.col_range - pos.y;
        pen.y = delta * delta * o->p.k_col;
    }
    return pen;
}

/* ------------------------------------------------------------------ hash grid
 * wp.HashGrid(dim_x, dim_y, dim_z).build(points, radius) and hash_grid_query,
 * sim.py:123-127 and the query call sites sim.py:161,178,203,224.           */
static inline int hg_index(const Orc *o, int x, int y, int z) {
    const int origin = 1 << 20;
    x += origin; y += origin; z += origin;
    if (x < 0) x = 0; if (y < 0) y = 0; if (z < 0) z = 0;
    int cx = x % o->p.grid_x, cy = y % o->p.grid_y, cz = z % o->p.grid_z;
    return cz * (o->p.grid_x * o->p.grid_y) + cy * o->p.grid_x + cx;
}
static inline void hg_cell_coords(const Orc *o, v3 p, int c[3]) {
    c[0] = (int)(p.x * o->cell_width_inv);
    c[1] = (int)(p.y * o->cell_width_inv);
    c[2] = (int)(p.z * o->cell_width_inv);
}
static void hg_build(Orc *o) {
    int n = o->n;
    o->cell_width = 2.f * o->p.h;                 /* real(2.) * h */
    o->cell_width_inv = 1.f / o->cell_width;
    o->ncells = o->p.grid_x * o->p.grid_y * o->p.grid_z;
    o->cell_start = (int *)calloc((size_t)o->ncells, sizeof(int));
    o->cell_end = (int *)calloc((size_t)o->ncells, sizeof(int));
    int *count = (int *)calloc((size_t)o->ncells + 1, sizeof(int));
    for (int i = 0; i < n; i++) {
        int c[3]; hg_cell_coords(o, o->x0[i], c);
        o->point_cell[i] = hg_index(o, c[0], c[1], c[2]);
        count[o->point_cell[i] + 1]++;
    }
    for (int c = 0; c < o->ncells; c++) count[c + 1] += count[c];
    for (int c = 0; c < o->ncells; c++) { o->cell_start[c] = count[c]; o->cell_end[c] = count[c + 1]; }
    /* stable counting sort: ascending particle id inside a cell */
    for (int i = 0; i < n; i++) o->point_ids[count[o->point_cell[i]]++] = i;
    free(count);
}

typedef struct { int xs, ys, zs, xe, ye, ze; } HgQuery;
static inline HgQuery hg_query(const Orc *o, v3 pos, real radius) {
    HgQuery q;
    q.xs = (int)((pos.x - radius) * o->cell_width_inv);
    q.ys = (int)((pos.y - radius) * o->cell_width_inv);
    q.zs = (int)((pos.z - radius) * o->cell_width_inv);
    q.xe = (int)((pos.x + radius) * o->cell_width_inv);
    q.ye = (int)((pos.y + radius) * o->cell_width_inv);
    q.ze = (int)((pos.z + radius) * o->cell_width_inv);
    /* never visit a physical cell twice */
    if (q.xe > q.xs + o->p.grid_x - 1) q.xe = q.xs + o->p.grid_x - 1;
    if (q.ye > q.ys + o->p.grid_y - 1) q.ye = q.ys + o->p.grid_y - 1;
    if (q.ze > q.zs + o->p.grid_z - 1) q.ze = q.zs + o->p.grid_z - 1;
    return q;
}
/* materialise the candidate walk (x fastest, then y, then z; ascending sorted
 * slot inside a cell) into cand[]; returns the count                        */
static int hg_candidates(const Orc *o, v3 pos, int *cand, int cap) {
    HgQuery q = hg_query(o, pos, 2.f * o->p.h);
    int m = 0;
    for (int z = q.zs; z <= q.ze; z++)
        for (int y = q.ys; y <= q.ye; y++)
            for (int x = q.xs; x <= q.xe; x++) {
                int c = hg_index(o, x, y, z);
                for (int s = o->cell_start[c]; s < o->cell_end[c]; s++) {
                    if (m < cap) cand[m] = o->point_ids[s];
                    m++;
                }
            }
    if (m > cap) { fprintf(stderr, "mis_oracle: candidate overflow (%d > %d)\n", m, cap); abort(); }
    if (o->order == 1)
        for (int a = 0, b = m - 1; a < b; a++, b--) { int t = cand[a]; cand[a] = cand[b]; cand[b] = t; }
    return m;
}

#define ORC_MAX_CAND 16384

/* ------------------------------------------------------------------ passes */
/* sim.py:154-167 */
static void compute_v_i(Orc *o) {
    int n = o->n; real h = o->p.h;
#pragma omp parallel num_threads(o->threads)
    {
        int *cand = (int *)malloc(ORC_MAX_CAND * sizeof(int));
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {
            v3 x = o->x0[i];
            real r = 0.f;
            int m = hg_candidates(o, x, cand, ORC_MAX_CAND);
            for (int c = 0; c < m; c++) {
                int index = cand[c];
                if (index != i || o->p.self_density)
                    r += o->mass[index] * W_kernel(v3_sub(x, o->x0[index]), h);
            }
            o->rho[i] = r;
            o->vol[i] = o->mass[i] / r;
        }
        free(cand);
    }
}

/* sim.py:170-183 */
static void compute_A_pq(Orc *o, const v3 *pos, m33 *Aout) {
    int n = o->n; real h = o->p.h;
#pragma omp parallel num_threads(o->threads)
    {
        int *cand = (int *)malloc(ORC_MAX_CAND * sizeof(int));
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {
            v3 x = pos[i], x0 = o->x0[i];
            m33 a = m33_zero();
            int m = hg_candidates(o, x0, cand, ORC_MAX_CAND);
            for (int c = 0; c < m; c++) {
                int j = cand[c];
                if (j != i) {
                    real w = W_kernel(v3_sub(x0, o->x0[j]), h);
                    a = m33_add(a, m33_scale(w * o->mass[j], m33_outer(v3_sub(pos[j], x), v3_sub(o->x0[j], x0))));
                }
            }
            Aout[i] = a;
        }
        free(cand);
    }
}

/* sim.py:193-209 */
static void compute_nabla_u(Orc *o, const v3 *pos, const m33 *Ain, m33 *Fout) {
    int n = o->n; real h = o->p.h;
#pragma omp parallel num_threads(o->threads)
    {
        int *cand = (int *)malloc(ORC_MAX_CAND * sizeof(int));
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {
            v3 x0 = o->x0[i], x = pos[i];
            m33 n_u = m33_zero();
            m33 R = compute_R_i(o, Ain[i]);
            m33 Rt = m33_transpose(R);
            int m = hg_candidates(o, x0, cand, ORC_MAX_CAND);
            for (int c = 0; c < m; c++) {
                int j = cand[c];
                if (j != i) {
                    v3 n_w = nabla_W_kernel(v3_sub(x0, o->x0[j]), h);
                    v3 u = v3_sub(m33_mulv(Rt, v3_sub(pos[j], x)), v3_sub(o->x0[j], x0));
                    n_u = m33_add(n_u, m33_scale(o->vol[j], m33_outer(u, n_w)));
                }
            }
            Fout[i] = m33_add(m33_identity(), m33_transpose(n_u));
        }
        free(cand);
    }
}

/* sim.py:218-235.  mode 0 = faithful (as written), mode 1 = cached R_j, S_j. */
static void compute_elastic_forces(Orc *o, const m33 *Ain, const m33 *Fin, v3 *fout, int mode) {
    int n = o->n; real h = o->p.h;
    if (mode == 1) {
#pragma omp parallel for schedule(static) num_threads(o->threads)
        for (int i = 0; i < n; i++) {
            o->Rc[i] = compute_R_i(o, Ain[i]);
            o->Sc[i] = compute_sigma(o, Fin[i], o->mu[i], o->lam[i], o->ratio[i]);
        }
    }
#pragma omp parallel num_threads(o->threads)
    {
        int *cand = (int *)malloc(ORC_MAX_CAND * sizeof(int));
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < n; i++) {
            v3 x0 = o->x0[i];
            v3 force = v3_make(0.f, 0.f, 0.f);
            int m = hg_candidates(o, x0, cand, ORC_MAX_CAND);
            m33 R_i = compute_R_i(o, Ain[i]);
            m33 s_i = compute_sigma(o, Fin[i], o->mu[i], o->lam[i], o->ratio[i]);
            for (int c = 0; c < m; c++) {
                int j = cand[c];
                if (j == i) continue;
                m33 s_j, R_j;
                v3 n_w = nabla_W_kernel(v3_sub(o->x0[i], o->x0[j]), h);
                if (mode == 1) {
                    /* q >= 2: n_w == 0 exactly, the pair adds exact zeros */
                    if (n_w.x == 0.f && n_w.y == 0.f && n_w.z == 0.f) continue;
                    s_j = o->Sc[j]; R_j = o->Rc[j];
                } else {
                    s_j = compute_sigma(o, Fin[j], o->mu[j], o->lam[j], o->ratio[j]);
                    R_j = compute_R_i(o, Ain[j]);
                }
                /* f_ji = -volume[i] * def_grad[i] @ s_i @ (volume[j] * n_w) */
                v3 f_ji = m33_mulv(m33_mul(m33_scale(-o->vol[i], Fin[i]), s_i), v3_scale(o->vol[j], n_w));
                /* f_ij = volume[j] * def_grad[i] @ s_j @ (volume[i] * n_w): F_i, not F_j (sim.py:233) */
                const m33 *Fpair = o->p.symmetric_pair ? &Fin[j] : &Fin[i];
                v3 f_ij = m33_mulv(m33_mul(m33_scale(o->vol[j], *Fpair), s_j), v3_scale(o->vol[i], n_w));
                force = v3_add(force, v3_scale(0.5f, v3_sub(m33_mulv(R_j, f_ij), m33_mulv(R_i, f_ji))));
            }
            fout[i] = force;
        }
        free(cand);
    }
}

/* sim.py:247-251 */
static void part_1(Orc *o) {
    real dt = o->p.dt, damping = o->p.damping;
#pragma omp parallel for schedule(static) num_threads(o->threads)
    for (int i = 0; i < o->n; i++) {
        v3 force = v3_add(v3_sub(v3_add(o->fext[i], o->fel[i]), v3_scale(damping, o->v[i])), collision_penalty(o, o->x[i]));
        v3 dx = v3_add(v3_scale(dt, o->v[i]), v3_div(v3_scale(0.5f * dt * dt, force), o->mass[i]));
        o->xn[i] = v3_add(o->x[i], v3_cwmul(dx, o->free_[i]));
    }
}
/* sim.py:253-258 */
static void part_2(Orc *o) {
    real dt = o->p.dt, damping = o->p.damping;
#pragma omp parallel for schedule(static) num_threads(o->threads)
    for (int i = 0; i < o->n; i++) {
        v3 f1 = v3_add(v3_sub(v3_add(o->fext[i], o->fel[i]), v3_scale(damping, o->v[i])), collision_penalty(o, o->x[i]));
        v3 f2 = v3_add(v3_sub(v3_add(o->fext[i], o->feln[i]), v3_scale(damping, o->v[i])), collision_penalty(o, o->xn[i]));
        v3 dv = v3_div(v3_scale(dt, v3_add(f1, f2)), 2.f * o->mass[i]);
        o->vn[i] = v3_add(o->v[i], v3_cwmul(dv, o->free_[i]));
    }
}
/* sim_taichi.py:167-172 (advance; damping force folded in, sim_taichi.py:161-164) */
static void advance_euler(Orc *o) {
    real dt = o->p.dt, damping = o->p.damping;
#pragma omp parallel for schedule(static) num_threads(o->threads)
    for (int i = 0; i < o->n; i++) {
        v3 force = v3_add(v3_add(o->fext[i], o->fel[i]), v3_scale(-damping, o->v[i]));
        o->vn[i] = v3_add(o->v[i], v3_cwmul(v3_div(v3_scale(dt, force), o->mass[i]), o->free_[i]));
        o->xn[i] = v3_add(o->x[i], v3_cwmul(v3_scale(dt, o->vn[i]), o->free_[i]));
    }
}

/* ------------------------------------------------------------------ C API */
#define ORC_ALLOC(T, cnt) ((T *)calloc((size_t)(cnt), sizeof(T)))

Orc *orc_create(int n, const real *x0, const OrcParams *p, int threads) {
    Orc *o = ORC_ALLOC(Orc, 1);
    o->n = n; o->p = *p;
    if (o->p.grid_x < 1) o->p.grid_x = 1;
    if (o->p.grid_y < 1) o->p.grid_y = 1;
    if (o->p.grid_z < 1) o->p.grid_z = 1;
    if (o->p.stiff_a == 0.f && o->p.stiff_b == 0.f) { o->p.stiff_a = 200.f; o->p.stiff_b = 199.f; }
#ifdef _OPENMP
    o->threads = threads > 0 ? threads : omp_get_max_threads();
#else
    o->threads = 1; (void)threads;
#endif
    o->x0 = ORC_ALLOC(v3, n); memcpy(o->x0, x0, sizeof(v3) * (size_t)n);
    o->mass = ORC_ALLOC(real, n); o->rho = ORC_ALLOC(real, n); o->vol = ORC_ALLOC(real, n);
    o->E = ORC_ALLOC(real, n); o->nu = ORC_ALLOC(real, n); o->mu = ORC_ALLOC(real, n); o->lam = ORC_ALLOC(real, n);
    o->design = ORC_ALLOC(real, n); o->ratio = ORC_ALLOC(real, n);
    o->fext = ORC_ALLOC(v3, n); o->free_ = ORC_ALLOC(v3, n);
    o->x = ORC_ALLOC(v3, n); o->v = ORC_ALLOC(v3, n); o->xn = ORC_ALLOC(v3, n); o->vn = ORC_ALLOC(v3, n);
    o->fel = ORC_ALLOC(v3, n); o->feln = ORC_ALLOC(v3, n);
    o->A = ORC_ALLOC(m33, n); o->F = ORC_ALLOC(m33, n); o->Rc = ORC_ALLOC(m33, n); o->Sc = ORC_ALLOC(m33, n);
    o->point_cell = ORC_ALLOC(int, n); o->point_ids = ORC_ALLOC(int, n);
    for (int i = 0; i < n; i++) o->free_[i] = v3_make(1.f, 1.f, 1.f);   /* sim.py:81 */
    hg_build(o);
    return o;
}

void orc_destroy(Orc *o) {
    if (!o) return;
    free(o->x0); free(o->mass); free(o->rho); free(o->vol); free(o->E); free(o->nu); free(o->mu); free(o->lam);
    free(o->design); free(o->ratio); free(o->fext); free(o->free_); free(o->x); free(o->v); free(o->xn); free(o->vn);
    free(o->fel); free(o->feln); free(o->A); free(o->F); free(o->Rc); free(o->Sc);
    free(o->point_cell); free(o->point_ids); free(o->cell_start); free(o->cell_end); free(o);
}

void orc_set_threads(Orc *o, int threads) { if (threads > 0) o->threads = threads; }
void orc_set_order(Orc *o, int order) { o->order = order; }

/* sim.py:288-300: mu, lam from per-particle E, nu */
static void lame(Orc *o) {
    for (int i = 0; i < o->n; i++) {
        real E = o->E[i], nu = o->nu[i];
        o->mu[i] = E / (2.f * (1.f + nu));
        o->lam[i] = E * nu / ((1.f + nu) * (1.f - 2.f * nu));
    }
}
void orc_set_youngs_modulus(Orc *o, const real *E) { memcpy(o->E, E, sizeof(real) * (size_t)o->n); lame(o); }
void orc_set_poisson_ratio(Orc *o, const real *nu) { memcpy(o->nu, nu, sizeof(real) * (size_t)o->n); lame(o); }
/* sim.py:306-308 */
void orc_set_mass(Orc *o, const real *m) { memcpy(o->mass, m, sizeof(real) * (size_t)o->n); compute_v_i(o); }
/* sim.py:279-286 */
void orc_set_external_forces(Orc *o, const real *f) { memcpy(o->fext, f, sizeof(v3) * (size_t)o->n); }
void orc_set_free_points(Orc *o, const real *d) { memcpy(o->free_, d, sizeof(v3) * (size_t)o->n); }
/* sim.py:107-110 (tanh_k = 3; sim_taichi.py:81 uses 5) */
void orc_set_design(Orc *o, const real *x, real tanh_k) {
    memcpy(o->design, x, sizeof(real) * (size_t)o->n);
    for (int i = 0; i < o->n; i++) o->ratio[i] = 0.5f * R_TANH(tanh_k * x[i]) + 0.5f;
}
void orc_set_ratio(Orc *o, const real *ratio) { memcpy(o->ratio, ratio, sizeof(real) * (size_t)o->n); }

/* sim.py:261-266 startup + sim.py:349-351 frame-0 priming */
void orc_startup(Orc *o, const real *v0, int mode) {
    for (int i = 0; i < o->n; i++) { o->x[i] = o->x0[i]; o->v[i] = v3_make(v0[0], v0[1], v0[2]); }
    compute_A_pq(o, o->x, o->A);
    compute_nabla_u(o, o->x, o->A, o->F);
    compute_elastic_forces(o, o->A, o->F, o->fel, mode);
}
/* restart from an arbitrary state (x, v): re-primes forces at x */
void orc_set_state(Orc *o, const real *x, const real *v, int mode) {
    memcpy(o->x, x, sizeof(v3) * (size_t)o->n); memcpy(o->v, v, sizeof(v3) * (size_t)o->n);
    compute_A_pq(o, o->x, o->A);
    compute_nabla_u(o, o->x, o->A, o->F);
    compute_elastic_forces(o, o->A, o->F, o->fel, mode);
}

/* sim.py:352-358 loop body, n_steps times */
void orc_step(Orc *o, int n_steps, int mode) {
    for (int s = 0; s < n_steps; s++) {
        if (o->p.euler) {
            /* sim_taichi.py:174-182: forces at frame f, then advance */
            advance_euler(o);
            v3 *t = o->x; o->x = o->xn; o->xn = t;
            t = o->v; o->v = o->vn; o->vn = t;
            compute_A_pq(o, o->x, o->A);
            compute_nabla_u(o, o->x, o->A, o->F);
            compute_elastic_forces(o, o->A, o->F, o->fel, mode);
            continue;
        }
        part_1(o);
        compute_A_pq(o, o->xn, o->A);
        compute_nabla_u(o, o->xn, o->A, o->F);
        compute_elastic_forces(o, o->A, o->F, o->feln, mode);
        part_2(o);
        v3 *t = o->x; o->x = o->xn; o->xn = t;
        t = o->v; o->v = o->vn; o->vn = t;
        t = o->fel; o->fel = o->feln; o->feln = t;
    }
}

/* one force evaluation at an arbitrary configuration (no integration) */
void orc_eval(Orc *o, const real *pos, int mode, real *A_out, real *R_out, real *F_out, real *S_out, real *f_out) {
    int n = o->n;
    v3 *P = ORC_ALLOC(v3, n); memcpy(P, pos, sizeof(v3) * (size_t)n);
    m33 *A = ORC_ALLOC(m33, n), *F = ORC_ALLOC(m33, n); v3 *f = ORC_ALLOC(v3, n);
    compute_A_pq(o, P, A);
    compute_nabla_u(o, P, A, F);
    compute_elastic_forces(o, A, F, f, mode);
    if (A_out) memcpy(A_out, A, sizeof(m33) * (size_t)n);
    if (F_out) memcpy(F_out, F, sizeof(m33) * (size_t)n);
    if (f_out) memcpy(f_out, f, sizeof(v3) * (size_t)n);
    for (int i = 0; i < n; i++) {
        if (R_out) { m33 R = compute_R_i(o, A[i]); memcpy(R_out + 9 * (size_t)i, &R, sizeof R); }
        if (S_out) { m33 S = compute_sigma(o, F[i], o->mu[i], o->lam[i], o->ratio[i]); memcpy(S_out + 9 * (size_t)i, &S, sizeof S); }
    }
    free(P); free(A); free(F); free(f);
}

void orc_get_state(const Orc *o, real *x, real *v) {
    if (x) memcpy(x, o->x, sizeof(v3) * (size_t)o->n);
    if (v) memcpy(v, o->v, sizeof(v3) * (size_t)o->n);
}
void orc_get_forces(const Orc *o, real *fel) { memcpy(fel, o->fel, sizeof(v3) * (size_t)o->n); }
void orc_get_fields(const Orc *o, real *A, real *F) {
    if (A) memcpy(A, o->A, sizeof(m33) * (size_t)o->n);
    if (F) memcpy(F, o->F, sizeof(m33) * (size_t)o->n);
}
void orc_get_volume(const Orc *o, real *rho, real *vol) {
    if (rho) memcpy(rho, o->rho, sizeof(real) * (size_t)o->n);
    if (vol) memcpy(vol, o->vol, sizeof(real) * (size_t)o->n);
}
void orc_get_lame(const Orc *o, real *mu, real *lam, real *ratio) {
    if (mu) memcpy(mu, o->mu, sizeof(real) * (size_t)o->n);
    if (lam) memcpy(lam, o->lam, sizeof(real) * (size_t)o->n);
    if (ratio) memcpy(ratio, o->ratio, sizeof(real) * (size_t)o->n);
}
/* hash-grid structures: per-particle linear cell index, unwrapped integer cell
 * coordinates, the cell-sorted particle order                               */
void orc_get_grid(const Orc *o, int *point_cell, int *cell_coords, int *point_ids) {
    if (point_cell) memcpy(point_cell, o->point_cell, sizeof(int) * (size_t)o->n);
    if (point_ids) memcpy(point_ids, o->point_ids, sizeof(int) * (size_t)o->n);
    if (cell_coords) for (int i = 0; i < o->n; i++) hg_cell_coords(o, o->x0[i], cell_coords + 3 * (size_t)i);
}
int orc_num_cells(const Orc *o) { return o->ncells; }
void orc_get_cell_ranges(const Orc *o, int *start, int *end) {
    memcpy(start, o->cell_start, sizeof(int) * (size_t)o->ncells);
    memcpy(end, o->cell_end, sizeof(int) * (size_t)o->ncells);
}

/* effective neighbour lists: candidates of the 27-cell walk with q < 2, j != i,
 * ascending j.  counts[n]; flat[] receives at most cap entries; returns total. */
static int cmp_int(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }
long long orc_neighbor_lists(const Orc *o, int *counts, long long *offsets, int *flat, long long cap) {
    int *cand = (int *)malloc(ORC_MAX_CAND * sizeof(int));
    long long total = 0;
    for (int i = 0; i < o->n; i++) {
        int m = hg_candidates(o, o->x0[i], cand, ORC_MAX_CAND), k = 0;
        for (int c = 0; c < m; c++) {
            int j = cand[c];
            if (j == i) continue;
            real q = v3_length(v3_sub(o->x0[i], o->x0[j])) / o->p.h;
            if (q < 2.f) cand[k++] = j;
        }
        qsort(cand, (size_t)k, sizeof(int), cmp_int);
        if (counts) counts[i] = k;
        if (offsets) offsets[i] = total;
        for (int c = 0; c < k; c++) { if (flat && total + c < cap) flat[total + c] = cand[c]; }
        total += k;
    }
    if (offsets) offsets[o->n] = total;
    free(cand);
    return total;
}
/* candidates visited per particle by the 27-cell walk (for the over-visit figure) */
long long orc_candidate_count(const Orc *o) {
    int *cand = (int *)malloc(ORC_MAX_CAND * sizeof(int));
    long long total = 0;
    for (int i = 0; i < o->n; i++) total += hg_candidates(o, o->x0[i], cand, ORC_MAX_CAND);
    free(cand);
    return total;
}

/* stand-alone helpers exported for the known-answer tests */
real orc_W(real x, real y, real z, real h) { return W_kernel(v3_make(x, y, z), h); }
void orc_nabla_W(real x, real y, real z, real h, real *out) { v3 g = nabla_W_kernel(v3_make(x, y, z), h); out[0] = g.x; out[1] = g.y; out[2] = g.z; }
void orc_polar(const real *A, real *R) { m33 a, r; memcpy(&a, A, sizeof a); r = polar_rotation(a); memcpy(R, &r, sizeof r); }
void orc_sigma(const real *F, real mu, real lam, real ratio, real *S) {
    Orc tmp; memset(&tmp, 0, sizeof tmp); tmp.p.stiff_a = 200.f; tmp.p.stiff_b = 199.f;
    m33 f, s; memcpy(&f, F, sizeof f); s = compute_sigma(&tmp, f, mu, lam, ratio); memcpy(S, &s, sizeof s);
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
