// mis_ref_host.cuh -- host side of mis_ref.cuh: the precision-templated engine (fp64 forward path of the Taichi prototype,
// reverse pass of the rollout in fp32 / fp64).  Included by mis_api.cu after `struct MisSim`.
#pragma once
#include "mis_ref.cuh"

#include <vector>

namespace mis {

enum MisF64What {            // mis_set_f64 / mis_get_f64 (include/mis.h)
    F64_X0 = 0, F64_MASS = 1, F64_YOUNGS = 2, F64_POISSON = 3, F64_DESIGN = 4, F64_EXT_FORCE = 5, F64_DIRICHLET = 6,
    F64_POSITION = 7, F64_VELOCITY = 8, F64_ELASTIC_FORCE = 9, F64_VOLUME = 10, F64_RHO = 11, F64_DEF_GRAD = 12, F64_STRESS = 13,
    F64_ROTATION = 14, F64_A_PQ = 15
};

template <typename T> struct RefEngine {
    ref::RS_<T> s{};
    ref::RP<T> c{};
    T *E = nullptr, *nu = nullptr, *design = nullptr;                 // raw inputs, slot order
    T *xb = nullptr, *vb = nullptr, *felb = nullptr, *qbuf = nullptr, *Rb = nullptr, *Fb = nullptr, *Sb = nullptr, *ratio_b = nullptr;
    T *seg = nullptr;  size_t seg_frames = 0;                          // (x, v) of every frame of one segment
    T *ck = nullptr;   size_t ck_count = 0;                            // (x, v) at the checkpoints
    double* loss_dev = nullptr;
    bool statics_dirty = true;        // mu / lam / ratio / volume must be recomputed from E, nu, design, m
    bool primed = false;              // fel holds the elastic force at x (Verlet)
    long long frame = 0;
    std::vector<void*> owned;
    template <typename U> cudaError_t alloc(U** p, size_t count) {
        cudaError_t e = cudaMalloc((void**)p, count * sizeof(U) + 64);
        if (e == cudaSuccess) { owned.push_back((void*)*p); e = cudaMemset(*p, 0, count * sizeof(U) + 64); }
        return e;
    }
    void release() { for (void* p : owned) cudaFree(p); owned.clear(); }
};

template <typename T> static cudaError_t ref_create(RefEngine<T>& g, int n, const unsigned long long* nbr_start, const uint32_t* nbr, const MisParams& p) {
    cudaError_t e = cudaSuccess;
    const size_t N = (size_t)n;
    g.s.n = n; g.s.nbr_start = nbr_start; g.s.nbr = nbr;
#define RA(ptr, cnt) if (e == cudaSuccess) e = g.alloc(&(ptr), (cnt))
    RA(g.s.x0, 3 * N); RA(g.s.m, N); RA(g.s.vol, N); RA(g.s.rho, N); RA(g.s.mu, N); RA(g.s.lam, N); RA(g.s.ratio, N);
    RA(g.s.fext, 3 * N); RA(g.s.freem, 3 * N); RA(g.s.x, 3 * N); RA(g.s.xn, 3 * N); RA(g.s.v, 3 * N); RA(g.s.fel, 3 * N); RA(g.s.feln, 3 * N);
    RA(g.s.A, 9 * N); RA(g.s.R, 9 * N); RA(g.s.F, 9 * N); RA(g.s.S, 9 * N); RA(g.s.B, 9 * N);
    RA(g.E, N); RA(g.nu, N); RA(g.design, N);
#undef RA
    g.c.h = (T)p.h; g.c.dt = (T)p.dt; g.c.damping = (T)p.damping; g.c.k_col = (T)p.k_col; g.c.col_range = (T)p.col_range;
    g.c.stiff_a = (T)p.stiff_a; g.c.stiff_b = (T)p.stiff_b; g.c.tanh_k = (T)p.tanh_k;
    g.c.identity_rot = p.identity_rot; g.c.euler = p.euler; g.c.no_contact = p.no_contact; g.c.symmetric_pair = p.symmetric_pair;
    g.c.self_density = p.self_density;
    return e;
}

static inline int rblk(long long n, int t) { return (int)((n + t - 1) / t); }

// mu, lam, ratio, rho, V from the raw inputs (sim.py:288-308, 107-110, 154-167)
template <typename T> static void ref_statics(RefEngine<T>& g, cudaStream_t st) {
    if (!g.statics_dirty) return;
    const int n = g.s.n;
    ref::kr_material<T><<<rblk(n, 256), 256, 0, st>>>(n, g.E, g.nu, g.design, g.c.tanh_k, g.s.mu, g.s.lam, g.s.ratio);
    ref::kr_volume<T><<<rblk((long long)n * ref::RG, ref::RTHREADS), ref::RTHREADS, 0, st>>>(g.s, g.c);
    g.statics_dirty = false;
    g.primed = false;
}

// A, R, B, F, S at positions x, then the elastic force into fel_out
template <typename T> static void ref_eval(RefEngine<T>& g, const T* x, T* fel_out, cudaStream_t st) {
    const int blocks = rblk((long long)g.s.n * ref::RG, ref::RTHREADS);
    ref::kr_Apq<T><<<blocks, ref::RTHREADS, 0, st>>>(g.s, g.c, x);
    ref::kr_nabla_u<T><<<blocks, ref::RTHREADS, 0, st>>>(g.s, g.c, x);
    if (fel_out) ref::kr_force<T><<<blocks, ref::RTHREADS, 0, st>>>(g.s, g.c, fel_out);
}

// one step of the loop body (sim.py:352-358), or of forward() (sim_taichi.py:174-182) when c.euler
template <typename T> static void ref_step(RefEngine<T>& g, cudaStream_t st) {
    const int n = g.s.n;
    ref_statics(g, st);
    if (g.c.euler) {
        ref_eval(g, g.s.x, g.s.fel, st);
        ref::kr_euler<T><<<rblk(n, 256), 256, 0, st>>>(g.s, g.c);
        T* t = g.s.x; g.s.x = g.s.xn; g.s.xn = t;
    } else {
        if (!g.primed) { ref_eval(g, g.s.x, g.s.fel, st); g.primed = true; }       // frame-0 evaluation, sim.py:349-351
        ref::kr_part1<T><<<rblk(n, 256), 256, 0, st>>>(g.s, g.c);
        ref_eval(g, g.s.xn, g.s.feln, st);
        ref::kr_part2<T><<<rblk(n, 256), 256, 0, st>>>(g.s, g.c);
        T* t = g.s.x; g.s.x = g.s.xn; g.s.xn = t;
        t = g.s.fel; g.s.fel = g.s.feln; g.s.feln = t;
    }
    g.frame++;
}

template <typename T> static void ref_startup(RefEngine<T>& g, const double v0[3], cudaStream_t st) {
    const size_t N3 = 3 * (size_t)g.s.n;
    cudaMemcpyAsync(g.s.x, g.s.x0, N3 * sizeof(T), cudaMemcpyDeviceToDevice, st);
    ref::kr_fill3<T><<<rblk(g.s.n, 256), 256, 0, st>>>(g.s.n, (T)v0[0], (T)v0[1], (T)v0[2], g.s.v);
    g.primed = false;
    g.frame = 0;
}

// Reverse pass of the rollout (sim.py:341-372 with compute_grad): startup, `frames` steps, compute_loss against target t at frame
// (frames / n_targets) (t + 1), then the adjoint sweep.  Velocity-Verlet only.  ratio_b accumulates dL/d ratio.
template <typename T>
static cudaError_t ref_rollout_grad(RefEngine<T>& g, const double v0[3], int frames, int n_targets, const float* tx, const float* tv,
                                    const uint32_t* perm, int K, double* loss_host, cudaStream_t st) {
    const int n = g.s.n;
    const size_t N = (size_t)n, N3 = 3 * N;
    cudaError_t e = cudaSuccess;
#define RA(ptr, cnt) if (e == cudaSuccess && !(ptr)) e = g.alloc(&(ptr), (cnt))
    RA(g.xb, N3); RA(g.vb, N3); RA(g.felb, N3); RA(g.qbuf, N3); RA(g.Rb, 9 * N); RA(g.Fb, 9 * N); RA(g.Sb, 9 * N); RA(g.ratio_b, N); RA(g.loss_dev, 1);
#undef RA
    if (e != cudaSuccess) return e;
    if (K < 1) K = 1;
    const int n_seg = (frames + K - 1) / K;
    if (g.seg_frames < (size_t)K + 1) { T* p = nullptr; e = g.alloc(&p, (size_t)(K + 1) * 2 * N3); if (e != cudaSuccess) return e; g.seg = p; g.seg_frames = (size_t)K + 1; }
    if (g.ck_count < (size_t)n_seg) { T* p = nullptr; e = g.alloc(&p, (size_t)n_seg * 2 * N3); if (e != cudaSuccess) return e; g.ck = p; g.ck_count = (size_t)n_seg; }
    const int every = n_targets > 0 ? frames / n_targets : 0;
    auto target_of = [&](int f) { return (every > 0 && f > 0 && f % every == 0 && f / every <= n_targets) ? f / every - 1 : -1; };
    const int b256 = rblk(n, 256), bg = rblk((long long)n * ref::RG, ref::RTHREADS);
    cudaMemsetAsync(g.loss_dev, 0, sizeof(double), st);
    cudaMemsetAsync(g.ratio_b, 0, N * sizeof(T), st);
    cudaMemsetAsync(g.xb, 0, N3 * sizeof(T), st); cudaMemsetAsync(g.vb, 0, N3 * sizeof(T), st); cudaMemsetAsync(g.felb, 0, N3 * sizeof(T), st);
    // ---- forward: checkpoints + loss value
    ref_startup(g, v0, st);
    ref_statics(g, st);
    for (int f = 0; f < frames; f++) {
        if (f % K == 0) {
            T* c = g.ck + (size_t)(f / K) * 2 * N3;
            cudaMemcpyAsync(c, g.s.x, N3 * sizeof(T), cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(c + N3, g.s.v, N3 * sizeof(T), cudaMemcpyDeviceToDevice, st);
        }
        ref_step(g, st);
        const int t = target_of(f + 1);
        if (t >= 0) ref::kr_loss<T><<<148, 256, 0, st>>>(g.s, g.c, g.s.x, g.s.v, tx + (size_t)t * N3, tv + (size_t)t * N3, perm, (T*)nullptr, (T*)nullptr, g.loss_dev);
    }
    // ---- backward, one segment at a time: recompute its frames from the checkpoint, then sweep them in reverse
    for (int sg = n_seg - 1; sg >= 0; sg--) {
        const int f0 = sg * K, f1 = (f0 + K < frames) ? f0 + K : frames;          // frames f0 .. f1 of this segment
        const T* c = g.ck + (size_t)sg * 2 * N3;
        cudaMemcpyAsync(g.s.x, c, N3 * sizeof(T), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(g.s.v, c + N3, N3 * sizeof(T), cudaMemcpyDeviceToDevice, st);
        g.primed = false;
        for (int f = f0; f <= f1; f++) {
            T* slot = g.seg + (size_t)(f - f0) * 2 * N3;
            cudaMemcpyAsync(slot, g.s.x, N3 * sizeof(T), cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(slot + N3, g.s.v, N3 * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (f < f1) ref_step(g, st);
        }
        for (int f = f1 - 1; f >= f0; f--) {
            // (xb, vb) = adjoint of frame f + 1 without its own loss term / elastic term; felb = partial adjoint of fel_{f+1}
            const T* xn = g.seg + (size_t)(f + 1 - f0) * 2 * N3;
            const T* vn = xn + N3;
            const T* xp = g.seg + (size_t)(f - f0) * 2 * N3;
            const int t = target_of(f + 1);
            if (t >= 0) ref::kr_loss<T><<<148, 256, 0, st>>>(g.s, g.c, xn, vn, tx + (size_t)t * N3, tv + (size_t)t * N3, perm, g.xb, g.vb, (double*)nullptr);
            ref::kr_adj_stepA<T><<<b256, 256, 0, st>>>(g.s, g.c, xn, g.xb, g.vb, g.felb, g.qbuf);
            // elastic adjoint at frame f + 1: fields at x_{f+1}, then Rb/Fb/Sb -> Ab/Bb (+ design gradient) -> positions
            ref_eval(g, xn, (T*)nullptr, st);
            ref::kr_adj_force<T><<<bg, ref::RTHREADS, 0, st>>>(g.s, g.c, g.felb, g.Rb, g.Fb, g.Sb);
            ref::kr_adj_particle<T><<<rblk(n, 128), 128, 0, st>>>(g.s, g.c, g.Rb, g.Fb, g.Sb, g.ratio_b);
            ref::kr_adj_pos<T><<<bg, ref::RTHREADS, 0, st>>>(g.s, g.c, g.Rb, g.Fb, g.xb);
            ref::kr_adj_stepB<T><<<b256, 256, 0, st>>>(g.s, g.c, xp, g.xb, g.vb, g.felb, g.qbuf);
        }
    }
    // frame 0: fel_0 enters force_1 of step 0 only; its design gradient (zero at the rest state, kept for generality)
    {
        const T* x0f = g.seg;                                       // frame 0 of segment 0
        ref_eval(g, x0f, (T*)nullptr, st);
        ref::kr_adj_force<T><<<bg, ref::RTHREADS, 0, st>>>(g.s, g.c, g.felb, g.Rb, g.Fb, g.Sb);
        ref::kr_adj_particle<T><<<rblk(n, 128), 128, 0, st>>>(g.s, g.c, g.Rb, g.Fb, g.Sb, g.ratio_b);
    }
    g.primed = false;
    e = cudaMemcpyAsync(loss_host, g.loss_dev, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e;
}

}  // namespace mis
