// mis_sdf_host.cuh -- host side of the DeepSDF query: owns the packed weights and the activation
// tiles of one network and enqueues the layer chain of mis_sdf.cuh.
#pragma once
#include "mis_sdf.cuh"
#include <stdlib.h>
#include <string>
#include <vector>

struct MisSdf {
    int L = 0;                 // number of Linear layers (9 in deepsdf.py)
    int H = 0;                 // hidden width (network_size = 1024)
    int cap = 0;               // activation capacity in rows (multiple of 128)
    float *W0 = nullptr, *b0 = nullptr;              // first layer, plain [H,3], [H]
    std::vector<float*> Whi, Wlo, bh;                // hidden layers: UMMA tiles [H,H] hi / lo (both inside Wslab, hi then lo per layer), bias [H]
    float* Wslab = nullptr;                          // all hidden-layer weights, contiguous: one L2 access-policy window per layer
    float* bslab = nullptr;                          // hidden-layer biases, contiguous [L-2][H] (bh[l] point into it)
    unsigned* chain_sync = nullptr;                  // device-wide barrier / exit counters of k_sdf_chain_sk
    int chain_mode = 1;                              // 1: few-row queries run all hidden layers in one cooperative launch (MIS_SDF_CHAIN=0: one launch per layer)
    bool l2_pin = false;                             // per-step contact query: keep the weights in the persisting part of L2
    float *wl = nullptr, *bl = nullptr;              // last layer, plain [H], [1]
    float *act[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // ping-pong activations, [buf][hi/lo], cap x H tiled
    float *vals = nullptr;     // 4 x cap scalars: sdf at p, p+eps ex, p+eps ey, p+eps ez
    long long launches = 0;
    long long gemm_launches = 0;
    int num_sms = 148;         // persistent GEMM grid: one CTA per SM
    int force_path = 0;        // tests / tuning: 0 = automatic, 1 = split-K cluster kernel, 2 = persistent big-tile kernel
};

namespace mis {

// kernel launch, optionally as a programmatic dependent launch of the previous kernel in the stream (see pdl_wait / pdl_trigger)
// pin: optional [base, bytes) range this launch keeps in the persisting part of L2 (everything else it touches streams)
struct L2Pin { const void* base = nullptr; size_t bytes = 0; };
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kp(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, L2Pin pin, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    if (pin.base && pin.bytes) {
        at[na].id = cudaLaunchAttributeAccessPolicyWindow;
        at[na].val.accessPolicyWindow.base_ptr = const_cast<void*>(pin.base);
        at[na].val.accessPolicyWindow.num_bytes = pin.bytes;
        at[na].val.accessPolicyWindow.hitRatio = 1.f;
        at[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na++;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    return launch_kp(kern, grid, block, smem, st, pdl, L2Pin{}, args...);
}

inline void sdf_free(MisSdf* s) {
    if (!s) return;
    void* ptrs[] = {s->W0, s->b0, s->wl, s->bl, s->act[0][0], s->act[0][1], s->act[1][0], s->act[1][1], s->vals, s->Wslab, s->bslab, s->chain_sync};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete s;
}

inline cudaError_t sdf_reserve(MisSdf* s, int rows) {
    const int need = (rows + 127) / 128 * 128;
    if (need <= s->cap) return cudaSuccess;
    for (int b = 0; b < 2; b++) for (int h = 0; h < 2; h++) { if (s->act[b][h]) cudaFree(s->act[b][h]); s->act[b][h] = nullptr; }
    if (s->vals) cudaFree(s->vals);
    s->vals = nullptr; s->cap = 0;
    cudaError_t e;
    for (int b = 0; b < 2; b++) for (int h = 0; h < 2; h++) {
        e = cudaMalloc((void**)&s->act[b][h], (size_t)need * s->H * sizeof(float));
        if (e != cudaSuccess) return e;
    }
    e = cudaMalloc((void**)&s->vals, (size_t)4 * need * sizeof(float));
    if (e != cudaSuccess) return e;
    // arrival counters of the one-launch chain: one per (row-block, column tile) + 2
    if (s->chain_sync) cudaFree(s->chain_sync);
    s->chain_sync = nullptr;
    const size_t nsync = 2 + (size_t)(need / 128) * (size_t)((s->H + CH_BN - 1) / CH_BN);
    e = cudaMalloc((void**)&s->chain_sync, nsync * sizeof(unsigned));
    if (e != cudaSuccess) return e;
    e = cudaMemset(s->chain_sync, 0, nsync * sizeof(unsigned));
    if (e != cudaSuccess) return e;
    s->cap = need;
    return cudaSuccess;
}

// One forward pass of the chain for `rows` points (rows <= cap): points -> out[rows].
//   pts/idx: point r is pts[idx ? idx[r] : r]; xf/shift: p_model = A (p - t) + shift; m_count: optional device-side live-row count.
// Returns (through *final_buf) the index of the activation buffer holding the last hidden layer's output when out == nullptr
// (the caller runs its own fused last-layer kernel).
// pdl_first: layer 0 is a programmatic dependent launch of the caller's previous kernel (which must call pdl_trigger / exit).
inline cudaError_t sdf_forward(MisSdf* s, const float* pts, const int* idx, int rows, const int* m_count, const SdfXform& xf, float3 shift,
                               float* out, cudaStream_t st, int fd3 = 0, int* final_buf = nullptr, bool pdl_first = false,
                               cudaEvent_t after_layer0 = nullptr) {
    static bool attr_set_dev[64] = {};                 // the opt-in shared-memory size is a per-DEVICE function attribute
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& attr_set = attr_set_dev[dev_ & 63];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_sdf_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sdf_gemm_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sdf_chain_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        { const char* c = getenv("MIS_SK_CARVEOUT"); if (c && c[0]) cudaFuncSetAttribute(k_sdf_gemm_sk, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c)); }
        attr_set = true;
    }
    const int m_pad = (rows + 127) / 128 * 128;
    const int H = s->H;
    const long long threads0 = (long long)m_pad * (H / 4);
    const long long blocks0 = (threads0 + 255) / 256;
    const long long cap0 = m_count ? 148 * 4 : 148 * 16;            // device-side counts are small: a short grid lets the next kernel start early
    cudaError_t le = launch_k(k_sdf_layer0, dim3((unsigned)(blocks0 < cap0 ? blocks0 : cap0)), dim3(256), 0, st, pdl_first,
                              pts, idx, rows, m_pad, m_count, xf, shift, fd3, (const float*)s->W0, (const float*)s->b0, H, s->act[0][0], s->act[0][1]);
    if (le != cudaSuccess) return le;
    s->launches++;
    if (after_layer0) cudaEventRecord(after_layer0, st);            // the GEMM chain is the next launch on this stream
    // few rows (a device-side count = the per-step contact query, or a small host-side count): split-K over 8-CTA clusters;
    // bulk queries: persistent 128 x 256 tiles
    const bool skinny = s->force_path ? (s->force_path == 1 || s->force_path == 3) : (H <= SK_MAX_KBS * SK_SPLIT * SDF_BK && (m_count != nullptr || rows <= 1024));
    int cur = 0;
    // the cooperative launch takes at most 15 co-resident 8-CTA clusters on a B200 (cudaOccupancyMaxActiveClusters): H <= 15 * 128
    const bool chain = skinny && s->chain_mode && s->force_path != 1 && !s->Whi.empty() && H % CH_BN == 0 && H / CH_BN <= 15;
    if (chain) {
        // every hidden layer in one cooperative launch (all CTAs co-resident: they meet at a device-wide barrier per layer)
        SkChain c;
        c.W = s->Wslab; c.bias = s->bslab; c.n_layers = (int)s->Whi.size(); c.sync = s->chain_sync;
        for (int b = 0; b < 2; b++) for (int h = 0; h < 2; h++) c.act[b][h] = s->act[b][h];
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((H / CH_BN) * SK_SPLIT); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = CH_SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        le = cudaLaunchKernelEx(&cfg, k_sdf_chain_sk, c, H, rows, m_count);
        if (le != cudaSuccess) return le;
        s->launches++; s->gemm_launches++;
        cur = c.n_layers & 1;
    }
    for (size_t l = 0; l < s->Whi.size() && !chain; l++) {
        if (skinny) {
            L2Pin pin;
            if (s->l2_pin && m_count) { pin.base = s->Whi[l]; pin.bytes = 2 * (size_t)H * H * sizeof(float); }
            le = launch_kp(k_sdf_gemm_sk, dim3((H / SK_BN) * SK_SPLIT), dim3(SDF_THREADS), SK_SMEM_BYTES, st, true, pin,
                          (const float*)s->act[cur][0], (const float*)s->act[cur][1], (const float*)s->Whi[l], (const float*)s->Wlo[l], (const float*)s->bh[l],
                          H, H, s->act[cur ^ 1][0], s->act[cur ^ 1][1], rows, m_count);
            if (le != cudaSuccess) return le;
        } else {
            const int tiles = (H / SDF_BN) * (m_pad / SDF_BM);
            k_sdf_gemm<<<tiles < s->num_sms ? tiles : s->num_sms, SDF_THREADS, SDF_SMEM_BYTES, st>>>(
                s->act[cur][0], s->act[cur][1], s->Whi[l], s->Wlo[l], s->bh[l], H, H, s->act[cur ^ 1][0], s->act[cur ^ 1][1], rows, m_count);
        }
        s->launches++; s->gemm_launches++;
        cur ^= 1;
    }
    if (final_buf) *final_buf = cur;
    if (out) {
        const int blocksl = (rows + 7) / 8;
        le = launch_k(k_sdf_last, dim3(blocksl < 148 * 8 ? blocksl : 148 * 8), dim3(256), 0, st, skinny && !chain,
                      (const float*)s->act[cur][0], (const float*)s->act[cur][1], rows, m_count, (const float*)s->wl, (const float*)s->bl, H, out);
        if (le != cudaSuccess) return le;
        s->launches++;
    }
    return cudaGetLastError();
}

// forward differences in model space, returned in the caller's frame: grad = A^T g_model
__global__ void __launch_bounds__(256) k_sdf_fd_grad(const float* __restrict__ vals, int cap, int n, float inv_eps, SdfXform xf, float* __restrict__ sdf_out, float* __restrict__ grad_out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float s0 = vals[r];
    if (sdf_out) sdf_out[r] = s0;
    const float gx = (vals[cap + r] - s0) * inv_eps, gy = (vals[2 * (size_t)cap + r] - s0) * inv_eps, gz = (vals[3 * (size_t)cap + r] - s0) * inv_eps;
    grad_out[3 * (size_t)r + 0] = xf.A[0] * gx + xf.A[3] * gy + xf.A[6] * gz;
    grad_out[3 * (size_t)r + 1] = xf.A[1] * gx + xf.A[4] * gy + xf.A[7] * gz;
    grad_out[3 * (size_t)r + 2] = xf.A[2] * gx + xf.A[5] * gy + xf.A[8] * gz;
}

// ---------------------------------------------------------------- per-step contact (extension of sim.py:238-244)
// broad phase: cell-sorted particles whose model-space position lies in the obstacle's bounding box (+ margin)
// count[0] = candidates appended (capped at `cap` rows; count[3] is raised when the box holds more than the chain can evaluate)
__global__ void __launch_bounds__(256) k_contact_select(const float4* __restrict__ xcur, int n, SdfXform xf, float3 lo, float3 hi, int cap,
                                                        int* __restrict__ idx, int* __restrict__ count, float* __restrict__ pts_out,
                                                        float4* __restrict__ fcon) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < n) {
        fcon[s] = make_float4(0.f, 0.f, 0.f, 0.f);        // particles outside the contact band feel no obstacle force
        p = xcur[s];
        const float px = p.x - xf.t[0], py = p.y - xf.t[1], pz = p.z - xf.t[2];
        const float x = xf.A[0] * px + xf.A[1] * py + xf.A[2] * pz;
        const float y = xf.A[3] * px + xf.A[4] * py + xf.A[5] * pz;
        const float z = xf.A[6] * px + xf.A[7] * py + xf.A[8] * pz;
        in = x >= lo.x && x <= hi.x && y >= lo.y && y <= hi.y && z >= lo.z && z <= hi.z;
    }
    // warp-aggregated append
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int basep = 0;
    if (lane == __ffs(m) - 1) basep = atomicAdd(count, __popc(m));
    basep = __shfl_sync(0xffffffffu, basep, __ffs(m) - 1);
    if (in) {
        const int r = basep + __popc(m & ((1u << lane) - 1));
        if (r < cap) {
            idx[r] = s;
            pts_out[3 * (size_t)r] = p.x; pts_out[3 * (size_t)r + 1] = p.y; pts_out[3 * (size_t)r + 2] = p.z;
        } else {
            count[3] = 1;                                 // overflow: reported by mis_get_contact_count, never silent
        }
    }
}

// last layer of the value pass fused with the narrow phase: one warp per candidate row finishes sdf(p) (Linear(H,1)) and keeps
// the candidates inside the contact band (sdf < range); only they need a normal.  The order of the kept list depends on the
// atomics, the per-particle results do not (rows of the chain are independent).
__global__ void __launch_bounds__(256) k_contact_last_narrow(const float* __restrict__ Xhi, const float* __restrict__ Xlo, int m, const int* __restrict__ count,
                                                             const float* __restrict__ w, const float* __restrict__ b, int H, float range,
                                                             const int* __restrict__ idx, const float* __restrict__ pts,
                                                             int* __restrict__ idx2, float* __restrict__ pts2, float* __restrict__ s0c, int* __restrict__ count2) {
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    const int live = min(*count, m);
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < live; r += gridDim.x * 8) {
        const float v = sdf_last_row(Xhi, Xlo, r, w, H, lane) + b[0];
        if (lane == 0 && v < range) {
            const int q = atomicAdd(count2, 1);
            if (3 * q + 2 >= m) { count2[2] = 1; continue; }            // the forward-difference pass has m rows: overflow flag ([3] of the counter block)
            atomicAdd(count2 + 1, 3);                                   // [1] = rows of the forward-difference pass
            idx2[q] = idx[r];
            pts2[3 * (size_t)q] = pts[3 * (size_t)r]; pts2[3 * (size_t)q + 1] = pts[3 * (size_t)r + 1]; pts2[3 * (size_t)q + 2] = pts[3 * (size_t)r + 2];
            s0c[q] = v;
        }
    }
}

// last layer of the forward-difference pass (rows 3 q + axis) fused with the contact law:
// delta = range - sdf(p_model); f = delta^2 k n, n = grad / |grad| in world space.  One warp per in-contact particle.
__global__ void __launch_bounds__(256) k_contact_last_apply(const float* __restrict__ Xhi, const float* __restrict__ Xlo, int fd_rows, const int* __restrict__ count2,
                                                            const float* __restrict__ w, const float* __restrict__ b, int H,
                                                            const float* __restrict__ s0c, const int* __restrict__ idx2, float inv_eps, SdfXform xf,
                                                            float range, float k_col, float4* __restrict__ fcon) {
    const int lane = threadIdx.x & 31;
    pdl_wait();
    const int live = min(*count2, fd_rows / 3);                         // fd_rows = rows the FD pass could evaluate
    for (int q = blockIdx.x * 8 + (threadIdx.x >> 5); q < live; q += gridDim.x * 8) {
        const float sx = sdf_last_row(Xhi, Xlo, 3 * q, w, H, lane) + b[0];
        const float sy = sdf_last_row(Xhi, Xlo, 3 * q + 1, w, H, lane) + b[0];
        const float sz = sdf_last_row(Xhi, Xlo, 3 * q + 2, w, H, lane) + b[0];
        if (lane == 0) {
            const float s0 = s0c[q];
            const float gx = (sx - s0) * inv_eps, gy = (sy - s0) * inv_eps, gz = (sz - s0) * inv_eps;
            const float wx = xf.A[0] * gx + xf.A[3] * gy + xf.A[6] * gz;
            const float wy = xf.A[1] * gx + xf.A[4] * gy + xf.A[7] * gz;
            const float wz = xf.A[2] * gx + xf.A[5] * gy + xf.A[8] * gz;
            const float nn = sqrtf(wx * wx + wy * wy + wz * wz);
            if (nn > 1e-20f) {
                const float d = range - s0;
                const float f = d * d * k_col / nn;
                fcon[idx2[q]] = make_float4(f * wx, f * wy, f * wz, 0.f);
            }
        }
    }
}

}  // namespace mis
