// mis_sdf_host.cuh -- host side of the DeepSDF query: owns the packed weights and the activation
// tiles of one network and enqueues the layer chain of mis_sdf.cuh.
#pragma once
#include "mis_sdf.cuh"
#include <string>
#include <vector>

struct MisSdf {
    int L = 0;                 // number of Linear layers (9 in deepsdf.py)
    int H = 0;                 // hidden width (network_size = 1024)
    int cap = 0;               // activation capacity in rows (multiple of 128)
    float *W0 = nullptr, *b0 = nullptr;              // first layer, plain [H,3], [H]
    std::vector<float*> Whi, Wlo, bh;                // hidden layers: UMMA tiles [H,H] hi / lo, bias [H]
    float *wl = nullptr, *bl = nullptr;              // last layer, plain [H], [1]
    float *act[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // ping-pong activations, [buf][hi/lo], cap x H tiled
    float *vals = nullptr;     // 4 x cap scalars: sdf at p, p+eps ex, p+eps ey, p+eps ez
    long long launches = 0;
    long long gemm_launches = 0;
    int num_sms = 148;         // persistent GEMM grid: one CTA per SM
};

namespace mis {

inline void sdf_free(MisSdf* s) {
    if (!s) return;
    void* ptrs[] = {s->W0, s->b0, s->wl, s->bl, s->act[0][0], s->act[0][1], s->act[1][0], s->act[1][1], s->vals};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (float* p : s->Whi) if (p) cudaFree(p);
    for (float* p : s->Wlo) if (p) cudaFree(p);
    for (float* p : s->bh) if (p) cudaFree(p);
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
    s->cap = need;
    return cudaSuccess;
}

// One forward pass of the chain for `rows` points (rows <= cap): points -> out[rows].
//   pts/idx: point r is pts[idx ? idx[r] : r]; xf/shift: p_model = A (p - t) + shift; m_count: optional device-side live-row count.
inline cudaError_t sdf_forward(MisSdf* s, const float* pts, const int* idx, int rows, const int* m_count, const SdfXform& xf, float3 shift,
                               float* out, cudaStream_t st, int fd3 = 0) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_sdf_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int m_pad = (rows + 127) / 128 * 128;
    const int H = s->H;
    const long long threads0 = (long long)m_pad * (H / 4);
    const long long blocks0 = (threads0 + 255) / 256;
    k_sdf_layer0<<<(unsigned)(blocks0 < 148 * 16 ? blocks0 : 148 * 16), 256, 0, st>>>(pts, idx, rows, m_pad, m_count, xf, shift, fd3, s->W0, s->b0, H, s->act[0][0], s->act[0][1]);
    s->launches++;
    int cur = 0;
    for (size_t l = 0; l < s->Whi.size(); l++) {
        const int tiles = (H / SDF_BN) * (m_pad / SDF_BM);
        k_sdf_gemm<<<tiles < s->num_sms ? tiles : s->num_sms, SDF_THREADS, SDF_SMEM_BYTES, st>>>(
            s->act[cur][0], s->act[cur][1], s->Whi[l], s->Wlo[l], s->bh[l], H, H, s->act[cur ^ 1][0], s->act[cur ^ 1][1], rows, m_count);
        s->launches++; s->gemm_launches++;
        cur ^= 1;
    }
    const int blocksl = (rows + 7) / 8;
    k_sdf_last<<<blocksl < 148 * 8 ? blocksl : 148 * 8, 256, 0, st>>>(s->act[cur][0], s->act[cur][1], rows, m_count, s->wl, s->bl, H, out);
    s->launches++;
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
__global__ void __launch_bounds__(256) k_contact_select(const float4* __restrict__ xcur, int n, SdfXform xf, float3 lo, float3 hi,
                                                        int* __restrict__ idx, int* __restrict__ count, float* __restrict__ pts_out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < n) {
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
        idx[r] = s;
        pts_out[3 * (size_t)r] = p.x; pts_out[3 * (size_t)r + 1] = p.y; pts_out[3 * (size_t)r + 2] = p.z;
    }
}

// narrow phase: keep the candidates whose value is inside the contact band (sdf < range); only they need a normal
__global__ void __launch_bounds__(256) k_contact_narrow(const float* __restrict__ vals, const int* __restrict__ idx, const float* __restrict__ pts,
                                                        const int* __restrict__ count, float range,
                                                        int* __restrict__ idx2, float* __restrict__ pts2, float* __restrict__ s0c, int* __restrict__ count2) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = r < *count && vals[r] < range;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int basep = 0;
    if (lane == __ffs(m) - 1) { basep = atomicAdd(count2, __popc(m)); atomicAdd(count2 + 1, 3 * __popc(m)); }   // [1] = rows of the FD pass
    basep = __shfl_sync(0xffffffffu, basep, __ffs(m) - 1);
    if (in) {
        const int q = basep + __popc(m & ((1u << lane) - 1));
        idx2[q] = idx[r];
        pts2[3 * (size_t)q] = pts[3 * (size_t)r]; pts2[3 * (size_t)q + 1] = pts[3 * (size_t)r + 1]; pts2[3 * (size_t)q + 2] = pts[3 * (size_t)r + 2];
        s0c[q] = vals[r];
    }
}

// contact law: delta = range - sdf(p_model); f = delta^2 k n, n = grad / |grad| in world space.
// fd holds the three forward-difference evaluations of the in-contact particles, interleaved (3 r + axis).
__global__ void __launch_bounds__(256) k_contact_apply(const float* __restrict__ s0c, const float* __restrict__ fd, const int* __restrict__ idx2,
                                                       const int* __restrict__ count2, int fd_rows, float inv_eps, SdfXform xf, float range, float k_col,
                                                       float4* __restrict__ fcon) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *count2 || 3 * r + 2 >= fd_rows) return;      // fd_rows = rows the FD pass could evaluate
    const float s0 = s0c[r];
    const float gx = (fd[3 * (size_t)r] - s0) * inv_eps, gy = (fd[3 * (size_t)r + 1] - s0) * inv_eps, gz = (fd[3 * (size_t)r + 2] - s0) * inv_eps;
    float wx = xf.A[0] * gx + xf.A[3] * gy + xf.A[6] * gz;
    float wy = xf.A[1] * gx + xf.A[4] * gy + xf.A[7] * gz;
    float wz = xf.A[2] * gx + xf.A[5] * gy + xf.A[8] * gz;
    const float nn = sqrtf(wx * wx + wy * wy + wz * wz);
    if (!(nn > 1e-20f)) return;
    const float d = range - s0;
    const float f = d * d * k_col / nn;
    fcon[idx2[r]] = make_float4(f * wx, f * wy, f * wz, 0.f);
}

}  // namespace mis
