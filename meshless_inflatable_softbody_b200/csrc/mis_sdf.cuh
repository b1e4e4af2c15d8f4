// mis_sdf.cuh -- DeepSDF MLP (deepsdf.py:9-41, DeepSDFWithCode) as a tcgen05 tensor-core GEMM chain.
//
// The network is Linear(3,H)+ReLU, (L-2) x [Linear(H,H)+ReLU], Linear(H,1) with weight_norm on every
// Linear (W = g * v / ||v|| per output row, deepsdf.py:3,13-37) and Dropout(0.0) = identity.  The
// reference evaluates it in fp32 (cuBLAS SGEMM, sim.py:100); contact and the design field need
// fp32-class accuracy, so the hidden layers run as 3xTF32 on the 5th-generation tensor cores:
//     x = x_hi + x_lo,  x_hi = x with the low 13 mantissa bits cleared (exactly a TF32 number),
//     X W^T ~= X_hi W_hi^T + X_hi W_lo^T + X_lo W_hi^T        (error ~2^-21 relative, fp32 accumulate in TMEM)
//
// Data layout in HBM ("UMMA tiles"): every GEMM operand is stored as the exact shared-memory image the
// MMA reads, so a tile moves with ONE 1-D TMA bulk copy (cp.async.bulk) and needs no tensor map:
//   tile (rb, kb) = rows [128 rb, 128 rb+128) x k [32 kb, 32 kb+32) of fp32 = 16 KB at ((rb*KB + kb) * 4096) floats
//   inside a tile: off(r, k) = (r/8)*256 + (k/4)*32 + (r%8)*4 + (k%4) floats
//   = the canonical K-major no-swizzle UMMA layout: 8x16-byte core matrices, LBO (K step) 128 B, SBO (8-row step) 1024 B.
// Activations are produced in this layout (hi and lo planes) by the previous layer's epilogue.
//
// Kernel k_sdf_gemm (one 128 x 256 output tile per CTA, 192 threads):
//   warp 0   : TMA producer  -- per 32-wide k-block: A_hi, A_lo (16 KB each), B_hi, B_lo (2 x 16 KB each) into a 2-stage ring
//   warp 1   : TMEM allocation + MMA issuer -- 4 k-steps x 3 tcgen05.mma.kind::tf32 (128x256x8) per k-block, fp32 accumulators
//              in 256 TMEM columns; tcgen05.commit releases the smem stage / signals the epilogue
//   warps 2-5: epilogue -- tcgen05.ld 32x32b, + bias, ReLU, hi/lo split, 16-byte stores straight into the next layer's tiles
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mis {

constexpr int SDF_BM = 128, SDF_BN = 256, SDF_BK = 32;
constexpr int SDF_TILE_FLOATS = 128 * 32;                    // one 128 x 32 fp32 tile
constexpr int SDF_TILE_BYTES = SDF_TILE_FLOATS * 4;          // 16 KB
constexpr int SDF_STAGES = 2;
constexpr int SDF_STAGE_BYTES = 6 * SDF_TILE_BYTES;          // A_hi, A_lo, B_hi[2], B_lo[2]
constexpr int SDF_SMEM_BYTES = SDF_STAGES * SDF_STAGE_BYTES + 256 /*barriers*/ + 1024 /*align*/;
constexpr int SDF_THREADS = 192;

__host__ __device__ __forceinline__ size_t sdf_tile_off(int r, int k, int KB) {
    // float offset of element (r, k) in a tiled matrix with KB k-blocks per row-block
    return ((size_t)(r >> 7) * KB + (k >> 5)) * SDF_TILE_FLOATS + ((r & 127) >> 3) * 256 + ((k & 31) >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may start while its
// predecessor in the stream still runs; pdl_wait() blocks until the predecessor has completed and its writes are visible (a no-op
// for a normally launched kernel), pdl_trigger() lets the successor's blocks be scheduled early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, no swizzle: LBO = 128 B between core matrices along K, SBO = 1024 B between 8-row groups, descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t SDF_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(SDF_BN >> 3) << 17) | ((uint32_t)(SDF_BM >> 4) << 24);

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------- the hidden layer
// Yhi/Ylo[M, N] = split( relu( (Xhi + Xlo)[M, K] (Whi + Wlo)[N, K]^T + bias ) ), all matrices in UMMA tiles.
// Persistent: gridDim.x CTAs (one per SM) walk the 128 x 256 output tiles t = blockIdx.x, + gridDim.x, ... with
// t -> (row-block t / NT, column-block t % NT), so the CTAs running together share A row-blocks and W in L2.  The number
// of live rows may sit on the device (m_count): no host synchronisation is needed to size the work.
// The accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of tile i overlaps the MMAs of tile i+1.
__global__ void __launch_bounds__(SDF_THREADS, 1) k_sdf_gemm(const float* __restrict__ Xhi, const float* __restrict__ Xlo,
                                                             const float* __restrict__ Whi, const float* __restrict__ Wlo,
                                                             const float* __restrict__ bias, int K, int N,
                                                             float* __restrict__ Yhi, float* __restrict__ Ylo,
                                                             int m_rows, const int* __restrict__ m_count /* device-side live rows, may be null */) {
    extern __shared__ uint8_t smem_raw[];
    const int live = m_count ? min(*m_count, m_rows) : m_rows;
    const int row_blocks = (live + SDF_BM - 1) / SDF_BM;
    // few live rows (per-step contact queries): 64-wide column tiles spread one row-block over 16 CTAs instead of 4, which
    // cuts the latency of a layer ~4x; many rows: 256-wide tiles (best operand reuse).  Uniform across the grid.
    const int bn = (row_blocks * (N / SDF_BN) * 2 <= (int)gridDim.x) ? 64 : SDF_BN;
    const int NT = N / bn;
    const int total = row_blocks * NT;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(SDF_BM >> 4) << 24);
    const uint32_t b_bytes = (uint32_t)bn * SDF_BK * 4;              // one B plane (hi or lo) of this tile
    const uint32_t stage_bytes = 2 * SDF_TILE_BYTES + 2 * b_bytes;   // A_hi, A_lo, B_hi, B_lo
    const uint32_t stages = (bn == SDF_BN) ? 2u : 4u;                // 2 x 96 KB or 4 x 48 KB: deeper ring when the MMAs are short
    const uint32_t off_bhi = 2 * SDF_TILE_BYTES, off_blo = 2 * SDF_TILE_BYTES + b_bytes;
    if ((int)blockIdx.x >= total) return;                    // uniform per CTA: nothing allocated yet
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = K / SDF_BK;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + SDF_STAGES * SDF_STAGE_BYTES;
    // full[s] +8 s | empty[s] +32 + 8 s (s < 4) | tmem_full[b] +64 + 8 b | tmem_empty[b] +80 + 8 b | tmem pointer +96
    const uint32_t BAR_EMPTY = bars + 32, BAR_TFULL = bars + 64, BAR_TEMPTY = bars + 80, TMEM_SLOT = bars + 96;
    volatile uint32_t* tmem_ptr_sm = reinterpret_cast<volatile uint32_t*>(smem_raw + (TMEM_SLOT - smem_u32(smem_raw)));

    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; s++) { mbar_init(bars + 8 * s, 1); mbar_init(BAR_EMPTY + 8 * s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(BAR_TFULL + 8 * b, 1); mbar_init(BAR_TEMPTY + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(TMEM_SLOT), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_ptr_sm;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int mb = t / NT, nb = t - mb * NT;
                const float* a_hi = Xhi + (size_t)mb * KB * SDF_TILE_FLOATS;
                const float* a_lo = Xlo + (size_t)mb * KB * SDF_TILE_FLOATS;
                // W rows [bn nb, bn nb + bn): two whole 128-row tiles (bn = 256) or half of one (bn = 64: 8 row-groups = 8 KB contiguous)
                const int n0 = nb * bn;
                const size_t wrow = (size_t)(n0 >> 7) * KB * SDF_TILE_FLOATS + (size_t)((n0 & 127) >> 3) * 256;
                const float* b_hi0 = Whi + wrow;
                const float* b_lo0 = Wlo + wrow;
                const float* b_hi1 = b_hi0 + (size_t)KB * SDF_TILE_FLOATS;
                const float* b_lo1 = b_lo0 + (size_t)KB * SDF_TILE_FLOATS;
                for (int kb = 0; kb < KB; kb++, it++) {
                    const uint32_t s = it % stages, ph = (it / stages) & 1;
                    mbar_wait(BAR_EMPTY + 8 * s, ph ^ 1);               // slot free (first pass returns at once)
                    const uint32_t full = bars + 8 * s;
                    mbar_expect_tx(full, stage_bytes);
                    const uint32_t st = base + s * stage_bytes;
                    const size_t o = (size_t)kb * SDF_TILE_FLOATS;
                    tma_bulk_g2s(st + 0 * SDF_TILE_BYTES, a_hi + o, SDF_TILE_BYTES, full);
                    tma_bulk_g2s(st + 1 * SDF_TILE_BYTES, a_lo + o, SDF_TILE_BYTES, full);
                    if (bn == SDF_BN) {
                        tma_bulk_g2s(st + off_bhi, b_hi0 + o, SDF_TILE_BYTES, full);
                        tma_bulk_g2s(st + off_bhi + SDF_TILE_BYTES, b_hi1 + o, SDF_TILE_BYTES, full);
                        tma_bulk_g2s(st + off_blo, b_lo0 + o, SDF_TILE_BYTES, full);
                        tma_bulk_g2s(st + off_blo + SDF_TILE_BYTES, b_lo1 + o, SDF_TILE_BYTES, full);
                    } else {
                        tma_bulk_g2s(st + off_bhi, b_hi0 + o, b_bytes, full);
                        tma_bulk_g2s(st + off_blo, b_lo0 + o, b_bytes, full);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, lt = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, lt++) {
                const uint32_t b = lt & 1;
                mbar_wait(BAR_TEMPTY + 8 * b, ((lt >> 1) & 1) ^ 1);      // epilogue has drained this accumulator buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_acc + b * SDF_BN;
                for (int kb = 0; kb < KB; kb++, it++) {
                    const uint32_t s = it % stages, ph = (it / stages) & 1;
                    mbar_wait(bars + 8 * s, ph);                         // operands landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * stage_bytes;
#pragma unroll
                    for (int ks = 0; ks < SDF_BK / 8; ks++) {
                        const uint32_t ko = ks * 256;                    // 8 floats along K = 2 core matrices = 256 B
                        const uint64_t ahi = umma_desc(st + 0 * SDF_TILE_BYTES + ko), alo = umma_desc(st + 1 * SDF_TILE_BYTES + ko);
                        const uint64_t bhi = umma_desc(st + off_bhi + ko), blo = umma_desc(st + off_blo + ko);
                        umma_tf32(acc, alo, bhi, idesc, (kb | ks) ? 1u : 0u);
                        umma_tf32(acc, ahi, blo, idesc, 1u);
                        umma_tf32(acc, ahi, bhi, idesc, 1u);
                    }
                    umma_commit(BAR_EMPTY + 8 * s);                      // frees the smem stage when these MMAs retire
                }
                umma_commit(BAR_TFULL + 8 * b);                          // accumulator complete
            }
        }
    } else {
        // epilogue: warp w may touch TMEM lanes [32 (w % 4), 32 (w % 4) + 32)
        const int q = warp & 3;
        const int rl = 32 * q + lane;                                 // row inside the 128-row block
        const int KBn = N / SDF_BK;                                   // k-blocks of the NEXT layer's A operand
        const size_t row_off = (size_t)(rl >> 3) * 256 + (rl & 7) * 4;
        uint32_t lt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, lt++) {
            const int mb = t / NT, nb = t - mb * NT;
            const uint32_t b = lt & 1;
            mbar_wait(BAR_TFULL + 8 * b, (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < bn; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(32 * q) << 16) + b * SDF_BN + (uint32_t)c, v);
                const size_t tile = ((size_t)mb * KBn + (size_t)(nb * bn + c) / SDF_BK) * SDF_TILE_FLOATS + row_off;
                float4* dhi = reinterpret_cast<float4*>(Yhi + tile);
                float4* dlo = reinterpret_cast<float4*>(Ylo + tile);
                const float4* bp = reinterpret_cast<const float4*>(bias + nb * bn + c);
#pragma unroll
                for (int kc = 0; kc < 8; kc++) {
                    const float4 bb = __ldg(bp + kc);                 // same address in every lane: one broadcast load
                    const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
                    float h[4], l[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float y = fmaxf(__uint_as_float(v[4 * kc + e]) + bv[e], 0.f);
                        h[e] = tf32_hi(y);
                        l[e] = y - h[e];
                    }
                    dhi[kc * 8] = make_float4(h[0], h[1], h[2], h[3]);    // next core matrix along K: +32 floats = 8 float4
                    dlo[kc * 8] = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_TEMPTY + 8 * b);            // this warp is done reading the buffer
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------- the hidden layer for FEW rows (per-step contact queries)
// With a few hundred live rows the layer is bound by how fast the 8 MB of W (hi + lo planes) and the activations reach the SMs and by
// fixed latencies, not by the tensor pipe: one 128 x 256 tile per CTA would leave 144 SMs idle and make 4 SMs ingest 3 MB each.
// k_sdf_gemm_sk spreads one layer over (N / 64) clusters x 8 CTAs: cluster c owns output columns [64 c, 64 c + 64); CTA rank r of the
// cluster owns the K slice [K r / 8, K (r + 1) / 8) (split-K), i.e. 64 KB of W and 128 KB of activations per 128-row block.  The 8
// partial accumulators (TMEM -> shared memory) are reduced through distributed shared memory in FIXED rank order (deterministic): rank r
// finishes columns [8 r, 8 r + 8) of the tile (+ bias, ReLU, hi/lo split, stores into the next layer's UMMA tiles).
// Shared memory is kept under half an SM (2 x 48 KB stages; the fp32 partial tile aliases stage 0) and only 64 TMEM columns are
// allocated, so two CTAs fit on an SM: at most 15 clusters of 8 one-CTA-per-SM blocks are co-resident on a B200 (measured), and the
// 16 column tiles of a 1024-wide layer must not fall into two waves.  Row-blocks are processed one after the other (any row count is
// correct; the bulk kernel above is the efficient one for many rows).
constexpr int SK_BN = 64, SK_SPLIT = 8, SK_STAGES = 2;
constexpr int SK_W_HALF_BYTES = SK_BN * SDF_BK * 4;                        // 64 rows x 32 k of one plane = 8 KB
constexpr int SK_MAX_KBS = 64;                                             // k-blocks per rank (W is streamed with A, so only a sanity bound)
constexpr int SK_STAGE_BYTES = 2 * SDF_TILE_BYTES + 2 * SK_W_HALF_BYTES;   // A_hi, A_lo, W_hi, W_lo of one k-block: 48 KB
constexpr int SK_STAGING_BYTES = SDF_BM * SK_BN * 4;                       // 32 KB fp32 partial tile (aliases the A part of stage 0)
constexpr int SK_TMEM_COLS = 64;
constexpr int SK_SMEM_BYTES = SK_STAGES * SK_STAGE_BYTES + 256 + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t cluster_addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_smem_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

#ifdef MIS_SK_TIMING
__device__ unsigned long long sk_dbg[64];
__device__ unsigned long long sk_cta[4 * 256];
__device__ __forceinline__ unsigned long long sk_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SK_T(slot) do { if (blockIdx.x == MIS_SK_TIMING) sk_dbg[slot] = sk_now(); } while (0)
#define SK_CTA(k) do { if (threadIdx.x == 0) { sk_cta[4 * blockIdx.x + (k)] = sk_now(); if ((k) == 0) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); sk_cta[4 * blockIdx.x + 2] = sm_; } } } while (0)
#else
#define SK_T(slot) do { } while (0)
#define SK_CTA(k) do { } while (0)
#endif

__global__ void __cluster_dims__(SK_SPLIT, 1, 1) __launch_bounds__(SDF_THREADS, 2)
k_sdf_gemm_sk(const float* __restrict__ Xhi, const float* __restrict__ Xlo, const float* __restrict__ Whi, const float* __restrict__ Wlo,
              const float* __restrict__ bias, int K, int N, float* __restrict__ Yhi, float* __restrict__ Ylo,
              int m_rows, const int* __restrict__ m_count) {
    extern __shared__ uint8_t smem_raw[];
    if (threadIdx.x == 0) SK_T(0);
    SK_CTA(0);
    pdl_trigger();                                             // the next layer may set itself up while this one runs
    const int nb = blockIdx.x / SK_SPLIT;                      // cluster id = column tile
    if (nb >= N / SK_BN) return;                               // uniform per cluster
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = K / SDF_BK;
    const int kbs = KB / SK_SPLIT;                             // k-blocks of this rank (<= SK_MAX_KBS)
    const int k0 = (int)rank * kbs;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stg_sm = base;                              // partial tile: the A part of stage 0, free once the row-block's MMAs retired
    const uint32_t bars = base + SK_STAGES * SK_STAGE_BYTES;
    // full[s] +8 s | empty[s] +16 + 8 s | tfull +32 | stg_done +40 | ready +48 | freeb +56 | tmem pointer +64
    const uint32_t BAR_F = bars, BAR_E = bars + 16, BAR_TF = bars + 32, BAR_SD = bars + 40, BAR_RDY = bars + 48, BAR_FREE = bars + 56,
                   TMEM_SLOT = bars + 64;
    volatile uint32_t* tmem_ptr_sm = reinterpret_cast<volatile uint32_t*>(smem_raw + (TMEM_SLOT - smem_u32(smem_raw)));
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(SK_BN >> 3) << 17) | ((uint32_t)(SDF_BM >> 4) << 24);

    if (threadIdx.x == 0) {
        for (int s = 0; s < SK_STAGES; s++) { mbar_init(BAR_F + 8 * s, 1); mbar_init(BAR_E + 8 * s, 1); }
        mbar_init(BAR_TF, 1); mbar_init(BAR_SD, 4);
        mbar_init(BAR_RDY, SK_SPLIT); mbar_init(BAR_FREE, SK_SPLIT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(TMEM_SLOT), "r"((uint32_t)SK_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // the peers' barriers must be initialised before anyone arrives on them remotely
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_ptr_sm;
    if (threadIdx.x == 0) SK_T(1);
    // everything above is independent of the previous kernel; its outputs (activations, the device-side row count) are read below
    pdl_wait();
    const int live = m_count ? min(*m_count, m_rows) : m_rows;
    const int row_blocks = (live + SDF_BM - 1) / SDF_BM;       // uniform over the grid; 0: nothing to do but release TMEM

    if (warp == 0) {
        if (lane == 0) {
            // W slice: rows [64 nb, 64 nb + 64) = half a 128-row tile (8 KB contiguous per k-block and plane), k-blocks [k0, k0 + kbs)
            const int n0 = nb * SK_BN;
            const size_t wrow = (size_t)(n0 >> 7) * KB * SDF_TILE_FLOATS + (size_t)((n0 & 127) >> 3) * 256;
            uint32_t it = 0;
            for (int mb = 0; mb < row_blocks; mb++) {
                if (mb > 0) mbar_wait(BAR_SD, (uint32_t)(mb - 1) & 1);          // the previous partial tile (aliasing stage 0) has been consumed
                for (int kb = 0; kb < kbs; kb++, it++) {
                    const uint32_t s = it % SK_STAGES, ph = (it / SK_STAGES) & 1;
                    mbar_wait(BAR_E + 8 * s, ph ^ 1);
                    const uint32_t full = BAR_F + 8 * s;
                    mbar_expect_tx(full, SK_STAGE_BYTES);
                    const uint32_t st = base + s * SK_STAGE_BYTES;
                    const size_t oa = ((size_t)mb * KB + k0 + kb) * SDF_TILE_FLOATS;
                    const size_t ow = wrow + (size_t)(k0 + kb) * SDF_TILE_FLOATS;
                    tma_bulk_g2s(st, Xhi + oa, SDF_TILE_BYTES, full);
                    tma_bulk_g2s(st + SDF_TILE_BYTES, Xlo + oa, SDF_TILE_BYTES, full);
                    tma_bulk_g2s(st + 2 * SDF_TILE_BYTES, Whi + ow, SK_W_HALF_BYTES, full);
                    tma_bulk_g2s(st + 2 * SDF_TILE_BYTES + SK_W_HALF_BYTES, Wlo + ow, SK_W_HALF_BYTES, full);
                }
                if (mb == 0) SK_T(3);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int mb = 0; mb < row_blocks; mb++) {
                // the accumulator is free: the loads of this row-block were only issued after the previous epilogue finished (BAR_SD)
                for (int kb = 0; kb < kbs; kb++, it++) {
                    const uint32_t s = it % SK_STAGES, ph = (it / SK_STAGES) & 1;
                    mbar_wait(BAR_F + 8 * s, ph);
                    if (it == 0) SK_T(5);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t as = base + s * SK_STAGE_BYTES, ws = as + 2 * SDF_TILE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < SDF_BK / 8; ks++) {
                        const uint32_t ko = ks * 256;
                        const uint64_t ahi = umma_desc(as + ko), alo = umma_desc(as + SDF_TILE_BYTES + ko);
                        const uint64_t bhi = umma_desc(ws + ko), blo = umma_desc(ws + SK_W_HALF_BYTES + ko);
                        umma_tf32(tmem_acc, alo, bhi, idesc, (kb | ks) ? 1u : 0u);
                        umma_tf32(tmem_acc, ahi, blo, idesc, 1u);
                        umma_tf32(tmem_acc, ahi, bhi, idesc, 1u);
                    }
                    umma_commit(BAR_E + 8 * s);
                }
                umma_commit(BAR_TF);
                if (mb == 0) SK_T(6);
            }
        }
    } else {
        const int q = warp & 3;
        const int rl = 32 * q + lane;                               // row inside the 128-row block = TMEM lane
        const int et = threadIdx.x - 64;                            // 0..127 among the epilogue threads
        const int KBn = N / SDF_BK;
        const size_t row_off = (size_t)(rl >> 3) * 256 + (rl & 7) * 4;
        const int col0 = nb * SK_BN + 8 * (int)rank;                // the 8 output columns this CTA finishes
        const float4 bia0 = __ldg(reinterpret_cast<const float4*>(bias + col0));
        const float4 bia1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 4));
        const size_t out_off = (size_t)(col0 / SDF_BK) * SDF_TILE_FLOATS + (size_t)((col0 % SDF_BK) / 4) * 32 + row_off;
        for (int mb = 0; mb < row_blocks; mb++) {
            const uint32_t par = (uint32_t)mb & 1;
            mbar_wait(BAR_TF, par);                                 // every MMA of this row-block has retired: TMEM complete, stage 0 idle
            if (et == 0 && mb == 0) SK_T(7);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < SK_BN; c += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)c, v);
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++)                      // staging[c / 4][row][4]: consecutive lanes -> consecutive 16-byte words
                    st_smem_f4(stg_sm + (uint32_t)(((c >> 2) + c4) * SDF_BM + rl) * 16u, __uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]),
                               __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
            }
            if (et == 0 && mb == 0) SK_T(9);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0 && mb == 0) SK_T(10);
            if (et < SK_SPLIT) mbar_arrive_remote(mapa_u32(BAR_RDY, (uint32_t)et));
            mbar_wait_cluster(BAR_RDY, par);                        // all 8 partial tiles are in shared memory
            if (et == 0 && mb == 0) SK_T(11);
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (uint32_t r = 0; r < SK_SPLIT; r++) {               // fixed order: deterministic
                const uint32_t ra = mapa_u32(stg_sm + (uint32_t)((2 * rank) * SDF_BM + rl) * 16u, r);
                const float4 a0 = ld_dsmem_f4(ra), a1 = ld_dsmem_f4(ra + SDF_BM * 16u);
                acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w;
                acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
            }
            if (et == 0 && mb == 0) SK_T(12);
            asm volatile("bar.sync 1, 128;" ::: "memory");          // all reads of the peers' tiles are done
            if (et < SK_SPLIT) mbar_arrive_remote(mapa_u32(BAR_FREE, (uint32_t)et));
            const float bv[8] = {bia0.x, bia0.y, bia0.z, bia0.w, bia1.x, bia1.y, bia1.z, bia1.w};
            float h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float y = fmaxf(acc[e] + bv[e], 0.f);
                h[e] = tf32_hi(y); l[e] = y - h[e];
            }
            float* dh = Yhi + (size_t)mb * KBn * SDF_TILE_FLOATS + out_off;
            float* dl = Ylo + (size_t)mb * KBn * SDF_TILE_FLOATS + out_off;
            *reinterpret_cast<float4*>(dh) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(dh + 32) = make_float4(h[4], h[5], h[6], h[7]);
            *reinterpret_cast<float4*>(dl) = make_float4(l[0], l[1], l[2], l[3]);
            *reinterpret_cast<float4*>(dl + 32) = make_float4(l[4], l[5], l[6], l[7]);
            if (et == 0 && mb == 0) SK_T(13);
            mbar_wait_cluster(BAR_FREE, par);                       // no peer still reads this CTA's partial tile
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses to the tile before the next TMA write
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_SD);                     // stage 0 and the accumulator may be reused
        }
        if (et == 0) SK_T(14);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)SK_TMEM_COLS) : "memory");
    }
    if (threadIdx.x == 0) SK_T(15);
    SK_CTA(1);
}

// ---------------------------------------------------------------- ALL hidden layers of a few-row query in ONE launch
// k_sdf_gemm_sk pays a launch, a pipeline fill and a drain per layer; at a few hundred rows those fixed latencies are the whole cost
// (7 layers x ~14 us per pass of the per-step contact query).  k_sdf_chain_sk walks every hidden layer inside one launch:
//   * decomposition: cluster c (8 CTAs) owns output columns [128 c, 128 c + 128), CTA rank r the K slice [K r / 8, K (r + 1) / 8)
//     (split-K); every CTA PUSHES the 16 columns of its partial 128 x 128 accumulator that rank d finishes from TMEM straight into
//     rank d's shared memory (st.shared::cluster), and each CTA then sums its 8 slabs locally in FIXED rank order (deterministic, and
//     the same order as k_sdf_gemm_sk: bit-identical results).  Pulling the slices with ld.shared::cluster cost 3.5 us per tile.
//     H / 128 clusters = 64 CTAs at H = 1024: a cooperative launch accepts at most 15 co-resident 8-CTA clusters on a B200 (measured:
//     cudaOccupancyMaxActiveClusters), which rules out the 16 clusters of 64-column tiles;
//   * dependencies between layers are tracked per (row-block, column tile): the K slice a CTA reads is produced by the CTAs of the
//     cluster(s) owning those columns, which count in (fence + release add) when their part of a row-block is stored; the TMA producer
//     waits only for the tiles it is about to load.  Row-blocks therefore flow through the layers as a wavefront instead of meeting
//     at a device-wide barrier.  Cooperative launch => all CTAs are co-resident, the waits cannot deadlock.  The two ping-pong
//     activation buffers stay safe: a column tile of a row-block is overwritten (layer l + 1) only after the whole writing cluster
//     holds the layer-l outputs of ALL clusters for that row-block, i.e. after every reader of the old contents has finished;
//   * the W tiles of the next layer's first k-blocks do not depend on the activations: they are issued into the ring BEFORE the
//     dependency wait (same stage barrier; the A halves complete the transaction count afterwards);
//   * the slabs have their own 64 KB of shared memory and the accumulator is double-buffered in TMEM (2 x 128 columns), so loads and
//     MMAs of the next tile run under the drain / push / sum / store of the current one; a seventh warp publishes finished tiles
//     (release add on the dependency counter), which keeps the wait for the global stores off the epilogue's path;
//   * the kernel ends with a cluster barrier: no CTA may exit while a peer can still arrive on one of its barriers.
constexpr int CH_BN = 128;
constexpr int CH_STAGES = 2;
constexpr int CH_STAGE_BYTES = 4 * SDF_TILE_BYTES;                         // A_hi, A_lo, W_hi, W_lo of one k-block: 64 KB
constexpr int CH_STAGING_BYTES = SDF_BM * CH_BN * 4;                       // 64 KB: 8 slabs (one per source rank) of 128 rows x 16 columns fp32
constexpr int CH_TMEM_COLS = 256;                                         // two 128-column accumulators: the MMAs of tile t + 1 run while tile t is drained
constexpr int CH_SMEM_BYTES = CH_STAGES * CH_STAGE_BYTES + CH_STAGING_BYTES + 256 + 1024;
constexpr int CH_THREADS = SDF_THREADS + 32;                             // + one warp that publishes finished tiles (fence + release) off the epilogue's path

#ifdef MIS_CHAIN_TIMING
// phase timestamps of CTA MIS_CHAIN_TIMING, 16 slots per tile (scripts/ubench/chain_timing.cu)
__device__ unsigned long long ch_dbg[16 * 64];
#define CH_T(tile, k) do { if (blockIdx.x == MIS_CHAIN_TIMING && (tile) < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ch_dbg[(tile) * 16 + (k)] = t_; } } while (0)
#else
#define CH_T(tile, k) do { } while (0)
#endif

struct SkChain {
    const float* W;          // hidden-layer weights, contiguous: layer l = [hi plane H*H | lo plane H*H] at W + l * 2 H H (UMMA tiles)
    const float* bias;       // [n_layers][H]
    float* act[2][2];        // ping-pong activations [buffer][hi / lo]; layer l reads buffer l & 1, writes (l + 1) & 1
    int n_layers;
    unsigned* sync;          // [0] exit count, [1] unused, [2 + mb * NT + tile] arrivals of (row-block mb, column tile); the last CTA out zeroes what was used
};

__global__ void __cluster_dims__(SK_SPLIT, 1, 1) __launch_bounds__(CH_THREADS, 1)
k_sdf_chain_sk(SkChain c, int H, int m_rows, const int* __restrict__ m_count) {
    extern __shared__ uint8_t smem_raw[];
    const int nb = blockIdx.x / SK_SPLIT;                      // cluster id = column tile (gridDim.x = (H / 128) * 8 exactly)
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = H / SDF_BK;
    const int kbs = KB / SK_SPLIT;                             // k-blocks of this rank
    const int k0 = (int)rank * kbs;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stg_sm = base + CH_STAGES * CH_STAGE_BYTES; // partial tile
    const uint32_t bars = stg_sm + CH_STAGING_BYTES;
    // full[s] +8 s | empty[s] +16 + 8 s | tfull[b] +32 + 8 b | drained[b] +48 + 8 b | ready +64 | consumed +72 | stored[b] +80 + 8 b |
    // published[b] +96 + 8 b | tmem pointer +112
    const uint32_t BAR_F = bars, BAR_E = bars + 16, BAR_TF = bars + 32, BAR_SD = bars + 48, BAR_RDY = bars + 64, BAR_CONS = bars + 72,
                   BAR_ST = bars + 80, BAR_STA = bars + 96, TMEM_SLOT = bars + 112;
    volatile uint32_t* tmem_ptr_sm = reinterpret_cast<volatile uint32_t*>(smem_raw + (TMEM_SLOT - smem_u32(smem_raw)));
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CH_BN >> 3) << 17) | ((uint32_t)(SDF_BM >> 4) << 24);
    const unsigned nctas = gridDim.x;
    const int NT = H / CH_BN;                                  // column tiles = clusters

    if (threadIdx.x == 0) {
        for (int s = 0; s < CH_STAGES; s++) { mbar_init(BAR_F + 8 * s, 1); mbar_init(BAR_E + 8 * s, 1); }
        for (int b2 = 0; b2 < 2; b2++) { mbar_init(BAR_TF + 8 * b2, 1); mbar_init(BAR_SD + 8 * b2, 4); }
        mbar_init(BAR_RDY, SK_SPLIT); mbar_init(BAR_CONS, SK_SPLIT);
        for (int b2 = 0; b2 < 2; b2++) { mbar_init(BAR_ST + 8 * b2, 1); mbar_init(BAR_STA + 8 * b2, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(TMEM_SLOT), "r"((uint32_t)CH_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_ptr_sm;
    const int live = m_count ? min(*m_count, m_rows) : m_rows;
    const int row_blocks = (live + SDF_BM - 1) / SDF_BM;       // uniform over the grid; 0: nothing to do but release TMEM
    const int nl = row_blocks > 0 ? c.n_layers : 0;
    const size_t HH = (size_t)H * H;

    if (warp == 0) {
        if (lane == 0) {
            // W tile (rows [128 nb, 128 nb + 128), k-block kb) of layer l into the stage of ring position `pos`: waits for the stage, arms its barrier
            auto issue_w = [&](int l, int kb, uint32_t pos) {
                const uint32_t s = pos % CH_STAGES, ph = (pos / CH_STAGES) & 1;
                mbar_wait(BAR_E + 8 * s, ph ^ 1);
                const uint32_t full = BAR_F + 8 * s;
                mbar_expect_tx(full, CH_STAGE_BYTES);
                const uint32_t st = base + s * CH_STAGE_BYTES;
                const float* whi = c.W + (size_t)l * 2 * HH;
                const size_t ow = ((size_t)nb * KB + k0 + kb) * SDF_TILE_FLOATS;
                tma_bulk_g2s(st + 2 * SDF_TILE_BYTES, whi + ow, SDF_TILE_BYTES, full);
                tma_bulk_g2s(st + 3 * SDF_TILE_BYTES, whi + HH + ow, SDF_TILE_BYTES, full);
            };
            const int npre = kbs < CH_STAGES ? kbs : CH_STAGES;
            uint32_t it = 0;
            for (int l = 0; l < nl; l++) {
                const float* xhi = c.act[l & 1][0];
                const float* xlo = c.act[l & 1][1];
                for (int mb = 0; mb < row_blocks; mb++) {
                    const bool first = l > 0;
                    CH_T(l * row_blocks + mb, 0);
                    if (first) {
                        for (int j = 0; j < npre; j++) issue_w(l, j, it + (uint32_t)j);      // independent of the activations
                        // this rank's K slice of row-block mb = columns [32 k0, 32 (k0 + kbs)) of layer l - 1: wait for the column tiles holding them
                        const int t_lo = (32 * k0) / CH_BN, t_hi = (32 * (k0 + kbs) - 1) / CH_BN;
                        const unsigned need = (unsigned)l * SK_SPLIT;                         // 8 CTAs of the owning cluster, l layers so far
                        for (int t = t_lo; t <= t_hi; t++) {
                            const unsigned* cnt = c.sync + 2 + (size_t)mb * NT + t;
                            unsigned seen;
                            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory"); } while (seen < need);
                        }
                        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of the other CTAs before this CTA's bulk copies
                    }
                    CH_T(l * row_blocks + mb, 1);
                    for (int kb = 0; kb < kbs; kb++, it++) {
                        if (!(first && kb < npre)) issue_w(l, kb, it);
                        const uint32_t s = it % CH_STAGES;
                        const uint32_t full = BAR_F + 8 * s;
                        const uint32_t st = base + s * CH_STAGE_BYTES;
                        const size_t oa = ((size_t)mb * KB + k0 + kb) * SDF_TILE_FLOATS;
                        tma_bulk_g2s(st, xhi + oa, SDF_TILE_BYTES, full);
                        tma_bulk_g2s(st + SDF_TILE_BYTES, xlo + oa, SDF_TILE_BYTES, full);
                    }
                    CH_T(l * row_blocks + mb, 2);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0;
            int tile = 0;
            for (int l = 0; l < nl; l++) {
                for (int mb = 0; mb < row_blocks; mb++, tile++) {
                    const uint32_t ab = (uint32_t)tile & 1;              // accumulator buffer of this tile
                    const uint32_t acc = tmem_acc + ab * CH_BN;
                    if (tile > 1) {                                      // the epilogue has drained this buffer (tile - 2)
                        mbar_wait(BAR_SD + 8 * ab, (uint32_t)((tile >> 1) - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int kb = 0; kb < kbs; kb++, it++) {
                        const uint32_t s = it % CH_STAGES, ph = (it / CH_STAGES) & 1;
                        mbar_wait(BAR_F + 8 * s, ph);
                        if (kb == 0) CH_T(tile, 3);
                        if (kb == kbs - 1) CH_T(tile, 4);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t as = base + s * CH_STAGE_BYTES, ws = as + 2 * SDF_TILE_BYTES;
#pragma unroll
                        for (int ks = 0; ks < SDF_BK / 8; ks++) {
                            const uint32_t ko = ks * 256;
                            const uint64_t ahi = umma_desc(as + ko), alo = umma_desc(as + SDF_TILE_BYTES + ko);
                            const uint64_t bhi = umma_desc(ws + ko), blo = umma_desc(ws + SDF_TILE_BYTES + ko);
                            umma_tf32(acc, alo, bhi, idesc, (kb | ks) ? 1u : 0u);
                            umma_tf32(acc, ahi, blo, idesc, 1u);
                            umma_tf32(acc, ahi, bhi, idesc, 1u);
                        }
                        umma_commit(BAR_E + 8 * s);
                    }
                    umma_commit(BAR_TF + 8 * ab);
                }
            }
        }
    } else if (warp == 6) {
        // publisher: makes a finished tile visible device-wide and counts it in, off the epilogue's critical path.  The epilogue threads'
        // global stores happen-before its fence through the CTA barrier + mbarrier (release / acquire), the fence is cumulative.
        if (lane == 0) {
            int tile = 0;
            for (int l = 0; l + 1 < nl; l++) {
                for (int mb = 0; mb < row_blocks; mb++, tile++) {
                    const uint32_t sb = (uint32_t)tile & 1;
                    mbar_wait(BAR_ST + 8 * sb, (uint32_t)(tile >> 1) & 1);
                    asm volatile("fence.proxy.async;" ::: "memory");
                    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c.sync + 2 + (size_t)mb * NT + nb), "r"(1u) : "memory");   // release: cumulative over the tile's stores
                    mbar_arrive(BAR_STA + 8 * sb);
                }
            }
        }
    } else {
        // epilogue.  Split-K reduce-scatter by PUSH: every CTA drains its 128 x 128 partial accumulator from TMEM and stores the 16 columns
        // that rank d finishes straight into rank d's shared memory (st.shared::cluster: posted writes, no round trip), slab [source rank];
        // after the cluster-wide "pushed" barrier each CTA sums its 8 slabs locally in fixed rank order (deterministic; the order of
        // k_sdf_gemm_sk: bit-identical results).  Pulling the slices with ld.shared::cluster instead cost 3.5 us per tile.
        const int q = warp & 3;
        const int rl = 32 * q + lane;                               // row inside the 128-row block = TMEM lane
        const int et = threadIdx.x - 64;                            // 0..127 among the epilogue threads
        const size_t row_off = (size_t)(rl >> 3) * 256 + (rl & 7) * 4;
        const int col0 = nb * CH_BN + 16 * (int)rank;               // the 16 output columns this CTA finishes
        const size_t out_off = (size_t)(col0 / SDF_BK) * SDF_TILE_FLOATS + (size_t)((col0 % SDF_BK) / 4) * 32 + row_off;
        // slab of source rank r, column group g (4 columns), row: ((r * 4 + g) * 128 + row) * 16 bytes inside the 64 KB staging area
        const uint32_t my_slab = stg_sm + (uint32_t)((rank * 4) * SDF_BM + rl) * 16u;     // where THIS rank's values land in a destination CTA
        int tile = 0;
        for (int l = 0; l < nl; l++) {
            float bv[16];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(c.bias + (size_t)l * H + col0 + 4 * g));
                bv[4 * g] = b4.x; bv[4 * g + 1] = b4.y; bv[4 * g + 2] = b4.z; bv[4 * g + 3] = b4.w;
            }
            float* yhi = c.act[(l + 1) & 1][0];
            float* ylo = c.act[(l + 1) & 1][1];
            for (int mb = 0; mb < row_blocks; mb++, tile++) {
                const uint32_t par = (uint32_t)tile & 1;                // phase of the cluster barriers (one use per tile)
                const uint32_t ab = (uint32_t)tile & 1;                 // accumulator buffer
                if (et == 0) CH_T(tile, 5);
                mbar_wait(BAR_TF + 8 * ab, (uint32_t)(tile >> 1) & 1);  // every MMA of this tile has retired: the accumulator is complete
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tile > 0) mbar_wait_cluster(BAR_CONS, (uint32_t)(tile - 1) & 1);   // every destination has summed the previous tile's slabs
                if (et == 0) CH_T(tile, 6);
#pragma unroll
                for (int cc = 0; cc < CH_BN; cc += 64) {                // two 32-column loads in flight per wait
                    uint32_t v[32], v2[32];
                    tmem_ld32_nowait(tmem_acc + ((uint32_t)(32 * q) << 16) + ab * CH_BN + (uint32_t)cc, v);
                    tmem_ld32_nowait(tmem_acc + ((uint32_t)(32 * q) << 16) + ab * CH_BN + (uint32_t)cc + 32u, v2);
                    tmem_ld_wait();
#pragma unroll
                    for (int h2 = 0; h2 < 4; h2++) {                    // columns [cc + 16 h2, + 16) belong to rank cc / 16 + h2
                        const uint32_t* vv = h2 < 2 ? v : v2;
                        const int hh = h2 & 1;
                        const uint32_t dst = mapa_u32(my_slab, (uint32_t)(cc >> 4) + h2);
#pragma unroll
                        for (int g = 0; g < 4; g++)
                            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)g * SDF_BM * 16u),
                                         "f"(__uint_as_float(vv[16 * hh + 4 * g])), "f"(__uint_as_float(vv[16 * hh + 4 * g + 1])),
                                         "f"(__uint_as_float(vv[16 * hh + 4 * g + 2])), "f"(__uint_as_float(vv[16 * hh + 4 * g + 3])) : "memory");
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et < SK_SPLIT) mbar_arrive_remote(mapa_u32(BAR_RDY, (uint32_t)et));     // "my slab is in your staging area"
                if (lane == 0) mbar_arrive(BAR_SD + 8 * ab);            // this accumulator has been read: tile + 2 may overwrite it
                if (et == 0) CH_T(tile, 7);
                mbar_wait_cluster(BAR_RDY, par);                        // all 8 slabs have landed here
                if (et == 0) CH_T(tile, 8);
                float acc[16];
#pragma unroll
                for (int e = 0; e < 16; e++) acc[e] = 0.f;
#pragma unroll
                for (uint32_t r = 0; r < SK_SPLIT; r++) {               // fixed order: deterministic
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        float4 a;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                                     : "r"(stg_sm + (uint32_t)((r * 4 + g) * SDF_BM + rl) * 16u) : "memory");
                        acc[4 * g] += a.x; acc[4 * g + 1] += a.y; acc[4 * g + 2] += a.z; acc[4 * g + 3] += a.w;
                    }
                }
                if (et == 0) CH_T(tile, 9);
                float* dh = yhi + (size_t)mb * KB * SDF_TILE_FLOATS + out_off;
                float* dl = ylo + (size_t)mb * KB * SDF_TILE_FLOATS + out_off;
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float h[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float y = fmaxf(acc[4 * g + e] + bv[4 * g + e], 0.f);
                        h[e] = tf32_hi(y); lo[e] = y - h[e];
                    }
                    *reinterpret_cast<float4*>(dh + 32 * g) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(dl + 32 * g) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");          // staging read by everyone here; this tile's global stores are issued
                if (et < SK_SPLIT) mbar_arrive_remote(mapa_u32(BAR_CONS, (uint32_t)et));    // "I have summed your slab": the source may push again
                if (et == 0 && l + 1 < nl) {
                    // hand the tile to the publisher warp (two tiles may be pending: double-buffered handshake)
                    const uint32_t sb = (uint32_t)tile & 1;
                    if (tile > 1) mbar_wait(BAR_STA + 8 * sb, (uint32_t)((tile >> 1) - 1) & 1);
                    mbar_arrive(BAR_ST + 8 * sb);
                }
                if (et == 0) CH_T(tile, 10);
                if (et == 0) CH_T(tile, 11);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // No CTA may leave while a peer can still touch its shared memory: after the last tile the destinations still arrive on this CTA's
    // "consumed" barrier (remote mbarrier arrive) once they have summed its slab.  Leaving early is an access to the shared memory of an
    // exited CTA -- a sporadic launch failure under load.  The cluster barrier closes the window.
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)CH_TMEM_COLS) : "memory");
    }
    if (threadIdx.x == 0) {
        // every CTA has passed its last barrier wait before it gets here: the last one out re-arms the counters
        const unsigned old = atomicAdd(c.sync, 1u);
        if (old == nctas - 1) {
            for (int k = 0; k < row_blocks * NT; k++) c.sync[2 + k] = 0u;
            c.sync[0] = 0u;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------- weights
// weight_norm (deepsdf.py:3): W[o, :] = g[o] * v[o, :] / ||v[o, :]||.  One block per output row; writes the
// plain fp32 row (first / last layer) and/or the hi / lo UMMA tiles (hidden layers).
__global__ void __launch_bounds__(256) k_sdf_pack_weights(const float* __restrict__ g, const float* __restrict__ v, int n_out, int n_in,
                                                          float* __restrict__ plain, float* __restrict__ Whi, float* __restrict__ Wlo) {
    __shared__ float red[8];
    __shared__ float scale_s;
    const int o = blockIdx.x;
    float ss = 0.f;
    for (int k = threadIdx.x; k < n_in; k += 256) { float x = v[(size_t)o * n_in + k]; ss += x * x; }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; w++) t += red[w];
        scale_s = g[o] / sqrtf(t);
    }
    __syncthreads();
    const float sc = scale_s;
    const int KB = n_in / SDF_BK;
    for (int k = threadIdx.x; k < n_in; k += 256) {
        const float w = v[(size_t)o * n_in + k] * sc;
        if (plain) plain[(size_t)o * n_in + k] = w;
        if (Whi) {
            const size_t off = sdf_tile_off(o, k, KB);
            const float h = tf32_hi(w);
            Whi[off] = h; Wlo[off] = w - h;
        }
    }
}

// ---------------------------------------------------------------- first layer: Linear(3, H) + ReLU on CUDA cores (K = 3)
// points are taken through p_model = A (p - t) first (A row-major 3x3; identity / zero for model-space input).
// One thread per (row, 4 consecutive outputs) = one 16-byte chunk of the hi and lo tiles.
struct SdfXform { float A[9]; float t[3]; };

// fd3 != 0: row r evaluates point r / 3 displaced by `shift.x` along model axis r % 3 (the three forward differences in one pass).
__global__ void __launch_bounds__(256) k_sdf_layer0(const float* __restrict__ pts, const int* __restrict__ idx, int m, int m_pad,
                                                    const int* __restrict__ m_count, SdfXform xf, float3 shift0, int fd3,
                                                    const float* __restrict__ W0 /* [H,3] */, const float* __restrict__ b0, int H,
                                                    float* __restrict__ Yhi, float* __restrict__ Ylo) {
    const int chunks = H / 4;
    pdl_trigger();
    pdl_wait();
    const int live = m_count ? min(*m_count, m) : m;
    const long long work = (long long)((live + 127) & ~127) * chunks;        // whole row-blocks: the GEMM reads 128-row tiles
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < work; gid += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(gid / chunks), ch = (int)(gid % chunks);
        float x = 0.f, y = 0.f, z = 0.f;
        if (r < live) {
            const int pr = fd3 ? r / 3 : r;
            const int src = idx ? idx[pr] : pr;
            float3 shift = shift0;
            if (fd3) { const int a = r - 3 * pr; shift = make_float3(a == 0 ? shift0.x : 0.f, a == 1 ? shift0.x : 0.f, a == 2 ? shift0.x : 0.f); }
            const float px = pts[3 * (size_t)src] - xf.t[0], py = pts[3 * (size_t)src + 1] - xf.t[1], pz = pts[3 * (size_t)src + 2] - xf.t[2];
            x = xf.A[0] * px + xf.A[1] * py + xf.A[2] * pz + shift.x;
            y = xf.A[3] * px + xf.A[4] * py + xf.A[5] * pz + shift.y;
            z = xf.A[6] * px + xf.A[7] * py + xf.A[8] * pz + shift.z;
        }
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int o = 4 * ch + e;
            // nn.Linear: x W^T + b, accumulated in k order like a dot product
            float acc = x * W0[3 * o];
            acc = fmaf(y, W0[3 * o + 1], acc);
            acc = fmaf(z, W0[3 * o + 2], acc);
            acc = fmaxf(acc + b0[o], 0.f);
            h[e] = tf32_hi(acc); l[e] = acc - h[e];
        }
        const size_t off = sdf_tile_off(r, 4 * ch, H / SDF_BK);
        *reinterpret_cast<float4*>(Yhi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(Ylo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// ---------------------------------------------------------------- last layer: Linear(H, 1) on CUDA cores
// One warp per row; lane = k-block, fixed summation order (deterministic).  Returns the full dot product in every lane.
__device__ __forceinline__ float sdf_last_row(const float* __restrict__ Xhi, const float* __restrict__ Xlo, int r,
                                              const float* __restrict__ w, int H, int lane) {
    const int KB = H / SDF_BK;
    float acc = 0.f;
    for (int kb = lane; kb < KB; kb += 32) {
        const size_t base = ((size_t)(r >> 7) * KB + kb) * SDF_TILE_FLOATS + ((r & 127) >> 3) * 256 + (r & 7) * 4;
#pragma unroll
        for (int kc = 0; kc < 8; kc++) {
            const float4 h = *reinterpret_cast<const float4*>(Xhi + base + kc * 32);
            const float4 l = *reinterpret_cast<const float4*>(Xlo + base + kc * 32);
            const float4 ww = *reinterpret_cast<const float4*>(w + kb * 32 + kc * 4);
            acc = fmaf(h.x + l.x, ww.x, acc); acc = fmaf(h.y + l.y, ww.y, acc);
            acc = fmaf(h.z + l.z, ww.z, acc); acc = fmaf(h.w + l.w, ww.w, acc);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    return acc;
}

__global__ void __launch_bounds__(256) k_sdf_last(const float* __restrict__ Xhi, const float* __restrict__ Xlo, int m, const int* __restrict__ m_count,
                                                  const float* __restrict__ w /* [H] */, const float* __restrict__ b, int H, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    pdl_wait();
    const int live = m_count ? min(*m_count, m) : m;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < live; r += gridDim.x * 8) {
        const float acc = sdf_last_row(Xhi, Xlo, r, w, H, lane);
        if (lane == 0) out[r] = acc + b[0];
    }
}

}  // namespace mis
