// Phase timestamps (%globaltimer) of one CTA of k_sdf_gemm_sk (diagnostic tooling; build: see scripts/ubench/README).
#ifndef MIS_SK_TIMING
#define MIS_SK_TIMING 0
#endif
#include "../../meshless_inflatable_softbody_b200/csrc/mis_sdf.cuh"
#include <cstdio>
#include <cstdlib>
using namespace mis;
int main(int argc, char** argv) {
    const int H = 1024, rows = argc > 1 ? atoi(argv[1]) : 128;
    const int m_pad = (rows + 127) / 128 * 128;
    float *a[2][2], *whi, *wlo, *b;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { cudaMalloc(&a[i][j], (size_t)m_pad * H * 4); cudaMemset(a[i][j], 0, (size_t)m_pad * H * 4); }
    cudaMalloc(&whi, (size_t)H * H * 4); cudaMalloc(&wlo, (size_t)H * H * 4); cudaMalloc(&b, H * 4);
    cudaMemset(whi, 0, (size_t)H * H * 4); cudaMemset(wlo, 0, (size_t)H * H * 4); cudaMemset(b, 0, H * 4);
    cudaFuncSetAttribute(k_sdf_gemm_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, st);
        for (int l = 0; l < 20; l++)
            k_sdf_gemm_sk<<<(H / SK_BN) * SK_SPLIT, SDF_THREADS, SK_SMEM_BYTES, st>>>(a[l & 1][0], a[l & 1][1], whi, wlo, b, H, H, a[(l & 1) ^ 1][0], a[(l & 1) ^ 1][1], rows, nullptr);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rows %d: %.2f us / launch (%s)\n", rows, 1e3 * ms / 20, cudaGetErrorString(cudaGetLastError()));
    }
    // the same chain as programmatic dependent launches (stream, then captured into a graph)
    auto chain = [&](bool pdl) {
        for (int l = 0; l < 20; l++) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((H / SK_BN) * SK_SPLIT); cfg.blockDim = dim3(SDF_THREADS); cfg.dynamicSmemBytes = SK_SMEM_BYTES; cfg.stream = st;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
            cudaLaunchKernelEx(&cfg, k_sdf_gemm_sk, (const float*)a[l & 1][0], (const float*)a[l & 1][1], (const float*)whi, (const float*)wlo, (const float*)b, H, H,
                               a[(l & 1) ^ 1][0], a[(l & 1) ^ 1][1], rows, (const int*)nullptr);
        }
    };
    for (int pdl = 0; pdl < 2; pdl++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0, st); chain(pdl); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("stream, pdl=%d: %.2f us / launch (%s)\n", pdl, 1e3 * ms / 20, cudaGetErrorString(cudaGetLastError()));
        }
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal); chain(pdl); cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("graph,  pdl=%d: %.2f us / launch (%s)\n", pdl, 1e3 * ms / 20, cudaGetErrorString(cudaGetLastError()));
        }
    }
    unsigned long long t[64];
    cudaMemcpyFromSymbol(t, sk_dbg, sizeof t);
    const char* names[] = {"start", "prologue done", "W issued", "A all issued", "W landed", "A stage 0 landed", "all MMAs issued", "accumulators complete",
                           "free wait", "staging written", "fence+bar", "ready (all partials)", "dsmem reads done", "stores done", "epilogue done", "end", "A stage 3 landed", "2nd rb free wait"};
    for (int k = 0; k < 16; k++) printf("  %-24s %8.2f us\n", names[k], (double)(long long)(t[k] - t[0]) * 1e-3);
    static unsigned long long c[4 * 256];
    cudaMemcpyFromSymbol(c, sk_cta, sizeof c);
    unsigned long long t0 = ~0ull;
    for (int b = 0; b < 128; b++) if (c[4 * b] < t0) t0 = c[4 * b];
    printf("per-cluster (cluster: first start .. last end us, SMs):\n");
    for (int cl = 0; cl < 16; cl++) {
        unsigned long long s0 = ~0ull, e1 = 0;
        for (int r = 0; r < 8; r++) { int b = cl * 8 + r; if (c[4 * b] < s0) s0 = c[4 * b]; if (c[4 * b + 1] > e1) e1 = c[4 * b + 1]; }
        printf("  cluster %2d: %7.2f .. %7.2f   sm", cl, (double)(s0 - t0) * 1e-3, (double)(e1 - t0) * 1e-3);
        for (int r = 0; r < 8; r++) printf(" %3d", (int)c[4 * (cl * 8 + r) + 2]);
        printf("\n");
    }
    int nclus = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(SDF_THREADS); cfg.dynamicSmemBytes = SK_SMEM_BYTES;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclus, k_sdf_gemm_sk, &cfg);
    printf("cudaOccupancyMaxActiveClusters = %d (%s)\n", nclus, cudaGetErrorString(e));
    return 0;
}
