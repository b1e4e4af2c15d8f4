// Phase timestamps (%globaltimer) of one CTA of k_sdf_chain_sk per (layer, row-block) tile (diagnostic tooling).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMIS_CHAIN_TIMING=9 -o chain_timing chain_timing.cu
#ifndef MIS_CHAIN_TIMING
#define MIS_CHAIN_TIMING 9
#endif
#include "../../meshless_inflatable_softbody_b200/csrc/mis_sdf.cuh"
#include <cstdio>
#include <cstdlib>
using namespace mis;
int main(int argc, char** argv) {
    const int H = 1024, rows = argc > 1 ? atoi(argv[1]) : 256, NL = 7;
    const int m_pad = (rows + 127) / 128 * 128, rb = m_pad / 128;
    SkChain c = {};
    float *w, *b;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { cudaMalloc(&c.act[i][j], (size_t)m_pad * H * 4); cudaMemset(c.act[i][j], 0, (size_t)m_pad * H * 4); }
    cudaMalloc(&w, (size_t)NL * 2 * H * H * 4); cudaMemset(w, 0, (size_t)NL * 2 * H * H * 4);
    cudaMalloc(&b, (size_t)NL * H * 4); cudaMemset(b, 0, (size_t)NL * H * 4);
    cudaMalloc(&c.sync, (2 + rb * 8) * 4); cudaMemset(c.sync, 0, (2 + rb * 8) * 4);
    c.W = w; c.bias = b; c.n_layers = NL;
    cudaFuncSetAttribute(k_sdf_chain_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((H / CH_BN) * SK_SPLIT); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = CH_SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaEventRecord(e0, st);
        cudaError_t le = cudaLaunchKernelEx(&cfg, k_sdf_chain_sk, c, H, rows, (const int*)nullptr);
        cudaEventRecord(e1, st); cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rows %d (%d row-blocks): %.2f us for %d layers (%s / %s)\n", rows, rb, 1e3 * ms, NL, cudaGetErrorString(le), cudaGetErrorString(cudaGetLastError()));
    }
    static unsigned long long t[16 * 64];
    cudaMemcpyFromSymbol(t, ch_dbg, sizeof t);
    const char* names[] = {"dep wait start", "dep wait done", "A issued", "stage0 landed", "last stage landed", "epi: wait TF", "epi: TF seen", "drained+SD",
                           "RDY (all partials)", "reduced", "stored+counted", "FREE"};
    const unsigned long long t0 = t[0];
    printf("tile (layer,rb):");
    for (int k = 0; k < 12; k++) printf(" %s |", names[k]);
    printf("\n");
    for (int tile = 0; tile < NL * rb && tile < 64; tile++) {
        printf("(%d,%d)", tile / rb, tile % rb);
        for (int k = 0; k < 12; k++) printf(" %7.2f", (double)(long long)(t[tile * 16 + k] - t0) * 1e-3);
        printf("\n");
    }
    return 0;
}
