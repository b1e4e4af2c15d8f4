// Probe: how many 8-CTA clusters of k_sdf_chain_sk / k_sdf_gemm_sk are co-resident, and does a cooperative launch accept them?
#include "../../meshless_inflatable_softbody_b200/csrc/mis_sdf.cuh"
#include <cstdio>
using namespace mis;
int main() {
    cudaFuncSetAttribute(k_sdf_chain_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sdf_chain_sk, SDF_THREADS, SK_SMEM_BYTES);
    printf("chain blocks/SM = %d\n", per_sm);
    cudaFuncSetAttribute(k_sdf_gemm_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sdf_gemm_sk, SDF_THREADS, SK_SMEM_BYTES);
    printf("gemm_sk blocks/SM = %d\n", per_sm);
    for (int sm : {0, 49152, 65536, 99584, 110000}) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sdf_chain_sk, SDF_THREADS, sm); printf("chain smem %d -> %d\n", sm, per_sm); }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_sdf_chain_sk); printf("chain regs %d static smem %zu local %zu maxdyn %d\n", fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxDynamicSharedSizeBytes);
    cudaFuncGetAttributes(&fa, k_sdf_gemm_sk); printf("gemm_sk regs %d static smem %zu local %zu maxdyn %d\n", fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxDynamicSharedSizeBytes);
    { cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(128); cfg.blockDim = dim3(SDF_THREADS); cfg.dynamicSmemBytes = SK_SMEM_BYTES;
      int nc = -1; cudaOccupancyMaxActiveClusters(&nc, k_sdf_gemm_sk, &cfg); printf("gemm_sk max active clusters %d\n", nc); }
    for (int coop = 0; coop < 2; coop++)
        for (int grid : {128, 120, 64, 8}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(SDF_THREADS); cfg.dynamicSmemBytes = SK_SMEM_BYTES;
            cudaLaunchAttribute at[2];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
            cfg.attrs = at; cfg.numAttrs = 1 + coop;
            int nclusters = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_sdf_chain_sk, &cfg);
            printf("coop=%d grid=%d: max active clusters = %d (%s)\n", coop, grid, nclusters, cudaGetErrorString(e));
            cudaGetLastError();
        }
    // actual cooperative launches with zero layers (kernel exits at once)
    unsigned* sync; cudaMalloc(&sync, 64); cudaMemset(sync, 0, 64);
    SkChain c = {}; c.n_layers = 0; c.sync = sync;
    for (int grid : {128, 120, 112, 64, 8}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(SDF_THREADS); cfg.dynamicSmemBytes = SK_SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_sdf_chain_sk, c, 1024, 0, (const int*)nullptr);
        cudaError_t e2 = cudaDeviceSynchronize();
        printf("coop launch grid=%d: %s / %s\n", grid, cudaGetErrorString(e), cudaGetErrorString(e2));
        cudaGetLastError();
    }
    return 0;
}
