// pxb_env.cu -- instantiations and launchers of the environment path's kernels (pxb_env.cuh: k_env_bp, k_env_solve).
#include "pxb_launch.h"

#ifndef PXB_ENV_BP_CTA_MAX_ENVS
#define PXB_ENV_BP_CTA_MAX_ENVS 148   // up to one CTA per SM; beyond that a warp per environment keeps more environments in flight
#endif

cudaError_t pxb_env_set_attributes(int solveSmemMax, int bpSmemMax) {
  cudaError_t e = cudaSuccess;
#define ENV_ATTR(T) do { if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_solve<T, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, solveSmemMax); \
                         if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_solve<T, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, solveSmemMax); \
                         if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_solve<T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, solveSmemMax); \
                         if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_solve<T, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, solveSmemMax); } while (0)
  ENV_ATTR(32); ENV_ATTR(64); ENV_ATTR(128); ENV_ATTR(256);
#undef ENV_ATTR
#define BP_ATTR(H, L) do { if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_bp<H, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, bpSmemMax); \
                           if (e == cudaSuccess) e = cudaFuncSetAttribute(k_env_bp_cta<H, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, bpSmemMax); } while (0)
  BP_ATTR(false, false); BP_ATTR(true, false); BP_ATTR(false, true); BP_ATTR(true, true);
#undef BP_ATTR
  return e;
}
void pxb_launch_env_bp(cudaStream_t st, const EnvBpArgs& A, bool hulls, size_t smem) {
  const bool local = A.L.s2bP != nullptr || A.shapeOff != nullptr || A.aggId != nullptr;
#define BP_PICK(K, ...) do { if (hulls) { if (local) K<true, true>__VA_ARGS__; else K<true, false>__VA_ARGS__; } else { if (local) K<false, true>__VA_ARGS__; else K<false, false>__VA_ARGS__; } } while (0)
  if (A.nEnv <= PXB_ENV_BP_CTA_MAX_ENVS) {   // few environments: a CTA per environment (the warps share the rows), else the step is one warp's latency
    const size_t ctaSmem = (size_t)A.maxList * (2 * sizeof(float4) + 2 * sizeof(uint32_t));
    BP_PICK(k_env_bp_cta, <<<A.nEnv, ENV_BP_CTA_THREADS, ctaSmem, st>>>(A));
    return;
  }
  const uint32_t grid = (A.nEnv + ENV_BP_WARPS - 1) / ENV_BP_WARPS;
  BP_PICK(k_env_bp, <<<grid, 32 * ENV_BP_WARPS, smem, st>>>(A));
#undef BP_PICK
}
void pxb_launch_env_solve(cudaStream_t st, const EnvSolveArgs& A, uint32_t threads, bool pgs, bool ext, size_t smem) {
#define ENV_LAUNCH(T) do { if (pgs) { if (ext) k_env_solve<T, true, true><<<A.nEnv, T, smem, st>>>(A); else k_env_solve<T, true, false><<<A.nEnv, T, smem, st>>>(A); } \
                           else { if (ext) k_env_solve<T, false, true><<<A.nEnv, T, smem, st>>>(A); else k_env_solve<T, false, false><<<A.nEnv, T, smem, st>>>(A); } } while (0)
  if (threads == 32) ENV_LAUNCH(32); else if (threads == 64) ENV_LAUNCH(64); else if (threads == 128) ENV_LAUNCH(128); else ENV_LAUNCH(256);
#undef ENV_LAUNCH
}
