// pxb_solve.cu -- the device-wide solver kernels (pxb_pgs.cuh: k_prep_rows, k_solve_tgs / k_solve_pgs, k_writeback_rows, k_finalize_bodies_pgs)
// and their launchers.
#define PXB_SOLVE_KERNELS
#include "pxb_launch.h"

cudaError_t pxb_solve_occupancy(int* tgs, int* pgs) {
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(pgs, k_solve_pgs, 256, 0);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(tgs, k_solve_tgs<false>, 256, 0);
  int kin = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&kin, k_solve_tgs<true>, 256, 0);   // one grid size serves both instantiations
  if (e == cudaSuccess && kin < *tgs) *tgs = kin;
  return e;
}
void pxb_launch_prep_rows(cudaStream_t st, bool pgs, uint32_t capPairs, const PrepArgs& A) {
  const uint32_t grid = (capPairs + 127) / 128;
  if (pgs) k_prep_rows<true><<<grid, 128, 0, st>>>(A.counters, A.ordered, A.conPair, A.pairSlots, A.pairBodies, A.geomFlags, A.cHdr, A.cPts, A.pos, A.quat, A.linVel, A.sbOrigAng, A.invInertia, A.sbIA, A.sbIB, A.frictions, A.P, A.R, A.M, A.angVel, A.kinFtv);
  else k_prep_rows<false><<<grid, 128, 0, st>>>(A.counters, A.ordered, A.conPair, A.pairSlots, A.pairBodies, A.geomFlags, A.cHdr, A.cPts, A.pos, A.quat, A.linVel, A.sbOrigAng, A.invInertia, A.sbIA, A.sbIB, A.frictions, A.P, A.R, A.M, A.angVel, A.kinFtv);
}
cudaError_t pxb_launch_solve(cudaStream_t st, bool pgs, int blocks, SolveArgs& A) {
  if (pgs) {
    void* args[] = {&A.counters, &A.partStart, &A.posIters, &A.velIters, &A.R, &A.sbLin, &A.sbAng, &A.sbDLin, &A.sbDAng, &A.nDyn, &A.dynActor};
    return cudaLaunchCooperativeKernel((void*)k_solve_pgs, dim3(blocks), dim3(256), args, 0, st);
  }
  void* args[] = {&A.counters, &A.partStart, &A.posIters, &A.velIters, &A.stepDt, &A.R, &A.sbLin, &A.sbAng, &A.sbDLin, &A.sbDAng, &A.sbIA, &A.sbIB, &A.sbP, &A.sbQ, &A.bodyHasCon, &A.nDyn, &A.dynActor, &A.kinFtv};
  if (A.kinFtv) return cudaLaunchCooperativeKernel((void*)k_solve_tgs<true>, dim3(blocks), dim3(256), args, 0, st);   // scenes with kinematic bodies
  return cudaLaunchCooperativeKernel((void*)k_solve_tgs<false>, dim3(blocks), dim3(256), args, 0, st);
}
void pxb_launch_writeback_rows(cudaStream_t st, uint32_t capPairs, const uint32_t* counters, Rows R, const uint32_t* pairSlots, float* cForce, float4* frictions, float4* frReport, const uint2* pairBodies,
                               const float4* pos, const float4* quat) {
  k_writeback_rows<<<(capPairs + 255) / 256, 256, 0, st>>>(counters, R, pairSlots, cForce, frictions, frReport, pairBodies, pos, quat);
}
void pxb_launch_finalize_bodies_pgs(cudaStream_t st, uint32_t nDyn, const uint32_t* dynActor, float dt, float4* pos, float4* quat, float4* linVel, float4* angVel, const float4* sbLin, const float4* sbAng,
                                    const float4* sbDLin, const float4* sbDAng, const float4* sbIA, const float4* sbIB, const float4* invInertia, SleepArgs S, const uint32_t* geomFlags) {
  k_finalize_bodies_pgs<<<(nDyn + 255) / 256, 256, 0, st>>>(nDyn, dynActor, dt, pos, quat, linVel, angVel, sbLin, sbAng, sbDLin, sbDAng, sbIA, sbIB, invInertia, S, geomFlags);
}
