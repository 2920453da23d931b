// pxb_np_launch.h -- arguments and launchers of the narrowphase kernels (pxb_narrowphase.cu).  Kept apart from pxb_launch.h so that the GJK / EPA
// translation unit -- by far the slowest to compile -- does not depend on the environment-path or solver headers.
#pragma once
#include "pxb_common.cuh"

struct NpArgs {
  const uint64_t* pairKeys; const uint32_t* pairSlots; const uint32_t* nPairsP; uint32_t bitsA;
  const float4 *pos, *quat, *dims; const uint32_t* geomFlags; float contactDist, toleranceLength;
  float4 *manifolds, *cHdr, *cPts; uint2* pairBodies; uint32_t* conFlag; float* cForce; uint32_t *counters, *gjkList, *gjkQuery, *gjkFull, *gjkEpa, *boxList /* null = box-box regeneration inside k_narrowphase (environment path) */; const uint32_t* pairOrder; HullArrays hulls; TouchLists touch; FilterArgs filter /* data == null: no filter shader */;
};
void pxb_launch_narrowphase(cudaStream_t st, uint32_t capPairs, const NpArgs& A);
void pxb_launch_narrowphase_gjk(cudaStream_t st, uint32_t ctas, const NpArgs& A);
// the same work in four phases (refresh -> GJK query -> EPA -> full manifold generation), each over a compacted worklist: scenes with convex meshes
void pxb_launch_narrowphase_gjk_phases(cudaStream_t st, uint32_t ctas, const NpArgs& A);

