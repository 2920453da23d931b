// pxb_launch.h -- launchers of the kernels that live in their own translation units (pxb_narrowphase.cu, pxb_env.cu, pxb_solve.cu), called by
// the host side in pxb_engine.cu.  Splitting the library into four units keeps a kernel edit from recompiling the GJK / EPA family.
#pragma once
#include "pxb_common.cuh"
#include "pxb_np_launch.h"
#include "pxb_env.cuh"   // EnvBpArgs / EnvSolveArgs / Rows (plain structs; the kernels in there are templates instantiated by pxb_env.cu only)

cudaError_t pxb_env_set_attributes(int solveSmemMax, int bpSmemMax);
void pxb_launch_env_bp(cudaStream_t st, const EnvBpArgs& A, bool hulls, size_t smem);
void pxb_launch_env_solve(cudaStream_t st, const EnvSolveArgs& A, uint32_t threads, bool pgs, bool ext, size_t smem);

struct PrepArgs {
  const uint32_t *counters, *ordered, *conPair, *pairSlots; const uint2* pairBodies; const uint32_t* geomFlags; const float4 *cHdr, *cPts, *pos, *quat, *linVel, *sbOrigAng, *invInertia, *sbIA, *sbIB;
  float4* frictions; SolverParams P; Rows R; MaterialArgs M;
  const float4* angVel; float4* kinFtv;   // kinFtv: friction target velocities per pair, non-null in scenes with kinematic bodies
};
struct SolveArgs {
  uint32_t *counters, *partStart; uint32_t posIters, velIters; float stepDt; Rows R;
  float4 *sbLin, *sbAng, *sbDLin, *sbDAng, *sbIA, *sbIB, *sbP, *sbQ; uint32_t* bodyHasCon; uint32_t nDyn; uint32_t* dynActor; const float4* kinFtv;
};
cudaError_t pxb_solve_occupancy(int* tgsCtasPerSm, int* pgsCtasPerSm);
void pxb_launch_prep_rows(cudaStream_t st, bool pgs, uint32_t capPairs, const PrepArgs& A);
cudaError_t pxb_launch_solve(cudaStream_t st, bool pgs, int blocks, SolveArgs& A);   // ONE cooperative launch: every solver iteration
void pxb_launch_writeback_rows(cudaStream_t st, uint32_t capPairs, const uint32_t* counters, Rows R, const uint32_t* pairSlots, float* cForce, float4* frictions, float4* frReport, const uint2* pairBodies,
                               const float4* pos, const float4* quat);   // frReport: null unless contact reports are enabled (pxb_scene_enable_contact_data)
void pxb_launch_finalize_bodies_pgs(cudaStream_t st, uint32_t nDyn, const uint32_t* dynActor, float dt, float4* pos, float4* quat, float4* linVel, float4* angVel, const float4* sbLin, const float4* sbAng,
                                    const float4* sbDLin, const float4* sbDAng, const float4* sbIA, const float4* sbIB, const float4* invInertia, SleepArgs S, const uint32_t* geomFlags);
