// pxb_engine.cu -- the B200-native rigid-body step: bounds -> grid broadphase -> PCM narrowphase ->
// order-preserving graph colouring -> TGS contact solve -> integration, plus the C ABI (include/physx_b200.h).
//
// Everything between pxb_scene_simulate() and pxb_scene_fetch_results() runs on one CUDA stream with
// all counts kept in device memory: there is no host synchronisation inside a step (the reference GPU
// pipeline has three).  Stage <-> reference mapping:
//   k_bounds            a1/a2  updateTransformCacheAndBoundArrayLaunch + translateAABBsLaunch  (oracle: Gu::computeBounds)
//   radix sort + k_bp_* a3-a5  radixSort*/performIncrementalSAP/region kernels                  (oracle: ABP pair set)
//   k_pair_*            a7     removeContactManagers_Stage*, prepareLostFoundPairs_*           (found/lost + persistent slots)
//   k_narrowphase       a8-a11 sphereNphase/boxBoxNphase/convexConvex... kernels               (oracle: CPU PCM)
//   k_colour_partition  a13    PxgIncrementalPartition (host) -> here on device, first-fit in solver input order
//   k_preintegrate      a12    preIntegrationLaunchTGS
//   k_prep_rows         a14    constraintContactBlockPrePrepLaunch + contactConstraintBlockPrepareParallelLaunch[TGS]   (pxb_pgs.cuh)
//   k_solve_tgs/_pgs    a15/a16 solveBlockUnified + propagateAverageSolverBodyVelocityTGS loop  (ONE cooperative launch)
//   k_finalize          a17/a18 writebackBlocksTGS + integrateCoreParallelLaunchTGS
//   k_rd_get/k_rd_set   a19    get/setRigidDynamic* (PxDirectGPUAPI)
#include <vector>
#include <string>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include "pxb_launch.h"
#include "pxb_sort.cuh"



struct PxbScene {
  PxbSceneDesc desc; int device = 0; cudaStream_t stream = nullptr; bool abort = false; bool stepping = false;
  uint32_t nA = 0, nDyn = 0, capA = 0, capPairs = 0, bitsA = 1, nLarge = 0;
  std::vector<ActorRec> recs; std::vector<int> dynIndex; std::vector<uint32_t> dynActor; std::vector<uint32_t> largeHost;
  GridParams grid; bool gridDirty = true;
  int numSMs = 148, coopBlocksColour = 0, coopBlocksSolve = 0, coopBlocksSolvePgs = 0;
  // per actor
  float4 *pos = 0, *quat = 0, *linVel = 0, *angVel = 0, *invInertia = 0, *damp = 0, *dims = 0, *aabbMin = 0, *aabbMax = 0;
  uint32_t *geomFlags = 0, *envId = 0, *dynActorDev = 0, *largeList = 0;
  float* tight = 0;
  // solver body state (per actor)
  float4 *sbLin = 0, *sbAng = 0, *sbDLin = 0, *sbDAng = 0, *sbIA = 0, *sbIB = 0, *sbP = 0, *sbQ = 0, *sbOrigAng = 0;
  uint32_t *bodyCnt = 0, *bodyStart = 0, *bodyCursor = 0, *bodyNext = 0, *bodyHasCon = 0; unsigned long long* bodyMask = 0;   // bodyMask: the 64 dynamic colours a body's constraints hold
  // broadphase
  uint64_t *cellKey = 0, *cellKeyAlt = 0; uint32_t *cellVal = 0, *cellValAlt = 0; float4 *sMin = 0, *sMax = 0;
  uint64_t* pairKeys[2] = {0, 0}; uint32_t* pairSlots[2] = {0, 0}; uint64_t* pairKeyAlt = 0; uint32_t *pairValTmp = 0, *pairValAlt = 0;
  uint32_t* nPairsDev = 0;  // [2]
  int cur = 0;
  uint32_t *freeList = 0; uint64_t *createdKeys = 0, *deletedKeys = 0;
  float4 *manifolds = 0, *frictions = 0;
  // per pair (this frame)
  float4 *cHdr = 0, *cPts = 0; uint2* pairBodies = 0; float* cForce = 0;
  uint32_t *pairOrder = 0, *npClassCount = 0; uint8_t* npClass = 0; bool binPairs = false;   // mixed-type scenes: pairs binned by type pair before the narrowphase
  float4 *tcPos = 0, *tcQuat = 0, *s2bP = 0, *s2bQ = 0, *b2aP = 0, *b2aQ = 0, *actorPos = 0, *actorQuat = 0; bool hasLocal = false, hasCom = false;   // local poses (pxb_scene_set_local_poses)
  std::vector<float4> hS2aP, hS2aQ, hB2aP, hB2aQ;
  uint4* filterData = 0; FilterConfig filterCfg; bool hasFilter = false;
  float2* shapeOff = 0; bool hasShapeOff = false; float maxContactOffset = 0.f;   // PxShape::setContactOffset / setRestOffset per actor (pxb_scene_set_shape_offsets)   // f1: default simulation filter shader (pxb_scene_set_filter_shader / _data)
  float4* frReport = 0; uint32_t *ccIdx = 0, *ccOff = 0, *ccCount = 0, *ccTotal = 0, *actorDyn = 0; uint8_t *ccPatches = 0, *ccPoints = 0, *ccFriction = 0; float* ccForces = 0; bool contactData = false;
  uint32_t *gjkList = 0, *gjkQuery = 0, *gjkFull = 0, *gjkEpa = 0, *boxList = 0; bool boxPhases = true, boxPhasesEnv = false; bool gjkPhases = true; bool hasGjkPairs = false, anyLocks = false, anyConvex = false;
  float4 *extForce = 0, *extTorque = 0; bool forcesUsed = false;
  uint4* hullMeta = 0; float4 *hullVerts = 0, *hullPolys = 0; uint8_t *hullRefs = 0, *hullEdges = 0; uint32_t nHulls = 0; std::vector<float> hullDiam;   // cooked convex hulls (pxb_scene_set_convex_meshes)   // PxDirectGPUAPI eFORCE / eTORQUE writes pending for the next step   // a10: worklist of GJK-family pairs (filled by k_narrowphase)
  uint32_t *conFlag = 0, *conIdx = 0, *conPair = 0, *rankOfPair = 0; uint64_t *conSortKey = 0, *conSortKeyAlt = 0; uint32_t* conPairAlt = 0;
  uint64_t* orderKeys = 0; uint32_t nOrder = 0, capOrder = 0;
  uint32_t *conB0 = 0, *conB1 = 0, *conPos0 = 0, *conPos1 = 0, *conColour = 0, *conDone = 0, *bodyList = 0, *ordered = 0;
  uint32_t *partCnt = 0, *partStart = 0, *partCursor = 0, *colourTicket = 0, *prevB0 = 0, *prevB1 = 0, *prevColour = 0, *prevNCon = 0; bool colourLegacy = false, colourPrefix = true; uint32_t colourBackoffNs = 400, colourWindow = 0;
  // rows (solve order)
  float4 *ptA = 0, *ptB = 0, *ptC = 0, *frA = 0, *frB = 0, *frC = 0, *frD = 0;
  uint32_t* counters = 0; uint32_t* hostCounters = 0;  // pinned mirror
  RadixSortTemp rsTmp; uint32_t* scanSums = 0;
  uint32_t launches = 0;
  float* stage = 0; uint32_t* stageIdx = 0;
  // [parity][0 = whole step, 1 = bounds + broadphase + narrowphase, 2 = the rest]: the split pair is replayed when a stream-ordered velocity
  // write is pending on the copy stream, so that its host-to-device copy overlaps the first part (which reads poses only)
  bool useGraph = true; cudaGraphExec_t graphExec[2][3] = {{0, 0, 0}, {0, 0, 0}}; float graphDt = 0.f; uint32_t graphLaunches[2][3] = {{0, 0, 0}, {0, 0, 0}};
  cudaStream_t copyStream = nullptr; cudaEvent_t velEvent = nullptr, orderEvent = nullptr; bool velPending = false;
  uint32_t *candKeys = 0, *candCount = 0, candCap = 0; float4 *candRefMin = 0, *candRefMax = 0; bool candOn = true;   // k_env_bp temporal coherence: candidate pairs per environment + the bounds they were built from
  bool anyAggregate = false; uint32_t* aggId = 0;   // PxAggregate membership per actor (ActorRec::aggregate)
  bool anyKinematic = false; uint32_t nKin = 0; uint32_t* kinList = 0; float4 *kinP = 0, *kinQ = 0, *kinFtv = 0; uint32_t* kinHas = 0; std::vector<uint32_t> kinHost;   // kinematic bodies: actor list, pending targets (body frame), friction target velocities per pair
  bool bodyAccel = false; float4 *prevLin = 0, *prevAng = 0; float accelInvDt = 0.f;   // PxSceneFlag::eENABLE_BODY_ACCELERATIONS: velocities the last step started from
  bool profiling = false; cudaEvent_t ev[8] = {0, 0, 0, 0, 0, 0, 0, 0}; float stageMs[7] = {0, 0, 0, 0, 0, 0, 0};
  uint32_t hNPairs = 0, hNCreated = 0, hNDeleted = 0, hNCon = 0, hNPart = 0, hErr = 0;
  // environment-partitioned path (pxb_env.cuh)
  bool envEligible = false, envActive = false, envDisabled = false, everStepped = false; uint32_t ringMask = 0;
  uint32_t nEnv = 0, envMaxList = 0, envConCap = 0, envConCapForced = 0, envThreadsForced = 0, hMaxConEnv = 0, hMaxPairEnv = 0, envSolveThreads = 64;
  uint32_t *envStart = 0, *envList = 0, *actorLocal = 0, *slotColour = 0; unsigned long long* bodyBest = 0; bool relaxedPartitioning = false;
  uint32_t* actorMat = 0; float4* matTab = 0; uint32_t nMaterials = 0;   // a11 material table (pxb_scene_set_materials)
  uint32_t* touchState = 0; uint64_t *touchFound = 0, *touchLost = 0; uint32_t hNTouchFound = 0, hNTouchLost = 0;   // a7 touch found / lost events
  ExportTable* exportTab = 0; uint2* envDyn = 0; bool exportOn = false, envDynContiguous = false;   // fused state export (pxb_scene_set_state_export)
  float sleepThreshold = 0.f; float* wake = 0; float4 *accLin = 0, *accAng = 0; uint32_t *asleep = 0, *nInter = 0, *islandLabel = 0, *islandAwake = 0; int coopBlocksSleep = 0; uint2* envSeg[2] = {0, 0}; unsigned long long* envTiming = 0;
};

#define ACTOR_REMOVED 0x80000000u   // host-side bit of ActorRec::flags: the actor was taken out by pxb_scene_remove_actors (its index stays)
static thread_local std::string g_err;
// Every entry point runs with the scene's device current and restores the caller's on return: scenes on different GPUs can be driven from
// one thread, and torch / other libraries may change the current device between calls.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != device) { if (cudaSetDevice(device) == cudaSuccess) prev = cur; } else if (cur < 0) cudaGetLastError(); }
  explicit DeviceGuard(const PxbScene* s) : DeviceGuard(s ? s->device : 0) {}
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static LocalPoses local_poses(const PxbScene* s) { LocalPoses L; L.s2bP = s->hasLocal ? s->s2bP : nullptr; L.s2bQ = s->s2bQ; L.tcPos = s->tcPos; L.tcQuat = s->tcQuat; return L; }
static MaterialArgs material_args(const PxbScene* s) { MaterialArgs M; M.actorMat = s->actorMat; M.matTab = s->nMaterials ? s->matTab : nullptr; M.shapeOff = s->hasShapeOff ? s->shapeOff : nullptr; return M; }
static TouchLists touch_lists(const PxbScene* s) { TouchLists T; T.state = s->touchState; T.found = s->touchFound; T.lost = s->touchLost; return T; }
static HullArrays hull_arrays(const PxbScene* s) { HullArrays H; H.meta = s->hullMeta; H.verts = s->hullVerts; H.polys = s->hullPolys; H.refs = s->hullRefs; H.edges = s->hullEdges; return H; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (s) s->abort = true; return fail(PXB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)

// ---------------------------------------------------------------------------------------------
// kernels
// a1/a2: world AABB of every actor (tight, then inflated by the contact offset, BpBroadPhaseABP.cpp:1187-1197) + grid cell key.
__global__ void k_bounds(uint32_t nA, const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ dims,
                         const uint32_t* __restrict__ geomFlags, const uint32_t* __restrict__ envId, float contactOffset, float* __restrict__ tight,
                         int externalTight, float4* __restrict__ aabbMin, float4* __restrict__ aabbMax, GridParams g, uint32_t envCount,
                         uint64_t* __restrict__ cellKey, uint32_t* __restrict__ cellVal, HullArrays hulls, LocalPoses L, const float2* __restrict__ shapeOff) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nA) return;
  const uint32_t gf = geomFlags[a];
  if (gf & 0x400u) {   // removed actor (pxb_scene_remove_actors): empty bounds, in no grid cell -- its pairs are reported deleted by this step's lifecycle
    aabbMin[a] = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, __uint_as_float(NONE32)); aabbMax[a] = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, __uint_as_float(gf));
    cellKey[a] = ~0ull; cellVal[a] = a;
    return;
  }
  float mn[3], mx[3];
  const xf shape = shape_world_pose(L, a, pos[a], quat[a]);   // (also refreshes the transform cache the narrowphase reads when the scene has local poses)
  if (externalTight) { for (int k = 0; k < 3; ++k) { mn[k] = tight[a * 6 + k]; mx[k] = tight[a * 6 + 3 + k]; } }
  else {
    tight_bounds(gf & 0xff, shape.p, shape.q, dims[a], mn, mx, &hulls);
    for (int k = 0; k < 3; ++k) { tight[a * 6 + k] = mn[k]; tight[a * 6 + 3 + k] = mx[k]; }
  }
  const float co = shapeOff ? shapeOff[a].x : contactOffset;   // every bound is inflated by its own shape's contact offset
  const uint32_t env = envId[a];
  aabbMin[a] = make_float4(mn[0] - co, mn[1] - co, mn[2] - co, __uint_as_float(env));
  aabbMax[a] = make_float4(mx[0] + co, mx[1] + co, mx[2] + co, __uint_as_float(gf));
  uint64_t key = ~0ull;
  if (!(gf & 0x200u)) {  // not a global/large object: grid key (env, cz, cy, cx), cx least significant
    int cx = (int)floorf((mn[0] - co - g.ox) * g.invCell), cy = (int)floorf((mn[1] - co - g.oy) * g.invCell), cz = (int)floorf((mn[2] - co - g.oz) * g.invCell);
    cx = max(0, min(g.nx - 1, cx)); cy = max(0, min(g.ny - 1, cy)); cz = max(0, min(g.nz - 1, cz));
    const uint64_t e = (env == NONE32) ? 0ull : (uint64_t)min(env, envCount - 1);
    key = ((e * (uint64_t)g.nz + (uint64_t)cz) * (uint64_t)g.ny + (uint64_t)cy) * (uint64_t)g.nx + (uint64_t)cx;
  }
  cellKey[a] = key; cellVal[a] = a;
}

__global__ void k_bp_gather(uint32_t nA, const uint32_t* __restrict__ sortedActor, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                            float4* __restrict__ sMin, float4* __restrict__ sMax) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA) return;
  const uint32_t a = sortedActor[i];
  sMin[i] = aabbMin[a]; sMax[i] = aabbMax[a];
}

// a4/a5: every grid object looks "forward" in (env, cz, cy, cx) order: its own row from itself on, and the
// 4 following neighbour rows; because the cell edge is >= every object extent, overlapping objects differ
// by at most one cell per axis, so each unordered pair is visited exactly once.
// Pair filters.  EngineFilter: the scene engine's (bp_test: at least one dynamic actor, environment ids).  GroupFilter: the reference's
// Bp::FilterGroup semantics for the standalone broadphase object behind Bp::BroadPhase (groupFiltering, BpFiltering.h:99-114: equal groups never
// pair, otherwise the 7 x 7 BpFilter table indexed by the 3-bit type in the group's low bits -- passed as a 49-bit mask; environment ids as in
// broadphase.cu:62-80): amin.w = environment id, amax.w = filter group.
struct EngineFilter {
  const uint32_t* agg;   // per-actor PxAggregate id (NULL: the scene has none); members of one aggregate without self collisions never pair
  __device__ __forceinline__ bool operator()(const float4& amin, const float4& amax, const float4& bmin, const float4& bmax) const { return bp_test(amin, amax, bmin, bmax); }
  __device__ __forceinline__ bool pairOk(uint32_t a, uint32_t b) const { if (!agg) return true; const uint32_t ga = agg[a]; return !(ga && ga == agg[b] && !(ga & 0x80000000u)); }
  __device__ __forceinline__ bool isLarge(const float4& amax, uint32_t) const { return (__float_as_uint(amax.w) & 0x200u) != 0; }
};
struct GroupFilter {
  unsigned long long lut; const uint32_t* large;
  __device__ __forceinline__ bool operator()(const float4& amin, const float4& amax, const float4& bmin, const float4& bmax) const {
    if (amin.x > bmax.x || bmin.x > amax.x || amin.y > bmax.y || bmin.y > amax.y || amin.z > bmax.z || bmin.z > amax.z) return false;
    const uint32_t ga = __float_as_uint(amax.w), gb = __float_as_uint(bmax.w);
    if (ga == gb || !((lut >> ((ga & 7u) * 7u + (gb & 7u))) & 1ull)) return false;
    const uint32_t ea = __float_as_uint(amin.w), eb = __float_as_uint(bmin.w);
    return ea == NONE32 || eb == NONE32 || ea == eb;
  }
  __device__ __forceinline__ bool isLarge(const float4&, uint32_t a) const { return large[a] != 0u; }
  __device__ __forceinline__ bool pairOk(uint32_t, uint32_t) const { return true; }
};
template <class F>
__global__ void k_bp_pairs(uint32_t nA, const uint64_t* __restrict__ key, const uint32_t* __restrict__ sortedActor, const float4* __restrict__ sMin,
                           const float4* __restrict__ sMax, GridParams g, uint32_t bitsA, uint64_t* __restrict__ pairKeys, uint32_t* __restrict__ counters, uint32_t cap, const F filter) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA) return;
  const uint64_t k = key[i];
  if (k == ~0ull) return;
  const uint32_t a = sortedActor[i];
  const float4 amin = sMin[i], amax = sMax[i];
  const uint64_t nx = (uint64_t)g.nx, ny = (uint64_t)g.ny, nz = (uint64_t)g.nz;
  const int cx = (int)(k % nx); const uint64_t r1 = k / nx; const int cy = (int)(r1 % ny); const uint64_t r2 = r1 / ny; const int cz = (int)(r2 % nz); const uint64_t e = r2 / nz;
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
  {  // own row, forward
    const uint64_t last = r1 * nx + (uint64_t)x1;
    for (uint32_t j = i + 1; j < nA && key[j] <= last; ++j)
      if (filter(amin, amax, sMin[j], sMax[j]) && filter.pairOk(a, sortedActor[j])) bp_emit(a, sortedActor[j], bitsA, pairKeys, &counters[C_NPAIRS_NEW], cap, &counters[C_ERROR]);
  }
  const int dzs[4] = {0, 1, 1, 1}, dys[4] = {1, -1, 0, 1};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = cy + dys[r], zz = cz + dzs[r];
    if (yy < 0 || yy >= g.ny || zz >= g.nz) continue;
    const uint64_t rowBase = ((e * nz + (uint64_t)zz) * ny + (uint64_t)yy) * nx;
    const uint64_t first = rowBase + (uint64_t)x0, last = rowBase + (uint64_t)x1;
    for (uint32_t j = lower_bound_u64(key, nA, first); j < nA && key[j] <= last; ++j)
      if (filter(amin, amax, sMin[j], sMax[j]) && filter.pairOk(a, sortedActor[j])) bp_emit(a, sortedActor[j], bitsA, pairKeys, &counters[C_NPAIRS_NEW], cap, &counters[C_ERROR]);
  }
}
// global / oversize objects (planes, shapes larger than a cell, env-less actors in env scenes) against everything
template <class F>
__global__ void k_bp_large(uint32_t nA, uint32_t nLarge, const uint32_t* __restrict__ largeList, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax,
                           uint32_t bitsA, uint64_t* __restrict__ pairKeys, uint32_t* __restrict__ counters, uint32_t cap, const F filter) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nA) return;
  const float4 amin = aabbMin[a], amax = aabbMax[a];
  const bool aLarge = filter.isLarge(amax, a);
  for (uint32_t l = 0; l < nLarge; ++l) {
    const uint32_t b = largeList[l];
    if (b == a || (aLarge && a > b)) continue;
    if (filter(amin, amax, aabbMin[b], aabbMax[b]) && filter.pairOk(a, b)) bp_emit(a, b, bitsA, pairKeys, &counters[C_NPAIRS_NEW], cap, &counters[C_ERROR]);
  }
}
__global__ void k_clamp_count(uint32_t* __restrict__ counters, uint32_t cap, uint32_t* __restrict__ nPairsCur) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { const uint32_t n = min(counters[C_NPAIRS_NEW], cap); counters[C_NPAIRS_NEW] = n; *nPairsCur = n; }
}

// a7: pair lifecycle.  Lost pairs give their persistent slot back, surviving pairs keep theirs, new pairs
// take one from the free list and start with an empty manifold (PersistentContactManifold::initialize).
__global__ void k_pair_lost(const uint64_t* __restrict__ oldKeys, const uint32_t* __restrict__ oldSlots, const uint32_t* __restrict__ nOldP,
                            const uint64_t* __restrict__ newKeys, const uint32_t* __restrict__ nNewP, uint32_t* __restrict__ freeList, uint32_t ringMask,
                            uint64_t* __restrict__ deletedKeys, uint32_t* __restrict__ counters, const TouchLists touch) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nOld = *nOldP, nNew = *nNewP;
  if (j >= nOld) return;
  const uint64_t k = oldKeys[j];
  const uint32_t p = lower_bound_u64(newKeys, nNew, k);
  if (p < nNew && newKeys[p] == k) return;
  freeList[atomicAdd(&counters[C_FREE_TAIL], 1u) & ringMask] = oldSlots[j];
  deletedKeys[atomicAdd(&counters[C_NDELETED], 1u)] = k;
  touch_event(touch, counters, oldSlots[j], k, false);   // a pair that leaves the broadphase while touching loses its touch
}
__global__ void k_pair_found(const uint64_t* __restrict__ oldKeys, const uint32_t* __restrict__ oldSlots, const uint32_t* __restrict__ nOldP,
                             const uint64_t* __restrict__ newKeys, uint32_t* __restrict__ newSlots, const uint32_t* __restrict__ nNewP,
                             const uint32_t* __restrict__ freeList, uint32_t ringMask, uint64_t* __restrict__ createdKeys, uint32_t* __restrict__ counters,
                             float4* __restrict__ manifolds, float4* __restrict__ frictions, uint32_t* __restrict__ touchState) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nOld = *nOldP, nNew = *nNewP;
  if (i >= nNew) return;
  const uint64_t k = newKeys[i];
  const uint32_t p = lower_bound_u64(oldKeys, nOld, k);
  if (p < nOld && oldKeys[p] == k) { newSlots[i] = oldSlots[p]; return; }
  const uint32_t h = atomicAdd(&counters[C_FREE_HEAD], 1u);   // free slots live in a ring: pops advance the head, pushes the tail
  uint32_t slot = 0;
  if ((int32_t)(counters[C_FREE_TAIL] - h) <= 0) atomicOr(&counters[C_ERROR], (uint32_t)E_PAIR_OVERFLOW); else slot = freeList[h & ringMask];
  newSlots[i] = slot;
  createdKeys[atomicAdd(&counters[C_NCREATED], 1u)] = k;
  float4* m = manifolds + (size_t)slot * PXB_MANIFOLD_F4;
  m[0] = make_float4(__int_as_float(0), FLT_MAX, FLT_MAX, FLT_MAX); m[1] = make_float4(0, 0, 0, 1); m[2] = make_float4(0, 0, 0, 1); m[3] = make_float4(0, 0, 0, 1); m[14] = make_float4(0, 0, 0, 0);
  float4* f = frictions + (size_t)slot * PXB_FRICTION_F4;
  f[0] = make_float4(0, 0, 0, __int_as_float(0)); f[1] = make_float4(0, 0, 0, __int_as_float(0)); f[2] = make_float4(0, 0, 0, __int_as_float(0));
  touchState[slot] = 0u;
}

// Scenes that mix geometry types: pairs are binned by (type0, type1) before the narrowphase so that a warp runs ONE contact function
// (in pair-key order the types alternate at random: ncu measured 4.5 of 32 threads active per instruction on BASELINE config 3).
// Counting sort in two passes (histogram, then scatter with warp-aggregated cursors); the order inside a bin is arbitrary, every pair
// writes only its own outputs, so the result does not depend on it.
#define NP_CLASSES 75   // 6 x 6 geometry types x {no contacts last frame, contacts last frame} + 1 bin for dropped keys (+ padding)
// The second key bit -- did the pair's persistent manifold hold contacts last frame -- separates the (cheap) separated pairs from the touching
// ones that run the full contact generation, so warps are uniform in work as well as in code path.
__device__ __forceinline__ uint32_t np_pair_class(uint64_t key, uint32_t bitsA, const uint32_t* __restrict__ geomFlags, uint32_t slot, const float4* __restrict__ manifolds) {
  if (key == ~0ull) return NP_CLASSES - 1;
  const uint32_t lo = (uint32_t)(key >> bitsA), hi = (uint32_t)(key & ((1ull << bitsA) - 1ull));
  const uint32_t a = geomFlags[lo] & 0xff, b = geomFlags[hi] & 0xff;
  const uint32_t touching = __float_as_int(manifolds[(size_t)slot * PXB_MANIFOLD_F4].x) > 0 ? 1u : 0u;
  return 2u * (a < b ? a * 6 + b : b * 6 + a) + touching;
}
__global__ void k_np_class_count(const uint64_t* __restrict__ pairKeys, const uint32_t* __restrict__ nPairsP, uint32_t bitsA, const uint32_t* __restrict__ geomFlags, uint8_t* __restrict__ cls,
                                 uint32_t* __restrict__ classCount, const uint32_t* __restrict__ pairSlots, const float4* __restrict__ manifolds) {
  __shared__ uint32_t hist[NP_CLASSES];
  if (threadIdx.x < NP_CLASSES) hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *nPairsP) { const uint32_t c = np_pair_class(pairKeys[i], bitsA, geomFlags, pairSlots[i], manifolds); cls[i] = (uint8_t)c; atomicAdd(&hist[c], 1u); }
  __syncthreads();
  if (threadIdx.x < NP_CLASSES && hist[threadIdx.x]) atomicAdd(&classCount[threadIdx.x], hist[threadIdx.x]);
}
__global__ void k_np_class_scatter(const uint32_t* __restrict__ nPairsP, const uint8_t* __restrict__ cls, const uint32_t* __restrict__ classCount, uint32_t* __restrict__ classCursor,
                                   uint32_t* __restrict__ pairOrder) {
  __shared__ uint32_t base[NP_CLASSES];
  if (threadIdx.x == 0) { uint32_t acc = 0; for (int c = 0; c < NP_CLASSES; ++c) { base[c] = acc; acc += classCount[c]; } }
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < *nPairsP;
  const uint32_t c = valid ? cls[i] : 0xffu;
  const uint32_t active = __activemask();
  const uint32_t peers = __match_any_sync(active, c);
  if (!valid) return;
  const uint32_t lane = threadIdx.x & 31, leader = __ffs(peers) - 1, rank = __popc(peers & ((1u << lane) - 1u));
  uint32_t start = 0;
  if (lane == leader) start = atomicAdd(&classCursor[c], (uint32_t)__popc(peers));
  start = __shfl_sync(peers, start, leader);
  pairOrder[base[c] + start + rank] = i;
}

__global__ void k_compact(const uint32_t* __restrict__ nPairsP, const uint32_t* __restrict__ conFlag, const uint32_t* __restrict__ conIdx, uint32_t* __restrict__ conPair,
                          const uint32_t* __restrict__ pairSlots, const float4* __restrict__ cHdr, float4* __restrict__ frictions) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *nPairsP) return;
  if (conFlag[i]) conPair[conIdx[i]] = i;
  if (__float_as_int(cHdr[i].w) == 0) frictions[(size_t)pairSlots[i] * PXB_FRICTION_F4 + 2].w = __int_as_float(0);  // no contacts: friction patch state is dropped
}
// Pairs the host's island manager still lists stay in the constraint list even without contacts this frame: they are
// empty constraints that still take a colour (lost-touch edges leave the island only after the solve).
__global__ void k_flag_ordered(const uint32_t* __restrict__ nPairsP, const uint32_t* __restrict__ rankOfPair, uint32_t* __restrict__ conFlag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *nPairsP && rankOfPair[i] < 0x80000000u) conFlag[i] = 1u;
}
// host-provided solver input order -> per-pair rank
__global__ void k_rank_init(const uint32_t* __restrict__ nPairsP, uint32_t* __restrict__ rankOfPair) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *nPairsP) rankOfPair[i] = 0x80000000u + i;
}
__global__ void k_rank_map(uint32_t nOrder, const uint64_t* __restrict__ orderKeys, const uint64_t* __restrict__ pairKeys, const uint32_t* __restrict__ nPairsP, uint32_t* __restrict__ rankOfPair) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nOrder) return;
  const uint32_t n = *nPairsP; const uint64_t key = orderKeys[k];
  const uint32_t p = lower_bound_u64(pairKeys, n, key);
  if (p < n && pairKeys[p] == key) rankOfPair[p] = k;
}
__global__ void k_con_sortkeys(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ conPair, const uint32_t* __restrict__ rankOfPair, uint64_t* __restrict__ keys) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < counters[C_NCON]) keys[c] = rankOfPair[conPair[c]];
}

// a13 (1/3): constraint -> bodies, per-body constraint counts
__global__ void k_con_bodies(const uint32_t* __restrict__ counters_, uint32_t* __restrict__ counters, const uint32_t* __restrict__ conPair, const uint2* __restrict__ pairBodies,
                             const uint32_t* __restrict__ geomFlags, uint32_t* __restrict__ conB0, uint32_t* __restrict__ conB1, uint32_t* __restrict__ conDone, uint32_t* __restrict__ bodyCnt) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= counters_[C_NCON]) return;
  const uint2 b = pairBodies[conPair[c]];
  const bool dyn1 = gf_dynamic(geomFlags[b.y]);   // static and kinematic bodies sit outside the partitioned body range
  conB0[c] = b.x; conB1[c] = dyn1 ? b.y : NONE32; conDone[c] = dyn1 ? 0u : 1u;
  atomicAdd(&bodyCnt[b.x], 1u);
  if (dyn1) { atomicAdd(&bodyCnt[b.y], 1u); atomicAdd(&counters[C_REMAINING], 1u); }
}
__global__ void k_con_fill(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ conB0, const uint32_t* __restrict__ conB1, const uint32_t* __restrict__ bodyStart,
                           uint32_t* __restrict__ bodyCursor, uint32_t* __restrict__ bodyList) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= counters[C_NCON]) return;
  const uint32_t a = conB0[c], b = conB1[c];
  bodyList[bodyStart[a] + atomicAdd(&bodyCursor[a], 1u)] = c;
  if (b != NONE32) bodyList[bodyStart[b] + atomicAdd(&bodyCursor[b], 1u)] = c;
}
// a13 (2/3): order each body's list by solver input order; position of each constraint among the body's
// dynamic (resp. static) constraints
__global__ void k_body_lists(uint32_t nA, const uint32_t* __restrict__ bodyStart, const uint32_t* __restrict__ bodyCnt, uint32_t* __restrict__ bodyList,
                             const uint32_t* __restrict__ conB0, const uint32_t* __restrict__ conB1, uint32_t* __restrict__ conPos0, uint32_t* __restrict__ conPos1,
                             uint32_t* __restrict__ bodyNext, unsigned long long* __restrict__ bodyMask, uint32_t* __restrict__ bodyHasCon) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nA) return;
  const uint32_t n = bodyCnt[a]; uint32_t* l = bodyList + bodyStart[a];
  for (uint32_t i = 1; i < n; ++i) { const uint32_t v = l[i]; uint32_t j = i; while (j > 0 && l[j - 1] > v) { l[j] = l[j - 1]; --j; } l[j] = v; }
  uint32_t dpos = 0, spos = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t c = l[i]; const bool isStatic = conB1[c] == NONE32;
    if (conB0[c] == a) conPos0[c] = isStatic ? spos : dpos; else conPos1[c] = dpos;
    if (isStatic) spos++; else dpos++;
  }
  bodyNext[a] = 0; bodyMask[a] = 0ull; bodyHasCon[a] = n > 0 ? 1u : 0u;
}
// a13 prefix reuse.  First fit is sequential in solver input order, so a constraint's colour depends only on the constraints BEFORE it: when this
// frame's constraint list agrees with last frame's up to index F, the first F colours are last frame's.  k_colour_prefix_find computes F (first
// difference), k_colour_prefix_apply re-installs those colours and the body masks they imply, and the dataflow below only runs for c >= F.  A pile
// at rest keeps its list, so its colouring costs two light passes (BASELINE config 4: 3.6 ms -> ~0.1 ms once settled).
__global__ void k_colour_prefix_find(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ conB0, const uint32_t* __restrict__ conB1, const uint32_t* __restrict__ prevB0,
                                     const uint32_t* __restrict__ prevB1, const uint32_t* __restrict__ prevN, uint32_t* __restrict__ ticket) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = min(counters[C_NCON], *prevN);
  if (c == 0) atomicMin(&ticket[2], n);
  if (c >= n) return;
  if (conB0[c] != prevB0[c] || conB1[c] != prevB1[c]) atomicMin(&ticket[2], c);
}
__global__ void k_colour_prefix_apply(const uint32_t* __restrict__ conB0, const uint32_t* __restrict__ conB1, const uint32_t* __restrict__ prevColour, uint32_t* __restrict__ conColour,
                                      unsigned long long* __restrict__ bodyMask, const uint32_t* __restrict__ ticket) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ticket[2]) return;
  const uint32_t b = conB1[c];
  if (b == NONE32) return;
  const uint32_t col = prevColour[c];
  conColour[c] = col;
  atomicOr(&bodyMask[conB0[c]], 1ull << col); atomicOr(&bodyMask[b], 1ull << col);
}
__global__ void k_colour_remember(const uint32_t* __restrict__ counters, uint32_t* __restrict__ prevN) { if (threadIdx.x == 0 && blockIdx.x == 0) *prevN = (counters[C_ERROR] & E_COLOUR_OVERFLOW) ? 0u : counters[C_NCON]; }
// a13 (3/3, exact): the reference's sequential first-fit (classifyConstraintDesc, DyConstraintPartition.cpp:475-568) as a DATAFLOW over the
// constraint list instead of grid-wide rounds.  One thread per constraint; a dynamic constraint may take its colour once every earlier
// constraint of both its bodies has one.  Colours on a body are distinct, so popcount(bodyMask[body]) IS the number of the body's coloured
// dynamic constraints: a single 64-bit word per body carries readiness and data, and one volatile load per body is the whole handshake (no
// fence, no atomic: the only thread allowed to write a body's mask is the one whose turn it is).  Blocks take a ticket for their position in
// the list, so every constraint a thread waits for belongs to a block that has already started -- the wait cannot deadlock.  The length
// of the longest dependency chain (2 091 links on BASELINE config 4's 2.4 M constraints) times one L2 round trip bounds the time, instead
// of chain length x grid barrier + list sweep (405 ms -> a few ms on config 4).
__global__ void __launch_bounds__(128) k_colour_firstfit(uint32_t* __restrict__ counters, const uint32_t* __restrict__ conB0, const uint32_t* __restrict__ conB1, const uint32_t* __restrict__ conPos0,
                                                         const uint32_t* __restrict__ conPos1, uint32_t* __restrict__ conColour, unsigned long long* __restrict__ bodyMask, uint32_t* __restrict__ ticket,
                                                         uint32_t backoffNs) {
  __shared__ uint32_t sBlock;
  if (threadIdx.x == 0) sBlock = atomicAdd(&ticket[0], 1u);
  __syncthreads();
  const uint32_t c = sBlock * blockDim.x + threadIdx.x;
  if (c >= counters[C_NCON] || c < ticket[2]) return;   // ticket[2]: constraints before it kept last frame's colours (k_colour_prefix_*)
  const uint32_t b = conB1[c];
  if (b == NONE32) return;   // static contacts take their partition after the dynamic colours are known (k_colour_partition's ordering pass)
  const uint32_t a = conB0[c], pa = conPos0[c], pb = conPos1[c];
  uint32_t spins = 0;
  for (;;) {   // the colour is published INSIDE the loop body: no thread of the warp ever waits for code after a divergent loop exit
    const unsigned long long ma = ld_volatile64(&bodyMask[a]);
    if ((uint32_t)__popcll(ma) == pa) {
      const unsigned long long mb = ld_volatile64(&bodyMask[b]);
      if ((uint32_t)__popcll(mb) == pb) {
        const unsigned long long comb = ~ma & ~mb;   // first fit over 64 colours = the reference's second 32-colour round for the constraints that overflow the first
        if (!comb) { atomicOr(&counters[C_ERROR], (uint32_t)E_COLOUR_OVERFLOW); st_volatile(&ticket[1], 1u); conColour[c] = 63; return; }   // reported by fetchResults; everybody stops waiting
        const uint32_t col = __ffsll((long long)comb) - 1;
        conColour[c] = col;
        st_volatile64(&bodyMask[a], ma | (1ull << col)); st_volatile64(&bodyMask[b], mb | (1ull << col));
        return;
      }
    }
    if ((++spins & 63u) == 0 && ld_volatile(&ticket[1])) return;
    if (backoffNs && spins > 2) __nanosleep(backoffNs);
  }
}
// a13 (3/3): first-fit colouring in solver input order, identical to the sequential
// classifyConstraintDesc (DyConstraintPartition.cpp:475-568): a constraint takes the lowest colour free on
// both bodies once every earlier constraint of both bodies is coloured.  Static contacts of a body go to
// partitions maxDynamicColour(body)+k (:203-262).  Then partition-major ordering of the constraints.
__global__ void __launch_bounds__(256) k_colour_partition(uint32_t* __restrict__ counters, const uint32_t* __restrict__ conB0, const uint32_t* __restrict__ conB1,
                                   const uint32_t* __restrict__ conPos0, const uint32_t* __restrict__ conPos1, uint32_t* __restrict__ conColour, uint32_t* __restrict__ conDone,
                                   uint32_t* __restrict__ bodyNext, unsigned long long* __restrict__ bodyMask, uint32_t* __restrict__ partCnt, uint32_t* __restrict__ partStart,
                                   uint32_t* __restrict__ partCursor, uint32_t* __restrict__ ordered, unsigned long long* __restrict__ bodyBest, int relaxed) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t nCon = counters[C_NCON];
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  for (uint32_t p = gtid; p < MAX_PARTITIONS + 1; p += gsize) { partCnt[p] = 0; }
  const uint32_t perCta = (nCon + gridDim.x - 1) / gridDim.x;
  const uint32_t cb = min(nCon, blockIdx.x * perCta), ce = min(nCon, cb + perCta);
  if (relaxed == 2) {   // colours already assigned by k_colour_firstfit
  } else if (relaxed) {
    // PXB_FLAG_RELAXED_PARTITIONING: Jones-Plassmann style rounds.  Every uncoloured constraint bids its priority (a fixed
    // bijective hash of its index) on both bodies; the constraint that holds the highest bid on BOTH bodies takes the lowest
    // colour free on both.  Winners of a round are body-disjoint, the result is a valid partitioning that depends only on the
    // constraint list (deterministic; oracle/pxb_oracle.c runs the same rounds), but it is not the reference's first-fit.
    for (uint32_t round = 1;; ++round) {
      const uint32_t remaining = ld_volatile(&counters[C_REMAINING]);
      grid.sync();
      if (remaining == 0) break;
      for (uint32_t c = cb + threadIdx.x; c < ce; c += blockDim.x) {
        if (conDone[c]) continue;
        const unsigned long long bid = ((unsigned long long)round << 32) | (c * 2654435761u + 1u);
        atomicMax(&bodyBest[conB0[c]], bid); atomicMax(&bodyBest[conB1[c]], bid);
      }
      grid.sync();
      uint32_t won = 0;
      for (uint32_t c = cb + threadIdx.x; c < ce; c += blockDim.x) {
        if (conDone[c]) continue;
        const unsigned long long bid = ((unsigned long long)round << 32) | (c * 2654435761u + 1u);
        const uint32_t a = conB0[c], b = conB1[c];
        if (bodyBest[a] != bid || bodyBest[b] != bid) continue;
        const unsigned long long ma = bodyMask[a], mb = bodyMask[b]; const unsigned long long comb = ~ma & ~mb;
        uint32_t col = 63;
        if (comb) col = __ffsll((long long)comb) - 1; else atomicOr(&counters[C_ERROR], (uint32_t)E_COLOUR_OVERFLOW);
        conColour[c] = col; conDone[c] = 1u; bodyMask[a] = ma | (1ull << col); bodyMask[b] = mb | (1ull << col);
        ++won;
      }
      if (won) atomicSub(&counters[C_REMAINING], won);
      grid.sync();
    }
  } else
  for (;;) {
    // every thread samples the counter between two grid barriers, so all of them take the same branch
    const uint32_t remaining = ld_volatile(&counters[C_REMAINING]);
    grid.sync();
    if (remaining == 0) break;
    // Each CTA owns a contiguous range of constraints (solver input order keeps an island's constraints
    // together), and runs the dependency fixed point locally with CTA barriers; only chains that cross a
    // CTA boundary need another grid-wide round.
    for (;;) {
      int progress = 0;
      for (uint32_t c = cb + threadIdx.x; c < ce; c += blockDim.x) {
        if (conDone[c]) continue;
        const uint32_t a = conB0[c], b = conB1[c];
        if (ld_volatile(&bodyNext[a]) != conPos0[c] || ld_volatile(&bodyNext[b]) != conPos1[c]) continue;
        __threadfence();
        const unsigned long long ma = ld_volatile64(&bodyMask[a]), mb = ld_volatile64(&bodyMask[b]);
        const unsigned long long comb = ~ma & ~mb;   // first fit over 64 colours = the reference's second 32-colour round for the constraints that overflow the first
        uint32_t col = 63;
        if (comb) col = __ffsll((long long)comb) - 1; else atomicOr(&counters[C_ERROR], (uint32_t)E_COLOUR_OVERFLOW);
        conColour[c] = col; conDone[c] = 1u;
        st_volatile64(&bodyMask[a], ma | (1ull << col)); st_volatile64(&bodyMask[b], mb | (1ull << col));
        __threadfence();
        st_volatile(&bodyNext[a], conPos0[c] + 1); st_volatile(&bodyNext[b], conPos1[c] + 1);
        atomicSub(&counters[C_REMAINING], 1u);
        progress = 1;
      }
      if (!__syncthreads_or(progress)) break;
    }
    grid.sync();
  }
  grid.sync();
  // partition-major ordering: per-CTA histogram in shared memory, one global atomic per (CTA, partition)
  __shared__ uint32_t shCnt[MAX_PARTITIONS];
  __shared__ uint32_t shBase[MAX_PARTITIONS];
  for (uint32_t p = threadIdx.x; p < MAX_PARTITIONS; p += blockDim.x) shCnt[p] = 0;
  __syncthreads();
  for (uint32_t c = cb + threadIdx.x; c < ce; c += blockDim.x) {
    uint32_t col;
    if (conB1[c] == NONE32) { const unsigned long long m = bodyMask[conB0[c]]; col = (m ? 64u - __clzll((long long)m) : 0u) + conPos0[c]; }
    else col = conColour[c];
    if (col >= MAX_PARTITIONS) { col = MAX_PARTITIONS - 1; atomicOr(&counters[C_ERROR], (uint32_t)E_PARTITION_OVERFLOW); }
    conColour[c] = col;
    atomicAdd(&shCnt[col], 1u);
  }
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < MAX_PARTITIONS; p += blockDim.x) if (shCnt[p]) atomicAdd(&partCnt[p], shCnt[p]);
  grid.sync();
  if (gtid == 0) {
    uint32_t s = 0, np = 0;
    for (uint32_t p = 0; p < MAX_PARTITIONS; ++p) { const uint32_t c = partCnt[p]; partStart[p] = s; partCursor[p] = s; s += c; if (c) np = p + 1; }
    partStart[MAX_PARTITIONS] = s; counters[C_NPART] = np;
  }
  grid.sync();
  for (uint32_t p = threadIdx.x; p < MAX_PARTITIONS; p += blockDim.x) { const uint32_t n = shCnt[p]; shBase[p] = n ? atomicAdd(&partCursor[p], n) : 0u; shCnt[p] = 0; }
  __syncthreads();
  for (uint32_t c = cb + threadIdx.x; c < ce; c += blockDim.x) { const uint32_t col = conColour[c]; ordered[shBase[col] + atomicAdd(&shCnt[col], 1u)] = c; }
}

// ---------------------------------------------------------------------------------------------
// islands = connected components over touching dynamic-dynamic pairs (min-label hooking + pointer jumping, cooperative)
__global__ void __launch_bounds__(256) k_sleep_islands(uint32_t nA, const uint32_t* __restrict__ nPairsP, const uint2* __restrict__ pairBodies, const float4* __restrict__ cHdr,
                                                       const uint32_t* __restrict__ geomFlags, SleepArgs S, uint32_t* __restrict__ label, uint32_t* __restrict__ islandAwake,
                                                       uint32_t* __restrict__ counters, float4* __restrict__ linVel, float4* __restrict__ angVel, uint32_t* __restrict__ conFlag,
                                                       const uint64_t* __restrict__ deletedKeys, uint32_t bitsA) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  const uint32_t nPairs = *nPairsP;
  for (uint32_t a = gtid; a < nA; a += gsize) { label[a] = a; S.nInter[a] = 0; islandAwake[a] = 0; }
  grid.sync();
  for (uint32_t i = gtid; i < nPairs; i += gsize) {   // numCountedInteractions: every broadphase pair of the body (ScBodySim.h:137)
    const uint2 b = pairBodies[i];
    if (geomFlags[b.x] & 0x100u) atomicAdd(&S.nInter[b.x], 1u);
    if (geomFlags[b.y] & 0x100u) atomicAdd(&S.nInter[b.y], 1u);
  }
  for (uint32_t i = gtid; i < counters[C_NDELETED]; i += gsize) {   // pairs lost this frame still count: their interaction is destroyed after the solver
    const uint64_t key = deletedKeys[i]; const uint32_t lo = (uint32_t)(key >> bitsA), hi = (uint32_t)(key & ((1ull << bitsA) - 1ull));
    if (geomFlags[lo] & 0x100u) atomicAdd(&S.nInter[lo], 1u);
    if (geomFlags[hi] & 0x100u) atomicAdd(&S.nInter[hi], 1u);
  }
  for (;;) {
    if (gtid == 0) counters[C_REMAINING] = 0;
    grid.sync();
    uint32_t changed = 0;
    for (uint32_t i = gtid; i < nPairs; i += gsize) {
      if (__float_as_int(cHdr[i].w) <= 0) continue;
      const uint2 b = pairBodies[i];
      if (!(geomFlags[b.x] & 0x100u) || !(geomFlags[b.y] & 0x100u)) continue;
      const uint32_t la = label[b.x], lb = label[b.y];
      if (la == lb) continue;
      const uint32_t mn = min(la, lb), mx = max(la, lb);
      atomicMin(&label[mx], mn); atomicMin(&label[b.x], mn); atomicMin(&label[b.y], mn);
      changed = 1;
    }
    if (changed) counters[C_REMAINING] = 1;
    grid.sync();
    for (uint32_t a = gtid; a < nA; a += gsize) { uint32_t l = label[a]; while (label[l] != l) l = label[l]; label[a] = l; }
    grid.sync();
    if (counters[C_REMAINING] == 0) break;
    grid.sync();
  }
  for (uint32_t a = gtid; a < nA; a += gsize) if ((geomFlags[a] & 0x100u) && !S.asleep[a] && S.wake[a] != 0.0f) islandAwake[label[a]] = 1u;   // a body that is not ready keeps its island awake
  grid.sync();
  for (uint32_t a = gtid; a < nA; a += gsize) {
    if (!(geomFlags[a] & 0x100u)) continue;
    if (islandAwake[label[a]]) S.asleep[a] = 0u;   // woken with the island; the wake counter is re-armed by this step's sleep check
    else if (!S.asleep[a]) { S.asleep[a] = 1u; linVel[a] = make_float4(0, 0, 0, 0); angVel[a] = make_float4(0, 0, 0, 0); S.accLin[a] = make_float4(0, 0, 0, 0); S.accAng[a] = make_float4(0, 0, 0, 0); }
  }
  grid.sync();
  for (uint32_t i = gtid; i < nPairs; i += gsize) {   // sleeping islands are not solved
    const uint2 b = pairBodies[i];
    const bool act0 = (geomFlags[b.x] & 0x100u) && !S.asleep[b.x], act1 = (geomFlags[b.y] & 0x100u) && !S.asleep[b.y];
    if (!act0 && !act1) conFlag[i] = 0u;
  }
}

// a12: unconstrained velocities + solver body setup (preIntegrateBodies, DyTGSDynamics.cpp:992-1021)
__global__ void k_preintegrate(uint32_t nDyn, const uint32_t* __restrict__ dynActor, const float4* __restrict__ pos, const float4* __restrict__ quat,
                               float4* __restrict__ linVel, float4* __restrict__ angVel, const float4* __restrict__ invInertia, const float4* __restrict__ damp,
                               float gx, float gy, float gz, float dt, float4* __restrict__ sbLin, float4* __restrict__ sbAng, float4* __restrict__ sbDLin,
                               float4* __restrict__ sbDAng, float4* __restrict__ sbIA, float4* __restrict__ sbIB, float4* __restrict__ sbP, float4* __restrict__ sbQ,
                               float4* __restrict__ sbOrigAng, int pgs, SleepArgs S, const uint32_t* __restrict__ geomFlags, float4* __restrict__ extForce, float4* __restrict__ extTorque) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nDyn) return;
  const uint32_t a = dynActor[d];
  if (!gf_dynamic(geomFlags[a])) return;   // removed actor, kinematic body (KinematicCopyTGSTask: no gravity, no damping)
  const bool asleep = body_asleep(S, a);
  const float4 dm = damp[a]; const float4 ii = invInertia[a]; const float4 p4 = pos[a];
  v3 lv = V3(linVel[a]), av = V3(angVel[a]);
  if (extForce) {   // pending eFORCE / eTORQUE writes: consumed by this step
    const float4 F = extForce[a], T = extTorque[a];
    if (F.x != 0.f || F.y != 0.f || F.z != 0.f || T.x != 0.f || T.y != 0.f || T.z != 0.f) {
      if (!asleep) apply_external_force(V3(F), V3(T), p4.w, ii, Q4(quat[a]), dt, lv, av);
      extForce[a] = make_float4(0, 0, 0, 0); extTorque[a] = make_float4(0, 0, 0, 0);
    }
  }
  const uint32_t gfl = geomFlags[a];
  if (!asleep) unconstrained_velocity((gfl & 0x1000u) ? V3(0, 0, 0) : V3(gx, gy, gz), dt, dm.x, dm.y, dm.z, dm.w, lv, av);   // eDISABLE_GRAVITY: no gravity term (adding the zero vector changes nothing)
  if ((gfl & 0x2000u) && !asleep) av = gyroscopic(av, V3(ii.x, ii.y, ii.z), Q4(quat[a]), dt);
  // lock flags: TGS locks both velocities (copyToSolverBodyDataStep, DyTGSDynamics.cpp:195-222); PGS only the angular one (copyToSolverBodyData, DyRigidBodyToSolverBody.cpp:72-98)
  const uint32_t lock = (geomFlags[a] >> 16) & 0x3fu;
  if (lock && !asleep) { if (!pgs) lv = lock3(lv, lock & 7u); av = lock3(av, (lock >> 3) & 7u); }
  linVel[a] = F4(lv, 0.f); angVel[a] = F4(av, 0.f);
  const m33 rot = amfromq(Q4(quat[a]));
  const v3 sqrtInvI = V3(ii.x == 0.f ? 0.f : sqrtf(ii.x), ii.y == 0.f ? 0.f : sqrtf(ii.y), ii.z == 0.f ? 0.f : sqrtf(ii.z));
  const v3 sqrtI = V3(sqrtInvI.x == 0.f ? 0.f : 1.0f / sqrtInvI.x, sqrtInvI.y == 0.f ? 0.f : 1.0f / sqrtInvI.y, sqrtInvI.z == 0.f ? 0.f : 1.0f / sqrtInvI.z);
  m33 sI, sInertia; transform_inertia(sqrtInvI, rot, sI); transform_inertia(sqrtI, rot, sInertia);
  if (pgs) { sbLin[a] = make_float4(0, 0, 0, 0); sbAng[a] = make_float4(0, 0, 0, 0); }   // PGS solver bodies hold velocity deltas
  else { sbLin[a] = F4(lv, 0.f); sbAng[a] = F4(mmul(sInertia, av), 0.f); }
  sbDLin[a] = make_float4(0, 0, 0, 0); sbDAng[a] = make_float4(0, 0, 0, 0);
  sbIA[a] = make_float4(sI.c0.x, sI.c0.y, sI.c0.z, sI.c1.y); sbIB[a] = make_float4(sI.c1.z, sI.c2.z, __uint_as_float(lock), 0.f);
  sbP[a] = make_float4(p4.x, p4.y, p4.z, 0.f); sbQ[a] = make_float4(0, 0, 0, 1); sbOrigAng[a] = F4(av, 0.f);
}
// a17/a18: copyBackBodies (DyTGSDynamics.cpp:1549-1580); bodies without constraints take one full-dt step (:2573-2577)
__global__ void k_finalize_bodies(uint32_t nDyn, const uint32_t* __restrict__ dynActor, float dt, float4* __restrict__ pos, float4* __restrict__ quat, float4* __restrict__ linVel,
                                  float4* __restrict__ angVel, const float4* __restrict__ sbLin, const float4* __restrict__ sbAng, const float4* __restrict__ sbIA,
                                  const float4* __restrict__ sbIB, const float4* __restrict__ sbP, const float4* __restrict__ sbQ, const uint32_t* __restrict__ bodyHasCon,
                                  const float4* __restrict__ sbDLin, const float4* __restrict__ sbDAng, const float4* __restrict__ invInertia, SleepArgs S, const uint32_t* __restrict__ geomFlags) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nDyn) return;
  const uint32_t a = dynActor[d];
  if (body_asleep(S, a) || !gf_dynamic(geomFlags[a])) return;   // (removed actors lose their dynamic bit; kinematic bodies move in k_kin_finalize)
  const float4 ib = sbIB[a];
  const m33 sI = load_sym(sbIA[a], ib);
  v3 p = V3(sbP[a]); q4 dq = Q4(sbQ[a]);
  v3 lv = V3(sbLin[a]), as = V3(sbAng[a]);
  v3 dl = V3(sbDLin[a]), da = V3(sbDAng[a]);
  if (!bodyHasCon[a]) { dl = V3(0, 0, 0); da = V3(0, 0, 0); integrate_core_step(lv, as, sI, dt, p, dq, dl, da, __float_as_uint(ib.z)); }
  const float invMass = pos[a].w;
  const q4 q = qnormalized(qmul(dq, Q4(quat[a])));
  pos[a] = make_float4(p.x, p.y, p.z, invMass); quat[a] = F4(q);
  linVel[a] = F4(lv, 0.f); angVel[a] = F4(mmul(sI, as), 0.f);
  if (S.threshold > 0.f) { const float invDt = 1.0f / dt; sleep_check_dev(S, a, q, invInertia[a], invMass, dl * invDt, mmul(sI, da * invDt)); }   // motionVel of copyBackBodies
}
// Kinematic bodies (PxRigidBodyFlag::eKINEMATIC).  k_kin_set_targets = PxRigidDynamic::setKinematicTarget (NpRigidDynamic.cpp:129-158: normalised, moved into the
// body frame); k_kin_setup = Sc::Scene::kinematicsSetup / BodySim::calculateKinematicVelocity (ScKinematics.cpp:44-97) at the start of the solver part of the step;
// k_kin_finalize = BodySim::updateKinematicPose (ScKinematics.cpp:181-204) after it.  Bounds, contacts and solver rows of a step see the pose BEFORE the move.
__global__ void k_kin_set_targets(uint32_t n, const uint32_t* __restrict__ idx, const float* __restrict__ poses, const uint32_t* __restrict__ dynActor, uint32_t nDyn, const uint32_t* __restrict__ geomFlags,
                                  float4* __restrict__ kinP, float4* __restrict__ kinQ, uint32_t* __restrict__ kinHas, const float4* __restrict__ b2aP, const float4* __restrict__ b2aQ, uint32_t* __restrict__ counters) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
  const uint32_t d = idx[t];
  if (d >= nDyn) { atomicOr(&counters[C_ERROR], (uint32_t)E_BAD_INDEX); return; }
  const uint32_t a = dynActor[d];
  if (!(geomFlags[a] & 0x800u)) { atomicOr(&counters[C_ERROR], (uint32_t)E_BAD_INDEX); return; }   // "Body must be kinematic!"
  const float* o = poses + (size_t)t * 7;
  q4 q = qnormalized(Q4(o[0], o[1], o[2], o[3])); v3 p = V3(o[4], o[5], o[6]);
  if (b2aP) { const float4 bp = b2aP[a]; if (bp.w != 0.f) { p = qrot(q, V3(bp.x, bp.y, bp.z)) + p; q = qmul(q, Q4(b2aQ[a])); } }
  kinP[a] = F4(p, 0.f); kinQ[a] = F4(q); kinHas[a] = 1u;
}
__global__ void k_kin_setup(uint32_t nKin, const uint32_t* __restrict__ kinList, const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ kinP, const float4* __restrict__ kinQ,
                            const uint32_t* __restrict__ kinHas, float4* __restrict__ linVel, float4* __restrict__ angVel, float oneOverDt) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= nKin) return;
  const uint32_t a = kinList[t];
  if (!kinHas[a]) { linVel[a] = make_float4(0, 0, 0, 0); angVel[a] = make_float4(0, 0, 0, 0); return; }   // no target this step: the body stands still
  const float4 c = pos[a]; const v3 tp = V3(kinP[a]);
  linVel[a] = F4((tp - V3(c.x, c.y, c.z)) * oneOverDt, 0.f);
  const q4 cq = Q4(quat[a]);
  q4 q = qmul(Q4(kinQ[a]), Q4(-cq.x, -cq.y, -cq.z, cq.w));
  if (q.w < 0.f) q = Q4(-q.x, -q.y, -q.z, -q.w);   // shortest arc
  float angle; v3 axis;   // PxQuat::toRadiansAndUnitAxis (PxQuat.h:158-173)
  const float s2 = q.x * q.x + q.y * q.y + q.z * q.z;
  if (s2 < 1.0e-8f * 1.0e-8f) { angle = 0.f; axis = V3(1, 0, 0); }
  else { const float rs = 1.0f / sqrtf(s2); axis = V3(q.x, q.y, q.z) * rs; angle = fabsf(q.w) < 1.0e-8f ? 3.14159265358979323846f : atan2f(s2 * rs, q.w) * 2.0f; }
  angVel[a] = F4((axis * angle) * oneOverDt, 0.f);
}
__global__ void k_kin_finalize(uint32_t nKin, const uint32_t* __restrict__ kinList, float4* __restrict__ pos, float4* __restrict__ quat, const float4* __restrict__ kinP, const float4* __restrict__ kinQ,
                               uint32_t* __restrict__ kinHas) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= nKin) return;
  const uint32_t a = kinList[t];
  if (!kinHas[a]) return;
  const float4 p = kinP[a]; pos[a] = make_float4(p.x, p.y, p.z, 0.f); quat[a] = kinQ[a]; kinHas[a] = 0u;   // the velocity of the move stays readable until the next step
}
// a19: PxDirectGPUAPI get/set (gather/scatter by dynamic-body index)
__global__ void k_rd_get(uint32_t nb, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ dynActor, int type, const float4* __restrict__ pos, const float4* __restrict__ quat,
                         const float4* __restrict__ linVel, const float4* __restrict__ angVel, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const uint32_t a = dynActor[idx ? idx[i] : i];
  if (type == PXB_RD_GLOBAL_POSE) { const float4 q = quat[a], p = pos[a]; float* o = out + (size_t)i * 7; o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; o[4] = p.x; o[5] = p.y; o[6] = p.z; }
  else { const float4 v = type == PXB_RD_LINEAR_VELOCITY ? linVel[a] : angVel[a]; float* o = out + (size_t)i * 3; o[0] = v.x; o[1] = v.y; o[2] = v.z; }
}
// getRigidDynamicLinearAcceleration / AngularAcceleration (updateBodiesAndShapes.cu:1063-1106): (velocity now - velocity the step started from) * (1 / dt),
// computed for the requested bodies only, as the reference does (the CPU path has the same expression, NpSceneFetchResults.cpp:170-186)
__global__ void k_rd_get_accel(uint32_t nb, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ dynActor, const float4* __restrict__ vel, const float4* __restrict__ prev,
                               float oneOverDt, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nb) return;
  const uint32_t a = dynActor[idx ? idx[i] : i];
  const float4 v = vel[a], p = prev[a]; float* o = out + (size_t)i * 3;
  o[0] = __fmul_rn(__fsub_rn(v.x, p.x), oneOverDt); o[1] = __fmul_rn(__fsub_rn(v.y, p.y), oneOverDt); o[2] = __fmul_rn(__fsub_rn(v.z, p.z), oneOverDt);
}
__global__ void k_rd_set(uint32_t nb, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ dynActor, int type, float4* __restrict__ pos, float4* __restrict__ quat,
                         float4* __restrict__ linVel, float4* __restrict__ angVel, const float* __restrict__ in, float* __restrict__ wake, uint32_t* __restrict__ asleep,
                         float4* __restrict__ extForce, float4* __restrict__ extTorque) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const uint32_t a = dynActor[idx ? idx[i] : i];
  if (type == PXB_RD_FORCE || type == PXB_RD_TORQUE) {   // applied by the next step only; a non-zero force wakes the body (autowake)
    const float* o = in + (size_t)i * 3; const float4 v = make_float4(o[0], o[1], o[2], 0.f);
    if (type == PXB_RD_FORCE) extForce[a] = v; else extTorque[a] = v;
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f) { if (wake[a] < 20.0f * 0.02f) wake[a] = 20.0f * 0.02f; asleep[a] = 0u; }
    return;
  }
  wake[a] = 20.0f * 0.02f; asleep[a] = 0u;   // setting pose / velocity wakes the body (PxRigidDynamic autowake, wakeCounterResetValue)
  if (type == PXB_RD_GLOBAL_POSE) { const float* o = in + (size_t)i * 7; quat[a] = make_float4(o[0], o[1], o[2], o[3]); const float w = pos[a].w; pos[a] = make_float4(o[4], o[5], o[6], w); }
  else { const float* o = in + (size_t)i * 3; const float4 v = make_float4(o[0], o[1], o[2], 0.f); if (type == PXB_RD_LINEAR_VELOCITY) linVel[a] = v; else angVel[a] = v; }
}
// f3: tensor front end in the reference's ovphysx wire formats (ovphysx/python/ovphysx/types.py TensorType): pose [N,7] = (p.xyz, q.xyzw), velocity
// [N,6] = (linear, angular), mass / inverse mass [N], force [N,3], wrench [N,9] = (force, torque, application point in the world frame).
__global__ void k_tensor_read(uint32_t nb, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ dynActor, int type, const float4* __restrict__ pos, const float4* __restrict__ quat,
                              const float4* __restrict__ linVel, const float4* __restrict__ angVel, float* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const uint32_t a = dynActor[idx ? idx[i] : i];
  if (type == PXB_TENSOR_RIGID_BODY_POSE) { const float4 q = quat[a], p = pos[a]; float* o = out + (size_t)i * 7; o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w; }
  else if (type == PXB_TENSOR_RIGID_BODY_VELOCITY) { const float4 v = linVel[a], w = angVel[a]; float* o = out + (size_t)i * 6; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = w.x; o[4] = w.y; o[5] = w.z; }
  else { const float im = pos[a].w; out[i] = type == PXB_TENSOR_RIGID_BODY_INV_MASS ? im : (im > 0.f ? 1.0f / im : 0.f); }
}
__global__ void k_tensor_write(uint32_t nb, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ dynActor, int type, float4* __restrict__ pos, float4* __restrict__ quat,
                               float4* __restrict__ linVel, float4* __restrict__ angVel, const float* __restrict__ in, float* __restrict__ wake, uint32_t* __restrict__ asleep,
                               float4* __restrict__ extForce, float4* __restrict__ extTorque) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const uint32_t a = dynActor[idx ? idx[i] : i];
  if (type == PXB_TENSOR_RIGID_BODY_FORCE || type == PXB_TENSOR_RIGID_BODY_WRENCH) {   // applied by the next step only (PxRigidBody::addForce / addTorque, eFORCE)
    const float* o = in + (size_t)i * (type == PXB_TENSOR_RIGID_BODY_FORCE ? 3 : 9);
    const v3 F = V3(o[0], o[1], o[2]); v3 T = V3(0, 0, 0);
    if (type == PXB_TENSOR_RIGID_BODY_WRENCH) { const float4 p = pos[a]; T = V3(o[3], o[4], o[5]) + cross(V3(o[6], o[7], o[8]) - V3(p.x, p.y, p.z), F); extTorque[a] = F4(T, 0.f); }
    extForce[a] = F4(F, 0.f);
    if (F.x != 0.f || F.y != 0.f || F.z != 0.f || T.x != 0.f || T.y != 0.f || T.z != 0.f) { if (wake[a] < 20.0f * 0.02f) wake[a] = 20.0f * 0.02f; asleep[a] = 0u; }
    return;
  }
  wake[a] = 20.0f * 0.02f; asleep[a] = 0u;
  if (type == PXB_TENSOR_RIGID_BODY_POSE) { const float* o = in + (size_t)i * 7; const float w = pos[a].w; pos[a] = make_float4(o[0], o[1], o[2], w); quat[a] = make_float4(o[3], o[4], o[5], o[6]); }
  else { const float* o = in + (size_t)i * 6; linVel[a] = make_float4(o[0], o[1], o[2], 0.f); angVel[a] = make_float4(o[3], o[4], o[5], 0.f); }
}
__global__ void k_states_get(uint32_t nDyn, const uint32_t* __restrict__ dynActor, const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ linVel,
                             const float4* __restrict__ angVel, float* __restrict__ out) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nDyn) return;
  const uint32_t a = dynActor[d]; float* o = out + (size_t)d * 13;
  const float4 p = pos[a], q = quat[a], l = linVel[a], w = angVel[a];
  o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w; o[7] = l.x; o[8] = l.y; o[9] = l.z; o[10] = w.x; o[11] = w.y; o[12] = w.z;
}
__global__ void k_states_set(uint32_t nDyn, const uint32_t* __restrict__ dynActor, float4* __restrict__ pos, float4* __restrict__ quat, float4* __restrict__ linVel,
                             float4* __restrict__ angVel, const float* __restrict__ in, float* __restrict__ wake, uint32_t* __restrict__ asleep) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nDyn) return;
  const uint32_t a = dynActor[d]; const float* o = in + (size_t)d * 13;
  wake[a] = 20.0f * 0.02f; asleep[a] = 0u;
  const float w = pos[a].w;
  pos[a] = make_float4(o[0], o[1], o[2], w); quat[a] = make_float4(o[3], o[4], o[5], o[6]); linVel[a] = make_float4(o[7], o[8], o[9], 0.f); angVel[a] = make_float4(o[10], o[11], o[12], 0.f);
}
// Multi-GPU state exchange: one launch pushes a packed block into up to 8 peer buffers over NVLink (P2P stores to peer-mapped
// addresses, e.g. torch symmetric memory); each 16-byte word is loaded once and stored to every destination.
struct PeerPtrs { float4* p[8]; };
__global__ void __launch_bounds__(256) k_scatter_to_peers(const float4* __restrict__ src, size_t n16, PeerPtrs dst, uint32_t nDst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < (int)nDst) dst.p[k][i] = v;
  }
}
// a19: packed state export for scenes whose step does not end in k_env_solve (device-wide path; environments whose dynamic bodies are not
// contiguous in dynamic-body order): one CTA per 256 bodies, coalesced stores into every target.
__global__ void __launch_bounds__(256) k_states_export(uint32_t nDyn, const uint32_t* __restrict__ dynActor, const float4* pos, const float4* quat, const float4* linVel, const float4* angVel,
                                                       const ExportTable* __restrict__ tab) {
  const uint32_t nT = tab->n;
  if (!nT) return;
  const uint32_t d0 = blockIdx.x * 256u;
  export_packed_range(tab, nT, dynActor, d0, min(256u, nDyn - d0), pos, quat, linVel, angVel, threadIdx.x, 256u);
}
__global__ void k_set_export(ExportTable* __restrict__ tab, ExportTable v) { if (threadIdx.x == 0 && blockIdx.x == 0) *tab = v; }
// Cross-GPU flags for the exchange built on the export (physx_b200/multi_gpu.py FusedStateGather): a rank raises its flag in every peer's
// signal pad after its step (the step's P2P stores are complete at the kernel boundary; the fence orders them before the flag at system
// scope), consumers spin on their own pad.
struct FlagPtrs { uint32_t* f[PXB_MAX_EXPORT]; };
__global__ void k_peer_signal(FlagPtrs p, uint32_t n, uint32_t value) {
  if (threadIdx.x < n) { __threadfence_system(); *reinterpret_cast<volatile uint32_t*>(p.f[threadIdx.x]) = value; }
}
__global__ void k_peer_wait(const uint32_t* __restrict__ flags, uint32_t n, uint32_t value) {
  if (threadIdx.x < n) { while ((int32_t)(*reinterpret_cast<const volatile uint32_t*>(flags + threadIdx.x) - value) < 0) __nanosleep(64); __threadfence_system(); }
}
__global__ void k_init_freelist(uint32_t cap, uint32_t* __restrict__ freeList) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) freeList[i] = i;
}

__global__ void k_env_begin(uint32_t* __restrict__ counters) {   // per-step counter reset of the environment path
  if (threadIdx.x == 0) { counters[C_NPAIRS_NEW] = 0; counters[C_NCREATED] = 0; counters[C_NDELETED] = 0; counters[C_NCON] = 0; counters[C_NPART] = 0; counters[C_NGJK] = 0; counters[C_NGJK_QUERY] = 0; counters[C_NGJK_FULL] = 0; counters[C_NGJK_EPA] = 0; counters[C_NBOXGEN] = 0; counters[C_MAXCONENV] = 0; counters[C_MAXPAIRENV] = 0; counters[C_NTOUCH_FOUND] = 0; counters[C_NTOUCH_LOST] = 0;
                          counters[C_FREE_SNAP] = counters[C_FREE_TAIL]; }
}

// ---------------------------------------------------------------------------------------------
// host side
static const size_t ENV_SMEM_MAX = 227 * 1024 - 2048;   // dynamic shared memory budget of k_env_solve (static part: partition tables)
static const size_t ENV_CON_BYTES = 5 * sizeof(uint32_t);   // 5 u32 lists per pair slot (rows live in registers)
static size_t env_solve_smem(uint32_t maxList, uint32_t conCap, uint32_t threads = 256) { return (size_t)maxList * (8 * sizeof(float4) + 4 * sizeof(uint32_t)) + (size_t)conCap * ENV_CON_BYTES + 16 + (size_t)threads * 12 * sizeof(float4); }
static uint32_t env_con_cap_limit(uint32_t maxList) { return (uint32_t)((ENV_SMEM_MAX - env_solve_smem(maxList, 0)) / ENV_CON_BYTES); }
template <typename T> static cudaError_t dalloc(T*& p, size_t n) { return cudaMalloc((void**)&p, sizeof(T) * (n ? n : 1)); }
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static uint32_t bits_for(uint64_t n) { uint32_t b = 1; while ((1ull << b) < n) ++b; return b; }

extern "C" {

PXB_API const char* pxb_last_error(void) { return g_err.c_str(); }
PXB_API int pxb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

static uint32_t env_threads_for(uint32_t nCon) { return nCon <= 32 ? 32u : (nCon <= 64 ? 64u : (nCon <= 128 ? 128u : 256u)); }   // CTA size of k_env_solve: one thread per constraint when possible
static void drop_graphs(PxbScene* s);
static int scene_alloc(PxbScene* s) {
  const size_t A = s->capA, Pn = s->capPairs;
  CK(dalloc(s->pos, A)); CK(dalloc(s->quat, A)); CK(dalloc(s->linVel, A)); CK(dalloc(s->angVel, A)); CK(dalloc(s->invInertia, A)); CK(dalloc(s->damp, A));
  CK(dalloc(s->dims, A)); CK(dalloc(s->aabbMin, A)); CK(dalloc(s->aabbMax, A)); CK(dalloc(s->geomFlags, A)); CK(dalloc(s->envId, A)); CK(dalloc(s->dynActorDev, A));
  CK(dalloc(s->largeList, A)); CK(dalloc(s->tight, A * 6));
  if (s->bodyAccel) { CK(dalloc(s->prevLin, A)); CK(dalloc(s->prevAng, A)); CK(cudaMemset(s->prevLin, 0, 16 * A)); CK(cudaMemset(s->prevAng, 0, 16 * A)); }
  CK(dalloc(s->sbLin, A)); CK(dalloc(s->sbAng, A)); CK(dalloc(s->sbDLin, A)); CK(dalloc(s->sbDAng, A)); CK(dalloc(s->sbIA, A)); CK(dalloc(s->sbIB, A)); CK(dalloc(s->sbP, A));
  CK(dalloc(s->sbQ, A)); CK(dalloc(s->sbOrigAng, A));
  CK(dalloc(s->bodyCnt, A)); CK(dalloc(s->bodyStart, A)); CK(dalloc(s->bodyCursor, A)); CK(dalloc(s->bodyNext, A)); CK(dalloc(s->bodyMask, A)); CK(dalloc(s->bodyHasCon, A));
  CK(dalloc(s->cellKey, A)); CK(dalloc(s->cellKeyAlt, A)); CK(dalloc(s->cellVal, A)); CK(dalloc(s->cellValAlt, A)); CK(dalloc(s->sMin, A)); CK(dalloc(s->sMax, A));
  for (int k = 0; k < 2; ++k) { CK(dalloc(s->pairKeys[k], Pn)); CK(dalloc(s->pairSlots[k], Pn)); }
  CK(dalloc(s->pairKeyAlt, Pn)); CK(dalloc(s->pairValTmp, Pn)); CK(dalloc(s->pairValAlt, Pn)); CK(dalloc(s->nPairsDev, 2));
  { uint32_t R = 1; while (R < Pn) R <<= 1; s->ringMask = R - 1; CK(dalloc(s->freeList, R)); }
  CK(dalloc(s->createdKeys, Pn)); CK(dalloc(s->deletedKeys, Pn));
  CK(dalloc(s->manifolds, Pn * PXB_MANIFOLD_F4)); CK(dalloc(s->frictions, Pn * PXB_FRICTION_F4));
  CK(dalloc(s->cHdr, Pn)); CK(dalloc(s->cPts, Pn * 4)); CK(dalloc(s->pairBodies, Pn)); CK(dalloc(s->cForce, Pn * 4));
  CK(dalloc(s->gjkList, Pn)); CK(dalloc(s->gjkQuery, Pn)); CK(dalloc(s->gjkFull, Pn)); CK(dalloc(s->gjkEpa, Pn)); CK(dalloc(s->boxList, Pn)); CK(dalloc(s->pairOrder, Pn)); CK(dalloc(s->npClass, Pn)); CK(dalloc(s->npClassCount, 2 * NP_CLASSES)); CK(dalloc(s->conFlag, Pn)); CK(dalloc(s->conIdx, Pn)); CK(dalloc(s->conPair, Pn)); CK(dalloc(s->rankOfPair, Pn)); CK(dalloc(s->conSortKey, Pn)); CK(dalloc(s->conSortKeyAlt, Pn));
  CK(dalloc(s->conPairAlt, Pn));
  CK(dalloc(s->conB0, Pn)); CK(dalloc(s->conB1, Pn)); CK(dalloc(s->conPos0, Pn)); CK(dalloc(s->conPos1, Pn)); CK(dalloc(s->conColour, Pn)); CK(dalloc(s->conDone, Pn));
  CK(dalloc(s->bodyList, Pn * 2)); CK(dalloc(s->ordered, Pn));
  CK(dalloc(s->colourTicket, 4)); CK(dalloc(s->prevB0, Pn)); CK(dalloc(s->prevB1, Pn)); CK(dalloc(s->prevColour, Pn)); CK(dalloc(s->prevNCon, 1)); CK(cudaMemsetAsync(s->prevNCon, 0, 4, s->stream)); CK(dalloc(s->partCnt, MAX_PARTITIONS + 1)); CK(dalloc(s->partStart, MAX_PARTITIONS + 1)); CK(dalloc(s->partCursor, MAX_PARTITIONS + 1));
  CK(dalloc(s->ptA, Pn * 28)); s->ptB = s->ptA + Pn * 4; s->ptC = s->ptA + Pn * 8;   // one allocation: the environment path views it as 25 x Pn (pxb_env.cuh Rows)
  s->frA = s->ptA + Pn * 12; s->frB = s->ptA + Pn * 16; s->frC = s->ptA + Pn * 20; s->frD = s->ptA + Pn * 24;
  CK(dalloc(s->stage, A * 38)); CK(dalloc(s->stageIdx, A));   // staging: {get, set} x {pose 7, linear 3, angular 3} + set {force 3, torque 3} + get {linear, angular acceleration 3} floats per actor
  CK(dalloc(s->extForce, A)); CK(dalloc(s->extTorque, A)); CK(cudaMemsetAsync(s->extForce, 0, sizeof(float4) * A, s->stream)); CK(cudaMemsetAsync(s->extTorque, 0, sizeof(float4) * A, s->stream));
  CK(dalloc(s->counters, C_COUNT)); CK(cudaMallocHost((void**)&s->hostCounters, sizeof(uint32_t) * (C_COUNT + 2)));
  CK(dalloc(s->rsTmp.blockHist, RS_MAX_CTAS * 256)); CK(dalloc(s->rsTmp.digitTotals, 256)); CK(dalloc(s->scanSums, RS_MAX_CTAS));
  CK(cudaMemsetAsync(s->counters, 0, sizeof(uint32_t) * C_COUNT, s->stream)); CK(cudaMemsetAsync(s->nPairsDev, 0, 8, s->stream));
  CK(cudaMemsetAsync(s->sbLin, 0, 16 * A, s->stream)); CK(cudaMemsetAsync(s->sbAng, 0, 16 * A, s->stream)); CK(cudaMemsetAsync(s->sbDLin, 0, 16 * A, s->stream));
  CK(cudaMemsetAsync(s->sbDAng, 0, 16 * A, s->stream)); CK(cudaMemsetAsync(s->sbIA, 0, 16 * A, s->stream)); CK(cudaMemsetAsync(s->sbIB, 0, 16 * A, s->stream));
  CK(cudaMemsetAsync(s->sbOrigAng, 0, 16 * A, s->stream)); CK(cudaMemsetAsync(s->bodyHasCon, 0, 4 * A, s->stream));
  CK(cudaMemsetAsync(s->manifolds, 0, sizeof(float4) * Pn * PXB_MANIFOLD_F4, s->stream)); CK(cudaMemsetAsync(s->frictions, 0, sizeof(float4) * Pn * PXB_FRICTION_F4, s->stream));
  k_init_freelist<<<cdiv((uint32_t)Pn, 256), 256, 0, s->stream>>>((uint32_t)Pn, s->freeList);
  const uint32_t top = (uint32_t)Pn;   // ring of free persistent slots: head = 0, tail = Pn
  CK(cudaMemcpyAsync(s->counters + C_FREE_TAIL, &top, 4, cudaMemcpyHostToDevice, s->stream));
  CK(dalloc(s->wake, A)); CK(dalloc(s->accLin, A)); CK(dalloc(s->accAng, A)); CK(dalloc(s->asleep, A)); CK(dalloc(s->nInter, A)); CK(dalloc(s->islandLabel, A)); CK(dalloc(s->islandAwake, A));
  CK(cudaMemsetAsync(s->accLin, 0, 16 * A, s->stream)); CK(cudaMemsetAsync(s->accAng, 0, 16 * A, s->stream)); CK(cudaMemsetAsync(s->asleep, 0, 4 * A, s->stream)); CK(cudaMemsetAsync(s->nInter, 0, 4 * A, s->stream));
  { std::vector<float> w(A, 20.0f * 0.02f); CK(cudaMemcpyAsync(s->wake, w.data(), 4 * A, cudaMemcpyHostToDevice, s->stream)); CK(cudaStreamSynchronize(s->stream)); }   // PxRigidDynamic default wake counter
  CK(dalloc(s->actorMat, A)); CK(cudaMemsetAsync(s->actorMat, 0, 4 * A, s->stream));
  CK(dalloc(s->touchState, Pn)); CK(cudaMemsetAsync(s->touchState, 0, 4 * Pn, s->stream)); CK(dalloc(s->touchFound, Pn)); CK(dalloc(s->touchLost, Pn));
  CK(dalloc(s->exportTab, 1)); CK(cudaMemsetAsync(s->exportTab, 0, sizeof(ExportTable), s->stream)); CK(dalloc(s->envDyn, A));
  CK(dalloc(s->bodyBest, A)); CK(dalloc(s->actorLocal, A)); CK(dalloc(s->slotColour, Pn)); CK(cudaMemsetAsync(s->slotColour, 0xff, 4 * Pn, s->stream)); for (int k = 0; k < 2; ++k) { CK(dalloc(s->envSeg[k], A)); CK(cudaMemsetAsync(s->envSeg[k], 0, sizeof(uint2) * A, s->stream)); }
  CK(cudaStreamSynchronize(s->stream));
  return PXB_OK;
}

PXB_API int pxb_scene_create(const PxbSceneDesc* desc, PxbScene** out) {
  PxbScene* s = nullptr;
  if (!desc || !out) return fail(PXB_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(PXB_ERR_NO_DEVICE, "no CUDA device: physx_b200 has no CPU fallback"); }
  if (desc->solverType != PXB_SOLVER_TGS && desc->solverType != PXB_SOLVER_PGS) return fail(PXB_ERR_INVALID, "unknown solver type");
  if (desc->device < 0 || desc->device >= ndev) return fail(PXB_ERR_INVALID, "bad device ordinal");
  DeviceGuard dg_(desc->device);
  s = new PxbScene(); s->desc = *desc; s->device = desc->device;
  CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&s->velEvent, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s->orderEvent, cudaEventDisableTiming));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, desc->device));
  s->numSMs = prop.multiProcessorCount;
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_colour_partition, 256, 0)); s->coopBlocksColour = std::max(1, std::min(occ, 4)) * s->numSMs;
  CK(cudaFuncSetAttribute(k_colour_firstfit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(pxb_env_set_attributes((int)ENV_SMEM_MAX, (int)(ENV_BP_WARPS * (ENV_MAX_LIST * 36 + ENV_BP_STAGE * 8))));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sleep_islands, 256, 0)); s->coopBlocksSleep = std::max(1, std::min(occ, 2)) * s->numSMs;
  { int occT = 0, occP = 0; CK(pxb_solve_occupancy(&occT, &occP));
    s->coopBlocksSolvePgs = std::max(1, std::min(occP, PXB_SOLVE_CTAS_PER_SM)) * s->numSMs; s->coopBlocksSolve = std::max(1, std::min(occT, PXB_SOLVE_CTAS_PER_SM)) * s->numSMs; }
  s->capA = std::max(16u, desc->maxActors);
  s->capPairs = desc->maxPairs ? desc->maxPairs : std::max(1024u, 8u * s->capA);
  s->bitsA = bits_for(s->capA);
  s->rsTmp.ctas = std::min<uint32_t>(RS_MAX_CTAS, (uint32_t)s->numSMs * 2);
  { const char* ng = getenv("PXB_NO_GRAPH"); if (ng && ng[0] == '1') s->useGraph = false; }
  { const char* cl = getenv("PXB_COLOUR_LEGACY"); if (cl && cl[0] == '1') s->colourLegacy = true; const char* cb = getenv("PXB_COLOUR_BACKOFF_NS"); if (cb) s->colourBackoffNs = (uint32_t)atoi(cb);
    const char* cp = getenv("PXB_COLOUR_PREFIX"); if (cp && cp[0] == '0') s->colourPrefix = false;
    const char* cw = getenv("PXB_COLOUR_WINDOW"); if (cw) s->colourWindow = (uint32_t)atoi(cw);
    const char* bp_ = getenv("PXB_BOX_PHASES"); if (bp_ && bp_[0] == '0') s->boxPhases = false; if (bp_ && bp_[0] == '2') s->boxPhasesEnv = true;   // 2: also on the environment path (A/B)
    const char* gp = getenv("PXB_GJK_PHASES"); if (gp && gp[0] == '0') s->gjkPhases = false; }   // A/B hooks of the exact colouring
  if (desc->reserved[1] & PXB_FLAG_NO_ENV_PATH) s->envDisabled = true;
  if (desc->reserved[1] & PXB_FLAG_RELAXED_PARTITIONING) { s->relaxedPartitioning = true; s->envDisabled = true; }
  if (desc->reserved[1] & PXB_FLAG_BODY_ACCELERATIONS) s->bodyAccel = true;
  s->envConCapForced = desc->reserved[2]; s->envThreadsForced = desc->reserved[3];
  memcpy(&s->sleepThreshold, &desc->reserved[4], 4); if (!(s->sleepThreshold > 0.f)) s->sleepThreshold = 0.f;
  const int rc = scene_alloc(s);
  if (rc != PXB_OK) { delete s; return rc; }
  *out = s;
  return PXB_OK;
}

PXB_API void pxb_scene_release(PxbScene* s) { DeviceGuard dg_(s);
  if (!s) return;
  cudaStreamSynchronize(s->stream);
  drop_graphs(s);
  void* ptrs[] = {s->candKeys, s->candCount, s->candRefMin, s->candRefMax, s->aggId, s->kinList, s->kinP, s->kinQ, s->kinHas, s->kinFtv, s->prevLin, s->prevAng, s->pos, s->quat, s->linVel, s->angVel, s->invInertia, s->damp, s->dims, s->aabbMin, s->aabbMax, s->geomFlags, s->envId, s->dynActorDev, s->largeList, s->tight,
                  s->sbLin, s->sbAng, s->sbDLin, s->sbDAng, s->sbIA, s->sbIB, s->sbP, s->sbQ, s->sbOrigAng, s->bodyCnt, s->bodyStart, s->bodyCursor, s->bodyNext, s->bodyMask, s->bodyHasCon,
                  s->cellKey, s->cellKeyAlt, s->cellVal, s->cellValAlt, s->sMin, s->sMax, s->pairKeys[0], s->pairKeys[1], s->pairSlots[0], s->pairSlots[1], s->pairKeyAlt, s->pairValTmp,
                  s->pairValAlt, s->nPairsDev, s->freeList, s->createdKeys, s->deletedKeys, s->manifolds, s->frictions, s->cHdr, s->cPts, s->pairBodies, s->cForce, s->gjkList, s->gjkQuery, s->gjkFull, s->gjkEpa, s->boxList, s->filterData, s->shapeOff, s->tcPos, s->tcQuat, s->s2bP, s->s2bQ, s->b2aP, s->b2aQ, s->actorPos, s->actorQuat, s->frReport, s->ccIdx, s->ccOff, s->ccCount, s->ccTotal, s->actorDyn, s->ccPatches, s->ccPoints, s->ccFriction, s->ccForces, s->pairOrder, s->npClass, s->npClassCount, s->conFlag, s->conIdx,
                  s->conPair, s->rankOfPair, s->conSortKey, s->conSortKeyAlt, s->conPairAlt, s->orderKeys, s->conB0, s->conB1, s->conPos0, s->conPos1, s->conColour, s->conDone, s->bodyList,
                  s->ordered, s->partCnt, s->partStart, s->partCursor, s->colourTicket, s->prevB0, s->prevB1, s->prevColour, s->prevNCon, s->ptA, s->counters, s->rsTmp.blockHist, s->rsTmp.digitTotals, s->scanSums, s->stage, s->stageIdx, s->extForce, s->extTorque, s->hullMeta, s->hullVerts, s->hullPolys, s->hullRefs, s->hullEdges,
                  s->envStart, s->envList, s->actorLocal, s->exportTab, s->envDyn, s->actorMat, s->matTab, s->touchState, s->touchFound, s->touchLost, s->slotColour, s->bodyBest, s->wake, s->accLin, s->accAng, s->asleep, s->nInter, s->islandLabel, s->islandAwake, s->envSeg[0], s->envSeg[1]};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (s->hostCounters) cudaFreeHost(s->hostCounters);
  cudaStreamDestroy(s->stream); if (s->copyStream) cudaStreamDestroy(s->copyStream); if (s->velEvent) cudaEventDestroy(s->velEvent); if (s->orderEvent) cudaEventDestroy(s->orderEvent);
  delete s;
}

PXB_API uint32_t pxb_scene_num_actors(const PxbScene* s) { return s ? s->nA : 0; }
PXB_API uint32_t pxb_scene_num_dynamic(const PxbScene* s) { return s ? s->nDyn : 0; }
PXB_API void* pxb_scene_stream(PxbScene* s) { DeviceGuard dg_(s); return s ? (void*)s->stream : nullptr; }
PXB_API void* pxb_scene_state_device_ptr(PxbScene* s, int which) { DeviceGuard dg_(s);
  if (!s) return nullptr;
  switch (which) { case 0: return s->pos; case 1: return s->quat; case 2: return s->linVel; case 3: return s->angVel; default: return nullptr; }
}

// bounding diameter of a shape (rotation independent), used to size the broadphase grid cell
static float shape_diameter(const ActorRec& r) {
  switch (r.geomType) {
    case PXB_GEOM_SPHERE: return 2.f * r.dims[0];
    case PXB_GEOM_CAPSULE: return 2.f * (r.dims[0] + r.dims[1]);
    case PXB_GEOM_BOX: return 2.f * std::sqrt(r.dims[0] * r.dims[0] + r.dims[1] * r.dims[1] + r.dims[2] * r.dims[2]);
    case PXB_GEOM_CONVEXMESH: return r.dims[3] > 0.f ? r.dims[3] : INFINITY;   // dims[3] = hull diameter, filled by pxb_scene_add_actors
    default: return INFINITY;
  }
}

// Environment path eligibility + per-environment actor lists (CSR).  Eligible: every dynamic actor carries an
// environment id, at most ENV_MAX_GLOBALS env-less actors (all static: the shared ground plane / walls), and no
// environment lists more than ENV_MAX_LIST actors.  Each list = the environment's actors merged with the env-less
// statics, ascending actor index, so row-major pair enumeration yields ascending pair keys.
static void rebuild_env(PxbScene* s, bool usesEnv, uint32_t maxEnv) {
  s->envEligible = false;
  const char* em = getenv("PXB_ENV_MODE");
  if (s->envDisabled || (em && em[0] == '0')) return;
  // A small scene without environment ids (BASELINE config 1: 100 boxes) is ONE environment: the whole step then runs on one SM in 5 launches
  // instead of ~36, which is what bounds a scene of that size.
  const bool single = !usesEnv && s->nA > 0 && s->nA <= ENV_MAX_LIST;
  if (!usesEnv && !single) return;
  if (single) maxEnv = 0;
  if (maxEnv >= s->capA) return;   // sparse environment ids: stay on the device-wide path
  std::vector<uint32_t> globals; const uint32_t nEnv = maxEnv + 1;
  std::vector<uint32_t> cnt(nEnv, 0);
  auto envOf = [&](uint32_t a) { return single ? 0u : s->recs[a].envId; };
  for (uint32_t a = 0; a < s->nA; ++a) {
    const ActorRec& r = s->recs[a];
    if (r.flags & ACTOR_REMOVED) continue;
    if (envOf(a) == NONE32) { if (r.flags & PXB_ACTOR_DYNAMIC) return; globals.push_back(a); if (globals.size() > ENV_MAX_GLOBALS) return; }
    else cnt[envOf(a)]++;
  }
  const uint32_t G = (uint32_t)globals.size();
  uint32_t maxList = 0;
  for (uint32_t e = 0; e < nEnv; ++e) maxList = std::max(maxList, cnt[e] + G);
  if (maxList > ENV_MAX_LIST) return;
  std::vector<uint32_t> start(nEnv + 1, 0), local(s->nA, NONE32);
  for (uint32_t e = 0; e < nEnv; ++e) start[e + 1] = start[e] + cnt[e] + G;
  std::vector<uint32_t> list(start[nEnv]), cur(start.begin(), start.end() - 1), gi(nEnv, 0);
  for (uint32_t a = 0; a < s->nA; ++a) {   // ascending a: merge the env-less statics in by index
    const uint32_t e = envOf(a);
    if (e == NONE32 || (s->recs[a].flags & ACTOR_REMOVED)) continue;
    while (gi[e] < G && globals[gi[e]] < a) list[cur[e]++] = globals[gi[e]++];
    local[a] = cur[e] - start[e]; list[cur[e]++] = a;
  }
  for (uint32_t e = 0; e < nEnv; ++e) while (gi[e] < G) list[cur[e]++] = globals[gi[e]++];
  if (s->envStart) cudaFree(s->envStart); if (s->envList) cudaFree(s->envList); s->envStart = s->envList = nullptr;
  if (dalloc(s->envStart, nEnv + 1) != cudaSuccess || dalloc(s->envList, list.size()) != cudaSuccess) { cudaGetLastError(); return; }
  cudaMemcpyAsync(s->envStart, start.data(), 4 * (nEnv + 1), cudaMemcpyHostToDevice, s->stream);
  cudaMemcpyAsync(s->envList, list.data(), 4 * list.size(), cudaMemcpyHostToDevice, s->stream);
  cudaMemcpyAsync(s->actorLocal, local.data(), 4 * s->nA, cudaMemcpyHostToDevice, s->stream);
  cudaStreamSynchronize(s->stream);
  {   // per environment: the range its dynamic bodies occupy in dynamic-body order (the fused state export writes whole blocks; contiguous in every
      // scene whose actors are added environment by environment)
    std::vector<uint2> ed(nEnv, make_uint2(0, 0)); bool contiguous = true;
    for (uint32_t e = 0; e < nEnv; ++e) {
      int first = -1, last = -1; uint32_t nd = 0;
      for (uint32_t k = start[e]; k < start[e + 1]; ++k) { const int d = s->dynIndex[list[k]]; if (d < 0 || envOf(list[k]) == NONE32) continue; if (first < 0) first = d; last = d; ++nd; }
      if (nd && (uint32_t)(last - first + 1) != nd) contiguous = false;
      ed[e] = make_uint2(first < 0 ? 0u : (uint32_t)first, nd);
    }
    cudaMemcpyAsync(s->envDyn, ed.data(), sizeof(uint2) * nEnv, cudaMemcpyHostToDevice, s->stream); cudaStreamSynchronize(s->stream);
    s->envDynContiguous = contiguous;
  }
  s->nEnv = nEnv; s->envMaxList = maxList; s->envEligible = true;
  {   // candidate lists of the broadphase (invalid until the first step of every environment builds them); without them k_env_bp enumerates all pairs every step
    if (s->candKeys) cudaFree(s->candKeys); if (s->candCount) cudaFree(s->candCount); if (s->candRefMin) cudaFree(s->candRefMin); if (s->candRefMax) cudaFree(s->candRefMax);
    s->candKeys = s->candCount = nullptr; s->candRefMin = s->candRefMax = nullptr;
    const char* cc = getenv("PXB_ENV_BP_CAND"); s->candOn = !(cc && cc[0] == '0');
    s->candCap = 8 * maxList;
    if (s->candOn && (dalloc(s->candKeys, (size_t)nEnv * s->candCap) != cudaSuccess || dalloc(s->candCount, nEnv) != cudaSuccess || dalloc(s->candRefMin, list.size()) != cudaSuccess ||
                      dalloc(s->candRefMax, list.size()) != cudaSuccess)) { cudaGetLastError(); if (s->candKeys) cudaFree(s->candKeys); s->candKeys = nullptr; }   // no memory: all pairs every step
    if (s->candKeys) { cudaMemsetAsync(s->candCount, 0xff, 4 * (size_t)nEnv, s->stream); cudaStreamSynchronize(s->stream); }
  }
#ifdef PXB_ENV_TIMING
  if (s->envTiming) cudaFree(s->envTiming); cudaMalloc((void**)&s->envTiming, (size_t)nEnv * 16 * 8); cudaMemset(s->envTiming, 0, (size_t)nEnv * 16 * 8);
#endif
  // first guess: about one constraint / two pairs per actor; adapted after every step from the device-side maxima (read_counters)
  if (s->envConCap == 0) { s->envConCap = std::max(32u, 2 * maxList); s->envSolveThreads = env_threads_for(maxList); }
  if (s->envConCapForced) s->envConCap = std::min(s->envConCapForced, env_con_cap_limit(maxList));
  if (s->envThreadsForced) s->envSolveThreads = env_threads_for(s->envThreadsForced);
}

static void rebuild_grid(PxbScene* s) {
  // cell edge = largest rotation-independent extent among regular shapes (+ inflation, + 2% slack); shapes more
  // than 8x the median are classified "large" (tested against everything) so they do not blow the cell up.
  std::vector<float> diam; diam.reserve(s->nA);
  for (auto& r : s->recs) { if (r.flags & ACTOR_REMOVED) continue; const float d = shape_diameter(r); if (std::isfinite(d)) diam.push_back(d); }
  float med = 1.f;
  if (!diam.empty()) { std::vector<float> t = diam; std::nth_element(t.begin(), t.begin() + t.size() / 2, t.end()); med = t[t.size() / 2]; }
  const float largeThresh = 8.f * med;
  bool usesEnv = false; uint32_t maxEnv = 0;
  for (auto& r : s->recs) if (r.envId != NONE32 && !(r.flags & ACTOR_REMOVED)) { usesEnv = true; maxEnv = std::max(maxEnv, r.envId); }
  float cell = 0.f; float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  s->largeHost.clear();
  std::vector<uint32_t> gf(s->nA);
  bool anyCapsule = false, anyBox = false, anyLocks = false, anyConvex = false; uint32_t typeMask = 0;
  for (uint32_t a = 0; a < s->nA; ++a) {
    const ActorRec& r = s->recs[a];
    if (r.flags & ACTOR_REMOVED) { gf[a] = (r.geomType & 0xff) | 0x400u; continue; }   // no dynamic bit, not in the grid, not in the large list
    anyCapsule |= r.geomType == PXB_GEOM_CAPSULE; anyBox |= r.geomType == PXB_GEOM_BOX; anyConvex |= r.geomType == PXB_GEOM_CONVEXMESH;
    if (r.geomType != PXB_GEOM_PLANE) typeMask |= 1u << (r.geomType & 31);
    anyLocks |= ((r.flags >> 8) & 0x3fu) != 0;
    const float d = shape_diameter(r);
    const bool global = !std::isfinite(d) || d > largeThresh || (usesEnv && r.envId == NONE32);
    gf[a] = (r.geomType & 0xff) | ((r.flags & PXB_ACTOR_DYNAMIC) ? 0x100u : 0u) | (global ? 0x200u : 0u) | ((r.flags & PXB_ACTOR_KINEMATIC) ? 0x800u : 0u) | ((r.flags & PXB_ACTOR_DISABLE_GRAVITY) ? 0x1000u : 0u) | ((r.flags & PXB_ACTOR_GYROSCOPIC) ? 0x2000u : 0u) | (((r.flags >> 8) & 0x3fu) << 16);
    anyLocks |= (r.flags & (PXB_ACTOR_DISABLE_GRAVITY | PXB_ACTOR_GYROSCOPIC)) != 0;   // the environment path serves these from its EXT instantiation, like lock flags   // bits 16..21: PxRigidDynamicLockFlags
    if (global) { s->largeHost.push_back(a); continue; }
    cell = std::max(cell, d);
    for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], r.pos[k]); mx[k] = std::max(mx[k], r.pos[k]); }
  }
  cell = (cell + 2.f * std::max(s->desc.contactOffset, s->maxContactOffset)) * 1.02f; if (!(cell > 0.f)) cell = 1.f;
  GridParams g; g.invCell = 1.0f / cell;
  int n[3];
  for (int k = 0; k < 3; ++k) {
    if (!std::isfinite(mn[k])) { mn[k] = 0.f; mx[k] = 0.f; }
    const float span = (mx[k] - mn[k]) + 16.f * cell;  // headroom: objects outside are clamped (still correct, just slower)
    n[k] = std::max(4, (int)std::ceil(span / cell) + 1);
  }
  g.ox = mn[0] - 8.f * cell; g.oy = mn[1] - 8.f * cell; g.oz = mn[2] - 8.f * cell; g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
  const uint64_t envCount = usesEnv ? (uint64_t)maxEnv + 1 : 1;
  // keep the key within 62 bits
  while ((long double)envCount * g.nx * g.ny * g.nz > 4.0e18L) { if (g.nx >= g.ny && g.nx >= g.nz) g.nx = (g.nx + 1) / 2; else if (g.ny >= g.nz) g.ny = (g.ny + 1) / 2; else g.nz = (g.nz + 1) / 2; }
  g.keyBits = bits_for((uint64_t)envCount * (uint64_t)g.nx * (uint64_t)g.ny * (uint64_t)g.nz + 1);
  s->anyLocks = anyLocks;
  { bool anyAgg = false; for (auto& r : s->recs) anyAgg |= r.aggregate != 0 && !(r.flags & ACTOR_REMOVED);
    if (anyAgg) {
      if (!s->aggId && dalloc(s->aggId, s->capA)) { s->abort = true; return; }
      std::vector<uint32_t> ag(s->nA); for (uint32_t a = 0; a < s->nA; ++a) ag[a] = s->recs[a].aggregate;
      cudaMemcpyAsync(s->aggId, ag.data(), 4 * (size_t)s->nA, cudaMemcpyHostToDevice, s->stream); cudaStreamSynchronize(s->stream);
    }
    if (anyAgg != s->anyAggregate) { s->anyAggregate = anyAgg; drop_graphs(s); } }
  s->kinHost.clear(); for (uint32_t a = 0; a < s->nA; ++a) if ((s->recs[a].flags & PXB_ACTOR_KINEMATIC) && !(s->recs[a].flags & ACTOR_REMOVED)) s->kinHost.push_back(a);
  s->anyKinematic = !s->kinHost.empty(); s->nKin = (uint32_t)s->kinHost.size();
  if (s->anyKinematic && !s->kinP) {
    if (dalloc(s->kinList, s->capA) || dalloc(s->kinP, s->capA) || dalloc(s->kinQ, s->capA) || dalloc(s->kinHas, s->capA) || dalloc(s->kinFtv, s->capPairs)) { s->abort = true; return; }
    cudaMemsetAsync(s->kinHas, 0, 4 * (size_t)s->capA, s->stream);
  }
  if (s->nKin) cudaMemcpyAsync(s->kinList, s->kinHost.data(), 4 * (size_t)s->nKin, cudaMemcpyHostToDevice, s->stream);
  s->hasGjkPairs = (anyCapsule && anyBox) || anyConvex;   // k_narrowphase_gjk: capsule-box and hull pairs
  s->anyConvex = anyConvex;
  s->binPairs = __builtin_popcount(typeMask) >= 2 && !getenv("PXB_NO_PAIR_BINS");
  s->grid = g; s->nLarge = (uint32_t)s->largeHost.size();
  s->desc.reserved[0] = (uint32_t)envCount;
  cudaMemcpyAsync(s->geomFlags, gf.data(), 4 * s->nA, cudaMemcpyHostToDevice, s->stream);
  if (s->nLarge) cudaMemcpyAsync(s->largeList, s->largeHost.data(), 4 * s->nLarge, cudaMemcpyHostToDevice, s->stream);
  cudaStreamSynchronize(s->stream);
  s->gridDirty = false;
  rebuild_env(s, usesEnv, maxEnv);
}

// Cooked convex hulls (Gu::ConvexHullData as produced by the host's PxCreateConvexMesh; the reference uploads the same data per shape through
// PxsSimulationController::addPxgShape -> PxgShape::hullOrMeshPtr, PxgConvexConvexShape.h:50-65).  Layout: include/physx_b200.h.
PXB_API int pxb_scene_set_convex_meshes(PxbScene* s, const void* cooked, size_t bytes, uint32_t nHulls) { DeviceGuard dg_(s);
  if (!s || (!cooked && nHulls)) return fail(PXB_ERR_INVALID, "null argument");
  if (s->nHulls) return fail(PXB_ERR_INVALID, "convex meshes are set once per scene, before the actors that use them");
  const uint8_t* q = (const uint8_t*)cooked; const uint8_t* end = q + bytes;
  std::vector<uint4> meta; std::vector<float4> verts, polys; std::vector<uint8_t> refs, edges; std::vector<float> diam;   // host state is committed only after the whole blob is validated
  for (uint32_t h = 0; h < nHulls; ++h) {
    if ((size_t)(end - q) < sizeof(PxbCookedHullHeader)) return fail(PXB_ERR_INVALID, "cooked hull data truncated");
    PxbCookedHullHeader ch; memcpy(&ch, q, sizeof(ch)); q += sizeof(ch);
    // sizes in size_t with every count bounded first (a hull of <= 255 vertices / polygons has at most 3 * 255 edges and 255 * 32 vertex references)
    if (ch.nVerts > 255 || ch.nPolys > 255 || ch.nEdges > 3u * 255u || ch.nIdx > 255u * 32u) return fail(PXB_ERR_INVALID, "cooked hull data out of range");
    const size_t idxBytes = ((size_t)ch.nIdx + 3) / 4 * 4, edgeBytes = (2 * (size_t)ch.nEdges + 3) / 4 * 4;
    const size_t need = (size_t)ch.nVerts * 12 + (size_t)ch.nPolys * sizeof(PxbCookedPoly) + idxBytes + edgeBytes;
    if ((size_t)(end - q) < need) return fail(PXB_ERR_INVALID, "cooked hull data truncated or out of range");
    // the reference's GPU pipeline takes hulls of <= 64 vertices and <= 64 polygons (PxConvexMeshDesc.h:139, cooking with buildGPUData); larger ones fall back to its CPU narrowphase
    if (ch.nVerts > 64 || ch.nPolys > 64) return fail(PXB_ERR_UNSUPPORTED, "convex hulls are limited to 64 vertices and 64 polygons (the reference's GPU-compatible limit)");
    const uint32_t bigSubdiv = ch.reserved[0] & 0xffffu, bigAdj = ch.reserved[0] >> 16;   // Gu::BigConvexRawData: hulls of more than 32 vertices
    const size_t bigBytes = bigSubdiv ? (6 * (size_t)bigSubdiv * bigSubdiv + 3) / 4 * 4 + (size_t)ch.nVerts * 4 + (bigAdj + 3) / 4 * 4 : 0;
    if ((size_t)(end - q) < need + bigBytes) return fail(PXB_ERR_INVALID, "cooked hull data truncated");
    if (ch.nVerts > 32 && !bigSubdiv) return fail(PXB_ERR_INVALID, "a hull of more than 32 vertices needs its hill-climbing data (Gu::BigConvexRawData)");
    meta.push_back(make_uint4((uint32_t)verts.size(), (uint32_t)(polys.size() / 2), (uint32_t)refs.size(), (uint32_t)edges.size()));
    meta.push_back(make_uint4(ch.nVerts, ch.nPolys, ch.nEdges, ch.nIdx));
    uint4 m2; memcpy(&m2.x, &ch.internalExtents[0], 4); memcpy(&m2.y, &ch.internalExtents[1], 4); memcpy(&m2.z, &ch.internalExtents[2], 4); memcpy(&m2.w, &ch.internalRadius, 4);
    meta.push_back(m2);
    uint4 m3; memcpy(&m3.x, &ch.centerOfMass[0], 4); memcpy(&m3.y, &ch.centerOfMass[1], 4); memcpy(&m3.z, &ch.centerOfMass[2], 4); m3.w = bigSubdiv; meta.push_back(m3);
    diam.push_back(2.f * (std::sqrt(ch.boundsCenter[0] * ch.boundsCenter[0] + ch.boundsCenter[1] * ch.boundsCenter[1] + ch.boundsCenter[2] * ch.boundsCenter[2]) +
                                 std::sqrt(ch.boundsExtents[0] * ch.boundsExtents[0] + ch.boundsExtents[1] * ch.boundsExtents[1] + ch.boundsExtents[2] * ch.boundsExtents[2])));
    const float* v = (const float*)q; for (uint32_t i = 0; i < ch.nVerts; ++i) verts.push_back(make_float4(v[i * 3], v[i * 3 + 1], v[i * 3 + 2], 0.f));
    q += (size_t)ch.nVerts * 12;
    for (uint32_t p = 0; p < ch.nPolys; ++p) {
      PxbCookedPoly cp; memcpy(&cp, q, sizeof(cp)); q += sizeof(cp);
      if (cp.nbVerts < 3 || cp.nbVerts > 32 || (size_t)cp.vref + cp.nbVerts > ch.nIdx || cp.minIndex >= ch.nVerts) return fail(PXB_ERR_INVALID, "hull polygon out of range (3..32 vertices per polygon)");
      polys.push_back(make_float4(cp.plane[0], cp.plane[1], cp.plane[2], cp.plane[3]));
      float4 m; memcpy(&m.x, &cp.vref, 4); memcpy(&m.y, &cp.nbVerts, 4); memcpy(&m.z, &cp.minIndex, 4); m.w = 0.f; polys.push_back(m);
    }
    for (uint32_t i = 0; i < ch.nIdx; ++i) if (q[i] >= ch.nVerts) return fail(PXB_ERR_INVALID, "hull vertex reference out of range");
    refs.insert(refs.end(), q, q + ch.nIdx); q += idxBytes;
    for (uint32_t i = 0; i < 2 * ch.nEdges; ++i) if (q[i] >= ch.nPolys) return fail(PXB_ERR_INVALID, "hull edge face out of range");
    edges.insert(edges.end(), q, q + edgeBytes); q += edgeBytes;   // padded: every hull starts 4-byte aligned
    if (bigSubdiv) {   // samples | valencies | adjacent vertices, as they lie in the cooked section (load_hull / gjk_hull_hill_climb)
      const uint8_t* val = q + (6 * (size_t)bigSubdiv * bigSubdiv + 3) / 4 * 4; const uint8_t* adj = val + (size_t)ch.nVerts * 4;
      for (uint32_t i = 0; i < 6 * bigSubdiv * bigSubdiv; ++i) if (q[i] >= ch.nVerts) return fail(PXB_ERR_INVALID, "hill-climbing sample out of range");
      for (uint32_t i = 0; i < ch.nVerts; ++i) { uint16_t v2[2]; memcpy(v2, val + 4 * i, 4); if ((uint32_t)v2[0] + v2[1] > bigAdj) return fail(PXB_ERR_INVALID, "hill-climbing valency out of range"); }
      for (uint32_t i = 0; i < bigAdj; ++i) if (adj[i] >= ch.nVerts) return fail(PXB_ERR_INVALID, "hill-climbing neighbour out of range");
      edges.insert(edges.end(), q, q + bigBytes); q += bigBytes;
    }
  }
  if (!nHulls) { s->hullDiam.clear(); return PXB_OK; }
  s->hullDiam = diam;
  CK(dalloc(s->hullMeta, meta.size())); CK(dalloc(s->hullVerts, verts.size())); CK(dalloc(s->hullPolys, polys.size())); CK(dalloc(s->hullRefs, refs.size() + 4)); CK(dalloc(s->hullEdges, edges.size() + 4));
  CK(cudaMemcpyAsync(s->hullMeta, meta.data(), 16 * meta.size(), cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->hullVerts, verts.data(), 16 * verts.size(), cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->hullPolys, polys.data(), 16 * polys.size(), cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->hullRefs, refs.data(), refs.size(), cudaMemcpyHostToDevice, s->stream));
  if (!edges.empty()) CK(cudaMemcpyAsync(s->hullEdges, edges.data(), edges.size(), cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  s->nHulls = nHulls;
  return PXB_OK;
}

PXB_API int pxb_scene_set_materials(PxbScene* s, const PxbMaterial* m, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || (nb && !m)) return fail(PXB_ERR_INVALID, "null argument");
  if (s->nA) return fail(PXB_ERR_INVALID, "the material table is set before the actors that refer to it");
  std::vector<float4> tab(nb);
  for (uint32_t i = 0; i < nb; ++i) {
    if (m[i].restitution < 0.f) return fail(PXB_ERR_UNSUPPORTED, "compliant contacts (negative restitution) are not supported");
    if ((m[i].bits & 15u) > 3u || ((m[i].bits >> 4) & 15u) > 3u || (m[i].bits >> 9)) return fail(PXB_ERR_UNSUPPORTED, "unknown combine mode or unsupported material flag (only eDISABLE_FRICTION)");
    float w; memcpy(&w, &m[i].bits, 4); tab[i] = make_float4(m[i].staticFriction, m[i].dynamicFriction, m[i].restitution, w);
  }
  if (s->matTab) { cudaFree(s->matTab); s->matTab = nullptr; }
  s->nMaterials = 0;
  if (!nb) return PXB_OK;
  CK(dalloc(s->matTab, nb)); CK(cudaMemcpyAsync(s->matTab, tab.data(), 16 * (size_t)nb, cudaMemcpyHostToDevice, s->stream)); CK(cudaStreamSynchronize(s->stream));
  s->nMaterials = nb; drop_graphs(s);
  return PXB_OK;
}
PXB_API int pxb_scene_add_actors(PxbScene* s, const void* recsIn, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || !recsIn) return fail(PXB_ERR_INVALID, "null argument");
  if (s->abort) return fail(PXB_ERR_CUDA, "scene is in abort mode");
  if (s->nA + nb > s->capA) return fail(PXB_ERR_CAPACITY, "maxActors exceeded");
  const ActorRec* in = (const ActorRec*)recsIn;
  const uint32_t base = s->nA;
  std::vector<float4> pos(nb), quat(nb), lin(nb), ang(nb), inv(nb), dmp(nb), dims(nb); std::vector<uint32_t> env(nb), mat(nb);
  for (uint32_t i = 0; i < nb; ++i) {   // first pass: every record is checked before any host state changes (a rejected batch leaves the scene untouched)
    const ActorRec& r = in[i];
    if (r.geomType == PXB_GEOM_CONVEXMESH) { if (r.hullIdx >= s->nHulls) return fail(PXB_ERR_UNSUPPORTED, "convex actor without a cooked hull: call pxb_scene_set_convex_meshes first"); }
    else if (r.geomType != PXB_GEOM_BOX && r.geomType != PXB_GEOM_PLANE && r.geomType != PXB_GEOM_SPHERE && r.geomType != PXB_GEOM_CAPSULE)
      return fail(PXB_ERR_UNSUPPORTED, "geometry type not supported yet");
    if (r.flags & PXB_ACTOR_KINEMATIC) {
      if (!(r.flags & PXB_ACTOR_DYNAMIC)) return fail(PXB_ERR_INVALID, "PXB_ACTOR_KINEMATIC is a flag of dynamic actors (PxRigidBodyFlag::eKINEMATIC)");
      if (s->sleepThreshold > 0.f) return fail(PXB_ERR_UNSUPPORTED, "kinematic bodies in scenes with sleeping enabled are not built");
      if (r.geomType == PXB_GEOM_PLANE) return fail(PXB_ERR_INVALID, "planes are static");
    }
    if ((r.aggregate & 0x7fffffffu) >= 0x40000000u) return fail(PXB_ERR_INVALID, "aggregate ids are 1 .. 2^30 - 1 (bit 31 = self collisions)");
  }
  for (uint32_t i = 0; i < nb; ++i) {
    const ActorRec& r = in[i];
    const bool kin = (r.flags & PXB_ACTOR_KINEMATIC) != 0;
    const bool dyn = (r.flags & PXB_ACTOR_DYNAMIC) && !kin;   // mass properties: a kinematic body has none (infinite mass and inertia)
    s->recs.push_back(r);
    if (r.geomType == PXB_GEOM_CONVEXMESH) s->recs.back().dims[3] = s->hullDiam[r.hullIdx];   // bounding diameter for the broadphase grid
    if (dyn || kin) { s->dynIndex.push_back((int)s->nDyn); s->dynActor.push_back(base + i); s->nDyn++; } else s->dynIndex.push_back(-1);   // kinematic bodies keep their place in the dynamic-body order (they are PxRigidDynamic)
    const float invMass = (dyn && r.mass > 0.f) ? 1.0f / r.mass : 0.f;
    pos[i] = make_float4(r.pos[0], r.pos[1], r.pos[2], invMass);
    { const float sN = 1.0f / sqrtf(r.quat[0] * r.quat[0] + r.quat[1] * r.quat[1] + r.quat[2] * r.quat[2] + r.quat[3] * r.quat[3]);   // NpPhysics::createRigidDynamic / createRigidStatic store globalPose.getNormalized()
      quat[i] = make_float4(r.quat[0] * sN, r.quat[1] * sN, r.quat[2] * sN, r.quat[3] * sN); }
    lin[i] = make_float4(r.linVel[0], r.linVel[1], r.linVel[2], 0.f); ang[i] = make_float4(r.angVel[0], r.angVel[1], r.angVel[2], 0.f);
    inv[i] = make_float4(dyn && r.inertia[0] > 0.f ? 1.0f / r.inertia[0] : 0.f, dyn && r.inertia[1] > 0.f ? 1.0f / r.inertia[1] : 0.f, dyn && r.inertia[2] > 0.f ? 1.0f / r.inertia[2] : 0.f, r.maxDepenetrationVel);
    dmp[i] = make_float4(r.linDamping, r.angDamping, r.maxLinVel * r.maxLinVel, r.maxAngVel * r.maxAngVel);
    dims[i] = make_float4(r.dims[0], r.dims[1], r.dims[2], r.dims[3]); env[i] = r.envId; mat[i] = (s->nMaterials && r.materialIndex < s->nMaterials) ? r.materialIndex : 0u;
    if (r.geomType == PXB_GEOM_CONVEXMESH) memcpy(&dims[i].x, &r.hullIdx, 4);   // convex actors carry their hull index where the primitives carry their size
  }
  s->nA += nb;
  CK(cudaMemcpyAsync(s->pos + base, pos.data(), 16 * nb, cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->quat + base, quat.data(), 16 * nb, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->linVel + base, lin.data(), 16 * nb, cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->angVel + base, ang.data(), 16 * nb, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->invInertia + base, inv.data(), 16 * nb, cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->damp + base, dmp.data(), 16 * nb, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->dims + base, dims.data(), 16 * nb, cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->envId + base, env.data(), 4 * nb, cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->actorMat + base, mat.data(), 4 * nb, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->dynActorDev, s->dynActor.data(), 4 * s->nDyn, cudaMemcpyHostToDevice, s->stream));
  CK(cudaMemcpyAsync(s->counters + C_NA, &s->nA, 4, cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  s->gridDirty = true;
  return PXB_OK;
}

// Bp::AABBManagerBase::removeBounds + PxsSimulationController::removeDynamic (BpAABBManagerBase.h:191): the actors leave the simulation at the next
// step -- their pairs are reported deleted (and touch-lost), they are no longer integrated -- while every actor / dynamic-body index stays valid
// (PxRigidDynamicGPUIndex values are stable node indices in the reference too).  Their last state remains readable.
PXB_API int pxb_scene_remove_actors(PxbScene* s, const uint32_t* actors, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || (nb && !actors)) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  for (uint32_t i = 0; i < nb; ++i) if (actors[i] >= s->nA) return fail(PXB_ERR_INVALID, "actor index out of range");
  for (uint32_t i = 0; i < nb; ++i) s->recs[actors[i]].flags |= ACTOR_REMOVED;
  if (nb) { s->gridDirty = true; drop_graphs(s); }
  return PXB_OK;
}
PXB_API int pxb_scene_set_constraint_order(PxbScene* s, const uint32_t* pairs, uint32_t n) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null scene");
  s->nOrder = 0;
  if (!n) return PXB_OK;
  if (!pairs) return fail(PXB_ERR_INVALID, "null pairs");
  if (n > s->capOrder) { if (s->orderKeys) cudaFree(s->orderKeys); s->capOrder = std::max(n, s->capPairs); CK(dalloc(s->orderKeys, s->capOrder)); }
  std::vector<uint64_t> keys(n);
  for (uint32_t k = 0; k < n; ++k) { const uint32_t a = std::min(pairs[2 * k], pairs[2 * k + 1]), b = std::max(pairs[2 * k], pairs[2 * k + 1]); keys[k] = ((uint64_t)a << s->bitsA) | b; }
  CK(cudaMemcpyAsync(s->orderKeys, keys.data(), 8 * n, cudaMemcpyHostToDevice, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  s->nOrder = n;
  return PXB_OK;
}

#define LAUNCH(kernel, grid, block, ...) do { kernel<<<(grid), (block), 0, st>>>(__VA_ARGS__); s->launches++; } while (0)

static void drop_graphs(PxbScene* s);
// Chooses between the environment path and the device-wide path for the coming step.  The pair list layout differs
// (per-environment segments vs one sorted list); env -> global is converted by one radix sort, the opposite direction
// is not needed (a scene that has stepped on the device-wide path stays there).
static int select_path(PxbScene* s) {
  const bool want = s->envEligible && !s->envDisabled && s->nOrder == 0;
  if (want == s->envActive) return PXB_OK;
  if (want) {
    if (s->everStepped) { s->envDisabled = true; return PXB_OK; }
    s->envActive = true; drop_graphs(s);
    return PXB_OK;
  }
  // env -> global: sort (key, slot) of the current list
  cudaStream_t st = s->stream; const int cur = s->cur;
  const int r = radix_sort_pairs(s->pairKeys[cur], s->pairSlots[cur], s->pairKeyAlt, s->pairValAlt, s->nPairsDev + cur, 2 * s->bitsA, s->rsTmp, st);
  if (r) { CK(cudaMemcpyAsync(s->pairKeys[cur], s->pairKeyAlt, 8 * (size_t)s->capPairs, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(s->pairSlots[cur], s->pairValAlt, 4 * (size_t)s->capPairs, cudaMemcpyDeviceToDevice, st)); }
  s->envActive = false; s->envDisabled = true; drop_graphs(s);
  return PXB_OK;
}

// stage 1: bounds + broadphase + pair lifecycle
static int run_broadphase(PxbScene* s, bool externalTight) {
  cudaStream_t st = s->stream;
  if (s->gridDirty) rebuild_grid(s);
  const uint32_t nA = s->nA, B = 256;
  const int prev = s->cur; s->cur ^= 1; const int cur = s->cur;
  if (s->envActive) {   // environment path: one warp per environment does bounds + pair finding + pair lifecycle
    LAUNCH(k_env_begin, 1, 32, s->counters);
    EnvBpArgs A;
    A.nEnv = s->nEnv; A.maxList = s->envMaxList; A.bitsA = s->bitsA; A.cap = s->capPairs; A.ringMask = s->ringMask; A.externalTight = externalTight ? 1 : 0; A.contactOffset = s->desc.contactOffset;
    A.envStart = s->envStart; A.envList = s->envList; A.pos = s->pos; A.quat = s->quat; A.dims = s->dims; A.geomFlags = s->geomFlags; A.envId = s->envId; A.tight = s->tight; A.hulls = hull_arrays(s); A.L = local_poses(s); A.aggId = s->anyAggregate ? s->aggId : nullptr; A.shapeOff = s->hasShapeOff ? s->shapeOff : nullptr;
    A.candKeys = s->candKeys; A.candCount = s->candCount; A.refMin = s->candRefMin; A.refMax = s->candRefMax; A.candCap = s->candCap;
    { const float co = std::max(s->desc.contactOffset, s->maxContactOffset); A.candMargin = 2.f * (co > 0.f ? co : 0.01f); }
    A.oldKeys = s->pairKeys[prev]; A.oldSlots = s->pairSlots[prev]; A.oldSeg = s->envSeg[prev]; A.newKeys = s->pairKeys[cur]; A.newSlots = s->pairSlots[cur]; A.newSeg = s->envSeg[cur];
    A.counters = s->counters; A.freeRing = s->freeList; A.createdKeys = s->createdKeys; A.deletedKeys = s->deletedKeys; A.manifolds = s->manifolds; A.frictions = s->frictions; A.slotColour = s->slotColour; A.touch = touch_lists(s);
    const size_t smem = (size_t)ENV_BP_WARPS * (s->envMaxList * (2 * sizeof(float4) + sizeof(uint32_t)) + ENV_BP_STAGE * sizeof(uint64_t));
    pxb_launch_env_bp(st, A, s->anyConvex, smem);
    s->launches++;
    LAUNCH(k_clamp_count, 1, 32, s->counters, s->capPairs, s->nPairsDev + cur);
    return PXB_OK;
  }
  CK(cudaMemsetAsync(s->counters + C_NPAIRS_NEW, 0, 4 * 3, st));  // NPAIRS_NEW, NCREATED, NDELETED
  CK(cudaMemsetAsync(s->counters + C_NTOUCH_FOUND, 0, 4 * 6, st));   // NTOUCH_FOUND, NTOUCH_LOST, NGJK_QUERY, NGJK_FULL, NGJK_EPA, NBOXGEN
  if (s->hasGjkPairs) CK(cudaMemsetAsync(s->counters + C_NGJK, 0, 4, st));
  LAUNCH(k_bounds, cdiv(nA, B), B, nA, s->pos, s->quat, s->dims, s->geomFlags, s->envId, s->desc.contactOffset, s->tight, externalTight ? 1 : 0, s->aabbMin, s->aabbMax, s->grid,
         s->desc.reserved[0], s->cellKey, s->cellVal, hull_arrays(s), local_poses(s), s->hasShapeOff ? s->shapeOff : (const float2*)nullptr);
  const int r = radix_sort_pairs(s->cellKey, s->cellVal, s->cellKeyAlt, s->cellValAlt, s->counters + C_NA, s->grid.keyBits, s->rsTmp, st);
  s->launches += 3 * ((s->grid.keyBits + 7) / 8);
  const uint64_t* sk = r ? s->cellKeyAlt : s->cellKey; const uint32_t* sv = r ? s->cellValAlt : s->cellVal;
  LAUNCH(k_bp_gather, cdiv(nA, B), B, nA, sv, s->aabbMin, s->aabbMax, s->sMin, s->sMax);
  // the emission buffer is chosen so that the last radix pass lands in pairKeys[cur] (no pointer swaps: the
  // launch sequence only depends on the parity `cur`, which keeps it CUDA-graph capturable)
  const bool oddPasses = (((2 * s->bitsA + 7) / 8) & 1u) != 0;
  uint64_t* emit = oddPasses ? s->pairKeyAlt : s->pairKeys[cur];
  uint64_t* other = oddPasses ? s->pairKeys[cur] : s->pairKeyAlt;
  LAUNCH(k_bp_pairs<EngineFilter>, cdiv(nA, 128), 128, nA, sk, sv, s->sMin, s->sMax, s->grid, s->bitsA, emit, s->counters, s->capPairs, EngineFilter{s->anyAggregate ? s->aggId : nullptr});
  if (s->nLarge) LAUNCH(k_bp_large<EngineFilter>, cdiv(nA, B), B, nA, s->nLarge, s->largeList, s->aabbMin, s->aabbMax, s->bitsA, emit, s->counters, s->capPairs, EngineFilter{s->anyAggregate ? s->aggId : nullptr});
  LAUNCH(k_clamp_count, 1, 32, s->counters, s->capPairs, s->nPairsDev + cur);
  radix_sort_pairs(emit, s->pairValTmp, other, s->pairValAlt, s->nPairsDev + cur, 2 * s->bitsA, s->rsTmp, st);
  s->launches += 3 * ((2 * s->bitsA + 7) / 8);
  const uint32_t gP = cdiv(s->capPairs, B);
  LAUNCH(k_pair_lost, gP, B, s->pairKeys[prev], s->pairSlots[prev], s->nPairsDev + prev, s->pairKeys[cur], s->nPairsDev + cur, s->freeList, s->ringMask, s->deletedKeys, s->counters, touch_lists(s));
  LAUNCH(k_pair_found, gP, B, s->pairKeys[prev], s->pairSlots[prev], s->nPairsDev + prev, s->pairKeys[cur], s->pairSlots[cur], s->nPairsDev + cur, s->freeList, s->ringMask, s->createdKeys, s->counters,
         s->manifolds, s->frictions, s->touchState);
  return PXB_OK;
}

static int read_counters(PxbScene* s) {
  CK(cudaMemcpyAsync(s->hostCounters, s->counters, 4 * C_COUNT, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaMemcpyAsync(s->hostCounters + C_COUNT, s->nPairsDev, 8, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  s->hNPairs = s->hostCounters[C_COUNT + s->cur]; s->hNCreated = s->hostCounters[C_NCREATED]; s->hNDeleted = s->hostCounters[C_NDELETED];
  s->hNCon = s->hostCounters[C_NCON]; s->hNPart = s->hostCounters[C_NPART]; s->hErr = s->hostCounters[C_ERROR];
  s->hNTouchFound = s->hostCounters[C_NTOUCH_FOUND]; s->hNTouchLost = s->hostCounters[C_NTOUCH_LOST];
  if (s->envActive) {   // fit k_env_solve to the largest environment seen: one thread per constraint (rows in registers), lists in shared memory
    s->hMaxConEnv = s->hostCounters[C_MAXCONENV]; s->hMaxPairEnv = s->hostCounters[C_MAXPAIRENV];
    if (!s->envThreadsForced) { const uint32_t t = env_threads_for(s->hMaxConEnv); if (t != s->envSolveThreads && s->hMaxConEnv) { s->envSolveThreads = t; drop_graphs(s); } }
    if (!s->envConCapForced) {
      const uint32_t want = std::min(env_con_cap_limit(s->envMaxList), std::max(32u, (s->hMaxPairEnv + 3u) & ~3u));
      if (s->hMaxPairEnv > s->envConCap ? want != s->envConCap : want * 2 <= s->envConCap) { s->envConCap = want; drop_graphs(s); }
    }
  }
  if (s->hErr) CK(cudaMemsetAsync(s->counters + C_ERROR, 0, 4, s->stream));   // an error is reported by the fetchResults of the step that raised it, not by every later one
  if (s->hErr & E_PAIR_OVERFLOW) return fail(PXB_ERR_CAPACITY, "broadphase pair capacity (maxPairs) exceeded");
  if (s->hErr & (E_COLOUR_OVERFLOW | E_PARTITION_OVERFLOW)) return fail(PXB_ERR_CAPACITY, "more than 64 dynamic colours / 160 partitions needed");
  if (s->hErr & E_UNSUPPORTED_PAIR) return fail(PXB_ERR_UNSUPPORTED, "a pair of an unsupported geometry type came into contact range");
  if (s->hErr & E_BAD_INDEX) return fail(PXB_ERR_INVALID, "a kinematic target named a body that does not exist or is not kinematic");
  return PXB_OK;
}

// phase 0: the whole step; 1: bounds + broadphase + narrowphase (+ island sleep decisions) only; 2: the rest (s->cur already switched by phase 1)
static int enqueue_step(PxbScene* s, float dt, int phase = 0) {
  cudaStream_t st = s->stream; const uint32_t B = 256;
  s->launches = 0;
#define MARK(i) do { if (s->profiling) CK(cudaEventRecord(s->ev[i], st)); } while (0)
  if (s->nKin && phase != 1) LAUNCH(k_kin_setup, cdiv(s->nKin, 128), 128, s->nKin, s->kinList, s->pos, s->quat, s->kinP, s->kinQ, s->kinHas, s->linVel, s->angVel, 1.0f / dt);
  if (s->bodyAccel && phase != 1) {   // velocities this step starts from (integrationTGS.cu:124-131 keeps them in mBodySimPrevVelocities); after the join of a stream-ordered velocity write
    CK(cudaMemcpyAsync(s->prevLin, s->linVel, 16 * (size_t)s->nA, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(s->prevAng, s->angVel, 16 * (size_t)s->nA, cudaMemcpyDeviceToDevice, st));
  }
  if (phase != 2) {
    MARK(0);
    int rc = run_broadphase(s, false); if (rc) return rc;
    MARK(1);
  }
  const int cur = s->cur; const uint32_t gP = cdiv(s->capPairs, B);
  const uint32_t* nP = s->nPairsDev + cur;
  const float contactDist = s->desc.contactOffset + s->desc.contactOffset;
  if (phase != 2) {
  if (s->binPairs) {   // several geometry types: counting sort of the pairs by type pair, so that warps do not diverge across contact functions
    CK(cudaMemsetAsync(s->npClassCount, 0, 4 * 2 * NP_CLASSES, st));
    LAUNCH(k_np_class_count, gP, B, s->pairKeys[cur], nP, s->bitsA, s->geomFlags, s->npClass, s->npClassCount, s->pairSlots[cur], s->manifolds);
    LAUNCH(k_np_class_scatter, gP, B, nP, s->npClass, s->npClassCount, s->npClassCount + NP_CLASSES, s->pairOrder);
  }
  NpArgs NA;
  NA.pairKeys = s->pairKeys[cur]; NA.pairSlots = s->pairSlots[cur]; NA.nPairsP = nP; NA.bitsA = s->bitsA; NA.pos = s->hasLocal ? s->tcPos : s->pos; NA.quat = s->hasLocal ? s->tcQuat : s->quat; /* shape world poses: the transform cache when the scene has local poses */ NA.dims = s->dims; NA.geomFlags = s->geomFlags;
  NA.contactDist = contactDist; NA.toleranceLength = s->desc.toleranceLength; NA.manifolds = s->manifolds; NA.cHdr = s->cHdr; NA.cPts = s->cPts; NA.pairBodies = s->pairBodies; NA.conFlag = s->conFlag;
  NA.cForce = s->cForce; NA.counters = s->counters; NA.gjkList = s->gjkList; NA.gjkQuery = s->gjkQuery; NA.gjkFull = s->gjkFull; NA.gjkEpa = s->gjkEpa; NA.boxList = (s->boxPhases && (!s->envActive || s->boxPhasesEnv)) ? s->boxList : nullptr; NA.pairOrder = s->binPairs ? s->pairOrder : (const uint32_t*)nullptr; NA.hulls = hull_arrays(s); NA.touch = touch_lists(s); NA.filter.data = s->hasFilter ? s->filterData : nullptr; NA.filter.cfg = s->filterCfg; NA.filter.shapeOff = s->hasShapeOff ? s->shapeOff : nullptr;
  pxb_launch_narrowphase(st, s->capPairs, NA); s->launches += NA.boxList ? 2 : 1;
  if (s->hasGjkPairs) {
    const uint32_t ctas = std::max(148u * 4u, std::min(cdiv(s->capPairs, 128), 148u * 64u));
    if (s->anyConvex && s->gjkPhases) { pxb_launch_narrowphase_gjk_phases(st, ctas, NA); s->launches += 4; }   // hull scenes: refresh -> query -> EPA -> manifold over compacted worklists
    else { pxb_launch_narrowphase_gjk(st, ctas, NA); s->launches++; }
  }
  }
  SleepArgs SA; SA.threshold = s->sleepThreshold; SA.dt = dt; SA.wake = s->wake; SA.accLin = s->accLin; SA.accAng = s->accAng; SA.asleep = s->asleep; SA.nInter = s->nInter;
  if (phase != 2 && s->sleepThreshold > 0.f) {   // island sleep / wake decisions for this step (needs this frame's touching pairs)
    uint32_t nA = s->nA;
    void* args[] = {&nA, &nP, &s->pairBodies, &s->cHdr, &s->geomFlags, &SA, &s->islandLabel, &s->islandAwake, &s->counters, &s->linVel, &s->angVel, &s->conFlag, &s->deletedKeys, &s->bitsA};
    CK(cudaLaunchCooperativeKernel((void*)k_sleep_islands, dim3(s->coopBlocksSleep), dim3(256), args, 0, st)); s->launches++;
  }
  if (phase == 1) { CK(cudaGetLastError()); return PXB_OK; }
  MARK(2);
  const float* g = s->desc.gravity; const bool pgs = s->desc.solverType == PXB_SOLVER_PGS;
  SolverParams P;
  P.dt = dt; P.stepDt = dt / (float)s->desc.posIters; P.invStepDt = 1.f / P.stepDt; P.invTotalDt = 1.0f / dt;
  P.biasCoefficient = 2.f * sqrtf(1.f / (float)s->desc.posIters);
  P.bounceThreshold = -s->desc.bounceThresholdVelocity;  // Sc::Scene hands the solver the negated threshold (ScScene.cpp:1002)
  P.frictionOffsetThreshold = s->desc.frictionOffsetThreshold; P.correlationDistance = s->desc.frictionCorrelationDistance;
  P.restDistance = s->desc.restOffset + s->desc.restOffset; P.staticFriction = s->desc.staticFriction; P.dynamicFriction = s->desc.dynamicFriction; P.restitution = s->desc.restitution;
  if (s->envActive) {   // environment path: colouring + prep + all solver iterations + integration in one launch, rows in shared memory
    MARK(3); MARK(4);
    EnvSolveArgs A;
    A.nEnv = s->nEnv; A.maxList = s->envMaxList; A.conCap = s->envConCap; A.cap = s->capPairs; A.posIters = s->desc.posIters; A.velIters = s->desc.velIters;
    A.dt = dt; A.gx = g[0]; A.gy = g[1]; A.gz = g[2]; A.P = P;
    A.envStart = s->envStart; A.envList = s->envList; A.actorLocal = s->actorLocal; A.seg = s->envSeg[cur];
    A.pos = s->pos; A.quat = s->quat; A.linVel = s->linVel; A.angVel = s->angVel; A.invInertia = s->invInertia; A.damp = s->damp; A.geomFlags = s->geomFlags; A.anyLocks = s->anyLocks ? 1u : 0u; A.extForce = s->forcesUsed ? s->extForce : nullptr; A.extTorque = s->forcesUsed ? s->extTorque : nullptr;
    A.pairSlots = s->pairSlots[cur]; A.pairBodies = s->pairBodies; A.cHdr = s->cHdr; A.cPts = s->cPts; A.cForce = s->cForce; A.frictions = s->frictions; A.frReport = s->contactData ? s->frReport : nullptr;
    A.conPair = s->conPair; A.conB0 = s->conB0; A.conB1 = s->conB1; A.conColour = s->conColour; A.ordered = s->ordered; A.broken = s->conDone;
    A.rowScratch = s->ptA;   // ptA|ptB|ptC|frA|frB|frC|frD are ONE allocation of 28 x cap float4 (scene_alloc); the environment path uses 25 of them
    A.counters = s->counters; A.timing = s->envTiming; A.slotColour = s->slotColour; A.S = SA;
    const bool fusedExport = s->exportOn && s->envDynContiguous && !s->nKin;   // kinematic bodies take their new pose after the kernel (k_kin_finalize): export afterwards
    A.exportTab = fusedExport ? s->exportTab : nullptr; A.envDyn = s->envDyn; A.dynActor = s->dynActorDev;
    size_t smem = env_solve_smem(s->envMaxList, s->envConCap, s->envSolveThreads);
    { static const char* pad = getenv("PXB_ENV_SMEM_PAD"); if (pad) smem += (size_t)atoi(pad); }   // experiment hook: padding the request lowers the CTAs per SM (wave quantisation A/B, profiles/README.md)
    A.M = material_args(s); A.kinFtv = (s->nKin && !pgs) ? s->kinFtv : nullptr; A.kinOn = s->nKin ? 1u : 0u;
    const bool ext = s->anyLocks || s->forcesUsed || s->nMaterials != 0 || s->hasShapeOff || s->nKin != 0;   // lock flags / external forces: the EXT instantiation; the plain one carries none of that code
    pxb_launch_env_solve(st, A, s->envSolveThreads, pgs, ext, smem);
    s->launches++;
    if (s->nKin) LAUNCH(k_kin_finalize, cdiv(s->nKin, 128), 128, s->nKin, s->kinList, s->pos, s->quat, s->kinP, s->kinQ, s->kinHas);
    if (s->exportOn && !fusedExport) LAUNCH(k_states_export, cdiv(s->nDyn, 256), 256, s->nDyn, s->dynActorDev, s->pos, s->quat, s->linVel, s->angVel, s->exportTab);
    MARK(5); MARK(6);
    CK(cudaGetLastError());
    return PXB_OK;
  }
  if (s->nOrder) {
    LAUNCH(k_rank_init, gP, B, nP, s->rankOfPair);
    LAUNCH(k_rank_map, cdiv(s->nOrder, B), B, s->nOrder, s->orderKeys, s->pairKeys[cur], nP, s->rankOfPair);
    LAUNCH(k_flag_ordered, gP, B, nP, s->rankOfPair, s->conFlag);
  }
  exclusive_scan_u32(s->conFlag, s->conIdx, nP, s->counters + C_NCON, s->scanSums, s->rsTmp.ctas, st); s->launches += 3;
  LAUNCH(k_compact, gP, B, nP, s->conFlag, s->conIdx, s->conPair, s->pairSlots[cur], s->cHdr, s->frictions);
  if (s->nOrder) {
    LAUNCH(k_con_sortkeys, gP, B, s->counters, s->conPair, s->rankOfPair, s->conSortKey);
    const int r = radix_sort_pairs(s->conSortKey, s->conPair, s->conSortKeyAlt, s->conPairAlt, s->counters + C_NCON, 32, s->rsTmp, st); s->launches += 12;
    if (r) std::swap(s->conPair, s->conPairAlt);
  }
  CK(cudaMemsetAsync(s->bodyCnt, 0, 4 * s->nA, st)); CK(cudaMemsetAsync(s->bodyCursor, 0, 4 * s->nA, st));
  CK(cudaMemsetAsync(s->counters + C_NPART, 0, 8, st));  // NPART, REMAINING
  LAUNCH(k_con_bodies, gP, B, s->counters, s->counters, s->conPair, s->pairBodies, s->geomFlags, s->conB0, s->conB1, s->conDone, s->bodyCnt);
  exclusive_scan_u32(s->bodyCnt, s->bodyStart, s->counters + C_NA, s->counters + C_NDYNCON, s->scanSums, s->rsTmp.ctas, st); s->launches += 3;
  LAUNCH(k_con_fill, gP, B, s->counters, s->conB0, s->conB1, s->bodyStart, s->bodyCursor, s->bodyList);
  LAUNCH(k_body_lists, cdiv(s->nA, B), B, s->nA, s->bodyStart, s->bodyCnt, s->bodyList, s->conB0, s->conB1, s->conPos0, s->conPos1, s->bodyNext, s->bodyMask, s->bodyHasCon);
  {
    int relaxed = (s->relaxedPartitioning && s->nOrder == 0) ? 1 : 0;   // a host-provided solver order always gets the exact first-fit
    if (relaxed) CK(cudaMemsetAsync(s->bodyBest, 0, 8 * (size_t)s->nA, st));
    else if (!s->colourLegacy) {   // exact first-fit as a dataflow (k_colour_firstfit); the cooperative kernel below then only orders the constraints partition-major
      CK(cudaMemsetAsync(s->colourTicket, 0, 8, st)); CK(cudaMemsetAsync(s->colourTicket + 2, s->colourPrefix ? 0xff : 0, 4, st));
      if (s->colourPrefix) {
        LAUNCH(k_colour_prefix_find, gP, B, s->counters, s->conB0, s->conB1, s->prevB0, s->prevB1, s->prevNCon, s->colourTicket);
        LAUNCH(k_colour_prefix_apply, gP, B, s->conB0, s->conB1, s->prevColour, s->conColour, s->bodyMask, s->colourTicket);
      }
      // resident window: every resident thread polls L2 while it waits, so the number of resident CTAs is capped with a dynamic shared-memory
      // request (colourWindow CTAs per SM); blocks behind the window wait in the hardware queue without polling
      k_colour_firstfit<<<cdiv(s->capPairs, 128), 128, s->colourWindow ? (size_t)(200 * 1024) / s->colourWindow : 0, st>>>(s->counters, s->conB0, s->conB1, s->conPos0, s->conPos1, s->conColour, s->bodyMask,
                                                                                                                              s->colourTicket, s->colourBackoffNs); s->launches++;
      relaxed = 2;
    }
    void* args[] = {&s->counters, &s->conB0, &s->conB1, &s->conPos0, &s->conPos1, &s->conColour, &s->conDone, &s->bodyNext, &s->bodyMask, &s->partCnt, &s->partStart, &s->partCursor, &s->ordered,
                    &s->bodyBest, &relaxed};
    CK(cudaLaunchCooperativeKernel((void*)k_colour_partition, dim3(s->coopBlocksColour), dim3(256), args, 0, st)); s->launches++;
    if (relaxed == 2 && s->colourPrefix) {   // remember this frame's list and colours for the next frame's prefix reuse
      CK(cudaMemcpyAsync(s->prevB0, s->conB0, 4 * (size_t)s->capPairs, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(s->prevB1, s->conB1, 4 * (size_t)s->capPairs, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(s->prevColour, s->conColour, 4 * (size_t)s->capPairs, cudaMemcpyDeviceToDevice, st));
      LAUNCH(k_colour_remember, 1, 32, s->counters, s->prevNCon);
    }
  }
  MARK(3);
  LAUNCH(k_preintegrate, cdiv(s->nDyn, B), B, s->nDyn, s->dynActorDev, s->pos, s->quat, s->linVel, s->angVel, s->invInertia, s->damp, g[0], g[1], g[2], dt, s->sbLin, s->sbAng, s->sbDLin, s->sbDAng,
         s->sbIA, s->sbIB, s->sbP, s->sbQ, s->sbOrigAng, pgs ? 1 : 0, SA, s->geomFlags, s->forcesUsed ? s->extForce : (float4*)nullptr, s->forcesUsed ? s->extTorque : (float4*)nullptr);
  {   // rows in the same 25-float4 record image as the environment path (pxb_env.cuh RegRows); PGS: velocity-delta solver bodies (pxb_pgs.cuh)
    Rows R; R.f = s->ptA; R.broken = s->conDone; R.stride = s->capPairs;
    PrepArgs PA;
    PA.counters = s->counters; PA.ordered = s->ordered; PA.conPair = s->conPair; PA.pairSlots = s->pairSlots[cur]; PA.pairBodies = s->pairBodies; PA.geomFlags = s->geomFlags; PA.cHdr = s->cHdr; PA.cPts = s->cPts;
    PA.pos = s->pos; PA.quat = s->quat; PA.linVel = s->linVel; PA.sbOrigAng = s->sbOrigAng; PA.invInertia = s->invInertia; PA.sbIA = s->sbIA; PA.sbIB = s->sbIB; PA.frictions = s->frictions; PA.P = P; PA.R = R; PA.M = material_args(s);
    PA.angVel = s->angVel; PA.kinFtv = (s->nKin && !pgs) ? s->kinFtv : nullptr;
    pxb_launch_prep_rows(st, pgs, s->capPairs, PA); s->launches++;
    MARK(4);
    SolveArgs VA;
    VA.counters = s->counters; VA.partStart = s->partStart; VA.posIters = s->desc.posIters; VA.velIters = s->desc.velIters; VA.stepDt = P.stepDt; VA.R = R;
    VA.sbLin = s->sbLin; VA.sbAng = s->sbAng; VA.sbDLin = s->sbDLin; VA.sbDAng = s->sbDAng; VA.sbIA = s->sbIA; VA.sbIB = s->sbIB; VA.sbP = s->sbP; VA.sbQ = s->sbQ; VA.bodyHasCon = s->bodyHasCon;
    VA.nDyn = s->nDyn; VA.dynActor = s->dynActorDev; VA.kinFtv = (s->nKin && !pgs) ? s->kinFtv : nullptr;
    CK(pxb_launch_solve(st, pgs, pgs ? s->coopBlocksSolvePgs : s->coopBlocksSolve, VA)); s->launches++;
    MARK(5);
    pxb_launch_writeback_rows(st, s->capPairs, s->counters, R, s->pairSlots[cur], s->cForce, s->frictions, s->contactData ? s->frReport : nullptr, s->pairBodies, s->pos, s->quat); s->launches++;
    if (pgs) {
      pxb_launch_finalize_bodies_pgs(st, s->nDyn, s->dynActorDev, dt, s->pos, s->quat, s->linVel, s->angVel, s->sbLin, s->sbAng, s->sbDLin, s->sbDAng, s->sbIA, s->sbIB, s->invInertia, SA, s->geomFlags); s->launches++;
      if (s->nKin) LAUNCH(k_kin_finalize, cdiv(s->nKin, 128), 128, s->nKin, s->kinList, s->pos, s->quat, s->kinP, s->kinQ, s->kinHas);
      if (s->exportOn) LAUNCH(k_states_export, cdiv(s->nDyn, 256), 256, s->nDyn, s->dynActorDev, s->pos, s->quat, s->linVel, s->angVel, s->exportTab);
      MARK(6);
      CK(cudaGetLastError());
      return PXB_OK;
    }
  }
  LAUNCH(k_finalize_bodies, cdiv(s->nDyn, B), B, s->nDyn, s->dynActorDev, dt, s->pos, s->quat, s->linVel, s->angVel, s->sbLin, s->sbAng, s->sbIA, s->sbIB, s->sbP, s->sbQ, s->bodyHasCon,
         s->sbDLin, s->sbDAng, s->invInertia, SA, s->geomFlags);
  if (s->nKin) LAUNCH(k_kin_finalize, cdiv(s->nKin, 128), 128, s->nKin, s->kinList, s->pos, s->quat, s->kinP, s->kinQ, s->kinHas);
  if (s->exportOn) LAUNCH(k_states_export, cdiv(s->nDyn, 256), 256, s->nDyn, s->dynActorDev, s->pos, s->quat, s->linVel, s->angVel, s->exportTab);
  MARK(6);
  CK(cudaGetLastError());
  return PXB_OK;
}

static void drop_graphs(PxbScene* s) {
  for (int k = 0; k < 2; ++k) for (int j = 0; j < 3; ++j) if (s->graphExec[k][j]) { cudaGraphExecDestroy(s->graphExec[k][j]); s->graphExec[k][j] = 0; }
}

// Captures one replayable piece of the step for buffer parity `par` (kind as enqueue_step's phase).  Returns false when capture is not possible.
static bool capture_graph(PxbScene* s, float dt, int par, int kind) {
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
  const int curBefore = s->cur;
  if (kind == 2) s->cur = par;   // the second part runs with the parity the first part switched to
  const int rc = enqueue_step(s, dt, kind);
  const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
  s->cur = curBefore;
  if (rc != PXB_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&s->graphExec[par][kind], graph, 0) != cudaSuccess) {
    cudaGetLastError(); if (graph) cudaGraphDestroy(graph);
    s->graphExec[par][kind] = 0; s->abort = false;
    return false;
  }
  cudaGraphDestroy(graph);
  s->graphLaunches[par][kind] = s->launches;
  return true;
}

// The launch sequence of a step depends only on the buffer parity (all counts live in device memory), so it is
// captured once per parity into a CUDA graph and replayed: ~55 kernel launches become one graph launch.
// Teacher-forced constraint order, stage profiling and PXB_NO_GRAPH=1 use direct launches.
PXB_API int pxb_scene_simulate(PxbScene* s, float dt) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null scene");
  if (s->abort) return fail(PXB_ERR_CUDA, "scene is in abort mode");
  if (s->nA == 0) return PXB_OK;
  if (!(dt > 0.f)) return fail(PXB_ERR_INVALID, "dt must be positive");
  if (s->gridDirty) { rebuild_grid(s); drop_graphs(s); }
  if (int rc = select_path(s)) return rc;
  if (!s->envActive) s->everStepped = true;
  if (s->bodyAccel) s->accelInvDt = 1.0f / dt;
  const bool graphOk = s->useGraph && s->nOrder == 0 && !s->profiling;
  if (!graphOk) {
    if (s->velPending) { CK(cudaStreamWaitEvent(s->stream, s->velEvent, 0)); s->velPending = false; }
    const int rc = enqueue_step(s, dt); if (rc) return rc; s->stepping = true; return PXB_OK;
  }
  if (s->graphDt != dt) { drop_graphs(s); s->graphDt = dt; }
  const int par = s->cur ^ 1;   // parity this step runs with
  if (s->velPending) {
    // a velocity write is in flight on the copy stream: replay the step in two pieces and join it between them (poses only before the join)
    if ((!s->graphExec[par][1] && !capture_graph(s, dt, par, 1)) || (!s->graphExec[par][2] && !capture_graph(s, dt, par, 2))) { s->useGraph = false; return pxb_scene_simulate(s, dt); }
    CK(cudaGraphLaunch(s->graphExec[par][1], s->stream));
    CK(cudaStreamWaitEvent(s->stream, s->velEvent, 0)); s->velPending = false;
    s->cur = par; s->launches = s->graphLaunches[par][1] + s->graphLaunches[par][2];
    CK(cudaGraphLaunch(s->graphExec[par][2], s->stream));
    s->stepping = true;
    return PXB_OK;
  }
  if (!s->graphExec[par][0] && !capture_graph(s, dt, par, 0)) { s->useGraph = false; return pxb_scene_simulate(s, dt); }   // capture not possible here: direct launches from now on
  s->cur = par; s->launches = s->graphLaunches[par][0];
  CK(cudaGraphLaunch(s->graphExec[par][0], s->stream));
  s->stepping = true;
  return PXB_OK;
}

// Per-stage device times of the last completed step (CUDA events on the scene stream), in ms:
// [0] bounds+broadphase+pair lifecycle, [1] narrowphase, [2] constraint ordering+colouring, [3] preintegrate+prep,
// [4] solve (the cooperative k_solve launch), [5] writeback+integration, [6] whole step.
PXB_API int pxb_scene_set_profiling(PxbScene* s, int enable) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null scene");
  if (enable && !s->ev[0]) for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&s->ev[i]));
  s->profiling = enable != 0;
  return PXB_OK;
}
PXB_API int pxb_scene_get_stage_times(PxbScene* s, float* ms7) { DeviceGuard dg_(s);
  if (!s || !ms7) return fail(PXB_ERR_INVALID, "null argument");
  if (!s->profiling) return fail(PXB_ERR_INVALID, "profiling is off");
  CK(cudaEventSynchronize(s->ev[6]));
  for (int i = 0; i < 6; ++i) CK(cudaEventElapsedTime(&ms7[i], s->ev[i], s->ev[i + 1]));
  CK(cudaEventElapsedTime(&ms7[6], s->ev[0], s->ev[6]));
  return PXB_OK;
}

PXB_API int pxb_scene_fetch_results(PxbScene* s, int block) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null scene");
  if (!block) { const cudaError_t e = cudaStreamQuery(s->stream); if (e == cudaErrorNotReady) return 1; }
  s->stepping = false;
  return read_counters(s);
}

PXB_API int pxb_scene_compute_bounds(PxbScene* s) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null scene");
  cudaStream_t st = s->stream;
  if (s->gridDirty) rebuild_grid(s);
  LAUNCH(k_bounds, cdiv(s->nA, 256), 256, s->nA, s->pos, s->quat, s->dims, s->geomFlags, s->envId, s->desc.contactOffset, s->tight, 0, s->aabbMin, s->aabbMax, s->grid, s->desc.reserved[0], s->cellKey, s->cellVal, hull_arrays(s), local_poses(s), s->hasShapeOff ? s->shapeOff : (const float2*)nullptr);
  CK(cudaStreamSynchronize(st));
  return PXB_OK;
}
PXB_API int pxb_scene_get_bounds(PxbScene* s, float* out6) { DeviceGuard dg_(s);
  if (!s || !out6) return fail(PXB_ERR_INVALID, "null argument");
  CK(cudaMemcpyAsync(out6, s->tight, 24 * (size_t)s->nA, cudaMemcpyDeviceToHost, s->stream)); CK(cudaStreamSynchronize(s->stream));
  return PXB_OK;
}
PXB_API int pxb_scene_broadphase(PxbScene* s, const float* tightBounds) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null scene");
  if (s->abort) return fail(PXB_ERR_CUDA, "scene is in abort mode");
  s->launches = 0;
  if (tightBounds) { CK(cudaMemcpyAsync(s->tight, tightBounds, 24 * (size_t)s->nA, cudaMemcpyHostToDevice, s->stream)); }
  if (s->gridDirty) rebuild_grid(s);
  if (int rc = select_path(s)) return rc;
  if (!s->envActive) s->everStepped = true;
  const int rc = run_broadphase(s, tightBounds != nullptr); if (rc) return rc;
  return read_counters(s);
}
static int copy_pairs(PxbScene* s, const uint64_t* dev, uint32_t n, uint32_t* out, bool sortThem) {
  std::vector<uint64_t> k(n);
  if (n) { CK(cudaMemcpyAsync(k.data(), dev, 8 * (size_t)n, cudaMemcpyDeviceToHost, s->stream)); CK(cudaStreamSynchronize(s->stream)); }
  if (sortThem) std::sort(k.begin(), k.end());
  for (uint32_t i = 0; i < n; ++i) { out[2 * i] = (uint32_t)(k[i] >> s->bitsA); out[2 * i + 1] = (uint32_t)(k[i] & ((1ull << s->bitsA) - 1ull)); }
  return PXB_OK;
}
PXB_API uint32_t pxb_scene_num_pairs(PxbScene* s) { DeviceGuard dg_(s); return s ? s->hNPairs : 0; }
PXB_API uint32_t pxb_scene_num_created(PxbScene* s) { DeviceGuard dg_(s); return s ? s->hNCreated : 0; }
PXB_API uint32_t pxb_scene_num_deleted(PxbScene* s) { DeviceGuard dg_(s); return s ? s->hNDeleted : 0; }
PXB_API int pxb_scene_get_pairs(PxbScene* s, uint32_t* out) { DeviceGuard dg_(s); if (!s || !out) return fail(PXB_ERR_INVALID, "null argument"); return copy_pairs(s, s->pairKeys[s->cur], s->hNPairs, out, s->envActive); }
PXB_API int pxb_scene_get_created(PxbScene* s, uint32_t* out) { DeviceGuard dg_(s); if (!s || !out) return fail(PXB_ERR_INVALID, "null argument"); return copy_pairs(s, s->createdKeys, s->hNCreated, out, true); }
PXB_API int pxb_scene_get_deleted(PxbScene* s, uint32_t* out) { DeviceGuard dg_(s); if (!s || !out) return fail(PXB_ERR_INVALID, "null argument"); return copy_pairs(s, s->deletedKeys, s->hNDeleted, out, true); }
PXB_API int pxb_scene_get_contacts(PxbScene* s, float* out24) { DeviceGuard dg_(s);
  if (!s || !out24) return fail(PXB_ERR_INVALID, "null argument");
  const uint32_t n = s->hNPairs; if (!n) return PXB_OK;
  std::vector<float4> h(n), p((size_t)n * 4); std::vector<float> f((size_t)n * 4);
  CK(cudaMemcpyAsync(h.data(), s->cHdr, 16 * (size_t)n, cudaMemcpyDeviceToHost, s->stream)); CK(cudaMemcpyAsync(p.data(), s->cPts, 64 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaMemcpyAsync(f.data(), s->cForce, 16 * (size_t)n, cudaMemcpyDeviceToHost, s->stream)); CK(cudaStreamSynchronize(s->stream));
  std::vector<uint32_t> perm(n); for (uint32_t i = 0; i < n; ++i) perm[i] = i;
  if (s->envActive) {   // the environment path keeps pairs per environment: report them in ascending key order like the device-wide path
    std::vector<uint64_t> k(n); CK(cudaMemcpyAsync(k.data(), s->pairKeys[s->cur], 8 * (size_t)n, cudaMemcpyDeviceToHost, s->stream)); CK(cudaStreamSynchronize(s->stream));
    std::sort(perm.begin(), perm.end(), [&](uint32_t x, uint32_t y) { return k[x] < k[y]; });
  }
  for (uint32_t r = 0; r < n; ++r) {
    const uint32_t i = perm[r];
    float* o = out24 + (size_t)r * 24; memset(o, 0, 96);
    int cnt; memcpy(&cnt, &h[i].w, 4);
    o[0] = (float)cnt;
    if (cnt) { o[1] = h[i].x; o[2] = h[i].y; o[3] = h[i].z; }
    for (int k = 0; k < cnt && k < 4; ++k) { const float4 q = p[(size_t)i * 4 + k]; float* r = o + 4 + k * 5; r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; r[4] = f[(size_t)i * 4 + k]; }
  }
  return PXB_OK;
}
// ---- a1: local poses (PxShape::setLocalPose, PxRigidBody::setCMassLocalPose) ----
// Host restatement of the scalar PxTransform algebra the API layer uses (foundation/PxTransform.h: a * b, getInverse, getNormalized) and of transformInvFast
// (CmTransformUtils.h:56-69) for the constant shape2Body = body2Actor^-1 * shape2Actor.  x86-64 host code without FMA, like the reference build.
namespace lp {
struct Q { float x, y, z, w; }; struct V { float x, y, z; }; struct T { V p; Q q; };
static inline V rot(Q q, V v) { const float vx = 2.0f * v.x, vy = 2.0f * v.y, vz = 2.0f * v.z; const float w2 = q.w * q.w - 0.5f; const float d2 = (q.x * vx + q.y * vy + q.z * vz);
  return V{(vx * w2 + (q.y * vz - q.z * vy) * q.w + q.x * d2), (vy * w2 + (q.z * vx - q.x * vz) * q.w + q.y * d2), (vz * w2 + (q.x * vy - q.y * vx) * q.w + q.z * d2)}; }
static inline V rotinv(Q q, V v) { const float vx = 2.0f * v.x, vy = 2.0f * v.y, vz = 2.0f * v.z; const float w2 = q.w * q.w - 0.5f; const float d2 = (q.x * vx + q.y * vy + q.z * vz);
  return V{(vx * w2 - (q.y * vz - q.z * vy) * q.w + q.x * d2), (vy * w2 - (q.z * vx - q.x * vz) * q.w + q.y * d2), (vz * w2 - (q.x * vy - q.y * vx) * q.w + q.z * d2)}; }
static inline Q qmul(Q a, Q b) { return Q{a.w * b.x + b.w * a.x + a.y * b.z - b.y * a.z, a.w * b.y + b.w * a.y + a.z * b.x - b.z * a.x, a.w * b.z + b.w * a.z + a.x * b.y - b.x * a.y, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z}; }
static inline T mul(const T& a, const T& b) { const V r = rot(a.q, b.p); return T{V{r.x + a.p.x, r.y + a.p.y, r.z + a.p.z}, qmul(a.q, b.q)}; }
static inline T inverse(const T& a) { return T{rotinv(a.q, V{-a.p.x, -a.p.y, -a.p.z}), Q{-a.q.x, -a.q.y, -a.q.z, a.q.w}}; }
static inline T normalized(const T& a) { const float s = 1.0f / sqrtf(a.q.x * a.q.x + a.q.y * a.q.y + a.q.z * a.q.z + a.q.w * a.q.w); return T{a.p, Q{a.q.x * s, a.q.y * s, a.q.z * s, a.q.w * s}}; }
static inline bool identity(const T& a) { return a.p.x == 0.f && a.p.y == 0.f && a.p.z == 0.f && a.q.x == 0.f && a.q.y == 0.f && a.q.z == 0.f && a.q.w == 1.f; }
static inline float adot(V a, V b) { return (a.x * b.x + a.z * b.z) + (a.y * b.y); }
static inline V cross(V a, V b) { return V{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline V scaleadd(V a, float s, V b) { return V{a.x * s + b.x, a.y * s + b.y, a.z * s + b.z}; }
static inline T inv_fast(const T& a, const T& b) {   // a^-1 * b, aos operation order
  const float wa = a.q.w, wb = b.q.w; const V va{a.q.x, a.q.y, a.q.z}, vb{b.q.x, b.q.y, b.q.z};
  const float wo = wa * wb + adot(va, vb);
  const V c = scaleadd(vb, wa, cross(vb, va)); const V vo{c.x - va.x * wb, c.y - va.y * wb, c.z - va.z * wb};
  const V pt{b.p.x - a.p.x, b.p.y - a.p.y, b.p.z - a.p.z};
  const float k = wa * wa + (-0.5f); const V t1{pt.x * k, pt.y * k, pt.z * k};
  const V t2 = scaleadd(cross(pt, va), wa, t1);
  const V t3 = scaleadd(va, adot(va, pt), t2);
  return T{V{t3.x + t3.x, t3.y + t3.y, t3.z + t3.z}, Q{vo.x, vo.y, vo.z, wo}};
}
}  // namespace lp
// actor pose = body2World * body2Actor^-1 (NpRigidDynamic::getGlobalPoseFast), body2World = actor pose * body2Actor (NpRigidDynamic::setGlobalPose), scalar PxTransform algebra
__global__ void k_actor_poses(uint32_t nA, const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ b2aP, const float4* __restrict__ b2aQ, float4* __restrict__ actorPos,
                              float4* __restrict__ actorQuat) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nA) return;
  const float4 p = pos[a], q = quat[a], bp = b2aP[a];
  if (bp.w == 0.f) { actorPos[a] = p; actorQuat[a] = q; return; }   // w of the body2Actor position: 1 = not the identity
  const q4 bq = Q4(b2aQ[a]);
  const v3 ip = qrotinv(bq, V3(-bp.x, -bp.y, -bp.z)); const q4 iq = conj(bq);
  const q4 Qw = Q4(q);
  actorPos[a] = F4(qrot(Qw, ip) + V3(p.x, p.y, p.z), p.w); actorQuat[a] = F4(qmul(Qw, iq));
}
__global__ void k_actor_to_body(uint32_t nb, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ dynActor, float4* __restrict__ pos, float4* __restrict__ quat, const float4* __restrict__ b2aP,
                                const float4* __restrict__ b2aQ) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb) return;
  const uint32_t a = dynActor[idx ? idx[t] : t];
  const float4 bp = b2aP[a];
  if (bp.w == 0.f) return;
  const float4 p = pos[a]; const q4 Qa = Q4(quat[a]);
  pos[a] = F4(qrot(Qa, V3(bp.x, bp.y, bp.z)) + V3(p.x, p.y, p.z), p.w); quat[a] = F4(qmul(Qa, Q4(b2aQ[a])));
}
// Poses handed out by the getters are ACTOR poses: with centre-of-mass local poses they are derived from the body frames first.
static int refresh_actor_poses(PxbScene* s, cudaStream_t st) {
  if (!s->hasCom || !s->nA) return PXB_OK;
  k_actor_poses<<<cdiv(s->nA, 256), 256, 0, st>>>(s->nA, s->pos, s->quat, s->b2aP, s->b2aQ, s->actorPos, s->actorQuat);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_scene_set_local_poses(PxbScene* s, uint32_t firstActor, uint32_t n, const float* shape2Actor, const float* body2Actor) { DeviceGuard dg_(s);
  if (!s || (n && (!shape2Actor || !body2Actor))) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if ((uint64_t)firstActor + n > s->nA) return fail(PXB_ERR_INVALID, "actor range out of bounds");
  if (s->exportOn) return fail(PXB_ERR_UNSUPPORTED, "the fused state export reports body frames: switch it off before setting local poses");
  if (!n) return PXB_OK;
  CK(cudaStreamSynchronize(s->stream));
  const size_t A = std::max<size_t>(s->capA, 1);
  if (!s->tcPos) {
    CK(dalloc(s->tcPos, A)); CK(dalloc(s->tcQuat, A)); CK(dalloc(s->s2bP, A)); CK(dalloc(s->s2bQ, A)); CK(dalloc(s->b2aP, A)); CK(dalloc(s->b2aQ, A)); CK(dalloc(s->actorPos, A)); CK(dalloc(s->actorQuat, A));
    s->hS2aP.assign(A, make_float4(0, 0, 0, 0)); s->hS2aQ.assign(A, make_float4(0, 0, 0, 1)); s->hB2aP.assign(A, make_float4(0, 0, 0, 0)); s->hB2aQ.assign(A, make_float4(0, 0, 0, 1));
    CK(cudaMemcpy(s->s2bP, s->hS2aP.data(), 16 * A, cudaMemcpyHostToDevice)); CK(cudaMemcpy(s->s2bQ, s->hS2aQ.data(), 16 * A, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->b2aP, s->hB2aP.data(), 16 * A, cudaMemcpyHostToDevice)); CK(cudaMemcpy(s->b2aQ, s->hB2aQ.data(), 16 * A, cudaMemcpyHostToDevice));
  }
  std::vector<float4> pos(n), quat(n), s2bP(n), s2bQ(n);
  CK(cudaMemcpy(pos.data(), s->pos + firstActor, 16 * (size_t)n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(quat.data(), s->quat + firstActor, 16 * (size_t)n, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t a = firstActor + i; const bool dyn = s->recs[a].flags & PXB_ACTOR_DYNAMIC;
    const float* sa = shape2Actor + 7 * (size_t)i; const float* ba = body2Actor + 7 * (size_t)i;
    const lp::T s2a = lp::normalized(lp::T{lp::V{sa[0], sa[1], sa[2]}, lp::Q{sa[3], sa[4], sa[5], sa[6]}});   // NpShape::setLocalPose / NpRigidDynamic::setCMassLocalPose keep the normalised transform
    lp::T b2a = lp::normalized(lp::T{lp::V{ba[0], ba[1], ba[2]}, lp::Q{ba[3], ba[4], ba[5], ba[6]}});
    if (!dyn) b2a = lp::T{lp::V{0, 0, 0}, lp::Q{0, 0, 0, 1}};
    const lp::T oldB2a{lp::V{s->hB2aP[a].x, s->hB2aP[a].y, s->hB2aP[a].z}, lp::Q{s->hB2aQ[a].x, s->hB2aQ[a].y, s->hB2aQ[a].z, s->hB2aQ[a].w}};
    lp::T b2w{lp::V{pos[i].x, pos[i].y, pos[i].z}, lp::Q{quat[i].x, quat[i].y, quat[i].z, quat[i].w}};
    if (dyn && !(lp::identity(oldB2a) && lp::identity(b2a))) {   // Sc::BodyCore::setCMassLocalPose (ScBodyCore.cpp:98-108): the actor keeps its pose
      const lp::T a2w = lp::mul(b2w, lp::inverse(oldB2a));
      b2w = lp::mul(a2w, b2a);
    }
    pos[i] = make_float4(b2w.p.x, b2w.p.y, b2w.p.z, pos[i].w); quat[i] = make_float4(b2w.q.x, b2w.q.y, b2w.q.z, b2w.q.w);
    const lp::T s2b = (dyn && !lp::identity(b2a)) ? lp::inv_fast(b2a, s2a) : s2a;   // Cm::getDynamicGlobalPoseAligned's first product, constant per shape
    s2bP[i] = make_float4(s2b.p.x, s2b.p.y, s2b.p.z, 0.f); s2bQ[i] = make_float4(s2b.q.x, s2b.q.y, s2b.q.z, s2b.q.w);
    s->hS2aP[a] = make_float4(s2a.p.x, s2a.p.y, s2a.p.z, 0.f); s->hS2aQ[a] = make_float4(s2a.q.x, s2a.q.y, s2a.q.z, s2a.q.w);
    s->hB2aP[a] = make_float4(b2a.p.x, b2a.p.y, b2a.p.z, lp::identity(b2a) ? 0.f : 1.f); s->hB2aQ[a] = make_float4(b2a.q.x, b2a.q.y, b2a.q.z, b2a.q.w);
  }
  CK(cudaMemcpy(s->pos + firstActor, pos.data(), 16 * (size_t)n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(s->quat + firstActor, quat.data(), 16 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->s2bP + firstActor, s2bP.data(), 16 * (size_t)n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(s->s2bQ + firstActor, s2bQ.data(), 16 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->b2aP + firstActor, s->hB2aP.data() + firstActor, 16 * (size_t)n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(s->b2aQ + firstActor, s->hB2aQ.data() + firstActor, 16 * (size_t)n, cudaMemcpyHostToDevice));
  s->hasLocal = true;
  s->hasCom = false; for (uint32_t a = 0; a < s->nA; ++a) if (s->hB2aP[a].w != 0.f) { s->hasCom = true; break; }
  drop_graphs(s);   // the narrowphase now reads the transform cache
  return PXB_OK;
}

// PxRigidBody::setMass / setMassSpaceInertiaTensor at run time (NpRigidBodyTemplate.h: the core keeps inverse mass / inverse inertia; 0 = infinite): domain randomisation
__global__ void k_set_mass(uint32_t n, const uint32_t* __restrict__ idx, const float* __restrict__ massInertia4, const uint32_t* __restrict__ dynActor, float4* __restrict__ pos, float4* __restrict__ invInertia) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
  const uint32_t a = dynActor[idx[t]]; const float* m = massInertia4 + (size_t)t * 4;
  pos[a].w = m[0] > 0.f ? 1.0f / m[0] : 0.f;
  const float w = invInertia[a].w;
  invInertia[a] = make_float4(m[1] > 0.f ? 1.0f / m[1] : 0.f, m[2] > 0.f ? 1.0f / m[2] : 0.f, m[3] > 0.f ? 1.0f / m[3] : 0.f, w);
}
PXB_API int pxb_scene_set_mass_properties(PxbScene* s, const uint32_t* indices, const float* massInertia4, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || (nb && (!indices || !massInertia4))) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if (!nb) return PXB_OK;
  if (nb > s->capA) return fail(PXB_ERR_INVALID, "nb exceeds the actor capacity");
  for (uint32_t i = 0; i < nb; ++i) {
    if (indices[i] >= s->nDyn) return fail(PXB_ERR_INVALID, "index out of range");
    if (s->recs[s->dynActor[indices[i]]].flags & PXB_ACTOR_KINEMATIC) return fail(PXB_ERR_INVALID, "a kinematic body has no mass properties");
    for (int k = 0; k < 4; ++k) if (!(massInertia4[4 * i + k] >= 0.f) || !std::isfinite(massInertia4[4 * i + k])) return fail(PXB_ERR_INVALID, "mass / inertia must be finite and >= 0 (0 = infinite)");
  }
  if (s->velPending) { CK(cudaStreamWaitEvent(s->stream, s->velEvent, 0)); s->velPending = false; }
  cudaStream_t st = s->stream;
  float* d = s->stage;   // staging: 4 floats per body fit the first region (7 per actor)
  CK(cudaMemcpyAsync(d, massInertia4, 16 * (size_t)nb, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(s->stageIdx, indices, 4 * (size_t)nb, cudaMemcpyHostToDevice, st));
  LAUNCH(k_set_mass, cdiv(nb, 128), 128, nb, s->stageIdx, d, s->dynActorDev, s->pos, s->invInertia);
  CK(cudaStreamSynchronize(st));
  for (uint32_t i = 0; i < nb; ++i) { ActorRec& r = s->recs[s->dynActor[indices[i]]]; r.mass = massInertia4[4 * i]; r.inertia[0] = massInertia4[4 * i + 1]; r.inertia[1] = massInertia4[4 * i + 2]; r.inertia[2] = massInertia4[4 * i + 3]; }
  return PXB_OK;
}

// PxScene::setGravity (NpScene.cpp:331-342 -> Sc::Scene::mGravity, read by the next step's pre-integration): domain randomisation changes it between steps
PXB_API int pxb_scene_set_gravity(PxbScene* s, const float* gravity3) { DeviceGuard dg_(s);
  if (!s || !gravity3) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "PxScene::setGravity() not allowed while simulation is running");
  if (!std::isfinite(gravity3[0]) || !std::isfinite(gravity3[1]) || !std::isfinite(gravity3[2])) return fail(PXB_ERR_INVALID, "gravity is not finite");
  if (gravity3[0] == s->desc.gravity[0] && gravity3[1] == s->desc.gravity[1] && gravity3[2] == s->desc.gravity[2]) return PXB_OK;
  s->desc.gravity[0] = gravity3[0]; s->desc.gravity[1] = gravity3[1]; s->desc.gravity[2] = gravity3[2];
  drop_graphs(s);   // the vector travels in the kernel arguments
  return PXB_OK;
}

// ---- f1: the default simulation filter shader on the device ----
PXB_API int pxb_scene_set_filter_shader(PxbScene* s, const PxbFilterShaderConfig* cfg) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if (cfg) { for (int k = 0; k < 3; ++k) if (cfg->ops[k] > 6u) return fail(PXB_ERR_INVALID, "filter op out of range (PxFilterOp: 0..6)"); }
  CK(cudaStreamSynchronize(s->stream));
  if (cfg && !s->filterData) { const size_t A = std::max<size_t>(s->capA, 1); CK(dalloc(s->filterData, A)); CK(cudaMemset(s->filterData, 0, 16 * A)); }
  if (cfg) { static_assert(sizeof(FilterConfig) == sizeof(PxbFilterShaderConfig), "filter config layout"); memcpy(&s->filterCfg, cfg, sizeof(FilterConfig)); }
  if ((cfg != nullptr) != s->hasFilter) { s->hasFilter = cfg != nullptr; }
  drop_graphs(s);   // the configuration travels in the kernel arguments
  return PXB_OK;
}
PXB_API int pxb_scene_set_filter_data(PxbScene* s, uint32_t firstActor, uint32_t n, const uint32_t* data4) { DeviceGuard dg_(s);
  if (!s || (n && !data4)) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if (!s->filterData) return fail(PXB_ERR_INVALID, "call pxb_scene_set_filter_shader first");
  if ((uint64_t)firstActor + n > s->capA) return fail(PXB_ERR_INVALID, "actor range out of bounds");
  if (!n) return PXB_OK;
  CK(cudaMemcpyAsync(s->filterData + firstActor, data4, 16 * (size_t)n, cudaMemcpyHostToDevice, s->stream)); CK(cudaStreamSynchronize(s->stream));
  return PXB_OK;
}

// ---- PxShape::setContactOffset / setRestOffset per shape ----
PXB_API int pxb_scene_set_shape_offsets(PxbScene* s, uint32_t firstActor, uint32_t n, const float* contactRest2) { DeviceGuard dg_(s);
  if (!s || (n && !contactRest2)) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if ((uint64_t)firstActor + n > s->capA) return fail(PXB_ERR_INVALID, "actor range out of bounds");
  for (uint32_t i = 0; i < n; ++i) {   // PxShape::setContactOffset / setRestOffset: contactOffset >= 0, contactOffset > restOffset (NpShape.cpp)
    const float c = contactRest2[2 * i], r = contactRest2[2 * i + 1];
    if (!(c >= 0.f) || !(c > r) || !std::isfinite(c) || !std::isfinite(r)) return fail(PXB_ERR_INVALID, "contactOffset must be >= 0 and larger than restOffset");
  }
  if (!n) return PXB_OK;
  CK(cudaStreamSynchronize(s->stream));
  if (!s->shapeOff) {
    const size_t A = std::max<size_t>(s->capA, 1);
    CK(dalloc(s->shapeOff, A));
    std::vector<float2> init(A, make_float2(s->desc.contactOffset, s->desc.restOffset));
    CK(cudaMemcpy(s->shapeOff, init.data(), 8 * A, cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(s->shapeOff + firstActor, contactRest2, 8 * (size_t)n, cudaMemcpyHostToDevice));
  for (uint32_t i = 0; i < n; ++i) s->maxContactOffset = std::max(s->maxContactOffset, contactRest2[2 * i]);
  s->hasShapeOff = true; s->gridDirty = true;
  drop_graphs(s);
  return PXB_OK;
}

// ---- f3: PxDirectGPUAPI::copyContactData (PxDirectGPUAPI.h:388-401; the reference's compressContactStage1/2, gpunarrowphase/src/CUDA/compressOutputContacts.cu) ----
// One PxGpuContactPair record per pair that has contacts, in pair order (stable compaction by two exclusive scans: touching flags -> record index, contact counts ->
// offset into the point / force streams), pointing into PxContactPatch / PxContact / force / PxFrictionPatch streams owned by the scene.
struct ContactPairRec { unsigned long long contactPatches, contactPoints, contactForces, frictionPatches; uint32_t ref0, ref1; unsigned long long node0, node1, actor0, actor1; uint16_t nbContacts, nbPatches; uint32_t pad; };
static_assert(sizeof(ContactPairRec) == 80 && sizeof(ContactPairRec) == sizeof(PxbGpuContactPair), "PxGpuContactPair layout");
__global__ void k_cc_counts(const uint32_t* __restrict__ nPairsP, const float4* __restrict__ cHdr, uint32_t* __restrict__ cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *nPairsP) cnt[i] = (uint32_t)__float_as_int(cHdr[i].w);
}
__global__ void k_cc_actor_dyn(uint32_t nDyn, const uint32_t* __restrict__ dynActor, uint32_t* __restrict__ actorDyn) { const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; if (d < nDyn) actorDyn[dynActor[d]] = d; }
__global__ void k_cc_write(const uint32_t* __restrict__ nPairsP, const float4* __restrict__ cHdr, const float4* __restrict__ cPts, const float* __restrict__ cForce, const uint2* __restrict__ pairBodies,
                           const uint32_t* __restrict__ pairSlots, const float4* __restrict__ frictions, const float4* __restrict__ frReport, const uint32_t* __restrict__ actorMat,
                           const uint32_t* __restrict__ actorDyn, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ off, uint8_t* __restrict__ patches, uint8_t* __restrict__ points,
                           float* __restrict__ forces, uint8_t* __restrict__ fric, ContactPairRec* __restrict__ out, uint32_t maxPairs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *nPairsP) return;
  const float4 h = cHdr[i]; const int n = __float_as_int(h.w);
  if (n <= 0) return;
  const uint32_t r = idx[i], o = off[i];
  const uint2 bb = pairBodies[i];
  const float4* fr = frictions + (size_t)pairSlots[i] * PXB_FRICTION_F4;
  // PxContactPatch (PxContact.h:56-137, 64 bytes): mass modification 1,1,1,1 | normal, restitution | dynamic friction, static friction, damping, (startContactIndex u16, nbContacts u8,
  // materialFlags u8) | (internalFlags u16, materialIndex0 u16), (materialIndex1 u16, pad) ...
  float4* P = reinterpret_cast<float4*>(patches + (size_t)r * 64);
  P[0] = make_float4(1.f, 1.f, 1.f, 1.f);
  P[1] = make_float4(h.x, h.y, h.z, fr[5].w);
  P[2] = make_float4(fr[4].w, fr[3].w, 0.f, __uint_as_float(((uint32_t)n & 0xffu) << 16));
  const uint32_t m0 = actorMat ? actorMat[bb.x] : 0u, m1 = actorMat ? actorMat[bb.y] : 0u;
  P[3] = make_float4(__uint_as_float((m0 & 0xffffu) << 16), __uint_as_float(m1 & 0xffffu), 0.f, 0.f);
  float4* C = reinterpret_cast<float4*>(points) + o;   // PxContact: point, separation
  for (int k = 0; k < n; ++k) { C[k] = cPts[(size_t)i * 4 + k]; forces[o + k] = cForce[(size_t)i * 4 + k]; }
  // PxFrictionPatch (PxContact.h:635-658, 52 bytes): anchorPositions[2], anchorImpulses[2], anchorCount
  float* F = reinterpret_cast<float*>(fric + (size_t)r * 52);
  const float4 q0 = frReport[(size_t)i * 4], q1 = frReport[(size_t)i * 4 + 1], w0 = frReport[(size_t)i * 4 + 2], w1 = frReport[(size_t)i * 4 + 3];
  F[0] = w0.x; F[1] = w0.y; F[2] = w0.z; F[3] = w1.x; F[4] = w1.y; F[5] = w1.z; F[6] = q0.x; F[7] = q0.y; F[8] = q0.z; F[9] = q1.x; F[10] = q1.y; F[11] = q1.z; F[12] = q0.w;
  if (r >= maxPairs) return;
  ContactPairRec c;
  c.contactPatches = (unsigned long long)(patches + (size_t)r * 64); c.contactPoints = (unsigned long long)(C); c.contactForces = (unsigned long long)(forces + o);
  c.frictionPatches = (unsigned long long)(fric + (size_t)r * 52); c.ref0 = bb.x; c.ref1 = bb.y;
  c.node0 = actorDyn[bb.x]; c.node1 = actorDyn[bb.y];   // PxNodeIndex: mID in the low word (0xffffffff = static actor), mLinkID 0
  c.actor0 = bb.x; c.actor1 = bb.y; c.nbContacts = (uint16_t)n; c.nbPatches = 1; c.pad = 0;
  out[r] = c;
}
PXB_API int pxb_scene_enable_contact_data(PxbScene* s, int enable) { DeviceGuard dg_(s);
  if (!s) return fail(PXB_ERR_INVALID, "null argument");
  if ((enable != 0) == s->contactData) return PXB_OK;
  CK(cudaStreamSynchronize(s->stream));
  if (enable) {
    const size_t Pn = s->capPairs;
    if (!s->frReport) {
      CK(dalloc(s->frReport, Pn * 4)); CK(dalloc(s->ccIdx, Pn)); CK(dalloc(s->ccOff, Pn)); CK(dalloc(s->ccCount, Pn)); CK(dalloc(s->ccTotal, 4)); CK(dalloc(s->actorDyn, std::max<size_t>(s->capA, 1)));
      CK(dalloc(s->ccPatches, Pn * 64)); CK(dalloc(s->ccPoints, Pn * 64)); CK(dalloc(s->ccFriction, Pn * 52)); CK(dalloc(s->ccForces, Pn * 4));
      CK(cudaMemsetAsync(s->frReport, 0, Pn * 64, s->stream));
    }
  }
  s->contactData = enable != 0;
  drop_graphs(s);   // the write-back kernels gain / lose the friction report
  return PXB_OK;
}
PXB_API int pxb_scene_copy_contact_data(PxbScene* s, void* data, uint32_t* nbContactPairs, uint32_t maxPairs) { DeviceGuard dg_(s);
  if (!s || !nbContactPairs || (maxPairs && !data)) return fail(PXB_ERR_INVALID, "null argument");
  if (!s->contactData) return fail(PXB_ERR_INVALID, "contact data is off: call pxb_scene_enable_contact_data before the step");
  cudaStream_t st = s->stream; const uint32_t B = 256; const uint32_t gP = cdiv(s->capPairs, B);
  const uint32_t* nP = s->nPairsDev + s->cur;
  CK(cudaMemsetAsync(s->actorDyn, 0xff, 4 * (size_t)std::max<uint32_t>(s->nA, 1), st));
  if (s->nDyn) k_cc_actor_dyn<<<cdiv(s->nDyn, B), B, 0, st>>>(s->nDyn, s->dynActorDev, s->actorDyn);
  k_cc_counts<<<gP, B, 0, st>>>(nP, s->cHdr, s->ccCount);
  exclusive_scan_u32(s->conFlag, s->ccIdx, nP, nbContactPairs, s->scanSums, s->rsTmp.ctas, st);
  exclusive_scan_u32(s->ccCount, s->ccOff, nP, s->ccTotal, s->scanSums, s->rsTmp.ctas, st);
  k_cc_write<<<gP, B, 0, st>>>(nP, s->cHdr, s->cPts, s->cForce, s->pairBodies, s->pairSlots[s->cur], s->frictions, s->frReport, s->matTab ? s->actorMat : (const uint32_t*)nullptr, s->actorDyn, s->ccIdx, s->ccOff,
                              s->ccPatches, s->ccPoints, s->ccForces, s->ccFriction, (ContactPairRec*)data, maxPairs);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API uint32_t pxb_scene_num_touch_found(PxbScene* s) { return s ? s->hNTouchFound : 0; }
PXB_API uint32_t pxb_scene_num_touch_lost(PxbScene* s) { return s ? s->hNTouchLost : 0; }
PXB_API int pxb_scene_get_touch_found(PxbScene* s, uint32_t* out) { DeviceGuard dg_(s); if (!s || !out) return fail(PXB_ERR_INVALID, "null argument"); return copy_pairs(s, s->touchFound, s->hNTouchFound, out, true); }
PXB_API int pxb_scene_get_touch_lost(PxbScene* s, uint32_t* out) { DeviceGuard dg_(s); if (!s || !out) return fail(PXB_ERR_INVALID, "null argument"); return copy_pairs(s, s->touchLost, s->hNTouchLost, out, true); }
PXB_API uint32_t pxb_scene_last_num_partitions(PxbScene* s) { DeviceGuard dg_(s); return s ? s->hNPart : 0; }
PXB_API uint32_t pxb_scene_last_num_constraints(PxbScene* s) { DeviceGuard dg_(s); return s ? s->hNCon : 0; }
PXB_API uint32_t pxb_scene_last_num_launches(PxbScene* s) { DeviceGuard dg_(s); return s ? s->launches : 0; }
PXB_API int pxb_scene_uses_env_path(PxbScene* s) { DeviceGuard dg_(s); return s && s->envActive ? 1 : 0; }
PXB_API int pxb_scene_get_sleep_data(PxbScene* s, float* wakeCounters, uint32_t* asleep) { DeviceGuard dg_(s);
  if (!s || !wakeCounters || !asleep) return fail(PXB_ERR_INVALID, "null argument");
  std::vector<float> w(s->nA); std::vector<uint32_t> f(s->nA);
  CK(cudaMemcpyAsync(w.data(), s->wake, 4 * (size_t)s->nA, cudaMemcpyDeviceToHost, s->stream)); CK(cudaMemcpyAsync(f.data(), s->asleep, 4 * (size_t)s->nA, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  for (uint32_t d = 0; d < s->nDyn; ++d) { wakeCounters[d] = w[s->dynActor[d]]; asleep[d] = f[s->dynActor[d]]; }
  return PXB_OK;
}
#ifdef PXB_ENV_TIMING
extern "C" PXB_API int pxb_debug_env_timing(PxbScene* s, unsigned long long* out) { DeviceGuard dg_(s); cudaStreamSynchronize(s->stream); cudaMemcpy(out, s->envTiming, (size_t)s->nEnv * 16 * 8, cudaMemcpyDeviceToHost); return (int)s->nEnv; }
#endif

// any other access to the body state is ordered after a velocity write still in flight on the copy stream
static int join_pending(PxbScene* s) { if (s->velPending) { CK(cudaStreamWaitEvent(s->stream, s->velEvent, 0)); s->velPending = false; } return PXB_OK; }
static int rd_common(PxbScene* s, void* devData, const uint32_t* devIdx, int type, uint32_t nb, bool set) {
  const bool accel = !set && (type == PXB_RD_LINEAR_ACCELERATION || type == PXB_RD_ANGULAR_ACCELERATION);
  if (!accel && (type < 0 || type > (set ? PXB_RD_TORQUE : PXB_RD_ANGULAR_VELOCITY))) return fail(PXB_ERR_INVALID, "bad dataType");
  if (accel && !s->bodyAccel) return fail(PXB_ERR_INVALID, "acceleration getters need PXB_FLAG_BODY_ACCELERATIONS (PxSceneFlag::eENABLE_BODY_ACCELERATIONS)");
  if (set && type >= PXB_RD_FORCE && !s->forcesUsed) { s->forcesUsed = true; drop_graphs(s); }   // the step kernels read the force arrays from now on
  if (!devIdx && nb > s->nDyn) return fail(PXB_ERR_INVALID, "nb exceeds the number of dynamic bodies");
  cudaStream_t st = s->stream;
  if (!nb) return PXB_OK;
  if (int rc = join_pending(s)) return rc;
  const bool poseIO = s->hasCom && type == PXB_RD_GLOBAL_POSE;   // PxRigidDynamic poses are ACTOR poses; the engine integrates body (centre-of-mass) frames
  if (set) {
    LAUNCH(k_rd_set, cdiv(nb, 256), 256, nb, devIdx, s->dynActorDev, type, s->pos, s->quat, s->linVel, s->angVel, (const float*)devData, s->wake, s->asleep, s->extForce, s->extTorque);
    if (poseIO) LAUNCH(k_actor_to_body, cdiv(nb, 256), 256, nb, devIdx, s->dynActorDev, s->pos, s->quat, s->b2aP, s->b2aQ);
  } else {
    if (poseIO) { if (int rc = refresh_actor_poses(s, st)) return rc; }
    if (accel) LAUNCH(k_rd_get_accel, cdiv(nb, 256), 256, nb, devIdx, s->dynActorDev, type == PXB_RD_LINEAR_ACCELERATION ? s->linVel : s->angVel, type == PXB_RD_LINEAR_ACCELERATION ? s->prevLin : s->prevAng, s->accelInvDt, (float*)devData);
    else LAUNCH(k_rd_get, cdiv(nb, 256), 256, nb, devIdx, s->dynActorDev, type, poseIO ? s->actorPos : s->pos, poseIO ? s->actorQuat : s->quat, s->linVel, s->angVel, (float*)devData);
  }
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_get_rigid_dynamic_data_device(PxbScene* s, void* devData, const uint32_t* devIdx, int type, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || !devData) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running (NpDirectGPUAPI.cpp:63-78)");
  return rd_common(s, devData, devIdx, type, nb, false);
}
PXB_API int pxb_set_rigid_dynamic_data_device(PxbScene* s, const void* devData, const uint32_t* devIdx, int type, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || !devData) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running (NpDirectGPUAPI.cpp:63-78)");
  return rd_common(s, const_cast<void*>(devData), devIdx, type, nb, true);
}
// start / finish events of the PxDirectGPUAPI calls (PxgSimulationCore.cpp:2736-2850): wait for the start event on the scene stream, record the finish event
// after the kernel, synchronise when the caller gave none
static int rd_events(PxbScene* s, void* devData, const uint32_t* devIdx, int type, uint32_t nb, bool set, void* startEvent, void* finishEvent) {
  if (!s || !devData) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running (NpDirectGPUAPI.cpp:63-78)");
  if (startEvent) CK(cudaStreamWaitEvent(s->stream, (cudaEvent_t)startEvent, 0));
  if (int rc = rd_common(s, devData, devIdx, type, nb, set)) return rc;
  if (finishEvent) CK(cudaEventRecord((cudaEvent_t)finishEvent, s->stream)); else CK(cudaStreamSynchronize(s->stream));
  return PXB_OK;
}
PXB_API int pxb_get_rigid_dynamic_data_device_ev(PxbScene* s, void* devData, const uint32_t* devIdx, int type, uint32_t nb, void* startEvent, void* finishEvent) { DeviceGuard dg_(s);
  return rd_events(s, devData, devIdx, type, nb, false, startEvent, finishEvent); }
PXB_API int pxb_set_rigid_dynamic_data_device_ev(PxbScene* s, const void* devData, const uint32_t* devIdx, int type, uint32_t nb, void* startEvent, void* finishEvent) { DeviceGuard dg_(s);
  return rd_events(s, const_cast<void*>(devData), devIdx, type, nb, true, startEvent, finishEvent); }
// PxRigidDynamic::setKinematicTarget for nb kinematic bodies: actor poses as PxTransform (q.xyzw, p.xyz); consumed by the next step.  A bad index is reported by that
// step's fetchResults (the check runs on the device).
static int kin_targets(PxbScene* s, const uint32_t* devIdx, const float* devPoses, uint32_t nb) {
  if (!s->nKin || !s->kinP) return fail(PXB_ERR_INVALID, "the scene has no kinematic body (or has not been stepped / rebuilt since they were added)");
  if (int rc = join_pending(s)) return rc;
  cudaStream_t st = s->stream;
  LAUNCH(k_kin_set_targets, cdiv(nb, 128), 128, nb, devIdx, devPoses, s->dynActorDev, s->nDyn, s->geomFlags, s->kinP, s->kinQ, s->kinHas, s->hasCom ? s->b2aP : (const float4*)nullptr, s->hasCom ? s->b2aQ : (const float4*)nullptr, s->counters);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_scene_set_kinematic_targets_device(PxbScene* s, const uint32_t* devIndices, const float* devPoses, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || (nb && (!devIndices || !devPoses))) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if (!nb) return PXB_OK;
  if (s->gridDirty) { rebuild_grid(s); drop_graphs(s); }
  return kin_targets(s, devIndices, devPoses, nb);
}
PXB_API int pxb_scene_set_kinematic_targets(PxbScene* s, const uint32_t* indices, const float* poses, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || (nb && (!indices || !poses))) return fail(PXB_ERR_INVALID, "null argument");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if (!nb) return PXB_OK;
  if (nb > s->capA) return fail(PXB_ERR_INVALID, "nb exceeds the actor capacity");
  if (s->gridDirty) { rebuild_grid(s); drop_graphs(s); }
  for (uint32_t i = 0; i < nb; ++i) { if (indices[i] >= s->nDyn || !(s->recs[s->dynActor[indices[i]]].flags & PXB_ACTOR_KINEMATIC)) return fail(PXB_ERR_INVALID, "setKinematicTarget: the body must be kinematic"); }
  float* d = s->stage + (size_t)s->capA * 13;   // the 'set pose' staging region (7 floats per actor)
  CK(cudaMemcpyAsync(d, poses, 28 * (size_t)nb, cudaMemcpyHostToDevice, s->stream)); CK(cudaMemcpyAsync(s->stageIdx, indices, 4 * (size_t)nb, cudaMemcpyHostToDevice, s->stream));
  const int rc = kin_targets(s, s->stageIdx, d, nb);
  CK(cudaStreamSynchronize(s->stream));   // the caller's buffers may be pageable: done with them on return
  return rc;
}
static int rd_host(PxbScene* s, void* data, const uint32_t* idx, int type, uint32_t nb, bool set, bool async) {
  if (!s || !data) return fail(PXB_ERR_INVALID, "null argument");
  const bool accel = !set && (type == PXB_RD_LINEAR_ACCELERATION || type == PXB_RD_ANGULAR_ACCELERATION);
  if (!accel && (type < 0 || type > (set ? PXB_RD_TORQUE : PXB_RD_ANGULAR_VELOCITY))) return fail(PXB_ERR_INVALID, "bad dataType");
  if (s->stepping && !async) return fail(PXB_ERR_INVALID, "illegal while the simulation is running (NpDirectGPUAPI.cpp:63-78)");
  if (async && idx) return fail(PXB_ERR_INVALID, "the stream-ordered host variants take no index list");
  if (!nb) return PXB_OK;
  const size_t bytes = (size_t)nb * (type == 0 ? 28 : 12);
  if (nb > s->capA) return fail(PXB_ERR_INVALID, "nb exceeds the actor capacity");
  // persistent staging, one region per (direction, data type): no allocation on the per-step path and stream-ordered calls never share a buffer
  float* d = s->stage + (size_t)s->capA * (accel ? 32 + 3 * (type - PXB_RD_LINEAR_ACCELERATION) : type >= PXB_RD_FORCE ? 26 + 3 * (type - PXB_RD_FORCE) : (set ? 13 : 0) + (type == 0 ? 0 : (type == 1 ? 7 : 10))); uint32_t* di = nullptr;
  if (idx) { for (uint32_t i = 0; i < nb; ++i) if (idx[i] >= s->nDyn) return fail(PXB_ERR_INVALID, "index out of range");
             di = s->stageIdx; CK(cudaMemcpyAsync(di, idx, 4 * (size_t)nb, cudaMemcpyHostToDevice, s->stream)); }
  if (set && async && (type == PXB_RD_LINEAR_VELOCITY || type == PXB_RD_ANGULAR_VELOCITY) && s->sleepThreshold == 0.f && !s->stepping && s->copyStream && nb <= s->nDyn) {
    // Stream-ordered velocity write: copy + scatter run on the copy stream, ordered after everything already enqueued on the scene stream; the next
    // pxb_scene_simulate joins them right before the solver, so the host-to-device copy overlaps bounds / broadphase / narrowphase (which read
    // poses only).  With sleeping enabled the write also wakes bodies, which the island pass of the first part reads: no overlap then.
    if (!s->velPending) { CK(cudaEventRecord(s->orderEvent, s->stream)); CK(cudaStreamWaitEvent(s->copyStream, s->orderEvent, 0)); }
    CK(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, s->copyStream));
    k_rd_set<<<cdiv(nb, 256), 256, 0, s->copyStream>>>(nb, nullptr, s->dynActorDev, type, s->pos, s->quat, s->linVel, s->angVel, (const float*)d, s->wake, s->asleep, s->extForce, s->extTorque);
    CK(cudaGetLastError());
    CK(cudaEventRecord(s->velEvent, s->copyStream)); s->velPending = true;
    return PXB_OK;
  }
  if (set) CK(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, s->stream));
  int rc = rd_common(s, d, di, type, nb, set);
  if (!rc && !set) CK(cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, s->stream));
  if (!async) CK(cudaStreamSynchronize(s->stream));
  return rc;
}
PXB_API int pxb_get_rigid_dynamic_data(PxbScene* s, void* data, const uint32_t* idx, int type, uint32_t nb) { DeviceGuard dg_(s); return rd_host(s, data, idx, type, nb, false, false); }
PXB_API int pxb_set_rigid_dynamic_data(PxbScene* s, const void* data, const uint32_t* idx, int type, uint32_t nb) { DeviceGuard dg_(s); return rd_host(s, const_cast<void*>(data), idx, type, nb, true, false); }
// Stream-ordered host variants (the PxDirectGPUAPI calls are asynchronous too: they take start / finish CUevents,
// PxDirectGPUAPI.h:311-463).  `data` must be PINNED host memory that stays valid until the next pxb_scene_fetch_results /
// pxb_scene_sync; a set takes effect for the next simulate, a get issued after pxb_scene_simulate returns that step's result.
PXB_API int pxb_get_rigid_dynamic_data_async(PxbScene* s, void* pinned, int type, uint32_t nb) { DeviceGuard dg_(s); return rd_host(s, pinned, nullptr, type, nb, false, true); }
PXB_API int pxb_set_rigid_dynamic_data_async(PxbScene* s, const void* pinned, int type, uint32_t nb) { DeviceGuard dg_(s); return rd_host(s, const_cast<void*>(pinned), nullptr, type, nb, true, true); }
// Pushes `bytes` (multiple of 16) from `devSrc` into nDst <= 8 peer-mapped device buffers with ONE kernel on `stream`
// (cudaStream_t; NULL = the scene stream).  Used by physx_b200/multi_gpu.py for the per-step all-gather of the state tensor.
PXB_API int pxb_scatter_to_peers(PxbScene* s, void* stream, const void* devSrc, size_t bytes, const uint64_t* devDstPtrs, uint32_t nDst, uint32_t ctas) { DeviceGuard dg_(s);
  if (!s || !devSrc || !devDstPtrs) return fail(PXB_ERR_INVALID, "null argument");
  if (nDst > 8 || (bytes & 15)) return fail(PXB_ERR_INVALID, "at most 8 destinations, size a multiple of 16 bytes");
  if (!nDst || !bytes) return PXB_OK;
  PeerPtrs P; for (uint32_t k = 0; k < 8; ++k) P.p[k] = k < nDst ? reinterpret_cast<float4*>(devDstPtrs[k]) : nullptr;
  cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
  k_scatter_to_peers<<<ctas ? ctas : 32, 256, 0, st>>>((const float4*)devSrc, bytes / 16, P, nDst);
  CK(cudaGetLastError());
  return PXB_OK;
}
// Fused state export (include/physx_b200.h).  The table is written by a one-thread kernel that takes it by value: stream-ordered before the
// next step, no host buffer has to outlive the call, and the captured step graph (which reads the table from device memory) stays valid.
PXB_API int pxb_scene_set_state_export(PxbScene* s, void* const* dst, uint32_t nDst, uint32_t rowOffset) { DeviceGuard dg_(s);
  if (!s || (nDst && !dst)) return fail(PXB_ERR_INVALID, "null argument");
  if (nDst > PXB_MAX_EXPORT) return fail(PXB_ERR_INVALID, "at most 9 export targets");
  if (nDst && s->hasCom) return fail(PXB_ERR_UNSUPPORTED, "the fused state export reports body frames: not available with centre-of-mass local poses (use pxb_scene_get_states_device)");
  ExportTable t; memset(&t, 0, sizeof(t));
  for (uint32_t k = 0; k < nDst; ++k) {
    if (!dst[k] || ((uintptr_t)dst[k] & 15u)) return fail(PXB_ERR_INVALID, "export targets must be non-null and 16-byte aligned");
    cudaPointerAttributes at; memset(&at, 0, sizeof(at));
    if (cudaPointerGetAttributes(&at, dst[k]) != cudaSuccess) { cudaGetLastError(); return fail(PXB_ERR_INVALID, "export target is not a CUDA-accessible pointer"); }
    if (at.type == cudaMemoryTypeUnregistered) return fail(PXB_ERR_INVALID, "export target is pageable host memory: pass device, peer-mapped or pinned (mapped) host memory");
    t.dst[k] = (float*)(at.type == cudaMemoryTypeHost && at.devicePointer ? at.devicePointer : dst[k]);   // mapped pinned host memory: the device-side alias
  }
  t.n = nDst; t.rowOffset = rowOffset;
  if ((nDst != 0) != s->exportOn) { s->exportOn = nDst != 0; drop_graphs(s); }   // the launch sequence gains / loses the export
  k_set_export<<<1, 32, 0, s->stream>>>(s->exportTab, t);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_peer_signal(PxbScene* s, void* stream, const uint64_t* flagPtrs, uint32_t n, uint32_t value) { DeviceGuard dg_(s);
  if (!s || (n && !flagPtrs)) return fail(PXB_ERR_INVALID, "null argument");
  if (n > PXB_MAX_EXPORT) return fail(PXB_ERR_INVALID, "at most 9 flags");
  if (!n) return PXB_OK;
  FlagPtrs P; for (uint32_t k = 0; k < PXB_MAX_EXPORT; ++k) P.f[k] = k < n ? reinterpret_cast<uint32_t*>(flagPtrs[k]) : nullptr;
  k_peer_signal<<<1, 32, 0, stream ? (cudaStream_t)stream : s->stream>>>(P, n, value);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_peer_wait(PxbScene* s, void* stream, const void* devFlags, uint32_t n, uint32_t value) { DeviceGuard dg_(s);
  if (!s || (n && !devFlags)) return fail(PXB_ERR_INVALID, "null argument");
  if (n > 32) return fail(PXB_ERR_INVALID, "at most 32 flags");
  if (!n) return PXB_OK;
  k_peer_wait<<<1, 32, 0, stream ? (cudaStream_t)stream : s->stream>>>((const uint32_t*)devFlags, n, value);
  CK(cudaGetLastError());
  return PXB_OK;
}
// ---------------------------------------------------------------------------------------------
// Standalone broadphase object (include/physx_b200.h pxb_bp_*): a2-a6 behind the reference's own Bp::BroadPhase interface.  The plugin shim's
// Bp::BroadPhase subclass (plugin/) forwards BroadPhaseUpdateData to pxb_bp_update and reads created / deleted pairs back with pxb_bp_fetch.
// Same kernels as the scene engine's device-wide broadphase (uniform grid keyed (env, cz, cy, cx), 8-bit LSD radix sorts, forward sweep,
// large objects against everything), with the reference's group / type-table / environment filter (GroupFilter).
struct PxbBroadPhase {
  int device = 0; cudaStream_t stream = nullptr; uint32_t cap = 0, capPairs = 0, bitsA = 1, nSlots = 0;
  float* bounds6 = 0; float* dist = 0; uint32_t *groups = 0, *envs = 0, *largeFlag = 0, *largeList = 0, *counters = 0, *cellVal = 0, *cellValAlt = 0, *valTmp = 0, *valAlt = 0;
  float4 *aabbMin = 0, *aabbMax = 0, *sMin = 0, *sMax = 0; uint64_t *cellKey = 0, *cellKeyAlt = 0, *keys[2] = {0, 0}, *keyAlt = 0, *created = 0, *deleted = 0;
  uint32_t* nPairsDev = 0; int cur = 0; uint32_t* hostCounters = 0; RadixSortTemp rs; GridParams grid; bool gridValid = false; uint32_t nLarge = 0, envCount = 1;
  std::vector<uint8_t> active; std::vector<uint32_t> outCreated, outDeleted; std::vector<uint64_t> tmpKeys; bool pending = false;
};
enum { BC_NPAIRS = 0, BC_NCREATED = 1, BC_NDELETED = 2, BC_N = 3, BC_ERROR = 4 /* = C_ERROR: bp_emit flags overflow there */, BC_COUNT = 8 };
__global__ void k_bpo_prepare(uint32_t n, const float* __restrict__ bounds6, const float* __restrict__ dist, const uint32_t* __restrict__ groups, const uint32_t* __restrict__ envs,
                              const uint32_t* __restrict__ largeFlag, GridParams g, uint32_t envCount, float4* __restrict__ aabbMin, float4* __restrict__ aabbMax, uint64_t* __restrict__ cellKey,
                              uint32_t* __restrict__ cellVal) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const uint32_t grp = groups[a]; const uint32_t env = envs ? envs[a] : NONE32;
  const float d = dist[a];   // inflation by the contact distance in float, as the reference's ABP does (BpBroadPhaseABP.cpp:1187-1197)
  const float mn0 = bounds6[a * 6 + 0] - d, mn1 = bounds6[a * 6 + 1] - d, mn2 = bounds6[a * 6 + 2] - d, mx0 = bounds6[a * 6 + 3] + d, mx1 = bounds6[a * 6 + 4] + d, mx2 = bounds6[a * 6 + 5] + d;
  aabbMin[a] = make_float4(mn0, mn1, mn2, __uint_as_float(env)); aabbMax[a] = make_float4(mx0, mx1, mx2, __uint_as_float(grp));
  uint64_t key = ~0ull;
  if (grp != NONE32 && !largeFlag[a]) {
    int cx = (int)floorf((mn0 - g.ox) * g.invCell), cy = (int)floorf((mn1 - g.oy) * g.invCell), cz = (int)floorf((mn2 - g.oz) * g.invCell);
    cx = max(0, min(g.nx - 1, cx)); cy = max(0, min(g.ny - 1, cy)); cz = max(0, min(g.nz - 1, cz));
    const uint64_t e = (env == NONE32) ? 0ull : (uint64_t)min(env, envCount - 1);
    key = ((e * (uint64_t)g.nz + (uint64_t)cz) * (uint64_t)g.ny + (uint64_t)cy) * (uint64_t)g.nx + (uint64_t)cx;
  }
  cellKey[a] = key; cellVal[a] = a;
}
__global__ void k_bpo_clamp(uint32_t* __restrict__ counters, uint32_t cap, uint32_t* __restrict__ nCur) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { if (counters[BC_NPAIRS] > cap) { counters[BC_ERROR] = 1u; counters[BC_NPAIRS] = cap; } *nCur = counters[BC_NPAIRS]; }
}
// created = new \ old; deleted = old \ new restricted to pairs whose two objects are still in the broadphase (lost overlaps caused by a
// removal are not reported, BpBroadPhase.h:181-192)
__global__ void k_bpo_diff(const uint64_t* __restrict__ a, const uint32_t* __restrict__ nAP, const uint64_t* __restrict__ b, const uint32_t* __restrict__ nBP, uint64_t* __restrict__ out,
                           uint32_t* __restrict__ outCount, const uint32_t* __restrict__ groups, uint32_t bitsA, int requireActive) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nA = *nAP, nB = *nBP;
  if (i >= nA) return;
  const uint64_t k = a[i];
  const uint32_t p = lower_bound_u64(b, nB, k);
  if (p < nB && b[p] == k) return;
  if (requireActive) { const uint32_t lo = (uint32_t)(k >> bitsA), hi = (uint32_t)(k & ((1ull << bitsA) - 1ull)); if (groups[lo] == NONE32 || groups[hi] == NONE32) return; }
  out[atomicAdd(outCount, 1u)] = k;
}
#define CKB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(PXB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
static void bpo_free(PxbBroadPhase* b) {
  void* ptrs[] = {b->bounds6, b->dist, b->groups, b->envs, b->largeFlag, b->largeList, b->counters, b->cellVal, b->cellValAlt, b->valTmp, b->valAlt, b->aabbMin, b->aabbMax, b->sMin, b->sMax, b->cellKey, b->cellKeyAlt,
                  b->keys[0], b->keys[1], b->keyAlt, b->created, b->deleted, b->nPairsDev, b->rs.blockHist, b->rs.digitTotals};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (b->hostCounters) cudaFreeHost(b->hostCounters);
  if (b->stream) cudaStreamDestroy(b->stream);
}
PXB_API int pxb_bp_create(uint32_t maxObjects, uint32_t maxPairs, int device, PxbBroadPhase** out) {
  if (!out) return fail(PXB_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(PXB_ERR_NO_DEVICE, "no CUDA device: physx_b200 has no CPU fallback"); }
  if (device < 0 || device >= ndev) return fail(PXB_ERR_INVALID, "bad device ordinal");
  DeviceGuard dg_(device);
  PxbBroadPhase* b = new PxbBroadPhase(); b->device = device;
  b->cap = std::max(64u, maxObjects); b->capPairs = maxPairs ? maxPairs : std::max(1024u, 8u * b->cap); b->bitsA = bits_for(b->cap);
  const size_t A = b->cap, P = b->capPairs;
  cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
  auto A_ = [&](auto*& p, size_t n) { if (e == cudaSuccess) e = dalloc(p, n); };
  A_(b->bounds6, A * 6); A_(b->dist, A); A_(b->groups, A); A_(b->envs, A); A_(b->largeFlag, A); A_(b->largeList, A); A_(b->counters, BC_COUNT); A_(b->cellVal, A); A_(b->cellValAlt, A); A_(b->valTmp, P); A_(b->valAlt, P);
  A_(b->aabbMin, A); A_(b->aabbMax, A); A_(b->sMin, A); A_(b->sMax, A); A_(b->cellKey, A); A_(b->cellKeyAlt, A); A_(b->keys[0], P); A_(b->keys[1], P); A_(b->keyAlt, P); A_(b->created, P); A_(b->deleted, P);
  A_(b->nPairsDev, 2); A_(b->rs.blockHist, RS_MAX_CTAS * 256); A_(b->rs.digitTotals, 256);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&b->hostCounters, 4 * (BC_COUNT + 2));
  cudaDeviceProp prop; if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { bpo_free(b); delete b; cudaGetLastError(); return fail(PXB_ERR_CUDA, std::string("pxb_bp_create: ") + cudaGetErrorString(e)); }
  b->rs.ctas = std::min<uint32_t>(RS_MAX_CTAS, (uint32_t)prop.multiProcessorCount * 2);
  cudaMemsetAsync(b->groups, 0xff, 4 * A, b->stream); cudaMemsetAsync(b->envs, 0xff, 4 * A, b->stream); cudaMemsetAsync(b->largeFlag, 0, 4 * A, b->stream); cudaMemsetAsync(b->nPairsDev, 0, 8, b->stream);
  cudaMemsetAsync(b->counters, 0, 4 * BC_COUNT, b->stream); cudaMemsetAsync(b->bounds6, 0, 24 * A, b->stream); cudaMemsetAsync(b->dist, 0, 4 * A, b->stream);
  cudaStreamSynchronize(b->stream);
  b->active.assign(A, 0);
  *out = b;
  return PXB_OK;
}
PXB_API void pxb_bp_release(PxbBroadPhase* b) { if (!b) return; DeviceGuard dg_(b->device); cudaStreamSynchronize(b->stream); bpo_free(b); delete b; }
// Grid for the objects currently in the broadphase: cell edge = the largest inflated extent among regular objects (x 1.02), objects more than
// 8 x the median extent (and unbounded ones: planes) are "large" and tested against everything.  Rebuilt on frames that add objects and when an
// updated object has grown past the cell; objects that wander outside are clamped to the border cells (still exact, only slower).
static int bpo_rebuild_grid(PxbBroadPhase* b, const float* bounds6, const float* dist, const uint32_t* envIds) {
  std::vector<float> ext; ext.reserve(b->nSlots);
  auto extent = [&](uint32_t a) { const float* q = bounds6 + 6 * (size_t)a; return std::max(q[3] - q[0], std::max(q[4] - q[1], q[5] - q[2])) + 2.f * dist[a]; };
  for (uint32_t a = 0; a < b->nSlots; ++a) if (b->active[a]) { const float x = extent(a); if (std::isfinite(x) && x < 1e18f) ext.push_back(x); }
  float med = 1.f;
  if (!ext.empty()) { std::nth_element(ext.begin(), ext.begin() + ext.size() / 2, ext.end()); med = ext[ext.size() / 2]; }
  const float largeThresh = 8.f * med;
  std::vector<uint32_t> lf(b->nSlots, 0), ll; float cell = 0.f; float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}; uint32_t maxEnv = 0; bool anyEnv = false;
  for (uint32_t a = 0; a < b->nSlots; ++a) {
    if (!b->active[a]) continue;
    const float x = extent(a);
    if (!(std::isfinite(x) && x < 1e18f) || x > largeThresh) { lf[a] = 1; ll.push_back(a); continue; }
    cell = std::max(cell, x);
    for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], bounds6[6 * (size_t)a + k]); mx[k] = std::max(mx[k], bounds6[6 * (size_t)a + k]); }
    if (envIds && envIds[a] != NONE32) { anyEnv = true; maxEnv = std::max(maxEnv, envIds[a]); }
  }
  cell *= 1.02f; if (!(cell > 0.f)) cell = 1.f;
  GridParams g; g.invCell = 1.0f / cell; int n[3];
  for (int k = 0; k < 3; ++k) { if (!std::isfinite(mn[k])) { mn[k] = 0.f; mx[k] = 0.f; } n[k] = std::max(4, (int)std::ceil(((mx[k] - mn[k]) + 16.f * cell) / cell) + 1); }
  g.ox = mn[0] - 8.f * cell; g.oy = mn[1] - 8.f * cell; g.oz = mn[2] - 8.f * cell; g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
  const uint64_t envCount = anyEnv ? (uint64_t)maxEnv + 1 : 1;
  while ((long double)envCount * g.nx * g.ny * g.nz > 4.0e18L) { if (g.nx >= g.ny && g.nx >= g.nz) g.nx = (g.nx + 1) / 2; else if (g.ny >= g.nz) g.ny = (g.ny + 1) / 2; else g.nz = (g.nz + 1) / 2; }
  g.keyBits = bits_for((uint64_t)envCount * (uint64_t)g.nx * (uint64_t)g.ny * (uint64_t)g.nz + 1);
  b->grid = g; b->envCount = (uint32_t)envCount; b->nLarge = (uint32_t)ll.size(); b->gridValid = true;
  CKB(cudaMemcpyAsync(b->largeFlag, lf.data(), 4 * (size_t)b->nSlots, cudaMemcpyHostToDevice, b->stream));
  if (b->nLarge) CKB(cudaMemcpyAsync(b->largeList, ll.data(), 4 * (size_t)b->nLarge, cudaMemcpyHostToDevice, b->stream));
  CKB(cudaStreamSynchronize(b->stream));   // lf / ll are locals
  return PXB_OK;
}
PXB_API int pxb_bp_update(PxbBroadPhase* b, const float* bounds6, const float* contactDist, const uint32_t* groups, const uint32_t* envIds, uint32_t capacity, const uint8_t* lut49,
                          const uint32_t* created, uint32_t nCreated, const uint32_t* updated, uint32_t nUpdated, const uint32_t* removed, uint32_t nRemoved) {
  if (!b || !bounds6 || !contactDist || !groups || !lut49) return fail(PXB_ERR_INVALID, "null argument");
  if (capacity > b->cap) return fail(PXB_ERR_CAPACITY, "more broadphase objects than pxb_bp_create was sized for");
  DeviceGuard dg_(b->device);
  cudaStream_t st = b->stream;
  for (uint32_t i = 0; i < nRemoved; ++i) { if (removed[i] >= capacity) return fail(PXB_ERR_INVALID, "removed handle out of range"); b->active[removed[i]] = 0; }
  for (uint32_t i = 0; i < nCreated; ++i) { if (created[i] >= capacity) return fail(PXB_ERR_INVALID, "created handle out of range"); b->active[created[i]] = 1; b->nSlots = std::max(b->nSlots, created[i] + 1); }
  const uint32_t n = b->nSlots;
  // device copies of the host arrays (the shim hands over pinned memory: Bp::BoundsArray and the AABB manager's arrays are allocated through the
  // plugin's host allocator).  Removed / never-added slots carry group eINVALID on the device whatever the host array holds.
  std::vector<uint32_t> g(n);
  for (uint32_t a = 0; a < n; ++a) g[a] = b->active[a] ? groups[a] : NONE32;
  bool needGrid = !b->gridValid || nCreated > 0;
  if (!needGrid) {   // an updated regular object that outgrew the cell edge would be missed by the one-cell neighbourhood
    const float cell = 1.0f / b->grid.invCell;
    for (uint32_t i = 0; i < nUpdated && !needGrid; ++i) { const uint32_t a = updated[i]; if (a < n && b->active[a]) { const float* q = bounds6 + 6 * (size_t)a; if (std::max(q[3] - q[0], std::max(q[4] - q[1], q[5] - q[2])) + 2.f * contactDist[a] > cell) needGrid = true; } }
  }
  if (needGrid) { if (int rc = bpo_rebuild_grid(b, bounds6, contactDist, envIds)) return rc; }
  if (!n) { b->pending = true; return PXB_OK; }
  CKB(cudaMemcpyAsync(b->bounds6, bounds6, 24 * (size_t)n, cudaMemcpyHostToDevice, st)); CKB(cudaMemcpyAsync(b->dist, contactDist, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  CKB(cudaMemcpyAsync(b->groups, g.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  if (envIds) CKB(cudaMemcpyAsync(b->envs, envIds, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  CKB(cudaStreamSynchronize(st));   // g is a local; the host arrays may change as soon as update() returns
  unsigned long long lut = 0; for (int i = 0; i < 49; ++i) if (lut49[i]) lut |= 1ull << i;
  const int prev = b->cur; b->cur ^= 1; const int cur = b->cur;
  CKB(cudaMemsetAsync(b->counters, 0, 4 * BC_COUNT, st));
  CKB(cudaMemcpyAsync(b->counters + BC_N, &n, 4, cudaMemcpyHostToDevice, st));
  k_bpo_prepare<<<cdiv(n, 256), 256, 0, st>>>(n, b->bounds6, b->dist, b->groups, envIds ? b->envs : nullptr, b->largeFlag, b->grid, b->envCount, b->aabbMin, b->aabbMax, b->cellKey, b->cellVal);
  const int r = radix_sort_pairs(b->cellKey, b->cellVal, b->cellKeyAlt, b->cellValAlt, b->counters + BC_N, b->grid.keyBits, b->rs, st);
  const uint64_t* sk = r ? b->cellKeyAlt : b->cellKey; const uint32_t* sv = r ? b->cellValAlt : b->cellVal;
  k_bp_gather<<<cdiv(n, 256), 256, 0, st>>>(n, sv, b->aabbMin, b->aabbMax, b->sMin, b->sMax);
  const bool oddPasses = (((2 * b->bitsA + 7) / 8) & 1u) != 0;
  uint64_t* emit = oddPasses ? b->keyAlt : b->keys[cur]; uint64_t* other = oddPasses ? b->keys[cur] : b->keyAlt;
  GroupFilter F; F.lut = lut; F.large = b->largeFlag;
  // the pair kernels count into counters[C_NPAIRS_NEW] and flag counters[C_ERROR]: the same slots as BC_NPAIRS / BC_ERROR
  k_bp_pairs<GroupFilter><<<cdiv(n, 128), 128, 0, st>>>(n, sk, sv, b->sMin, b->sMax, b->grid, b->bitsA, emit, b->counters, b->capPairs, F);
  if (b->nLarge) k_bp_large<GroupFilter><<<cdiv(n, 256), 256, 0, st>>>(n, b->nLarge, b->largeList, b->aabbMin, b->aabbMax, b->bitsA, emit, b->counters, b->capPairs, F);
  k_bpo_clamp<<<1, 32, 0, st>>>(b->counters, b->capPairs, b->nPairsDev + cur);
  radix_sort_pairs(emit, b->valTmp, other, b->valAlt, b->nPairsDev + cur, 2 * b->bitsA, b->rs, st);
  const uint32_t gP = cdiv(b->capPairs, 256);
  k_bpo_diff<<<gP, 256, 0, st>>>(b->keys[cur], b->nPairsDev + cur, b->keys[prev], b->nPairsDev + prev, b->created, b->counters + BC_NCREATED, b->groups, b->bitsA, 0);
  k_bpo_diff<<<gP, 256, 0, st>>>(b->keys[prev], b->nPairsDev + prev, b->keys[cur], b->nPairsDev + cur, b->deleted, b->counters + BC_NDELETED, b->groups, b->bitsA, 1);
  CKB(cudaMemcpyAsync(b->hostCounters, b->counters, 4 * BC_COUNT, cudaMemcpyDeviceToHost, st));
  CKB(cudaGetLastError());
  b->pending = true;
  return PXB_OK;
}
PXB_API int pxb_bp_fetch(PxbBroadPhase* b, const uint32_t** createdPairs, uint32_t* nCreated, const uint32_t** deletedPairs, uint32_t* nDeleted) {
  if (!b || !createdPairs || !nCreated || !deletedPairs || !nDeleted) return fail(PXB_ERR_INVALID, "null argument");
  DeviceGuard dg_(b->device);
  b->outCreated.clear(); b->outDeleted.clear();
  if (b->pending && b->nSlots) {
    CKB(cudaStreamSynchronize(b->stream));
    if (b->hostCounters[BC_ERROR]) return fail(PXB_ERR_CAPACITY, "broadphase pair capacity (maxPairs) exceeded");
    auto pull = [&](const uint64_t* dev, uint32_t n, std::vector<uint32_t>& out) -> int {
      b->tmpKeys.resize(n);
      if (n) { CKB(cudaMemcpyAsync(b->tmpKeys.data(), dev, 8 * (size_t)n, cudaMemcpyDeviceToHost, b->stream)); CKB(cudaStreamSynchronize(b->stream)); }
      std::sort(b->tmpKeys.begin(), b->tmpKeys.end());
      out.resize(2 * (size_t)n);
      for (uint32_t i = 0; i < n; ++i) { out[2 * i] = (uint32_t)(b->tmpKeys[i] >> b->bitsA); out[2 * i + 1] = (uint32_t)(b->tmpKeys[i] & ((1ull << b->bitsA) - 1ull)); }
      return PXB_OK;
    };
    if (int rc = pull(b->created, b->hostCounters[BC_NCREATED], b->outCreated)) return rc;
    if (int rc = pull(b->deleted, b->hostCounters[BC_NDELETED], b->outDeleted)) return rc;
  }
  b->pending = false;
  *createdPairs = b->outCreated.data(); *nCreated = (uint32_t)(b->outCreated.size() / 2); *deletedPairs = b->outDeleted.data(); *nDeleted = (uint32_t)(b->outDeleted.size() / 2);
  return PXB_OK;
}

// f3 tensor front end (include/physx_b200.h): device tensors in the ovphysx wire formats, stream-ordered on the scene stream.
PXB_API int pxb_tensor_read_device(PxbScene* s, int tensorType, void* devOut, const uint32_t* devIdx, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || !devOut) return fail(PXB_ERR_INVALID, "null argument");
  if (tensorType != PXB_TENSOR_RIGID_BODY_POSE && tensorType != PXB_TENSOR_RIGID_BODY_VELOCITY && tensorType != PXB_TENSOR_RIGID_BODY_MASS && tensorType != PXB_TENSOR_RIGID_BODY_INV_MASS)
    return fail(PXB_ERR_INVALID, "tensor type is not readable");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if (!devIdx && nb > s->nDyn) return fail(PXB_ERR_INVALID, "nb exceeds the number of dynamic bodies");
  if (!nb) return PXB_OK;
  if (int rc = join_pending(s)) return rc;
  const bool poseIO = s->hasCom && tensorType == PXB_TENSOR_RIGID_BODY_POSE;
  if (poseIO) { if (int rc = refresh_actor_poses(s, s->stream)) return rc; }
  k_tensor_read<<<cdiv(nb, 256), 256, 0, s->stream>>>(nb, devIdx, s->dynActorDev, tensorType, poseIO ? s->actorPos : s->pos, poseIO ? s->actorQuat : s->quat, s->linVel, s->angVel, (float*)devOut);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_tensor_write_device(PxbScene* s, int tensorType, const void* devIn, const uint32_t* devIdx, uint32_t nb) { DeviceGuard dg_(s);
  if (!s || !devIn) return fail(PXB_ERR_INVALID, "null argument");
  if (tensorType != PXB_TENSOR_RIGID_BODY_POSE && tensorType != PXB_TENSOR_RIGID_BODY_VELOCITY && tensorType != PXB_TENSOR_RIGID_BODY_FORCE && tensorType != PXB_TENSOR_RIGID_BODY_WRENCH)
    return fail(PXB_ERR_INVALID, "tensor type is not writable");
  if (s->stepping) return fail(PXB_ERR_INVALID, "illegal while the simulation is running");
  if (!devIdx && nb > s->nDyn) return fail(PXB_ERR_INVALID, "nb exceeds the number of dynamic bodies");
  if (!nb) return PXB_OK;
  if ((tensorType == PXB_TENSOR_RIGID_BODY_FORCE || tensorType == PXB_TENSOR_RIGID_BODY_WRENCH) && !s->forcesUsed) { s->forcesUsed = true; drop_graphs(s); }
  if (int rc = join_pending(s)) return rc;
  k_tensor_write<<<cdiv(nb, 256), 256, 0, s->stream>>>(nb, devIdx, s->dynActorDev, tensorType, s->pos, s->quat, s->linVel, s->angVel, (const float*)devIn, s->wake, s->asleep, s->extForce, s->extTorque);
  if (s->hasCom && tensorType == PXB_TENSOR_RIGID_BODY_POSE) k_actor_to_body<<<cdiv(nb, 256), 256, 0, s->stream>>>(nb, devIdx, s->dynActorDev, s->pos, s->quat, s->b2aP, s->b2aQ);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_scene_sync(PxbScene* s) { DeviceGuard dg_(s); if (!s) return fail(PXB_ERR_INVALID, "null scene"); CK(cudaStreamSynchronize(s->stream)); return PXB_OK; }

// Packed 13-float state of every dynamic body written straight into a DEVICE buffer (e.g. this rank's slice of
// the NCCL all-gather receive tensor); asynchronous on the scene stream.
PXB_API int pxb_scene_get_states_device(PxbScene* s, float* devOut) { DeviceGuard dg_(s);
  if (s) { if (int rc = join_pending(s)) return rc; }
  if (!s || !devOut) return fail(PXB_ERR_INVALID, "null argument");
  if (!s->nDyn) return PXB_OK;   // stream-ordered: legal right after pxb_scene_simulate, it reads the state that step produces
  cudaStream_t st = s->stream;
  if (int rc = refresh_actor_poses(s, st)) return rc;
  LAUNCH(k_states_get, cdiv(s->nDyn, 256), 256, s->nDyn, s->dynActorDev, s->hasCom ? s->actorPos : s->pos, s->hasCom ? s->actorQuat : s->quat, s->linVel, s->angVel, devOut);
  CK(cudaGetLastError());
  return PXB_OK;
}
PXB_API int pxb_scene_get_states(PxbScene* s, float* out) { DeviceGuard dg_(s);
  if (s) { if (int rc = join_pending(s)) return rc; }
  if (!s || !out) return fail(PXB_ERR_INVALID, "null argument");
  if (!s->nDyn) return PXB_OK;
  cudaStream_t st = s->stream; float* d = s->stage;   // persistent staging (26 floats per actor): no allocation per call
  if (int rc = refresh_actor_poses(s, st)) return rc;
  LAUNCH(k_states_get, cdiv(s->nDyn, 256), 256, s->nDyn, s->dynActorDev, s->hasCom ? s->actorPos : s->pos, s->hasCom ? s->actorQuat : s->quat, s->linVel, s->angVel, d);
  CK(cudaMemcpyAsync(out, d, 52 * (size_t)s->nDyn, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
  return PXB_OK;
}
PXB_API int pxb_scene_set_states(PxbScene* s, const float* in) { DeviceGuard dg_(s);
  if (s) { if (int rc = join_pending(s)) return rc; }
  if (!s || !in) return fail(PXB_ERR_INVALID, "null argument");
  if (!s->nDyn) return PXB_OK;
  cudaStream_t st = s->stream; float* d = s->stage + (size_t)s->capA * 13;
  CK(cudaMemcpyAsync(d, in, 52 * (size_t)s->nDyn, cudaMemcpyHostToDevice, st));
  LAUNCH(k_states_set, cdiv(s->nDyn, 256), 256, s->nDyn, s->dynActorDev, s->pos, s->quat, s->linVel, s->angVel, d, s->wake, s->asleep);
  if (s->hasCom) LAUNCH(k_actor_to_body, cdiv(s->nDyn, 256), 256, s->nDyn, (const uint32_t*)nullptr, s->dynActorDev, s->pos, s->quat, s->b2aP, s->b2aQ);
  CK(cudaStreamSynchronize(st));
  return PXB_OK;
}

}  // extern "C"
