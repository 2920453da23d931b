// pxb_common.cuh -- device helpers and constants shared by every translation unit of libphysx_b200.so (pxb_engine.cu: host side + small
// kernels; pxb_narrowphase.cu; pxb_env.cu; pxb_solve.cu).  Everything here is inline device code or plain data.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/physx_b200.h"
#include "pxb_math.cuh"
#include "pxb_np.cuh"
#include "pxb_gjk.cuh"
#include "pxb_solver.cuh"

namespace cg = cooperative_groups;

#define NONE32 0xffffffffu
#define MAX_PARTITIONS 160   // 64 dynamic colours (two rounds of the reference's 32, DyConstraintPartition.cpp:520-552) + static slots
#ifndef PXB_SOLVE_CTAS_PER_SM
#define PXB_SOLVE_CTAS_PER_SM 2   // measured on B200: 3 CTAs/SM (80 regs, spills, wider grid.sync) is 20% slower than 2
#endif

// ---------------------------------------------------------------------------------------------
// host-side record (layout of oracle/scene_format.h::PxbActorRec, 128 bytes)
struct ActorRec {
  uint32_t flags, geomType, envId, hullIdx;
  float pos[3], quat[4], dims[4], linVel[3], angVel[3], mass, inertia[3], linDamping, angDamping, maxLinVel, maxAngVel, maxDepenetrationVel; uint32_t materialIndex; uint32_t aggregate;
};
static_assert(sizeof(ActorRec) == 128, "actor record layout");

enum Counter { C_NPAIRS_NEW = 0, C_NCREATED, C_NDELETED, C_FREE_HEAD, C_ERROR, C_NCON, C_NPART, C_REMAINING, C_NA, C_NORDER, C_NDYNCON, C_FREE_TAIL, C_FREE_SNAP, C_MAXCONENV, C_MAXPAIRENV, C_NGJK, C_NTOUCH_FOUND, C_NTOUCH_LOST, C_NGJK_QUERY, C_NGJK_FULL, C_NGJK_EPA, C_NBOXGEN, C_COUNT = 24 };
enum ErrorBits { E_PAIR_OVERFLOW = 1, E_COLOUR_OVERFLOW = 2, E_PARTITION_OVERFLOW = 4, E_UNSUPPORTED_PAIR = 8, E_BAD_INDEX = 16 /* a device-side index list named a body that does not exist / is not kinematic */ };

struct GridParams { float ox, oy, oz, invCell; int nx, ny, nz; uint32_t keyBits; };
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
__device__ __forceinline__ unsigned long long ld_volatile64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
__device__ __forceinline__ void st_volatile64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// a1: tight world AABB of one shape.  Formulas: Gu::computeBounds (geomutils/src/GuBounds.cpp:354-400, plane :210-260).
// Gu::computeTightBounds (GuBounds.cpp:301-352; PxConvexMeshGeometry defaults to eTIGHT_BOUNDS): rotated vertices, last vertex first.  Out of line:
// the bounds kernels of box / sphere scenes keep their register budget.
static __device__ __noinline__ void hull_tight_bounds(const HullArrays* hulls, uint32_t hullIdx, v3 p, q4 q, float* mn, float* mx) {
  const DevHull h = load_hull(*hulls, hullIdx);
  const m33 b = amfromq(q);
  v3 lo = V3(0, 0, 0), hi = V3(0, 0, 0);
  for (uint32_t k = 0; k < h.nVerts; ++k) {
    const v3 v = h.vert(k == 0 ? h.nVerts - 1 : k - 1);
    const v3 w = (b.c0 * v.x + b.c1 * v.y) + b.c2 * v.z;
    if (k == 0) { lo = w; hi = w; } else { lo = vmin(lo, w); hi = vmax(hi, w); }
  }
  hi = hi + p; lo = lo + p;
  const v3 c = (hi + lo) * 0.5f, e = (hi - lo) * 0.5f;
  mn[0] = c.x - e.x; mn[1] = c.y - e.y; mn[2] = c.z - e.z; mx[0] = c.x + e.x; mx[1] = c.y + e.y; mx[2] = c.z + e.z;
}
__device__ __forceinline__ void tight_bounds(uint32_t type, v3 p, q4 q, float4 d, float* mn, float* mx, const HullArrays* hulls = nullptr) {
  if (hulls && type == PXB_GEOM_CONVEXMESH) { hull_tight_bounds(hulls, __float_as_uint(d.x), p, q, mn, mx); return; }
  v3 e = V3(0, 0, 0); bool plane = false;
  if (type == PXB_GEOM_SPHERE) e = V3(d.x, d.x, d.x);
  else if (type == PXB_GEOM_CAPSULE) { const v3 dd = qbasis0(q) * d.y; e = V3(fabsf(dd.x) + d.x, fabsf(dd.y) + d.x, fabsf(dd.z) + d.x); }
  else if (type == PXB_GEOM_BOX) {
    const m33 b = amfromq(q);
    const v3 c0 = b.c0 * d.x, c1 = b.c1 * d.y, c2 = b.c2 * d.z;
    e = V3((fabsf(c0.x) + fabsf(c1.x)) + fabsf(c2.x), (fabsf(c0.y) + fabsf(c1.y)) + fabsf(c2.y), (fabsf(c0.z) + fabsf(c1.z)) + fabsf(c2.z));
  } else if (type == PXB_GEOM_PLANE) plane = true;
  if (!plane) { mn[0] = p.x - e.x; mn[1] = p.y - e.y; mn[2] = p.z - e.z; mx[0] = p.x + e.x; mx[1] = p.y + e.y; mx[2] = p.z + e.z; }
  else {
    const float big = FLT_MAX * 0.25f;
    mn[0] = mn[1] = mn[2] = -big; mx[0] = mx[1] = mx[2] = big;
    const v3 n = qbasis0(q); const float dd = -dot(p, n);
    const float nx = fabsf(n.x), ny = fabsf(n.y), nz = fabsf(n.z); const float eps = 1e-6f, ome = 1.0f - eps;
    if (nx > ome && ny < eps && nz < eps) { if (n.x > 0.f) mx[0] = -dd; else mn[0] = dd; }
    else if (nx < eps && ny > ome && nz < eps) { if (n.y > 0.f) mx[1] = -dd; else mn[1] = dd; }
    else if (nx < eps && ny < eps && nz > ome) { if (n.z > 0.f) mx[2] = -dd; else mn[2] = dd; }
  }
}
// pair filter: closed-interval overlap on all axes (PxgIntegerAABB::intersects / ABP intersect2D semantics),
// at least one dynamic actor (BpFiltering.h:99-114 groups), equal-or-invalid environment ids (broadphase.cu:62-80)
// geomFlags word: bits 0..7 geometry type, 0x100 dynamic (PxRigidDynamic), 0x200 global / oversize object, 0x400 removed, 0x800 kinematic (PxRigidBodyFlag::eKINEMATIC),
// 0x1000 PxActorFlag::eDISABLE_GRAVITY, 0x2000 PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES,
// bits 16..21 PxRigidDynamicLockFlags.  A kinematic body is a PxRigidDynamic that the solver treats like a static one (infinite mass, second body of its pairs).
__device__ __forceinline__ bool gf_dynamic(uint32_t gf) { return (gf & 0x900u) == 0x100u; }
__device__ __forceinline__ bool bp_test(const float4& amin, const float4& amax, const float4& bmin, const float4& bmax) {
  if (amin.x > bmax.x || bmin.x > amax.x || amin.y > bmax.y || bmin.y > amax.y || amin.z > bmax.z || bmin.z > amax.z) return false;
  const uint32_t fa = __float_as_uint(amax.w), fb = __float_as_uint(bmax.w);
  if (!(gf_dynamic(fa) || gf_dynamic(fb))) return false;   // static-static, kinematic-static and kinematic-kinematic pairs are filtered (BpFiltering.cpp:36-48 with PxPairFilteringMode::eDEFAULT)
  const uint32_t ea = __float_as_uint(amin.w), eb = __float_as_uint(bmin.w);
  if (ea != NONE32 && eb != NONE32 && ea != eb) return false;
  return true;
}
__device__ __forceinline__ void bp_emit(uint32_t a, uint32_t b, uint32_t bitsA, uint64_t* __restrict__ keys, uint32_t* __restrict__ cnt, uint32_t cap, uint32_t* __restrict__ err) {
  const uint32_t lo = min(a, b), hi = max(a, b);
  const uint32_t idx = atomicAdd(cnt, 1u);   // (a warp-aggregated increment was measured and changes nothing: config 4 broadphase stage 0.577 vs 0.578 ms, profiles/README.md)
  if (idx < cap) keys[idx] = ((uint64_t)lo << bitsA) | hi; else atomicOr(err, (uint32_t)E_PAIR_OVERFLOW);
}
__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* __restrict__ k, uint32_t n, uint64_t v) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (k[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

// a18 sleeping.  Per body: Dy::sleepCheck / updateWakeCounter, non-stabilised branch (lowleveldynamics/src/DySleep.cpp:169-236;
// GPU reference: sleepCheck / updateWakeCounter in gpusolver integration.cuh:40-435).  Per island (the host island manager's
// job in the reference pipeline, IG::IslandSim): k_sleep_islands puts an island to sleep once every body in it is ready and
// wakes sleeping bodies that touch an awake island.  Disabled (threshold 0) for throughput runs, SURVEY.md 8d.
// a1, transform cache with local poses (PxShape::setLocalPose / PxRigidBody::setCMassLocalPose).  pos / quat hold body2World (the centre-of-mass frame the solver
// integrates); the shape's world pose is body2World * shape2Body with shape2Body = body2Actor^-1 * shape2Actor precomputed on the host, composed in the aos operation
// order of Cm::getDynamicGlobalPoseAligned / getStaticGlobalPoseAligned (common/src/CmTransformUtils.h:40-133, transformFast).  The narrowphase reads the cache.
struct LocalPoses { const float4 *s2bP, *s2bQ; float4 *tcPos, *tcQuat; };   // s2bP == nullptr: no local poses, shapes sit at pos / quat
__device__ __forceinline__ xf atransform_fast(const xf& a, const xf& b) {
  const float wa = a.q.w, wb = b.q.w; const v3 va = V3(a.q.x, a.q.y, a.q.z), vb = V3(b.q.x, b.q.y, b.q.z);
  const float wo = wa * wb - adot(va, vb);
  const v3 vo = scaleadd(va, wb, scaleadd(vb, wa, cross(va, vb)));
  const v3 t1 = b.p * (wa * wa + (-0.5f));
  const v3 t2 = scaleadd(cross(va, b.p), wa, t1);
  const v3 t3 = scaleadd(va, adot(va, b.p), t2);
  xf o; o.p = scaleadd(t3, 2.f, a.p); o.q = Q4(vo.x, vo.y, vo.z, wo); return o;
}
__device__ __forceinline__ xf shape_world_pose(const LocalPoses& L, uint32_t a, float4 p4, float4 q4_, bool writeCache = true) {
  xf t; t.p = V3(p4.x, p4.y, p4.z); t.q = Q4(q4_);
  if (!L.s2bP) return t;
  xf l; { const float4 lp = L.s2bP[a]; l.p = V3(lp.x, lp.y, lp.z); l.q = Q4(L.s2bQ[a]); }
  const xf w = atransform_fast(t, l);
  if (writeCache) { L.tcPos[a] = F4(w.p, p4.w); L.tcQuat[a] = F4(w.q); }
  return w;
}

// f1: PxDefaultSimulationFilterShader on the device (physxextensions/src/ExtDefaultSimulationFilterShader.cpp:238-280, without the trigger branch): collision-group
// table (PxSetGroupCollisionFlag), groups masks with PxSetFilterOps / PxSetFilterConstants / PxSetFilterBool.  data: PxFilterData (word0..3) per actor.
struct FilterConfig { uint32_t collisionTable[32]; uint32_t ops[3]; uint32_t filterBool; uint32_t constants[4]; };
struct FilterArgs { const uint4* data; FilterConfig cfg; const float2* shapeOff; };   // shapeOff: per-actor (contactOffset, restOffset) for the pair's contact distance, or null
__device__ __forceinline__ uint32_t filter_op16x2(uint32_t op, uint32_t x, uint32_t y) {   // two PxGroupsMask halves at once (16-bit lanes; the NOT forms stay inside 32 bits)
  return op == 1u ? (x | y) : (op == 2u ? (x ^ y) : (op == 3u ? ~(x & y) : (op == 4u ? ~(x | y) : (op == 5u ? ~(x ^ y) : (x & y)))));
}
__device__ __forceinline__ void filter_op(uint32_t op, uint32_t a2, uint32_t a3, uint32_t b2, uint32_t b3, uint32_t& r2, uint32_t& r3) {
  if (op == 6u) { const uint32_t t = b2; b2 = b3; b3 = t; }   // SWAP_AND: bits0 & bits2, bits1 & bits3, bits2 & bits0, bits3 & bits1 (word2 = bits0 | bits1 << 16, word3 = bits2 | bits3 << 16)
  r2 = filter_op16x2(op, a2, b2); r3 = filter_op16x2(op, a3, b3);
}
__device__ __forceinline__ bool filter_suppressed(const FilterArgs& F, uint32_t actor0, uint32_t actor1) {
  const uint4 f0 = F.data[actor0], f1 = F.data[actor1];
  if (!((F.cfg.collisionTable[f0.x & 31u] >> (f1.x & 31u)) & 1u)) return true;
  uint32_t a2, a3, b2, b3, r2, r3;
  filter_op(F.cfg.ops[0], f0.z, f0.w, F.cfg.constants[0], F.cfg.constants[1], a2, a3);
  filter_op(F.cfg.ops[1], f1.z, f1.w, F.cfg.constants[2], F.cfg.constants[3], b2, b3);
  filter_op(F.cfg.ops[2], a2, a3, b2, b3, r2, r3);
  return ((r2 | r3) != 0u) != (F.cfg.filterBool != 0u);
}

struct SleepArgs { float threshold, dt; float* wake; float4 *accLin, *accAng; uint32_t *asleep, *nInter; };
__device__ __forceinline__ bool body_asleep(const SleepArgs& S, uint32_t a) { return S.threshold > 0.f && S.asleep[a] != 0u; }
__device__ __forceinline__ void sleep_check_dev(const SleepArgs& S, uint32_t a, q4 q, float4 invInertia, float invMassIn, v3 motionLin, v3 motionAng) {
  const float wakeCounterResetTime = 20.0f * 0.02f;
  float wc = S.wake[a];
  if (wc < wakeCounterResetTime * 0.5f || wc < S.dt) {
    const v3 inertia = V3(invInertia.x > 0.f ? 1.0f / invInertia.x : 1.0f, invInertia.y > 0.f ? 1.0f / invInertia.y : 1.0f, invInertia.z > 0.f ? 1.0f / invInertia.z : 1.0f);
    const v3 accL = V3(S.accLin[a]) + motionLin, accA = V3(S.accAng[a]) + qrotinv(q, motionAng);
    const float invMass = invMassIn == 0.0f ? 1.0f : invMassIn;
    const float angular = dot(vmul(accA, accA), inertia) * invMass, linear = lensq(accL);
    const float normalizedEnergy = 0.5f * (angular + linear);
    const float clusterFactor = (float)(1u + S.nInter[a]);
    const float threshold = clusterFactor * S.threshold;
    if (normalizedEnergy >= threshold) {
      S.accLin[a] = make_float4(0, 0, 0, 0); S.accAng[a] = make_float4(0, 0, 0, 0);   // resetSleepFilter
      const float ratio = normalizedEnergy / threshold;
      const float factor = threshold == 0.0f ? 2.0f : (ratio < 2.0f ? ratio : 2.0f);
      S.wake[a] = factor * 0.5f * wakeCounterResetTime + S.dt * (clusterFactor - 1.0f);
      return;
    }
    S.accLin[a] = F4(accL, 0.f); S.accAng[a] = F4(accA, 0.f);
  }
  wc = wc - S.dt; if (!(wc > 0.0f)) wc = 0.0f;
  S.wake[a] = wc;
  if (wc == 0.0f) { S.accLin[a] = make_float4(0, 0, 0, 0); S.accAng[a] = make_float4(0, 0, 0, 0); }
}
__device__ __forceinline__ m33 load_sym(const float4 A, const float4 B) {
  m33 m; m.c0 = V3(A.x, A.y, A.z); m.c1 = V3(A.y, A.w, B.x); m.c2 = V3(A.z, B.x, B.y); return m;
}

// a14 inputs of one constraint: body frames, inverse masses, pre-solver velocities, world sqrt(inverse inertia)
struct PrepBodies { xf f0, f1; float invMass0, invMass1, pen0, pen1; v3 linVel0, linVel1, angVel0, angVel1; m33 sI0, sI1; };


// Fused state export (pxb_scene_set_state_export): the table lives in device memory so that the captured step graph stays valid while the
// targets alternate (double buffering); n = 0 switches the export off.  Targets are device-accessible buffers of packed 13-float records
// (pos3 quat4 linVel3 angVel3, dynamic-body order): this GPU's memory, peer-mapped memory (P2P stores over NVLink) or mapped pinned host memory.
#define PXB_MAX_EXPORT 9
struct ExportTable { float* dst[PXB_MAX_EXPORT]; uint32_t n, rowOffset; };
// One field of a body's packed record, read back from the per-actor float4 arrays (pos.w / vel.w are not part of the record).
__device__ __forceinline__ float packed_state_field(const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ linVel, const float4* __restrict__ angVel, uint32_t a, uint32_t f) {
  const float* src = f < 3 ? reinterpret_cast<const float*>(pos + a) + f : (f < 7 ? reinterpret_cast<const float*>(quat + a) + (f - 3)
                   : (f < 10 ? reinterpret_cast<const float*>(linVel + a) + (f - 7) : reinterpret_cast<const float*>(angVel + a) + (f - 10)));
  return *reinterpret_cast<const volatile float*>(src);   // volatile: the values were stored by other threads of this CTA (or an earlier kernel) a moment ago
}
// Coalesced export of the dynamic bodies [d0, d0 + nd): consecutive threads store consecutive floats of the packed block (full 128-byte
// lines per warp instruction) into every target.
__device__ __forceinline__ void export_packed_range(const ExportTable* __restrict__ tab, uint32_t nT, const uint32_t* __restrict__ dynActor, uint32_t d0, uint32_t nd, const float4* pos, const float4* quat,
                                                    const float4* linVel, const float4* angVel, uint32_t tid, uint32_t nThreads) {
  const uint32_t rowOffset = tab->rowOffset;
  const size_t base = (size_t)(rowOffset + d0) * 13u; const uint32_t nf = nd * 13u;
  if (((base | nf) & 3u) == 0) {   // 16-byte aligned block (e.g. 64-body environments): one float4 store per lane, 512 contiguous bytes per warp instruction
    for (uint32_t i4 = tid; i4 < nf / 4u; i4 += nThreads) {
      float v[4];
#pragma unroll
      for (uint32_t c = 0; c < 4; ++c) { const uint32_t i = i4 * 4u + c, j = i / 13u; v[c] = packed_state_field(pos, quat, linVel, angVel, dynActor[d0 + j], i - j * 13u); }
      const float4 w = make_float4(v[0], v[1], v[2], v[3]);
      for (uint32_t t = 0; t < nT; ++t) reinterpret_cast<float4*>(tab->dst[t] + base)[i4] = w;
    }
    return;
  }
  for (uint32_t i = tid; i < nf; i += nThreads) {
    const uint32_t j = i / 13u, f = i - j * 13u;
    const float v = packed_state_field(pos, quat, linVel, angVel, dynActor[d0 + j], f);
    for (uint32_t t = 0; t < nT; ++t) tab->dst[t][base + i] = v;
  }
}

// a7 touch events (prepareLostFoundPairs_Stage1 / 2, gpunarrowphase/src/CUDA/cudaGJKEPA.cu:1468,1532; consumed on the host by the island manager
// and the contact-report code): per persistent pair slot the narrowphase keeps whether the pair produced contacts last frame and appends the
// pair key to the touch-found / touch-lost list when that changes.  One patch per pair here, so "patch count changed" is the same event.
struct TouchLists { uint32_t* state; uint64_t *found, *lost; };
__device__ __forceinline__ void touch_event(const TouchLists& T, uint32_t* __restrict__ counters, uint32_t slot, uint64_t key, bool touching) {
  const uint32_t prev = T.state[slot];
  if ((prev != 0u) == touching) return;
  T.state[slot] = touching ? 1u : 0u;
  if (touching) T.found[atomicAdd(&counters[C_NTOUCH_FOUND], 1u)] = key; else T.lost[atomicAdd(&counters[C_NTOUCH_LOST], 1u)] = key;
}

// a11: PxsCombineMaterials (lowlevel/software/include/PxsMaterialCombiner.h:69-175; GPU: gpunarrowphase/src/CUDA/materialCombiner.cuh:35), rigid
// non-compliant branch.  matTab: one float4 per material (staticFriction, dynamicFriction, restitution, bits = frictionCombineMode |
// restitutionCombineMode << 4 | flags << 8); returns true when eDISABLE_FRICTION on either side removes the friction rows.
struct MaterialArgs { const uint32_t* actorMat; const float4* matTab; const float2* shapeOff; };   // shapeOff: per-actor (contactOffset, restOffset), null = the scene's uniform values
__device__ __forceinline__ float combine_scalars(float a, float b, uint32_t mode) { return mode == 0u ? 0.5f * (a + b) : (mode == 1u ? fminf(a, b) : (mode == 2u ? a * b : fmaxf(a, b))); }
__device__ __forceinline__ bool pair_material(const MaterialArgs& M, uint32_t actor0, uint32_t actor1, SolverParams& P) {
  const float4 m0 = M.matTab[M.actorMat[actor0]], m1 = M.matTab[M.actorMat[actor1]];
  const uint32_t b0 = __float_as_uint(m0.w), b1 = __float_as_uint(m1.w);
  P.restitution = combine_scalars(m0.z, m1.z, max((b0 >> 4) & 15u, (b1 >> 4) & 15u));
  if (((b0 | b1) >> 8) & 1u) { P.staticFriction = 0.f; P.dynamicFriction = 0.f; return true; }
  const uint32_t fm = max(b0 & 15u, b1 & 15u);
  const float dyn = combine_scalars(m0.y, m1.y, fm), sta = combine_scalars(m0.x, m1.x, fm);
  const float fDyn = fmaxf(dyn, 0.f);
  P.dynamicFriction = fDyn; P.staticFriction = (sta - fDyn) >= 0.f ? sta : fDyn;
  return false;
}
