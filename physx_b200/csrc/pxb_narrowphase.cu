// pxb_narrowphase.cu -- a8-a11: the contact-generation kernels (k_narrowphase: sphere family, plane-box, box-box PCM; k_narrowphase_gjk: the GJK /
// EPA family over a device-side worklist) in their own translation unit.
#include "pxb_np_launch.h"
#include <algorithm>

// a8-a11: one thread per pair.  Driver logic of PxcNpBatch.cpp:364-498: body0 is the dynamic actor (for
// two dynamics the later-created one, ScNPhaseCore.cpp:182-252), shapes are ordered by geometry type for
// the contact function and the normal is flipped back afterwards (flipContacts).
#ifndef PXB_NP_CTAS
#define PXB_NP_CTAS 5
#endif
template <bool BOXW, bool FILT>   // BOXW: box-box manifolds are only refreshed here, the invalidated ones go to k_boxbox_generate's worklist (device-wide path); FILT: default filter shader
__global__ void __launch_bounds__(128, PXB_NP_CTAS) k_narrowphase(const uint64_t* __restrict__ pairKeys, const uint32_t* __restrict__ pairSlots, const uint32_t* __restrict__ nPairsP, uint32_t bitsA,
                              const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ dims, const uint32_t* __restrict__ geomFlags,
                              float contactDistScene, float toleranceLength, float4* __restrict__ manifolds, float4* __restrict__ cHdr, float4* __restrict__ cPts,
                              uint2* __restrict__ pairBodies, uint32_t* __restrict__ conFlag, float* __restrict__ cForce, uint32_t* __restrict__ counters, uint32_t* __restrict__ gjkList,
                              const uint32_t* __restrict__ pairOrder, const TouchLists touch, uint32_t* __restrict__ boxList, const FilterArgs F) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *nPairsP) return;
  const uint32_t i = pairOrder ? pairOrder[t] : t;   // mixed-type scenes: pairs binned by type pair (k_np_class_*)
  const uint64_t key = pairKeys[i];
#ifndef PXB_NO_PREFETCH
  { const float4* r = manifolds + (size_t)pairSlots[i] * PXB_MANIFOLD_F4; prefetch_l2(r); prefetch_l2(r + 8); }   // the record is needed two dependent loads later (types -> poses -> manifold)
#endif
  if (key == ~0ull) { cHdr[i] = make_float4(0, 0, 0, __int_as_float(0)); conFlag[i] = 0u; pairBodies[i] = make_uint2(0, 0); return; }   // dropped segment (capacity error already flagged)
  const uint32_t lo = (uint32_t)(key >> bitsA), hi = (uint32_t)(key & ((1ull << bitsA) - 1ull));
  uint32_t a0 = hi, a1 = lo;
  const uint32_t gfHi = geomFlags[hi], gfLo = geomFlags[lo];
  if (!gf_dynamic(gfHi)) { a0 = lo; a1 = hi; }   // a static or kinematic actor is body B (shouldSwapBodies, ScNPhaseCore.cpp:182-252)
  const float contactDist = (FILT && F.shapeOff) ? F.shapeOff[a0].x + F.shapeOff[a1].x : contactDistScene;   // sum of the two shapes' contact offsets
  if (FILT && F.data && filter_suppressed(F, a0, a1)) {   // eSUPPRESS: the pair stays a broadphase pair, there is no contact manager behind it
    cHdr[i] = make_float4(0, 0, 0, __int_as_float(0)); conFlag[i] = 0u; pairBodies[i] = make_uint2(a0, a1);
    touch_event(touch, counters, pairSlots[i], key, false);
    return;
  }
  const uint32_t g0 = (a0 == hi) ? gfHi : gfLo, g1 = (a0 == hi) ? gfLo : gfHi;
  const uint32_t t0 = g0 & 0xff, t1 = g1 & 0xff;
  const bool flip = t1 < t0;
  const uint32_t s0 = flip ? a1 : a0, s1 = flip ? a0 : a1;
  const uint32_t ty0 = flip ? t1 : t0, ty1 = flip ? t0 : t1;
  const float4 p0 = pos[s0], p1 = pos[s1];
  xf tm0, tm1; tm0.p = V3(p0.x, p0.y, p0.z); tm0.q = Q4(quat[s0]); tm1.p = V3(p1.x, p1.y, p1.z); tm1.q = Q4(quat[s1]);
  const float4 d0 = dims[s0], d1 = dims[s1];
  float4* rec = manifolds + (size_t)pairSlots[i] * PXB_MANIFOLD_F4;
  // only the PCM pair types keep a persistent manifold (plane-box, box-box, plane-capsule); the closed-form sphere family does not
  const bool usesManifold = (ty0 == PXB_GEOM_PLANE && (ty1 == PXB_GEOM_BOX || ty1 == PXB_GEOM_CAPSULE)) || (ty0 == PXB_GEOM_BOX && ty1 == PXB_GEOM_BOX);   // (GJK-family pairs load theirs in k_narrowphase_gjk)
  Manifold man;
  if (usesManifold) manifold_load(man, rec); else { man.n = 0; man.dirty = 0; }
  Contacts out; out.count = 0; out.normal = V3(0, 0, 0);
  for (int k = 0; k < 4; ++k) { out.point[k] = V3(0, 0, 0); out.sep[k] = 0.f; }
  if (ty0 == PXB_GEOM_PLANE && ty1 == PXB_GEOM_BOX) pcm_plane_box(tm0, tm1, V3(d1.x, d1.y, d1.z), contactDist, toleranceLength, man, out);
  else if (BOXW && ty0 == PXB_GEOM_BOX && ty1 == PXB_GEOM_BOX) {   // device-wide path: only the refresh here, regeneration over a compacted worklist (k_boxbox_generate)
    if (pcm_box_box_refresh(tm0, tm1, V3(d0.x, d0.y, d0.z), V3(d1.x, d1.y, d1.z), contactDist, toleranceLength, man, out)) {
      manifold_store(man, rec); boxList[atomicAdd(&counters[C_NBOXGEN], 1u)] = i; return;
    }
  }
  else if (ty0 == PXB_GEOM_BOX && ty1 == PXB_GEOM_BOX) {
    if (pcm_box_box(tm0, tm1, V3(d0.x, d0.y, d0.z), V3(d1.x, d1.y, d1.z), contactDist, toleranceLength, man, out)) {
      // edge-edge / corner configuration (rare): the SAT passed but clipping found no point -> GJK / EPA single-point fallback, called out of
      // line so that this kernel keeps its register budget (measured: cheaper than handing the pair to a second, usually empty, launch).
      manifold_load_warm(man, rec);
      gjk_boxbox_gjk_fallback_outofline(&tm0, &tm1, V3(d0.x, d0.y, d0.z), V3(d1.x, d1.y, d1.z), contactDist, toleranceLength, &man, &out);
      manifold_store_warm(man, rec);
    }
  }
  else if (ty0 == PXB_GEOM_SPHERE && ty1 == PXB_GEOM_SPHERE) np_sphere_sphere(tm0.p, tm1.p, d0.x, d1.x, contactDist, out);
  else if (ty0 == PXB_GEOM_SPHERE && ty1 == PXB_GEOM_PLANE) np_sphere_plane(tm0.p, d0.x, tm1, contactDist, out);
  else if (ty0 == PXB_GEOM_SPHERE && ty1 == PXB_GEOM_CAPSULE) np_sphere_capsule(tm0.p, d0.x, tm1, d1.x, d1.y, contactDist, out);
  else if (ty0 == PXB_GEOM_SPHERE && ty1 == PXB_GEOM_BOX) np_sphere_box(tm0.p, d0.x, tm1, V3(d1.x, d1.y, d1.z), contactDist, out);
  else if (ty0 == PXB_GEOM_PLANE && ty1 == PXB_GEOM_CAPSULE) pcm_plane_capsule(tm0, tm1, d1.x, d1.y, contactDist, man, out);
  else if (ty0 == PXB_GEOM_CAPSULE && ty1 == PXB_GEOM_CAPSULE) np_capsule_capsule(tm0, tm1, d0.x, d0.y, d1.x, d1.y, contactDist, out);
  else if (ty0 == PXB_GEOM_CAPSULE && ty1 == PXB_GEOM_BOX) { gjkList[atomicAdd(&counters[C_NGJK], 1u)] = i; return; }   // GJK family (a10): k_narrowphase_gjk fills this pair's outputs
  else if (ty1 == PXB_GEOM_CONVEXMESH) { gjkList[atomicAdd(&counters[C_NGJK], 1u)] = i; return; }   // hull pairs also go through k_narrowphase_gjk
  else atomicOr(&counters[C_ERROR], (uint32_t)E_UNSUPPORTED_PAIR);   // unknown geometry type: reported by fetchResults, never silently skipped
  if (man.dirty) manifold_store(man, rec); else if (usesManifold && man.n > 0) manifold_store_pens(man, rec);   // steady state: only the penetrations change
  if (flip && out.count) out.normal = -out.normal;
  cHdr[i] = make_float4(out.normal.x, out.normal.y, out.normal.z, __int_as_float(out.count));
#pragma unroll
  for (int k = 0; k < 4; ++k) { cPts[(size_t)i * 4 + k] = make_float4(out.point[k].x, out.point[k].y, out.point[k].z, out.sep[k]); }   // (cForce: every pair with contacts is a constraint and gets its forces from write-back)
  pairBodies[i] = make_uint2(a0, a1);
  conFlag[i] = out.count > 0 ? 1u : 0u;
  touch_event(touch, counters, pairSlots[i], key, out.count > 0);
}

// a10: the GJK family (capsule-box), one thread per listed pair.  Kept out of k_narrowphase so that the box / sphere hot path keeps its register budget;
// the list order is arbitrary (atomic append) but every pair writes only its own outputs, so the result is deterministic.
#ifndef PXB_GJK_CTAS
#define PXB_GJK_CTAS 3   // 168 registers.  Measured on config 3 with hulls: 4 CTAs/SM (128 registers, +240 B of spills) is slower, 10.15 vs 10.0 ms/step -- the kernel is divergence bound (5 of 32 threads active), not residency bound
#endif
__global__ void __launch_bounds__(128, PXB_GJK_CTAS) k_narrowphase_gjk(const uint64_t* __restrict__ pairKeys, const uint32_t* __restrict__ pairSlots, uint32_t bitsA, const float4* __restrict__ pos, const float4* __restrict__ quat,
                              const float4* __restrict__ dims, const uint32_t* __restrict__ geomFlags, float contactDistScene, float toleranceLength, float4* __restrict__ manifolds, float4* __restrict__ cHdr,
                              float4* __restrict__ cPts, uint2* __restrict__ pairBodies, uint32_t* __restrict__ conFlag, uint32_t* __restrict__ counters, const uint32_t* __restrict__ gjkList, HullArrays hulls, const TouchLists touch, const float2* __restrict__ shapeOff) {
  const uint32_t n = counters[C_NGJK];
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const uint32_t i = gjkList[w];
    const uint64_t key = pairKeys[i];
    const uint32_t lo = (uint32_t)(key >> bitsA), hi = (uint32_t)(key & ((1ull << bitsA) - 1ull));
    uint32_t a0 = hi, a1 = lo;
    const uint32_t gfHi = geomFlags[hi], gfLo = geomFlags[lo];
    if (!gf_dynamic(gfHi)) { a0 = lo; a1 = hi; }
    const uint32_t t0 = ((a0 == hi) ? gfHi : gfLo) & 0xff, t1 = ((a0 == hi) ? gfLo : gfHi) & 0xff;
    const float contactDist = shapeOff ? shapeOff[a0].x + shapeOff[a1].x : contactDistScene;
    const bool flip = t1 < t0;
    const uint32_t s0 = flip ? a1 : a0, s1 = flip ? a0 : a1;   // s0 = capsule, s1 = box
    const float4 p0 = pos[s0], p1 = pos[s1];
    xf tm0, tm1; tm0.p = V3(p0.x, p0.y, p0.z); tm0.q = Q4(quat[s0]); tm1.p = V3(p1.x, p1.y, p1.z); tm1.q = Q4(quat[s1]);
    const float4 d0 = dims[s0], d1 = dims[s1];
    float4* rec = manifolds + (size_t)pairSlots[i] * PXB_MANIFOLD_F4;
    Manifold man; manifold_load(man, rec); manifold_load_warm(man, rec);
    Contacts out; out.count = 0; out.normal = V3(0, 0, 0);
    for (int k = 0; k < 4; ++k) { out.point[k] = V3(0, 0, 0); out.sep[k] = 0.f; }
    const uint32_t ty1 = flip ? t0 : t1;
    const uint32_t ty0 = flip ? t1 : t0;
    if (ty1 == PXB_GEOM_CONVEXMESH) {   // s1 = hull, s0 = plane, sphere, capsule, box or hull
      const DevHull h = load_hull(hulls, __float_as_uint(d1.x));
      if (ty0 == PXB_GEOM_PLANE) gjk_pcm_plane_convex(&tm0, &tm1, h, contactDist, toleranceLength, &man, &out);
      else if (ty0 == PXB_GEOM_SPHERE) gjk_pcm_sphere_convex(&tm0, &tm1, d0.x, &h, contactDist, toleranceLength, &man, &out);
      else if (ty0 == PXB_GEOM_CAPSULE) gjk_pcm_capsule_convex(&tm0, &tm1, d0.x, d0.y, &h, contactDist, toleranceLength, &man, &out);
      else {   // box-hull / hull-hull
        int sat;
        if (ty0 == PXB_GEOM_BOX) sat = gjk_pcm_box_convex(&tm0, &tm1, V3(d0.x, d0.y, d0.z), &h, contactDist, toleranceLength, &man, &out);
        else { const DevHull h0 = load_hull(hulls, __float_as_uint(d0.x)); sat = gjk_pcm_convex_convex(&tm0, &tm1, &h0, &h, contactDist, toleranceLength, &man, &out); }
        if (sat) atomicOr(&counters[C_ERROR], (uint32_t)E_UNSUPPORTED_PAIR);
      }
    }
    else gjk_pcm_capsule_box(&tm0, &tm1, d0.x, d0.y, V3(d1.x, d1.y, d1.z), contactDist, toleranceLength, &man, &out);
    if (man.dirty) { manifold_store(man, rec); manifold_store_warm(man, rec); } else if (man.n > 0) manifold_store_pens(man, rec);
    if (flip && out.count) out.normal = -out.normal;
    cHdr[i] = make_float4(out.normal.x, out.normal.y, out.normal.z, __int_as_float(out.count));
#pragma unroll
    for (int k = 0; k < 4; ++k) cPts[(size_t)i * 4 + k] = make_float4(out.point[k].x, out.point[k].y, out.point[k].z, out.sep[k]);
    pairBodies[i] = make_uint2(a0, a1);
    conFlag[i] = out.count > 0 ? 1u : 0u;
    touch_event(touch, counters, pairSlots[i], key, out.count > 0);
  }
}

// ---- a10 in four phases -------------------------------------------------------------------------------------------------------------
// A persistent-manifold pair does one of four very different amounts of work per step: its manifold is still valid (refresh only), it is
// invalidated and GJK finds it separated or only adds the GJK point, it is deep enough for EPA, or it regenerates the full manifold (polygon clipping / SAT).  One thread
// per pair through all of that left 5 of 32 lanes active (ncu, BASELINE config 3).  Here every phase is its own kernel over a compacted
// worklist, so a warp holds 32 pairs that all need that phase; what a pair carries between phases is its manifold record (stored by the
// phase that changed it) plus a few words parked in the pair's own, not yet written, output slots (conFlag: refresh flags, cPts: GjkCarry).
// The arithmetic per pair is the single-kernel variant's (same device functions, same order), so the contacts are bit-identical.
struct GjkPair { uint32_t i, a0, a1, ty0, ty1, slot; bool flip; uint64_t key; xf tm0, tm1; float4 d0, d1; float4* rec; float cd; };   // cd: the pair's contact distance
__device__ __forceinline__ void gjk_pair_setup(const NpArgs& A, uint32_t i, GjkPair& P) {
  P.i = i; P.key = A.pairKeys[i];
  const uint32_t lo = (uint32_t)(P.key >> A.bitsA), hi = (uint32_t)(P.key & ((1ull << A.bitsA) - 1ull));
  P.a0 = hi; P.a1 = lo;
  const uint32_t gfHi = A.geomFlags[hi], gfLo = A.geomFlags[lo];
  if (!gf_dynamic(gfHi)) { P.a0 = lo; P.a1 = hi; }
  const uint32_t t0 = ((P.a0 == hi) ? gfHi : gfLo) & 0xff, t1 = ((P.a0 == hi) ? gfLo : gfHi) & 0xff;
  P.flip = t1 < t0;
  const uint32_t s0 = P.flip ? P.a1 : P.a0, s1 = P.flip ? P.a0 : P.a1;
  const float4 p0 = A.pos[s0], p1 = A.pos[s1];
  P.tm0.p = V3(p0.x, p0.y, p0.z); P.tm0.q = Q4(A.quat[s0]); P.tm1.p = V3(p1.x, p1.y, p1.z); P.tm1.q = Q4(A.quat[s1]);
  P.d0 = A.dims[s0]; P.d1 = A.dims[s1];
  P.slot = A.pairSlots[i]; P.rec = A.manifolds + (size_t)P.slot * PXB_MANIFOLD_F4;
  P.ty1 = P.flip ? t0 : t1; P.ty0 = P.flip ? t1 : t0;
  P.cd = A.filter.shapeOff ? A.filter.shapeOff[P.a0].x + A.filter.shapeOff[P.a1].x : A.contactDist;
}
__device__ __forceinline__ void gjk_pair_finish(const NpArgs& A, const GjkPair& P, const Manifold& man, Contacts& out) {
  if (man.dirty) { manifold_store(man, P.rec); manifold_store_warm(man, P.rec); } else if (man.n > 0) manifold_store_pens(man, P.rec);
  if (P.flip && out.count) out.normal = -out.normal;
  A.cHdr[P.i] = make_float4(out.normal.x, out.normal.y, out.normal.z, __int_as_float(out.count));
#pragma unroll
  for (int k = 0; k < 4; ++k) A.cPts[(size_t)P.i * 4 + k] = make_float4(out.point[k].x, out.point[k].y, out.point[k].z, out.sep[k]);
  A.pairBodies[P.i] = make_uint2(P.a0, P.a1);
  A.conFlag[P.i] = out.count > 0 ? 1u : 0u;
  touch_event(A.touch, A.counters, P.slot, P.key, out.count > 0);
}
__device__ __forceinline__ void gjk_contacts_clear(Contacts& out) { out.count = 0; out.normal = V3(0, 0, 0); for (int k = 0; k < 4; ++k) { out.point[k] = V3(0, 0, 0); out.sep[k] = 0.f; } }
// all 32 lanes call: the lanes with `need` append `v` to the list in lane order behind one atomic
__device__ __forceinline__ void gjk_warp_append(bool need, uint32_t v, uint32_t* __restrict__ list, uint32_t* counter) {
  const uint32_t m = __ballot_sync(0xffffffffu, need);
  if (!m) return;
  const uint32_t lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (need) list[base + __popc(m & ((1u << lane) - 1u))] = v;
}
#define GJK_PHASE_LOOP(nExpr) const uint32_t n = (nExpr); const uint32_t lane = threadIdx.x & 31; \
  for (uint32_t w0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; w0 < n; w0 += gridDim.x * blockDim.x)

// phase 1: every GJK-family pair.  Plane / sphere vs hull and capsule-box (cheap, no second phase worth a launch) run to completion.
__global__ void __launch_bounds__(128, 4) k_gjk_refresh(const NpArgs A) {
  GJK_PHASE_LOOP(A.counters[C_NGJK]) {
    const uint32_t w = w0 + lane; bool need = false; uint32_t i = 0;
    if (w < n) {
      i = A.gjkList[w];
      GjkPair P; gjk_pair_setup(A, i, P);
      Manifold man; manifold_load(man, P.rec); manifold_load_warm(man, P.rec);
      Contacts out; gjk_contacts_clear(out);
      int flags = 0;
      if (P.ty1 == PXB_GEOM_CONVEXMESH) {
        const DevHull h = load_hull(A.hulls, __float_as_uint(P.d1.x));
        if (P.ty0 == PXB_GEOM_PLANE) gjk_pcm_plane_convex(&P.tm0, &P.tm1, h, P.cd, A.toleranceLength, &man, &out);
        else if (P.ty0 == PXB_GEOM_SPHERE) gjk_pcm_sphere_convex(&P.tm0, &P.tm1, P.d0.x, &h, P.cd, A.toleranceLength, &man, &out);
        else if (P.ty0 == PXB_GEOM_CAPSULE) need = gjk_capsule_convex_refresh(&P.tm0, &P.tm1, P.d0.x, &h, P.cd, A.toleranceLength, &man, &out, &flags);
        else if (P.ty0 == PXB_GEOM_BOX) { const v3 e = V3(P.d0.x, P.d0.y, P.d0.z); need = gjk_poly_convex_refresh(&P.tm0, &P.tm1, box_margin(e, A.toleranceLength), alen(e), &h, P.cd, A.toleranceLength, &man, &out, &flags); }
        else {
          const DevHull h0 = load_hull(A.hulls, __float_as_uint(P.d0.x));
          need = gjk_poly_convex_refresh(&P.tm0, &P.tm1, gjk_hull_pcm_margin(&h0, A.toleranceLength), alen(h0.internalExtents), &h, P.cd, A.toleranceLength, &man, &out, &flags);
        }
      }
      else gjk_pcm_capsule_box(&P.tm0, &P.tm1, P.d0.x, P.d0.y, V3(P.d1.x, P.d1.y, P.d1.z), P.cd, A.toleranceLength, &man, &out);
      if (need) { manifold_store(man, P.rec); A.conFlag[i] = (uint32_t)flags; }   // the refreshed manifold and the new relative frame; the warm-start simplex is untouched
      else gjk_pair_finish(A, P, man, out);
    }
    gjk_warp_append(need, i, A.gjkQuery, &A.counters[C_NGJK_QUERY]);
  }
}
// The shapes of an invalidated pair as the GJK / EPA query sees them: convexA = capsule (already in B's frame), box or hull (relative), convexB = hull.
struct GjkShapes { GjkConvex a, b; DevHull hA, hB; mxf aToB; v3 dir; float marginPcmA; };
__device__ __forceinline__ void gjk_pair_shapes(const NpArgs& A, const GjkPair& P, GjkShapes& S) {
  S.hB = load_hull(A.hulls, __float_as_uint(P.d1.x)); S.marginPcmA = 0.f;
  if (P.ty0 == PXB_GEOM_CAPSULE) gjk_capsule_convex_shapes(&P.tm0, &P.tm1, P.d0.x, P.d0.y, &S.hB, &S.a, &S.b, &S.aToB, &S.dir);
  else {
    const xf curRTrans = axfinvmul(&P.tm1, &P.tm0);
    S.aToB = amxffromxf(&curRTrans); S.dir = S.aToB.p; S.b = gjk_cvx_hull(&S.hB);
    if (P.ty0 == PXB_GEOM_BOX) { const v3 e = V3(P.d0.x, P.d0.y, P.d0.z); S.a = gjk_cvx_box(V3(0, 0, 0), e); S.marginPcmA = box_margin(e, A.toleranceLength); }
    else { S.hA = load_hull(A.hulls, __float_as_uint(P.d0.x)); S.a = gjk_cvx_hull(&S.hA); S.marginPcmA = gjk_hull_pcm_margin(&S.hA, A.toleranceLength); }
    gjk_cvx_make_relative(&S.a, &S.aToB);
  }
}
__device__ __forceinline__ bool gjk_pair_post(const NpArgs& A, const GjkPair& P, const GjkShapes& S, int flags, int status, int epaStatus, const GjkOutput& output, Manifold& man, Contacts& out) {
  GjkCarry carry; int need;
  if (P.ty0 == PXB_GEOM_CAPSULE) need = gjk_capsule_convex_post(&P.tm0, &P.tm1, P.d0.x, &S.hB, P.cd, A.toleranceLength, flags, status, epaStatus, &output, &S.aToB, &man, &out, &carry);
  else need = gjk_poly_convex_post(&P.tm1, S.a.center, S.b.center, S.marginPcmA, &S.hB, P.cd, A.toleranceLength, flags, status, epaStatus, &output, &S.aToB, &man, &out, &carry);
  if (need) {
    manifold_store(man, P.rec); manifold_store_warm(man, P.rec);
    float4* c = A.cPts + (size_t)P.i * 4;
    c[0] = make_float4(carry.normal.x, carry.normal.y, carry.normal.z, __int_as_float(carry.doOverlapTest));
    c[1] = make_float4(carry.closestA.x, carry.closestA.y, carry.closestA.z, 0.f); c[2] = make_float4(carry.closestB.x, carry.closestB.y, carry.closestB.z, 0.f);
  } else gjk_pair_finish(A, P, man, out);
  return need != 0;
}
// phase 2: GJK for the pairs whose manifold was invalidated -- ONE call site for every shape pair, so a warp that mixes capsule-hull, box-hull and hull-hull
// pairs still iterates together (the shapes differ only inside the support mapping).  Deep pairs (EPA_CONTACT) go on to k_gjk_epa.
__global__ void __launch_bounds__(128, 4) k_gjk_query(const NpArgs A) {
  GJK_PHASE_LOOP(A.counters[C_NGJK_QUERY]) {
    const uint32_t w = w0 + lane; bool needFull = false, needEpa = false; uint32_t i = 0;
    if (w < n) {
      i = A.gjkQuery[w];
      GjkPair P; gjk_pair_setup(A, i, P);
      Manifold man; manifold_load(man, P.rec); manifold_load_warm(man, P.rec); man.dirty = 1;
      Contacts out; gjk_contacts_clear(out);
      const int flags = (int)A.conFlag[i];
      GjkShapes S; gjk_pair_shapes(A, P, S);
      GjkOutput output; output.normal = output.closestA = output.closestB = output.searchDir = V3(0, 0, 0); output.penDep = 0.f;
      const int status = gjk_penetration(&S.a, &S.b, S.dir, P.cd, 1, man.aInd, man.bInd, &man.nWarm, &output);
      if (status == GJK_NON_INTERSECT) gjk_pair_finish(A, P, man, out);
      else if (status == EPA_CONTACT) {   // the GJK answer (EPA starts from the warm-start simplex and may leave parts of it in place) travels in the pair's output slots
        needEpa = true; manifold_store_warm(man, P.rec);
        float4* c = A.cPts + (size_t)i * 4;
        c[0] = make_float4(output.normal.x, output.normal.y, output.normal.z, output.penDep); c[1] = make_float4(output.closestA.x, output.closestA.y, output.closestA.z, 0.f);
        c[2] = make_float4(output.closestB.x, output.closestB.y, output.closestB.z, 0.f); c[3] = make_float4(output.searchDir.x, output.searchDir.y, output.searchDir.z, 0.f);
      }
      else needFull = gjk_pair_post(A, P, S, flags, status, 0, output, man, out);
    }
    gjk_warp_append(needEpa, i, A.gjkEpa, &A.counters[C_NGJK_EPA]);
    gjk_warp_append(needFull, i, A.gjkFull, &A.counters[C_NGJK_FULL]);
  }
}
// phase 2b: EPA for the deep pairs
__global__ void __launch_bounds__(128, PXB_GJK_CTAS) k_gjk_epa(const NpArgs A) {
  GJK_PHASE_LOOP(A.counters[C_NGJK_EPA]) {
    const uint32_t w = w0 + lane; bool needFull = false; uint32_t i = 0;
    if (w < n) {
      i = A.gjkEpa[w];
      GjkPair P; gjk_pair_setup(A, i, P);
      Manifold man; manifold_load(man, P.rec); manifold_load_warm(man, P.rec); man.dirty = 1;
      Contacts out; gjk_contacts_clear(out);
      const int flags = (int)A.conFlag[i];
      GjkShapes S; gjk_pair_shapes(A, P, S);
      GjkOutput output;
      { const float4* c = A.cPts + (size_t)i * 4; const float4 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
        output.normal = V3(c0.x, c0.y, c0.z); output.penDep = c0.w; output.closestA = V3(c1.x, c1.y, c1.z); output.closestB = V3(c2.x, c2.y, c2.z); output.searchDir = V3(c3.x, c3.y, c3.z); }
      const int epaStatus = gjk_epa_penetration(&S.a, &S.b, man.aInd, man.bInd, man.nWarm, 1, A.toleranceLength, &output);
      needFull = gjk_pair_post(A, P, S, flags, EPA_CONTACT, epaStatus, output, man, out);
    }
    gjk_warp_append(needFull, i, A.gjkFull, &A.counters[C_NGJK_FULL]);
  }
}
// phase 3: full manifold generation
__global__ void __launch_bounds__(128, PXB_GJK_CTAS) k_gjk_manifold(const NpArgs A) {
  GJK_PHASE_LOOP(A.counters[C_NGJK_FULL]) {
    const uint32_t w = w0 + lane;
    if (w < n) {
      const uint32_t i = A.gjkFull[w];
      GjkPair P; gjk_pair_setup(A, i, P);
      Manifold man; manifold_load(man, P.rec); manifold_load_warm(man, P.rec); man.dirty = 1;
      Contacts out; gjk_contacts_clear(out);
      GjkCarry carry;
      { const float4* c = A.cPts + (size_t)i * 4; const float4 c0 = c[0], c1 = c[1], c2 = c[2];
        carry.normal = V3(c0.x, c0.y, c0.z); carry.doOverlapTest = __float_as_int(c0.w); carry.closestA = V3(c1.x, c1.y, c1.z); carry.closestB = V3(c2.x, c2.y, c2.z); }
      const DevHull h = load_hull(A.hulls, __float_as_uint(P.d1.x));
      if (P.ty0 == PXB_GEOM_CAPSULE) gjk_capsule_convex_manifold(&P.tm0, &P.tm1, P.d0.x, P.d0.y, &h, P.cd, A.toleranceLength, &carry, &man, &out);
      else {
        int sat;
        if (P.ty0 == PXB_GEOM_BOX) {
          const v3 e = V3(P.d0.x, P.d0.y, P.d0.z); BoxAsHull bh; const DevHull* polyA = gjk_box_as_hull(&bh, e); const GjkConvex box = gjk_cvx_box(V3(0, 0, 0), e);
          sat = gjk_poly_convex_manifold(&P.tm0, &P.tm1, &box, polyA, &h, P.cd, A.toleranceLength, &carry, &man, &out);
        } else {
          const DevHull h0 = load_hull(A.hulls, __float_as_uint(P.d0.x)); const GjkConvex c0 = gjk_cvx_hull(&h0);
          sat = gjk_poly_convex_manifold(&P.tm0, &P.tm1, &c0, &h0, &h, P.cd, A.toleranceLength, &carry, &man, &out);
        }
        if (sat) atomicOr(&A.counters[C_ERROR], (uint32_t)E_UNSUPPORTED_PAIR);
      }
      gjk_pair_finish(A, P, man, out);
    }
  }
}

// box-box manifold regeneration (SAT + clipping + reduction, the GJK / EPA single-point fallback when clipping finds nothing) for the pairs whose
// manifold k_narrowphase found invalid: a dense pile regenerates a minority of its manifolds per step, and next to the cached majority those
// lanes ran alone (11 of 32 lanes active in k_narrowphase on BASELINE config 4).
__global__ void __launch_bounds__(128, PXB_NP_CTAS) k_boxbox_generate(const NpArgs A) {
  const uint32_t n = A.counters[C_NBOXGEN];
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
    const uint32_t i = A.boxList[w];
    GjkPair P; gjk_pair_setup(A, i, P);
    Manifold man; manifold_load(man, P.rec); manifold_load_warm(man, P.rec); man.dirty = 1;
    Contacts out; gjk_contacts_clear(out);
    const v3 e0 = V3(P.d0.x, P.d0.y, P.d0.z), e1 = V3(P.d1.x, P.d1.y, P.d1.z);
    if (pcm_box_box_generate(P.tm0, P.tm1, e0, e1, P.cd, A.toleranceLength, man, out))
      gjk_boxbox_gjk_fallback_outofline(&P.tm0, &P.tm1, e0, e1, P.cd, A.toleranceLength, &man, &out);
    gjk_pair_finish(A, P, man, out);
  }
}

void pxb_launch_narrowphase(cudaStream_t st, uint32_t capPairs, const NpArgs& A) {
#define NP_LAUNCH(B, F) k_narrowphase<B, F><<<(capPairs + 127) / 128, 128, 0, st>>>(A.pairKeys, A.pairSlots, A.nPairsP, A.bitsA, A.pos, A.quat, A.dims, A.geomFlags, A.contactDist, A.toleranceLength, A.manifolds, \
                                                                                A.cHdr, A.cPts, A.pairBodies, A.conFlag, A.cForce, A.counters, A.gjkList, A.pairOrder, A.touch, A.boxList, A.filter)
  if (A.filter.data || A.filter.shapeOff) { if (A.boxList) NP_LAUNCH(true, true); else NP_LAUNCH(false, true); }
  else { if (A.boxList) NP_LAUNCH(true, false); else NP_LAUNCH(false, false); }
#undef NP_LAUNCH
  if (A.boxList) k_boxbox_generate<<<std::max(148u * 4u, std::min((capPairs + 127) / 128, 148u * 64u)), 128, 0, st>>>(A);
}
void pxb_launch_narrowphase_gjk(cudaStream_t st, uint32_t ctas, const NpArgs& A) {
  k_narrowphase_gjk<<<ctas, 128, 0, st>>>(A.pairKeys, A.pairSlots, A.bitsA, A.pos, A.quat, A.dims, A.geomFlags, A.contactDist, A.toleranceLength, A.manifolds, A.cHdr, A.cPts, A.pairBodies, A.conFlag, A.counters,
                                          A.gjkList, A.hulls, A.touch, A.filter.shapeOff);
}
void pxb_launch_narrowphase_gjk_phases(cudaStream_t st, uint32_t ctas, const NpArgs& A) {
  k_gjk_refresh<<<ctas, 128, 0, st>>>(A);
  k_gjk_query<<<ctas, 128, 0, st>>>(A);
  k_gjk_epa<<<ctas, 128, 0, st>>>(A);
  k_gjk_manifold<<<ctas, 128, 0, st>>>(A);
}
