// pxb_gjk.cuh -- GJK penetration query and the PCM pair functions built on it, one thread per pair.
// Follows the reference's CPU PCM path so that contacts match it bit for bit (same operation order as its SSE2 vector layer, FMA contraction off):
//   simplex solver   physx/source/geomutils/src/gjk/GuGJKSimplex.h:50-448, GuGJKSimplex.cpp:38-213; barycentric coords common/GuBarycentricCoordinates.cpp:36-80
//   gjkPenetration   gjk/GuGJKPenetration.h:89-311; epaPenetration gjk/GuEPA.cpp:70-623, gjk/GuEPAFacet.h:61-301; supports gjk/GuVecCapsule.h:128-196, gjk/GuVecBox.h:56-68,148-205
//   capsule-box      pcm/GuPCMContactCapsuleBox.cpp:42-202, pcm/GuPCMContactGenSphereCapsule.cpp:43-420, pcm/GuPCMContactGenUtil.cpp:105-260, pcm/GuPCMShapeConvex.cpp:40-110
//   manifold helpers pcm/GuPersistentContactManifold.h:227-241, .cpp:179-187,289-360,783-807,1177-1310
// (replaces convexConvexNphase_stage1/2Kernel + gjk.cuh / epa.cuh of the reference GPU path for these pair types)
#pragma once
#include "pxb_np.cuh"

// pointer-style helpers so the simplex code reads like the reference's Vec3V code
PXB_D v3 v3add(v3 a, v3 b) { return a + b; }
PXB_D v3 v3sub(v3 a, v3 b) { return a - b; }
PXB_D v3 v3neg(v3 a) { return -a; }
PXB_D v3 v3abs(v3 a) { return vabs(a); }
PXB_D v3 v3scale(v3 a, float s) { return a * s; }
PXB_D v3 v3cross(v3 a, v3 b) { return cross(a, b); }
PXB_D float v3dot(v3 a, v3 b) { return dot(a, b); }
PXB_D v3 v3scaleadd(v3 a, float s, v3 b) { return scaleadd(a, s, b); }
PXB_D v3 v3negscalesub(v3 a, float s, v3 b) { return negscalesub(a, s, b); }
PXB_D v3 v3min(v3 a, v3 b) { return vmin(a, b); }
PXB_D v3 v3max(v3 a, v3 b) { return vmax(a, b); }
PXB_D v3 m33mul(const m33* m, v3 v) { return mmul(*m, v); }
PXB_D m33 m33transpose(const m33* m) { return mtranspose(*m); }
PXB_D v3 am33tmul(const m33* m, v3 v) { return amtmul(*m, v); }
PXB_D mxf amxfinvmul(const mxf* a, const mxf* b) { return amxfinvmul(*a, *b); }
PXB_D v3 amxftransform(const mxf* t, v3 v) { return amxftransform(*t, v); }
PXB_D v3 amxftransforminv(const mxf* t, v3 v) { return amtmul(t->r, v - t->p); }
PXB_D mxf amxffromxf(const xf* t) { return amxffromxf(*t); }
PXB_D xf axfinvmul(const xf* a, const xf* b) { return axfinvmul(*a, *b); }
PXB_D v3 axftransform(const xf* t, v3 v) { return axftransform(*t, v); }

/* ---------------- convex hulls: cooked Gu::ConvexHullData in device memory (uploaded by pxb_scene_set_convex_meshes) ---------------- */
struct HullArrays { const uint4* meta; const float4* verts; const float4* polys; const uint8_t* refs; const uint8_t* edges; };   // meta: 4 x 16 B per hull
struct DevHull {
  uint32_t nVerts, nPolys, nEdges; v3 internalExtents, centerOfMass; float internalRadius;
  const float4* verts; const float4* polys; const uint8_t* vertexRefs; const uint8_t* facesByEdges;
  uint32_t bigSubdiv; const uint8_t* big;   // Gu::BigConvexRawData of hulls with more than 32 vertices: samples[6 * subdiv^2] | u16 valencies[nVerts][2] | adjacentVerts (bigSubdiv == 0: none)
  PXB_D v3 vert(uint32_t i) const { return V3(verts[i]); }
  PXB_D v3 plane_n(uint32_t p) const { return V3(polys[2 * p]); }
  PXB_D float plane_d(uint32_t p) const { return polys[2 * p].w; }
  PXB_D uint4 poly_meta(uint32_t p) const { const float4 m = polys[2 * p + 1]; return make_uint4(__float_as_uint(m.x), __float_as_uint(m.y), __float_as_uint(m.z), 0u); }   // (vref, nbVerts, minIndex)
};
PXB_D DevHull load_hull(const HullArrays& H, uint32_t hullIdx) {
  const uint4 m0 = H.meta[4 * hullIdx], m1 = H.meta[4 * hullIdx + 1], m2 = H.meta[4 * hullIdx + 2], m3 = H.meta[4 * hullIdx + 3];
  DevHull h; h.nVerts = m1.x; h.nPolys = m1.y; h.nEdges = m1.z;
  h.internalExtents = V3(__uint_as_float(m2.x), __uint_as_float(m2.y), __uint_as_float(m2.z)); h.centerOfMass = V3(__uint_as_float(m3.x), __uint_as_float(m3.y), __uint_as_float(m3.z)); h.internalRadius = __uint_as_float(m2.w);
  h.verts = H.verts + m0.x; h.polys = H.polys + 2 * m0.y; h.vertexRefs = H.refs + m0.z; h.facesByEdges = H.edges + m0.w;
  h.bigSubdiv = m3.w; h.big = h.facesByEdges + ((2 * m1.z + 3) & ~3u);   // the hill-climbing data follows the hull's edge bytes (4-byte aligned)
  return h;
}
/* ComputeCubemapNearestOffset + CubemapLookup, GuCubeIndex.h:101-150: the cube-map texel a direction falls into */
PXB_D uint32_t gjk_cubemap_nearest_offset(v3 dir, uint32_t subdiv) {
  const float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);   // the reference compares the sign-stripped bit patterns: the same order for finite floats
  float d1 = dir.x, d2 = dir.y, d3 = dir.z; uint32_t i1 = 0;
  if ((ay > ax) & (ay > az)) { d1 = dir.y; d2 = dir.z; d3 = dir.x; i1 = 1; }
  else if (az > ax) { d1 = dir.z; d2 = dir.x; d3 = dir.y; i1 = 2; }
  const float c = __fdiv_rn(1.0f, fabsf(d1));
  float u = __fmul_rn(d2, c), v = __fmul_rn(d3, c);
  const uint32_t ci = (__float_as_uint(d1) >> 31) | (i1 + i1);
  const float coeff = __fmul_rn(0.5f, (float)(subdiv - 1));
  u = __fmul_rn(__fadd_rn(u, 1.0f), coeff); v = __fmul_rn(__fadd_rn(v, 1.0f), coeff);
  return ci * (subdiv * subdiv) + (uint32_t)__fadd_rn(u, 0.5f) * subdiv + (uint32_t)__fadd_rn(v, 0.5f);
}
/* ConvexHullV::hillClimbing GuVecConvexHull.h:321-375: start at the cube-map sample, walk to a better neighbour until none improves.
 * A neighbour is taken once only (the visited bits; <= 64 vertices on the GPU, PxConvexMeshDesc.h:139), and the start vertex is not marked -- both as in the reference. */
PXB_D uint32_t gjk_hull_hill_climb(const DevHull* h, v3 dir) {
  unsigned long long visited = 0ull;
  const uint16_t* valencies = (const uint16_t*)(h->big + ((6u * h->bigSubdiv * h->bigSubdiv + 3u) & ~3u));
  const uint8_t* adjacentVerts = (const uint8_t*)(valencies + 2 * h->nVerts);
  uint32_t index = h->big[gjk_cubemap_nearest_offset(dir, h->bigSubdiv)];
  float mx = adot(h->vert(index), dir);
  uint32_t initialIndex;
  do {
    initialIndex = index;
    const uint32_t numNeighbours = valencies[2 * index], offset = valencies[2 * index + 1];
    for (uint32_t a = 0; a < numNeighbours; ++a) {
      const uint32_t nb = adjacentVerts[offset + a];
      const float dist = adot(h->vert(nb), dir);
      if (dist > mx) {
        const unsigned long long mask = 1ull << nb;
        if ((visited & mask) == 0ull) { visited |= mask; mx = dist; index = nb; }
      }
    }
  } while (index != initialIndex);
  return index;
}
/* ConvexHullV::supportVertexIndex GuVecConvexHull.h:399-406: hill climbing when the hull has the data, else bruteForceSearch :377-397 */
PXB_D uint32_t gjk_hull_support_index(const DevHull* h, v3 dir) {
  if (h->bigSubdiv) return gjk_hull_hill_climb(h, dir);
  float mx = v3dot(h->vert(0), dir); uint32_t mi = 0;
  for (uint32_t i = 1; i < h->nVerts; ++i) { const float d = v3dot(h->vert(i), dir); if (d > mx) { mx = d; mi = i; } }
  return mi;
}

enum { GJK_CVX_CAPSULE = 0, GJK_CVX_BOX = 1, GJK_CVX_HULL = 2 };
enum { GJK_NON_INTERSECT = 0, GJK_CONTACT, GJK_UNDEFINED, GJK_DEGENERATE, EPA_CONTACT, EPA_DEGENERATE, EPA_FAIL };

typedef struct {
  int type;
  v3 center;                /* ConvexV::center (getCenter) */
  v3 p0, p1;                /* capsule segment */
  v3 ext;                   /* box half extents */
  float margin, minMargin;  /* ConvexV::margin / minMargin */
  int marginIsRadius;
  const DevHull* hull;      /* GJK_CVX_HULL: cooked hull (ConvexHullNoScaleV: identity mesh scale) */
  int relative;             /* RelativeConvex<T> (GuGJKType.h:110-150): the shape lives in A's frame, supports are returned in B's */
  mxf aToB; m33 aToBT;      /* mAToB and the precomputed transpose of its rotation */
} GjkConvex;

typedef struct { v3 normal, closestA, closestB, searchDir; float penDep; } GjkOutput;

/* CapsuleV(center, v, radius): GuVecCapsule.h:74-85 */
PXB_D GjkConvex gjk_cvx_capsule(v3 center, v3 v, float radius) {
  GjkConvex c; c.center = V3(0, 0, 0); c.p0 = c.p1 = c.ext = V3(0, 0, 0); c.relative = 0; c.hull = nullptr;
  c.type = GJK_CVX_CAPSULE; c.center = center; c.p0 = v3add(center, v); c.p1 = v3sub(center, v);
  c.margin = radius; c.minMargin = radius; c.marginIsRadius = 1;
  return c;
}
/* BoxV(origin, extent) + CalculateBoxMargin: GuVecBox.h:56-68,112-117 */
PXB_D GjkConvex gjk_cvx_box(v3 origin, v3 ext) {
  GjkConvex c; c.center = V3(0, 0, 0); c.p0 = c.p1 = c.ext = V3(0, 0, 0); c.relative = 0; c.hull = nullptr;
  const float mn = fmin_(ext.x, fmin_(ext.y, ext.z));
  c.type = GJK_CVX_BOX; c.center = origin; c.ext = ext; c.margin = mn * 0.15f; c.minMargin = mn * 0.05f; c.marginIsRadius = 0;
  return c;
}
/* ConvexHullV(hullData, centerOfMass, scale = 1, ...): GuVecConvexHull.h:202-215, CalculateConvexMargin :77-94 */
PXB_D GjkConvex gjk_cvx_hull(const DevHull* h) {
  GjkConvex c; c.center = V3(0, 0, 0); c.p0 = c.p1 = c.ext = V3(0, 0, 0); c.relative = 0; c.hull = nullptr;
  const float mn = fmin_(h->internalExtents.x, fmin_(h->internalExtents.y, h->internalExtents.z));
  c.type = GJK_CVX_HULL; c.hull = h; c.center = h->centerOfMass; c.margin = mn * 0.1f; c.minMargin = mn * 0.05f; c.marginIsRadius = 0;
  return c;
}
PXB_D void gjk_cvx_make_relative(GjkConvex* c, const mxf* aToB) { c->relative = 1; c->aToB = *aToB; c->aToBT = m33transpose(&aToB->r); }
/* LocalConvex<T>::support(dir, index) = T::supportLocal(dir, index); RelativeConvex<T>::support = T::supportRelative (GuVecBox.h:186-204) */
PXB_D v3 gjk_cvx_support_local(const GjkConvex* c, v3 dir, int* index);
PXB_D v3 gjk_cvx_support(const GjkConvex* c, v3 dir, int* index) {
  if (!c->relative) return gjk_cvx_support_local(c, dir, index);
  const v3 _dir = m33mul(&c->aToBT, dir);
  const v3 p = gjk_cvx_support_local(c, _dir, index);
  return amxftransform(&c->aToB, p);
}
PXB_D v3 gjk_cvx_support_local(const GjkConvex* c, v3 dir, int* index) {
  if (c->type == GJK_CVX_HULL) {   /* ConvexHullNoScaleV::supportLocal GuVecConvexHullNoScale.h:91-102 */
    const DevHull* h = c->hull;
    const uint32_t mi = gjk_hull_support_index(h, dir);
    *index = (int)mi;
    return h->vert(mi);
  }
  if (c->type == GJK_CVX_CAPSULE) {
    const float d0 = adot(c->p0, dir), d1 = adot(c->p1, dir);
    const int comp = d0 > d1;
    *index = comp ? 1 : 0;
    return comp ? c->p0 : c->p1;
  }
  const int bx = dir.x > 0.f, by = dir.y > 0.f, bz = dir.z > 0.f;
  *index = bx | (by << 1) | (bz << 2);
  return V3(bx ? c->ext.x : -c->ext.x, by ? c->ext.y : -c->ext.y, bz ? c->ext.z : -c->ext.z);
}
PXB_D v3 gjk_cvx_support_point_local(const GjkConvex* c, int index) {
  if (c->type == GJK_CVX_HULL) return c->hull->vert((uint32_t)index);
  if (c->type == GJK_CVX_CAPSULE) return index == 1 ? c->p0 : c->p1;   /* (&p0)[1-index] */
  return V3((index & 1) ? c->ext.x : -c->ext.x, (index & 2) ? c->ext.y : -c->ext.y, (index & 4) ? c->ext.z : -c->ext.z);
}
PXB_D v3 gjk_cvx_support_point(const GjkConvex* c, int index) {
  const v3 p = gjk_cvx_support_point_local(c, index);
  return c->relative ? amxftransform(&c->aToB, p) : p;
}

/* ---------------- simplex solver ---------------- */
PXB_D v3 gjk_v3div(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }   /* V3ScaleInv */

/* GuBarycentricCoordinates.cpp:36-48 */
PXB_D void gjk_bary2(v3 p, v3 a, v3 b, float* v) {
  const v3 v0 = v3sub(a, p), v1 = v3sub(b, p), d = v3sub(v1, v0);
  const float denominator = adot(d, d), numerator = adot(v3neg(v0), d);
  const float denom = denominator > 0.f ? 1.0f / denominator : 0.f;
  *v = numerator * denom;
}
/* GuBarycentricCoordinates.cpp:50-76 */
PXB_D void gjk_bary3(v3 p, v3 a, v3 b, v3 c, float* v, float* w) {
  const v3 ab = v3sub(b, a), ac = v3sub(c, a), n = v3cross(ab, ac);
  const v3 ca = v3sub(a, p), cb = v3sub(b, p), cc = v3sub(c, p);
  const v3 bCrossC = v3cross(cb, cc), cCrossA = v3cross(cc, ca), aCrossB = v3cross(ca, cb);
  const float va = adot(n, bCrossC), vb = adot(n, cCrossA), vc = adot(n, aCrossB);
  const float totalArea = va + (vb + vc);
  const float denom = totalArea == 0.f ? 0.f : 1.0f / totalArea;
  *v = vb * denom; *w = vc * denom;
}

/* GuGJKSimplex.h:91-118 */
PXB_D v3 gjk_closest_segment(const v3* Q, uint32_t* size) {
  const v3 a = Q[0], b = Q[1];
  const v3 ab = v3sub(b, a);
  const float denom = adot(ab, ab);
  const v3 ap = v3neg(a);
  const float nom = adot(ap, ab);
  if (FLT_EPSILON >= denom) { *size = 1; return Q[0]; }
  const float t = fmax_(fmin_(nom / denom, 1.f), 0.f);
  return v3scaleadd(ab, t, a);
}

/* GuGJKSimplex.h:120-160 */
PXB_D void gjk_closest_points(const v3* Q, const v3* A, const v3* B, v3 closest, v3* closestA, v3* closestB, uint32_t size) {
  switch (size) {
    case 1: *closestA = A[0]; *closestB = B[0]; break;
    case 2: {
      float v; gjk_bary2(closest, Q[0], Q[1], &v);
      *closestA = v3scaleadd(v3sub(A[1], A[0]), v, A[0]);
      *closestB = v3scaleadd(v3sub(B[1], B[0]), v, B[0]);
      break;
    }
    case 3: {
      float v, w; gjk_bary3(closest, Q[0], Q[1], Q[2], &v, &w);
      const v3 av0 = v3sub(A[1], A[0]), av1 = v3sub(A[2], A[0]), bv0 = v3sub(B[1], B[0]), bv1 = v3sub(B[2], B[0]);
      *closestA = v3add(A[0], v3add(v3scale(av0, v), v3scale(av1, w)));
      *closestB = v3add(B[0], v3add(v3scale(bv0, v), v3scale(bv1, w)));
      break;
    }
    default: break;
  }
}

/* GuGJKSimplex.h:162-283: closest point of the origin on triangle abc, the sub-simplex that holds it in indices / size; returns the squared distance */
PXB_D float gjk_closest_triangle_bary(v3 a, v3 b, v3 c, uint32_t* indices, uint32_t* size, v3* closestPt) {
  *size = 3;
  const float eps = FLT_EPSILON;
  const v3 ab = v3sub(b, a), ac = v3sub(c, a);
  const v3 n = v3cross(ab, ac);
  const float nn = adot(n, n);
  if (nn == 0.f) return FLT_MAX;
  const v3 bCrossC = v3cross(b, c), cCrossA = v3cross(c, a), aCrossB = v3cross(a, b);
  const float va = adot(n, bCrossC), vb = adot(n, cCrossA), vc = adot(n, aCrossB);
  if (va >= 0.f && vb >= 0.f && vc >= 0.f) {
    const float t = adot(n, a) / nn;
    const v3 q = v3scale(n, t);
    *closestPt = q; return adot(q, q);
  }
  const v3 ap = v3neg(a), bp = v3neg(b), cp = v3neg(c);
  const float d1 = adot(ab, ap), d2 = adot(ac, ap), d3 = adot(ab, bp), d4 = adot(ac, bp), d5 = adot(ab, cp), d6 = adot(ac, cp);
  const float unom = d4 - d3, udenom = d5 - d6;
  *size = 2;
  if (0.f >= vc && d1 >= 0.f && 0.f >= d3) {   /* edge AB */
    const float toRecip = d1 - d3;
    const float recip = fabsf(toRecip) > eps ? 1.0f / toRecip : 0.f;
    const float t = d1 * recip;
    const v3 q = v3scaleadd(ab, t, a);
    *closestPt = q; return adot(q, q);
  }
  if (0.f >= va && d4 >= d3 && d5 >= d6) {   /* edge BC */
    const v3 bc = v3sub(c, b);
    const float toRecip = unom + udenom;
    const float recip = fabsf(toRecip) > eps ? 1.0f / toRecip : 0.f;
    const float t = unom * recip;
    indices[0] = indices[1]; indices[1] = indices[2];
    const v3 q = v3scaleadd(bc, t, b);
    *closestPt = q; return adot(q, q);
  }
  if (0.f >= vb && d2 >= 0.f && 0.f >= d6) {   /* edge AC */
    const float toRecip = d2 - d6;
    const float recip = fabsf(toRecip) > eps ? 1.0f / toRecip : 0.f;
    const float t = d2 * recip;
    indices[1] = indices[2];
    const v3 q = v3scaleadd(ac, t, a);
    *closestPt = q; return adot(q, q);
  }
  *size = 1;
  if (0.f >= d1 && 0.f >= d2) { *closestPt = a; return adot(a, a); }
  if (d3 >= 0.f && d3 >= d4) { indices[0] = indices[1]; *closestPt = b; return adot(b, b); }
  indices[0] = indices[2]; *closestPt = c; return adot(c, c);
}

/* GuGJKSimplex.h:329-378 (the index-carrying overload) */
PXB_D v3 gjk_closest_triangle(v3* Q, v3* A, v3* B, int* aInd, int* bInd, uint32_t* size) {
  *size = 3;
  const float eps2 = FLT_EPSILON * FLT_EPSILON;
  const v3 a = Q[0], b = Q[1], c = Q[2];
  const v3 ab = v3sub(b, a), ac = v3sub(c, a);
  const v3 signArea = v3cross(ab, ac);
  const float area = adot(signArea, signArea);
  if (eps2 >= area) { *size = 2; return gjk_closest_segment(Q, size); }
  uint32_t _size; uint32_t indices[3] = {0, 1, 2};
  v3 closestPt = V3(0, 0, 0);
  gjk_closest_triangle_bary(a, b, c, indices, &_size, &closestPt);
  if (_size != 3) {
    const v3 q0 = Q[indices[0]], q1 = Q[indices[1]], a0 = A[indices[0]], a1 = A[indices[1]], b0 = B[indices[0]], b1 = B[indices[1]];
    const int ai0 = aInd[indices[0]], ai1 = aInd[indices[1]], bi0 = bInd[indices[0]], bi1 = bInd[indices[1]];
    Q[0] = q0; Q[1] = q1; A[0] = a0; A[1] = a1; B[0] = b0; B[1] = b1; aInd[0] = ai0; aInd[1] = ai1; bInd[0] = bi0; bInd[1] = bi1;
    *size = _size;
  }
  return closestPt;
}

/* GuGJKSimplex.h:56-89: per face (abc, acd, adb, bdc) -- is the origin outside (on the other side than the fourth vertex)? */
PXB_D void gjk_point_outside_of_plane4(v3 a, v3 b, v3 c, v3 d, int out[4]) {
  const v3 ab = v3sub(b, a), ac = v3sub(c, a), ad = v3sub(d, a), bd = v3sub(d, b), bc = v3sub(c, b);
  const v3 v0 = v3cross(ab, ac), v1 = v3cross(ac, ad), v2 = v3cross(ad, ab), v3_ = v3cross(bd, bc);
  const float signa0 = adot(v0, a), signa1 = adot(v1, a), signa2 = adot(v2, a), signd3 = adot(v3_, a);
  const float signd0 = adot(v0, d), signd1 = adot(v1, b), signd2 = adot(v2, c), signa3 = adot(v3_, b);
  out[0] = signa0 * signd0 >= 0.f; out[1] = signa1 * signd1 >= 0.f; out[2] = signa2 * signd2 >= 0.f; out[3] = signa3 * signd3 >= 0.f;
}

/* GuGJKSimplex.cpp:38-121 */
PXB_D v3 gjk_closest_of_faces(const v3* Q, const int outside[4], uint32_t* indices, uint32_t* size) {
  float bestSqDist = FLT_MAX;
  v3 closestPt = V3(0, 0, 0);
  if (outside[0]) bestSqDist = gjk_closest_triangle_bary(Q[0], Q[1], Q[2], indices, size, &closestPt);
  const uint32_t faces[3][3] = {{0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
  for (int f = 0; f < 3; ++f) {
    if (!outside[f + 1]) continue;
    uint32_t _size = 3; uint32_t _indices[3] = {faces[f][0], faces[f][1], faces[f][2]};
    v3 t = V3(0, 0, 0);
    const float sqDist = gjk_closest_triangle_bary(Q[faces[f][0]], Q[faces[f][1]], Q[faces[f][2]], _indices, &_size, &t);
    if (bestSqDist > sqDist) { closestPt = t; bestSqDist = sqDist; indices[0] = _indices[0]; indices[1] = _indices[1]; indices[2] = _indices[2]; *size = _size; }
  }
  return closestPt;
}

/* GuGJKSimplex.cpp:163-213 */
PXB_D v3 gjk_closest_tetrahedron(v3* Q, v3* A, v3* B, int* aInd, int* bInd, uint32_t* size) {
  const v3 a = Q[0], b = Q[1], c = Q[2], d = Q[3];
  const v3 ab = v3sub(b, a), ac = v3sub(c, a);
  const v3 n = anormalize(v3cross(ab, ac));
  const float signDist = adot(n, v3sub(d, a));
  if (1e-4f > fabsf(signDist)) { *size = 3; return gjk_closest_triangle(Q, A, B, aInd, bInd, size); }
  int outside[4]; gjk_point_outside_of_plane4(a, b, c, d, outside);
  if (!outside[0] && !outside[1] && !outside[2] && !outside[3]) return V3(0, 0, 0);   /* origin inside: size stays 4 */
  uint32_t indices[3] = {0, 1, 2};
  const v3 closest = gjk_closest_of_faces(Q, outside, indices, size);
  const v3 q0 = Q[indices[0]], q1 = Q[indices[1]], q2 = Q[indices[2]], a0 = A[indices[0]], a1 = A[indices[1]], a2 = A[indices[2]];
  const v3 b0 = B[indices[0]], b1 = B[indices[1]], b2 = B[indices[2]];
  const int ai0 = aInd[indices[0]], ai1 = aInd[indices[1]], ai2 = aInd[indices[2]], bi0 = bInd[indices[0]], bi1 = bInd[indices[1]], bi2 = bInd[indices[2]];
  Q[0] = q0; Q[1] = q1; Q[2] = q2; A[0] = a0; A[1] = a1; A[2] = a2; B[0] = b0; B[1] = b1; B[2] = b2;
  aInd[0] = ai0; aInd[1] = ai1; aInd[2] = ai2; bInd[0] = bi0; bInd[1] = bi1; bInd[2] = bi2;
  return closest;
}

/* GuGJKSimplex.h:413-443 */
PXB_D v3 gjk_do_simplex(v3* Q, v3* A, v3* B, int* aInd, int* bInd, v3 support, uint32_t* size) {
  switch (*size) {
    case 1: return support;
    case 2: return gjk_closest_segment(Q, size);
    case 3: return gjk_closest_triangle(Q, A, B, aInd, bInd, size);
    case 4: return gjk_closest_tetrahedron(Q, A, B, aInd, bInd, size);
    default: return support;
  }
}

/* ---------------- gjkPenetration: GuGJKPenetration.h:89-311 ----------------
 * aIndices / bIndices / warmStartSize: the manifold's warm-start cache (mAIndice, mBIndice, mNumWarmStartPoints). */
PXB_D int gjk_penetration(const GjkConvex* a, const GjkConvex* b, v3 initialSearchDir, float contactDist, int takeCoreShape,
                                      uint8_t* aIndices, uint8_t* bIndices, uint8_t* warmStartSize, GjkOutput* output) {
  const float minMargin = fmin_(a->minMargin, b->minMargin);
  const float eps = minMargin * 0.1f;
  const float epsRel = 0.000225f, relDif = 1.0f - epsRel;
  const float tMarginA = a->marginIsRadius ? a->margin : 0.f, tMarginB = b->marginIsRadius ? b->margin : 0.f;
  const float sumMargin = tMarginA + tMarginB;
  const float sumExpandedMargin = sumMargin + contactDist;
  float dist = FLT_MAX, prevDist = dist;
  v3 prevClos = V3(0, 0, 0);
  int notTerminated = 1, notDegenerated = 1;
  v3 closest, v;
  v3 Q[4], A[4], B[4]; int aInd[4], bInd[4];
  v3 supportA = V3(0, 0, 0), supportB = V3(0, 0, 0), support = V3(0, 0, 0);
  uint32_t size = 0;
#define GJK_ASSIGN_WARM(n) do { *warmStartSize = (uint8_t)(n); for (uint32_t i_ = 0; i_ < (uint32_t)(n); ++i_) { aIndices[i_] = (uint8_t)aInd[i_]; bIndices[i_] = (uint8_t)bInd[i_]; } } while (0)
  if (*warmStartSize != 0) {
    for (uint32_t i = 0; i < *warmStartSize; ++i) {
      aInd[i] = aIndices[i]; bInd[i] = bIndices[i];
      supportA = gjk_cvx_support_point(a, aIndices[i]); supportB = gjk_cvx_support_point(b, bIndices[i]);
      support = v3sub(supportA, supportB);
      A[size] = supportA; B[size] = supportB; Q[size++] = support;
    }
    closest = gjk_do_simplex(Q, A, B, aInd, bInd, support, &size);
    dist = alen(closest);
    v = gjk_v3div(closest, dist);
    prevDist = dist; prevClos = closest;
    notTerminated = dist > eps;
  } else {
    closest = adot(initialSearchDir, initialSearchDir) > 0.f ? initialSearchDir : V3(1, 0, 0);
    v = anormalize(closest);
  }
  while (notTerminated) {
    prevDist = dist; prevClos = closest;
    supportA = gjk_cvx_support(a, v3neg(closest), &aInd[size]);
    supportB = gjk_cvx_support(b, closest, &bInd[size]);
    support = v3sub(supportA, supportB);
    const float vw = adot(v, support);
    if (vw > sumExpandedMargin) { GJK_ASSIGN_WARM(size); return GJK_NON_INTERSECT; }
    if (vw > dist * relDif) {
      GJK_ASSIGN_WARM(size);
      output->normal = v;
      v3 closA = V3(0, 0, 0), closB = V3(0, 0, 0);
      gjk_closest_points(Q, A, B, closest, &closA, &closB, size);
      if (takeCoreShape) { output->closestA = closA; output->closestB = closB; output->penDep = dist; }
      else { output->closestA = v3negscalesub(v, tMarginA, closA); output->closestB = v3scaleadd(v, tMarginB, closB); output->penDep = dist - sumMargin; }
      return GJK_CONTACT;
    }
    A[size] = supportA; B[size] = supportB; Q[size++] = support;
    closest = gjk_do_simplex(Q, A, B, aInd, bInd, support, &size);
    dist = alen(closest);
    v = gjk_v3div(closest, dist);
    notDegenerated = prevDist > dist;
    notTerminated = (dist > eps) && notDegenerated;
  }
  if (!notDegenerated) {
    GJK_ASSIGN_WARM(size - 1);
    dist = prevDist; closest = prevClos;
    v3 closA = V3(0, 0, 0), closB = V3(0, 0, 0);
    gjk_closest_points(Q, A, B, closest, &closA, &closB, size);
    const v3 n = gjk_v3div(prevClos, prevDist);
    output->normal = n; output->searchDir = v;
    if (takeCoreShape) { output->closestA = closA; output->closestB = closB; output->penDep = dist; }
    else {
      output->closestA = v3negscalesub(n, tMarginA, closA); output->closestB = v3scaleadd(n, tMarginB, closB); output->penDep = dist - sumMargin;
      if (sumMargin >= dist) return GJK_CONTACT;
    }
    return GJK_DEGENERATE;
  }
  GJK_ASSIGN_WARM(size);
  return EPA_CONTACT;
#undef GJK_ASSIGN_WARM
}

/* ---------------- epaPenetration: GuEPA.cpp:70-623, GuEPAFacet.h:61-301 ----------------
 * Expanding polytope over the Minkowski difference, started from the GJK simplex (warm-start indices).  Facets live in a pool of 64
 * (Cm::InlineDeferredIDPool, CmIDPool.h:40-191), the open facets in a binary heap keyed by plane distance (CmPriorityQueue.h:77-118). */
#define EPA_MAX_FACETS 64
#define EPA_MAX_EDGES 32
#define EPA_MAX_SUPPORT 64
typedef struct { v3 n; float d; int8_t adjF[3], adjE[3], idx[3]; uint8_t obsolete, inHeap; } EpaFacet;
typedef struct {
  v3 aBuf[EPA_MAX_SUPPORT], bBuf[EPA_MAX_SUPPORT];
  EpaFacet f[EPA_MAX_FACETS];
  uint8_t heap[EPA_MAX_FACETS]; uint32_t heapSize;
  uint8_t edgeF[EPA_MAX_EDGES], edgeI[EPA_MAX_EDGES]; uint32_t edgeSize; int edgeOverflow;
  uint32_t curId, nFree, nDeferred; uint8_t freeIds[EPA_MAX_FACETS], deferred[EPA_MAX_FACETS];
} EpaScratch;

PXB_D void gjk_epa_heap_push(EpaScratch* e, uint8_t id) {
  uint32_t newIndex, parentIndex = (e->heapSize - 1) >> 1;
  for (newIndex = e->heapSize; newIndex > 0 && e->f[id].d < e->f[e->heap[parentIndex]].d; newIndex = parentIndex, parentIndex = (newIndex - 1) >> 1) e->heap[newIndex] = e->heap[parentIndex];
  e->heap[newIndex] = id; e->heapSize++;
}
PXB_D uint8_t gjk_epa_heap_pop(EpaScratch* e) {
  uint32_t i, child; const uint32_t tempHs = e->heapSize - 1;
  e->heapSize = tempHs;
  const uint8_t mn = e->heap[0], last = e->heap[tempHs];
  for (i = 0; (child = (i << 1) + 1) < tempHs; i = child) {
    const uint32_t rightChild = child + 1;
    child += ((rightChild < tempHs) && (e->f[e->heap[rightChild]].d < e->f[e->heap[child]].d)) ? 1 : 0;
    if (e->f[last].d < e->f[e->heap[child]].d) break;
    e->heap[i] = e->heap[child];
  }
  e->heap[i] = last;
  return mn;
}
PXB_D uint32_t gjk_epa_new_id(EpaScratch* e) { if (e->nFree) return e->freeIds[--e->nFree]; return e->curId++; }
PXB_D void gjk_epa_free_id(EpaScratch* e, uint32_t id) { if (id == e->curId - 1) --e->curId; else e->freeIds[e->nFree++] = (uint8_t)id; }
PXB_D void gjk_epa_process_deferred(EpaScratch* e) { for (uint32_t a = 0; a < e->nDeferred; ++a) gjk_epa_free_id(e, e->deferred[a]); e->nDeferred = 0; }
PXB_D uint32_t gjk_epa_remaining_ids(const EpaScratch* e) { return EPA_MAX_FACETS - (e->curId - e->nFree); }

/* Facet::isValid2 GuEPA.cpp:137-172 + EPA::addFacet :174-199 */
PXB_D int gjk_epa_add_facet(EpaScratch* e, uint32_t i0, uint32_t i1, uint32_t i2, float upper) {
  const uint32_t id = gjk_epa_new_id(e);
  EpaFacet* f = &e->f[id];
  f->obsolete = 0; f->inHeap = 0; f->idx[0] = (int8_t)i0; f->idx[1] = (int8_t)i1; f->idx[2] = (int8_t)i2;
  f->adjF[0] = f->adjF[1] = f->adjF[2] = -1; f->adjE[0] = f->adjE[1] = f->adjE[2] = -1;
  const v3 p0 = v3sub(e->aBuf[i0], e->bBuf[i0]), p1 = v3sub(e->aBuf[i1], e->bBuf[i1]), p2 = v3sub(e->aBuf[i2], e->bBuf[i2]);
  const v3 v0 = v3sub(p1, p0), v1 = v3sub(p2, p0);
  const v3 denormalizedNormal = v3cross(v0, v1);
  float norValue = adot(denormalizedNormal, denormalizedNormal);
  const int con = norValue > FLT_EPSILON;
  norValue = con ? norValue : 1.0f;
  const v3 planeNormal = v3scale(denormalizedNormal, 1.0f / sqrtf(norValue));
  const float planeDist = adot(planeNormal, p0);
  f->n = planeNormal; f->d = planeDist;
  if (con && upper >= planeDist) { gjk_epa_heap_push(e, (uint8_t)id); f->inHeap = 1; }
  return (int)id;
}
/* Facet::link GuEPAFacet.h:290-298 */
PXB_D void gjk_epa_link(EpaScratch* e, int f0, uint32_t edge0, int f1, uint32_t edge1) {
  e->f[f0].adjF[edge0] = (int8_t)f1; e->f[f0].adjE[edge0] = (int8_t)edge1; e->f[f1].adjF[edge1] = (int8_t)f0; e->f[f1].adjE[edge1] = (int8_t)edge0;
}
PXB_D float gjk_epa_plane_dist(const EpaScratch* e, const EpaFacet* f, v3 p) {
  const v3 p0 = v3sub(e->aBuf[f->idx[0]], e->bBuf[f->idx[0]]);
  return adot(f->n, v3sub(p, p0));
}
/* Facet::silhouette(index, w, ...) GuEPA.cpp:201-240 */
PXB_D void gjk_epa_silhouette_edge(EpaScratch* e, int facet, uint32_t _index, v3 w) {
  int stackF[EPA_MAX_FACETS]; uint32_t stackI[EPA_MAX_FACETS];
  stackF[0] = facet; stackI[0] = _index;
  int size = 1;
  while (size--) {
    EpaFacet* f = &e->f[stackF[size]]; const uint32_t index = stackI[size]; const int fid = stackF[size];
    if (!f->obsolete) {
      const float pointPlaneDist = gjk_epa_plane_dist(e, f, w);
      if (0.f > pointPlaneDist) {
        if (e->edgeSize < EPA_MAX_EDGES) { e->edgeF[e->edgeSize] = (uint8_t)fid; e->edgeI[e->edgeSize] = (uint8_t)index; e->edgeSize++; }
        else { e->edgeOverflow = 1; return; }
      } else {
        f->obsolete = 1;
        const uint32_t next = (index + 1) % 3, next2 = (next + 1) % 3;
        stackF[size] = f->adjF[next2]; stackI[size] = (uint32_t)f->adjE[next2]; size++;
        stackF[size] = f->adjF[next]; stackI[size] = (uint32_t)f->adjE[next]; size++;
        if (!f->inHeap) e->deferred[e->nDeferred++] = (uint8_t)fid;
      }
    }
  }
}
/* Facet::getClosestPoint GuEPAFacet.h:252-288 + calculateContactInformation GuEPA.cpp:306-337 */
PXB_D void gjk_epa_contact_info(const EpaScratch* e, const EpaFacet* f, const GjkConvex* a, const GjkConvex* b, int takeCoreShape, GjkOutput* out) {
  const v3 pa0 = e->aBuf[f->idx[0]], pa1 = e->aBuf[f->idx[1]], pa2 = e->aBuf[f->idx[2]], pb0 = e->bBuf[f->idx[0]], pb1 = e->bBuf[f->idx[1]], pb2 = e->bBuf[f->idx[2]];
  const v3 p0 = v3sub(pa0, pb0), p1 = v3sub(pa1, pb1), p2 = v3sub(pa2, pb2);
  const v3 v0 = v3sub(p1, p0), v1 = v3sub(p2, p0);
  const v3 closestP = v3scale(f->n, f->d);
  const v3 v2 = v3sub(closestP, p0);
  const float d00 = adot(v0, v0), d01 = adot(v0, v1), d11 = adot(v1, v1), d20 = adot(v2, v0), d21 = adot(v2, v1);
  const float det = d00 * d11 - d01 * d01;
  const float recip = det > FLT_EPSILON ? 1.0f / det : 0.f;
  const float lambda1 = (d11 * d20 - d01 * d21) * recip, lambda2 = (d00 * d21 - d01 * d20) * recip;
  const float u = 1.0f - (lambda1 + lambda2);
  const v3 _pa = v3scaleadd(pa0, u, v3scaleadd(pa1, lambda1, v3scale(pa2, lambda2)));
  const v3 _pb = v3scaleadd(pb0, u, v3scaleadd(pb1, lambda1, v3scale(pb2, lambda2)));
  const float dist = fabsf(f->d);
  const v3 planeNormal = v3neg(f->n);
  if (takeCoreShape) { out->closestA = _pa; out->closestB = _pb; out->normal = planeNormal; out->penDep = -dist; }
  else {
    const float marginA = a->marginIsRadius ? a->margin : 0.f, marginB = b->marginIsRadius ? b->margin : 0.f;
    const float sumMargin = marginA + marginB;
    out->closestA = v3negscalesub(planeNormal, marginA, _pa); out->closestB = v3scaleadd(planeNormal, marginB, _pb); out->normal = planeNormal; out->penDep = -(dist + sumMargin);
  }
}
PXB_D v3 gjk_cvx_support_noidx(const GjkConvex* c, v3 dir) { int i; return gjk_cvx_support(c, dir, &i); }
/* EPA::expandTriangle :293-304 */
PXB_D int gjk_epa_expand_triangle(EpaScratch* e, int* numVerts, float upper) {
  *numVerts = 3;
  const int f0 = gjk_epa_add_facet(e, 0, 1, 2, upper), f1 = gjk_epa_add_facet(e, 1, 0, 2, upper);
  if (e->heapSize == 0) return 0;
  gjk_epa_link(e, f0, 0, f1, 0); gjk_epa_link(e, f0, 1, f1, 2); gjk_epa_link(e, f0, 2, f1, 1);
  return 1;
}
/* EPA::expandSegment :257-291 */
PXB_D int gjk_epa_expand_segment(EpaScratch* e, const GjkConvex* a, const GjkConvex* b, int* numVerts, float upper) {
  const v3 q0 = v3sub(e->aBuf[0], e->bBuf[0]), q1 = v3sub(e->aBuf[1], e->bBuf[1]);
  const v3 v = v3sub(q1, q0), absV = v3abs(v);
  v3 axis = V3(1, 0, 0);
  if (absV.x > absV.y && absV.z > absV.y) axis = V3(0, 1, 0);
  else if (absV.x > absV.z) axis = V3(0, 0, 1);
  const v3 n = anormalize(v3cross(axis, v));
  e->aBuf[2] = gjk_cvx_support_noidx(a, v3neg(n)); e->bBuf[2] = gjk_cvx_support_noidx(b, n);   /* doSupport :83-90 */
  return gjk_epa_expand_triangle(e, numVerts, upper);
}
/* EPA::expandPoint :242-255 */
PXB_D int gjk_epa_expand_point(EpaScratch* e, const GjkConvex* a, const GjkConvex* b, int* numVerts, float upper) {
  const v3 x = V3(1, 0, 0);
  const v3 q0 = v3sub(e->aBuf[0], e->bBuf[0]);
  e->aBuf[1] = gjk_cvx_support_noidx(a, v3neg(x)); e->bBuf[1] = gjk_cvx_support_noidx(b, x);
  const v3 q1 = v3sub(e->aBuf[1], e->bBuf[1]);
  if (q0.x == q1.x && q0.y == q1.y && q0.z == q1.z) return 0;
  return gjk_epa_expand_segment(e, a, b, numVerts, upper);
}
/* epaPenetration (index overload) :92-110 + EPA::PenetrationDepth :339-621 */
PXB_D int gjk_epa_penetration(const GjkConvex* a, const GjkConvex* b, const uint8_t* aInd, const uint8_t* bInd, uint8_t size, int takeCoreShape, float toleranceLength, GjkOutput* output) {
  EpaScratch epa_;   // ~4 KB of per-thread local memory, touched only by the (rare) pairs that reach EPA
  EpaScratch* e = &epa_;
  e->heapSize = 0; e->edgeSize = 0; e->edgeOverflow = 0; e->curId = 0; e->nFree = 0; e->nDeferred = 0;
  for (int i = 0; i < 4; ++i) { e->aBuf[i] = V3(0, 0, 0); e->bBuf[i] = V3(0, 0, 0); }
  for (uint32_t i = 0; i < size; ++i) { e->aBuf[i] = gjk_cvx_support_point(a, aInd[i]); e->bBuf[i] = gjk_cvx_support_point(b, bInd[i]); }
  float upper_bound = FLT_MAX;
  int numVertsLocal = 0;
  switch (size) {
    case 1: if (!gjk_epa_expand_point(e, a, b, &numVertsLocal, upper_bound)) return EPA_FAIL; break;
    case 2: if (!gjk_epa_expand_segment(e, a, b, &numVertsLocal, upper_bound)) return EPA_FAIL; break;
    case 3: if (!gjk_epa_expand_triangle(e, &numVertsLocal, upper_bound)) return EPA_FAIL; break;
    case 4: {
      const v3 p0 = v3sub(e->aBuf[0], e->bBuf[0]), p1 = v3sub(e->aBuf[1], e->bBuf[1]), p2 = v3sub(e->aBuf[2], e->bBuf[2]), p3 = v3sub(e->aBuf[3], e->bBuf[3]);
      const v3 v1 = v3sub(p1, p0), v2 = v3sub(p2, p0);
      const v3 planeNormal = anormalize(v3cross(v1, v2));
      const float signDist = adot(planeNormal, v3sub(p3, p0));
      if (signDist > 0.f) { const v3 ta = e->aBuf[2], tb = e->bBuf[2]; e->aBuf[2] = e->aBuf[1]; e->bBuf[2] = e->bBuf[1]; e->aBuf[1] = ta; e->bBuf[1] = tb; }
      const int f0 = gjk_epa_add_facet(e, 0, 1, 2, upper_bound), f1 = gjk_epa_add_facet(e, 0, 3, 1, upper_bound), f2 = gjk_epa_add_facet(e, 0, 2, 3, upper_bound), f3 = gjk_epa_add_facet(e, 1, 3, 2, upper_bound);
      if (e->heapSize == 0) return EPA_FAIL;
      gjk_epa_link(e, f0, 0, f1, 2); gjk_epa_link(e, f0, 1, f3, 2); gjk_epa_link(e, f0, 2, f2, 0); gjk_epa_link(e, f1, 0, f2, 2); gjk_epa_link(e, f1, 1, f3, 0); gjk_epa_link(e, f2, 1, f3, 1);
      numVertsLocal = 4;
      break;
    }
    default: return EPA_FAIL;
  }
  const float minMargin = fmin_(a->minMargin, b->minMargin);
  const float eps = minMargin * 0.1f;
  int facetId = -1;
  do {
    gjk_epa_process_deferred(e);
    facetId = gjk_epa_heap_pop(e);
    EpaFacet* facet = &e->f[facetId];
    facet->inHeap = 0;
    if (!facet->obsolete) {
      const v3 planeNormal = facet->n; const float planeDist = facet->d;
      const v3 tempa = gjk_cvx_support_noidx(a, planeNormal), tempb = gjk_cvx_support_noidx(b, v3neg(planeNormal));
      const v3 q = v3sub(tempa, tempb);
      const float dist = adot(q, planeNormal);
      if (eps >= fabsf(dist - planeDist)) {
        gjk_epa_contact_info(e, facet, a, b, takeCoreShape, output);
        if (takeCoreShape) {
          const float toleranceEps = 1e-3f * toleranceLength;
          const v3 dif = v3sub(output->closestA, output->closestB);
          const float pen = fabsf(output->penDep) + toleranceEps;
          const float sqDif = adot(dif, dif);
          const float length = sqDif > 0.f ? sqrtf(sqDif) : 0.f;
          if (length > pen) return EPA_DEGENERATE;
        }
        return EPA_CONTACT;
      }
      upper_bound = fmin_(upper_bound, dist);
      e->aBuf[numVertsLocal] = tempa; e->bBuf[numVertsLocal] = tempb;
      const uint32_t index = (uint32_t)numVertsLocal++;
      e->edgeSize = 0; e->edgeOverflow = 0;
      facet->obsolete = 1;   /* Facet::silhouette(w, ...) :242-250 */
      for (uint32_t k = 0; k < 3; ++k) gjk_epa_silhouette_edge(e, facet->adjF[k], (uint32_t)facet->adjE[k], q);
      if (!(e->edgeSize > 0 && !e->edgeOverflow)) { gjk_epa_contact_info(e, facet, a, b, takeCoreShape, output); return EPA_DEGENERATE; }
      const uint32_t bufferSize = e->edgeSize;
      if (bufferSize > gjk_epa_remaining_ids(e)) { gjk_epa_contact_info(e, facet, a, b, takeCoreShape, output); return EPA_DEGENERATE; }
#define EPA_EDGE_SRC(k) ((uint32_t)e->f[e->edgeF[k]].idx[e->edgeI[k]])
#define EPA_EDGE_TGT(k) ((uint32_t)e->f[e->edgeF[k]].idx[(e->edgeI[k] + 1) % 3])
      const int firstFacet = gjk_epa_add_facet(e, EPA_EDGE_TGT(0), EPA_EDGE_SRC(0), index, upper_bound);
      gjk_epa_link(e, firstFacet, 0, e->edgeF[0], e->edgeI[0]);
      int lastFacet = firstFacet;
      for (uint32_t i = 1; i < bufferSize; ++i) {
        const int newFacet = gjk_epa_add_facet(e, EPA_EDGE_TGT(i), EPA_EDGE_SRC(i), index, upper_bound);
        gjk_epa_link(e, newFacet, 0, e->edgeF[i], e->edgeI[i]);
        gjk_epa_link(e, newFacet, 2, lastFacet, 1);
        lastFacet = newFacet;
      }
#undef EPA_EDGE_SRC
#undef EPA_EDGE_TGT
      gjk_epa_link(e, firstFacet, 2, lastFacet, 1);
    }
    gjk_epa_free_id(e, (uint32_t)facetId);
  } while (e->heapSize > 0 && upper_bound > e->f[e->heap[0]].d && numVertsLocal != EPA_MAX_SUPPORT);
  gjk_epa_contact_info(e, &e->f[facetId], a, b, takeCoreShape, output);
  return EPA_DEGENERATE;
}

/* ---------------- box-box: the GJK / EPA single-point fallback of pcmContactBoxBox ----------------
 * GuPersistentContactManifold.cpp:43-60 */
PXB_D float gjk_dist_point_segment_sq_local(v3 a, v3 b, v3 p) {
  const v3 ap = v3sub(p, a), ab = v3sub(b, a);
  const float nom = adot(ap, ab), denom = adot(ab, ab);
  const float tValue = fmax_(fmin_(nom / denom, 1.f), 0.f);
  const float t = denom == 0.f ? 0.f : tValue;
  const v3 v = v3negscalesub(ab, t, ap);
  return adot(v, v);
}
/* GuPersistentContactManifold.cpp:61-170 */
PXB_D float gjk_dist_point_triangle_sq_local(v3 p, v3 a, v3 b, v3 c) {
  const v3 ab = v3sub(b, a), ac = v3sub(c, a), bc = v3sub(c, b), ap = v3sub(p, a), bp = v3sub(p, b), cp = v3sub(p, c);
  const float d1 = adot(ab, ap), d2 = adot(ac, ap), d3 = adot(ab, bp), d4 = adot(ac, bp), d5 = adot(ab, cp), d6 = adot(ac, cp);
  const float unom = d4 - d3, udenom = d5 - d6;
  if (0.f > d1 && 0.f > d2) { const v3 vv = v3sub(p, a); return adot(vv, vv); }
  if (d3 >= 0.f && d3 >= d4) { const v3 vv = v3sub(p, b); return adot(vv, vv); }
  if (d6 >= 0.f && d6 >= d5) { const v3 vv = v3sub(p, c); return adot(vv, vv); }
  const float vc = d1 * d4 - d3 * d2;
  if (0.f > vc && d1 >= 0.f && 0.f > d3) { const float sScale = d1 / (d1 - d3); const v3 vv = v3sub(p, v3scaleadd(ab, sScale, a)); return adot(vv, vv); }
  const float va = d3 * d6 - d5 * d4;
  if (0.f > va && d4 >= d3 && d5 >= d6) { const float uScale = unom / (unom + udenom); const v3 vv = v3sub(p, v3scaleadd(bc, uScale, b)); return adot(vv, vv); }
  const float vb = d5 * d2 - d1 * d6;
  if (0.f > vb && d2 >= 0.f && 0.f > d6) { const float tScale = d2 / (d2 - d6); const v3 vv = v3sub(p, v3scaleadd(ac, tScale, a)); return adot(vv, vv); }
  const v3 n = v3cross(ab, ac);
  const float nn = adot(n, n);
  const float t = nn > 0.f ? adot(n, v3sub(a, p)) / nn : 0.f;
  const v3 closest6 = v3add(p, v3scale(n, t));
  const v3 vv = v3sub(p, closest6);
  return adot(vv, vv);
}
/* PersistentContactManifold::addManifoldPoint (.cpp:1246-1266) -> replaceManifoldPoint (:552-576) / reduceContactsForPCM (:603-737) */
PXB_D void gjk_add_manifold_point(Manifold* m, v3 la, v3 lb, v3 n, float pen, float replaceBreakingThreshold) {
  const float shortest = replaceBreakingThreshold * replaceBreakingThreshold;
  for (int i = 0; i < m->n; ++i) {
    const v3 dB = v3sub(m->pts[i].b, lb), dA = v3sub(m->pts[i].a, la);
    if (shortest > fmin_(adot(dB, dB), adot(dA, dA))) { m->pts[i].a = la; m->pts[i].b = lb; m->pts[i].n = n; m->pts[i].pen = pen; return; }
  }
  if (m->n < 4) { m->pts[m->n].a = la; m->pts[m->n].b = lb; m->pts[m->n].n = n; m->pts[m->n].pen = pen; m->n++; return; }
  int chosen[5] = {0, 0, 0, 0, 0};
  MPoint temp[5];
  for (int i = 0; i < 4; ++i) temp[i] = m->pts[i];
  temp[4].a = la; temp[4].b = lb; temp[4].n = n; temp[4].pen = pen;
  float maxDist = pen; int index = 4;
  for (int i = 0; i < 4; ++i) if (maxDist > temp[i].pen) { maxDist = temp[i].pen; index = i; }
  chosen[index] = 1; m->pts[0] = temp[index];
  v3 dir = v3sub(temp[0].b, m->pts[0].b);
  maxDist = adot(dir, dir); index = 0;
  for (int i = 1; i < 5; ++i) if (!chosen[i]) { dir = v3sub(temp[i].b, m->pts[0].b); const float d = adot(dir, dir); if (d > maxDist) { maxDist = d; index = i; } }
  chosen[index] = 1; m->pts[1] = temp[index];
  maxDist = -FLT_MAX;
  for (int i = 0; i < 5; ++i) if (!chosen[i]) { const float sq = gjk_dist_point_segment_sq_local(m->pts[0].b, m->pts[1].b, temp[i].b); if (sq > maxDist) { maxDist = sq; index = i; } }
  chosen[index] = 1; m->pts[2] = temp[index];
  maxDist = -FLT_MAX;
  for (int i = 0; i < 5; ++i) if (!chosen[i]) { const float sq = gjk_dist_point_triangle_sq_local(temp[i].b, m->pts[0].b, m->pts[1].b, m->pts[2].b); if (sq > maxDist) { maxDist = sq; index = i; } }
  if (chosen[index]) { m->n = 3; return; }
  chosen[index] = 1; m->pts[3] = temp[index];
  int notChosen = 0;
  for (int a = 0; a < 5; ++a) if (!chosen[a]) { notChosen = a; break; }
  float closest = FLT_MAX; index = 0;
  for (int a = 0; a < 4; ++a) { const v3 dif = v3sub(m->pts[a].a, temp[notChosen].a); const float d2 = adot(dif, dif); if (closest > d2) { closest = d2; index = a; } }
  if (m->pts[index].pen > temp[notChosen].pen) m->pts[index] = temp[notChosen];
}
/* GuPCMContactBoxBox.cpp:918-958: the SAT passed but face clipping produced no point (edge-edge / corner configurations):
 * one GJK (or EPA) point is merged into whatever the manifold still holds.  manifold->rel / quatA / quatB were already updated by the caller. */
PXB_D void gjk_boxbox_gjk_fallback(const xf* tm0, const xf* tm1, v3 ext0, v3 ext1, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out) {
  const xf curRTrans = axfinvmul(tm1, tm0);
  const mxf aToB = amxffromxf(&curRTrans);
  const float minMargin = fmin_(box_margin(ext0, toleranceLength), box_margin(ext1, toleranceLength));
  GjkConvex box0 = gjk_cvx_box(V3(0, 0, 0), ext0); gjk_cvx_make_relative(&box0, &aToB);
  const GjkConvex box1 = gjk_cvx_box(V3(0, 0, 0), ext1);
  manifold->nWarm = 0; manifold->dirty = 1;
  GjkOutput output; output.normal = output.closestA = output.closestB = output.searchDir = V3(0, 0, 0); output.penDep = 0.f;
  int status = gjk_penetration(&box0, &box1, aToB.p, contactDist, 1, manifold->aInd, manifold->bInd, &manifold->nWarm, &output);
  if (status == EPA_CONTACT) status = gjk_epa_penetration(&box0, &box1, manifold->aInd, manifold->bInd, manifold->nWarm, 1, toleranceLength, &output);
  out->count = 0;
  if (status == GJK_CONTACT || status == EPA_CONTACT) {
    const float replaceBreakingThreshold = minMargin * 0.05f;
    gjk_add_manifold_point(manifold, amxftransforminv(&aToB, output.closestA), output.closestB, output.normal, output.penDep, replaceBreakingThreshold);
    out->normal = anormalize(aqrot(tm1->q, output.normal));
    for (int i = 0; i < manifold->n; ++i) {   /* addManifoldContactsToContactBuffer(buffer, normal, transf1, contactOffset) .cpp:739-759 */
      const float dist = manifold->pts[i].pen;
      if (contactDist >= dist) { out->point[out->count] = axftransform(tm1, manifold->pts[i].b); out->sep[out->count] = dist; out->count++; }
    }
  }
}

static __device__ __noinline__ void gjk_boxbox_gjk_fallback_outofline(const xf* tm0, const xf* tm1, v3 ext0, v3 ext1, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out) {
  gjk_boxbox_gjk_fallback(tm0, tm1, ext0, ext1, contactDist, toleranceLength, manifold, out);
}

/* CalculatePCMConvexMargin GuVecConvexHull.h:55-65 (identity scale) */
PXB_D float gjk_hull_pcm_margin(const DevHull& h, float toleranceLength);
PXB_D float gjk_hull_pcm_margin(const DevHull* h, float toleranceLength) { return gjk_hull_pcm_margin(*h, toleranceLength); }
PXB_D float gjk_hull_pcm_margin(const DevHull& h, float toleranceLength) {
  const float mn = fmin_(h.internalExtents.x, fmin_(h.internalExtents.y, h.internalExtents.z));
  return fmin_(mn * 0.25f, toleranceLength * 0.05f);
}

/* pcmContactPlaneConvex: GuPCMContactPlaneConvex.cpp:36-227 (shape0 = plane, shape1 = convex mesh, identity mesh scale).
 * Note the reference's vertex loop visits mPolygons[closestFaceIndex] on BOTH passes (the second pass was meant for polyIndex2):
 * restated as is, the duplicated points are merged by the manifold reduction. */
PXB_D void gjk_pcm_plane_convex(const xf* planeTm, const xf* convexTm, const DevHull& hull, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out) {
  const xf* transf0 = convexTm; const xf* transf1 = planeTm;
  const xf curTransf = axfinvmul(transf1, transf0);
  const float convexMargin = gjk_hull_pcm_margin(hull, toleranceLength);
  const v3 planeNormal = anormalize(aqbasis0(transf1->q));
  const v3 negPlaneNormal = v3neg(planeNormal);
  const float projectBreakingThreshold = convexMargin * 0.2f;
  const int initialContacts = manifold->n;
  const mxf aToB = amxffromxf(&curTransf);
  manifold_refresh(*manifold, aToB, projectBreakingThreshold);
  const int bLostContacts = manifold->n != initialContacts;
  if (bLostContacts || invalidate_plane(*manifold, curTransf, convexMargin, 0.2f)) {
    const v3 localNormal = V3(1, 0, 0);
    manifold->n = 0; manifold->rel = curTransf; manifold->dirty = 1;
    const v3 n = anormalize(amtmul(aToB.r, localNormal));   /* vertex2Shape = identity: M33MulV3(I, v) = v */
    const v3 nnormal = v3neg(n);
    MPoint mc[64]; int numContacts = 0;
    float minProj = FLT_MAX; uint32_t closestFaceIndex = 0, polyIndex2 = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < hull.nPolys; ++i) { const float proj = adot(n, hull.plane_n(i)); if (minProj > proj) { minProj = proj; closestFaceIndex = i; } }
    uint32_t closestEdge = 0xffffffffu;
    minProj = minProj - 5e-4f;
    float maxDpSq = minProj * minProj;
    for (uint32_t i = 0; i < hull.nEdges; ++i) {
      const uint8_t f0 = hull.facesByEdges[i * 2], f1 = hull.facesByEdges[i * 2 + 1];
      const v3 edgeNormal = v3add(hull.plane_n(f0), hull.plane_n(f1));
      const float enMagSq = adot(edgeNormal, edgeNormal), dp = adot(edgeNormal, nnormal), sqDp = dp * dp;
      if (dp >= 0.f && sqDp > maxDpSq * enMagSq) { maxDpSq = sqDp / enMagSq; closestEdge = i; }
    }
    if (closestEdge != 0xffffffffu) {
      const uint32_t f0 = hull.facesByEdges[closestEdge * 2], f1 = hull.facesByEdges[closestEdge * 2 + 1];
      const float dp0 = adot(hull.plane_n(f0), nnormal), dp1 = adot(hull.plane_n(f1), nnormal);
      if (dp0 > dp1) { closestFaceIndex = f0; polyIndex2 = f1; } else { closestFaceIndex = f1; polyIndex2 = f0; }
    }
    for (uint32_t index = closestFaceIndex; index != 0xFFFFFFFFu; index = polyIndex2, polyIndex2 = 0xFFFFFFFFu) {
      const uint4 face = hull.poly_meta(closestFaceIndex);
      const uint8_t* vertInds = hull.vertexRefs + face.x;
      for (uint32_t i = 0; i < face.y; ++i) {
        const v3 pInVertexSpace = hull.vert(vertInds[i]);
        const v3 pInPlaneSpace = amxftransform(&aToB, pInVertexSpace);   /* aToBVertexSpace = (aToB.p, aToB.rot * I) */
        const float signDist = pInPlaneSpace.x;
        if (contactDist > signDist) {
          mc[numContacts].a = pInVertexSpace; mc[numContacts].b = v3negscalesub(localNormal, signDist, pInPlaneSpace); mc[numContacts].n = localNormal; mc[numContacts].pen = signDist; numContacts++;
          if (numContacts == 64) { reduce_cluster(*manifold, mc, numContacts); numContacts = PXB_MANIFOLD_CACHE; for (int c = 0; c < PXB_MANIFOLD_CACHE; ++c) mc[c] = manifold->pts[c]; }
        }
      }
    }
    /* addBatchManifoldContacts .cpp:812-831 */
    if (numContacts <= PXB_MANIFOLD_CACHE) { for (int i = 0; i < numContacts; ++i) manifold->pts[i] = mc[i]; manifold->n = numContacts; }
    else { reduce_batch(*manifold, mc, numContacts, toleranceLength); manifold->n = PXB_MANIFOLD_CACHE; }
  }
  out->count = 0; out->normal = negPlaneNormal;
  for (int i = 0; i < manifold->n; ++i) {   /* addManifoldContactsToContactBuffer(buffer, normal, transf1, contactOffset) .cpp:739-759 */
    const float dist = manifold->pts[i].pen;
    if (contactDist >= dist) { out->point[out->count] = axftransform(transf1, manifold->pts[i].b); out->sep[out->count] = dist; out->count++; }
  }
}

/* ---------------- polygonal box: GuPCMShapeConvex.cpp:40-110 ---------------- */
typedef struct { v3 n; float d; int minIndex; } Poly;
__device__ const uint8_t gjk_box_poly_refs[24] = {0, 3, 2, 1, 1, 2, 6, 5, 5, 6, 7, 4, 4, 7, 3, 0, 3, 7, 6, 2, 4, 0, 1, 5};
typedef struct { v3 verts[8]; Poly polys[6]; } PolyBox;
PXB_D void gjk_poly_box(PolyBox* pb, v3 h) {
  const v3 mn = v3neg(h), mx = h;
  pb->verts[0] = V3(mn.x, mn.y, mn.z); pb->verts[1] = V3(mx.x, mn.y, mn.z); pb->verts[2] = V3(mx.x, mx.y, mn.z); pb->verts[3] = V3(mn.x, mx.y, mn.z);
  pb->verts[4] = V3(mn.x, mn.y, mx.z); pb->verts[5] = V3(mx.x, mn.y, mx.z); pb->verts[6] = V3(mx.x, mx.y, mx.z); pb->verts[7] = V3(mn.x, mx.y, mx.z);
  pb->polys[1].n = V3(1, 0, 0);  pb->polys[1].d = -h.x; pb->polys[1].minIndex = 0;
  pb->polys[3].n = V3(-1, 0, 0); pb->polys[3].d = -h.x; pb->polys[3].minIndex = 1;
  pb->polys[4].n = V3(0, 1, 0);  pb->polys[4].d = -h.y; pb->polys[4].minIndex = 0;
  pb->polys[5].n = V3(0, -1, 0); pb->polys[5].d = -h.y; pb->polys[5].minIndex = 2;
  pb->polys[2].n = V3(0, 0, 1);  pb->polys[2].d = -h.z; pb->polys[2].minIndex = 0;
  pb->polys[0].n = V3(0, 0, -1); pb->polys[0].d = -h.z; pb->polys[0].minIndex = 4;
}

/* GuPCMContactGenUtil.cpp:105-135: box polygonal data has no edges (mNbEdges = 0), so only the face loop runs */
PXB_D int gjk_box_polygon_index(const PolyBox* pb, v3 normal) {
  const v3 n = normal;   /* vertex2Shape = identity */
  float minProj = adot(n, pb->polys[0].n);
  int closest = 0;
  for (int i = 1; i < 6; ++i) { const float proj = adot(n, pb->polys[i].n); if (minProj > proj) { minProj = proj; closest = i; } }
  return closest;
}
/* GuPCMContactGenUtil.cpp:204-260 (PxPlane::distance = n.dot(p) + d with PxVec3::dot = x+y+z order) */
PXB_D int gjk_box_witness_polygon_index(const PolyBox* pb, v3 normal, v3 closest, float tolerance) {
  float pd[6];
  const float eps = -tolerance;
  float dist = v3dot(closest, pb->polys[0].n) + pb->polys[0].d;
  float minDist = dist >= eps ? fabsf(dist) : FLT_MAX;
  pd[0] = minDist;
  float maxDist = dist; int maxFace = 0, closestFace = 0;
  for (int i = 1; i < 6; ++i) {
    dist = v3dot(closest, pb->polys[i].n) + pb->polys[i].d;
    pd[i] = dist >= eps ? fabsf(dist) : FLT_MAX;
    if (minDist > pd[i]) { minDist = pd[i]; closestFace = i; }
    if (dist > maxDist) { maxDist = dist; maxFace = i; }
  }
  if (minDist == FLT_MAX) return maxFace;
  float bestProj = adot(anormalize(pb->polys[closestFace].n), normal);
  const int first = closestFace;
  for (int i = 0; i < 6; ++i) {
    if ((tolerance > (pd[i] - minDist)) && first != i) {
      const float proj = adot(anormalize(pb->polys[i].n), normal);
      if (bestProj > proj) { closestFace = i; bestProj = proj; }
    }
  }
  return closestFace;
}

/* GuPersistentContactManifold.cpp:289-360 */
PXB_D m33 gjk_rotation_from_z(v3 to) {
  m33 m;
  const float e = to.z, f = fabsf(e);
  if (0.9999f > f) {
    const float vx = -to.y, vy = to.x;
    const float h = 1.0f / (1.0f + e);
    const float hvx = h * vx, hvxy = hvx * vy;
    m.c0 = V3(hvx * vx + e, hvxy, vy);
    m.c1 = V3(hvxy, h * (vy * vy) + e, -vx);
    m.c2 = V3(-vy, vx, e);
  } else {
    const v3 from = V3(0, 0, 1), absFrom = V3(0, 1, 0);
    const v3 u = v3sub(absFrom, from), v = v3sub(absFrom, to);
    const float dotU = adot(u, u), dotV = adot(v, v), dotUV = adot(u, v);
    const float c1 = -(2.f / dotU), c2 = -(2.f / dotV), c3 = c1 * (c2 * dotUV);
    const v3 c1u = v3scale(u, c1), c2v = v3scale(v, c2), c3v = v3scale(v, c3);
    m.c0 = v3scaleadd(u, c1u.x, v3scaleadd(v, c2v.x, v3scale(u, c3v.x))); m.c0.x = m.c0.x + 1.0f;
    m.c1 = v3scaleadd(u, c1u.y, v3scaleadd(v, c2v.y, v3scale(u, c3v.y))); m.c1.y = m.c1.y + 1.0f;
    m.c2 = v3scaleadd(u, c1u.z, v3scaleadd(v, c2v.z, v3scale(u, c3v.z))); m.c2.z = m.c2.z + 1.0f;
  }
  return m;
}

/* ---------------- capsule vs polygonal box full manifold: GuPCMContactGenSphereCapsule.cpp ---------------- */
/* :43-96 testPolyDataAxis (box: identity scaling) */
PXB_D int gjk_capbox_test_poly_axis(const GjkConvex* cap, const PolyBox* pb, float contactDist, float* minOverlap, v3* separatingAxis) {
  float _minOverlap = FLT_MAX; v3 tempAxis = V3(0, 1, 0);
  for (int i = 0; i < 6; ++i) {
    const Poly* poly = &pb->polys[i];
    const v3 minVert = pb->verts[poly->minIndex];
    const float magnitude = 1.0f / alen(poly->n);
    const v3 planeN = v3scale(poly->n, magnitude);
    const float min0 = adot(poly->n, minVert) * magnitude, max0 = (-poly->d) * magnitude;
    const float tempMin = adot(cap->p0, planeN), tempMax = adot(cap->p1, planeN);
    float min1 = fmin_(tempMin, tempMax), max1 = fmax_(tempMin, tempMax);
    min1 = min1 - cap->margin; max1 = max1 + cap->margin;
    if ((min1 > max0 + contactDist) || (min0 > max1 + contactDist)) return 0;
    const float tempOverlap = max0 - min1;
    if (_minOverlap > tempOverlap) { _minOverlap = tempOverlap; tempAxis = planeN; }
  }
  *separatingAxis = tempAxis; *minOverlap = _minOverlap;
  return 1;
}
/* :154-219 testSATCapsulePoly; map->doSupport(normal,min,max) for a box = BoxV::supportLocal(dir,min,max) GuVecBox.h:171-177 */
PXB_D int gjk_capbox_sat(const GjkConvex* cap, const PolyBox* pb, v3 ext, float contactDist, float* minOverlap, v3* separatingAxis) {
  float _minOverlap = FLT_MAX; v3 tempAxis = V3(0, 1, 0);
  if (!gjk_capbox_test_poly_axis(cap, pb, contactDist, &_minOverlap, &tempAxis)) return 0;
  const v3 capsuleAxis = v3sub(cap->p1, cap->p0);
  for (int i = 0; i < 6; ++i) {
    const uint8_t* inds = gjk_box_poly_refs + i * 4;
    for (int lStart = 0, lEnd = 3; lStart < 4; lEnd = lStart++) {
      const v3 p10 = pb->verts[inds[lStart]], p11 = pb->verts[inds[lEnd]];
      const v3 shapeSpaceV = v3sub(p11, p10);
      const v3 dir = v3cross(capsuleAxis, shapeSpaceV);
      const float lenSq = adot(dir, dir);
      if (FLT_EPSILON > lenSq) continue;
      const v3 normal = gjk_v3div(dir, sqrtf(lenSq));
      const v3 point = V3(normal.x > 0.f ? ext.x : -ext.x, normal.y > 0.f ? ext.y : -ext.y, normal.z > 0.f ? ext.z : -ext.z);
      const float max0 = adot(normal, point), min0 = -max0;
      const float tempMin = adot(cap->p0, normal), tempMax = adot(cap->p1, normal);
      float min1 = fmin_(tempMin, tempMax), max1 = fmax_(tempMin, tempMax);
      min1 = min1 - cap->margin; max1 = max1 + cap->margin;
      if ((min1 > max0 + contactDist) || (min0 > max1 + contactDist)) return 0;
      const float tempOverlap = max0 - min1;
      if (_minOverlap > tempOverlap) { _minOverlap = tempOverlap; tempAxis = normal; }
    }
  }
  *separatingAxis = tempAxis; *minOverlap = _minOverlap;
  return 1;
}
/* :221-284 generatedCapsuleBoxFaceContacts */
PXB_D void gjk_capbox_face_contacts(const GjkConvex* cap, const PolyBox* pb, int ref, const mxf* aToB, MPoint* mc, int* num, float contactDist, v3 normal) {
  const float radius = cap->margin + contactDist;
  const v3 planeNormal = anormalize(pb->polys[ref].n);
  const uint8_t* inds = gjk_box_poly_refs + ref * 4;
  const v3 a = pb->verts[inds[0]];
  const float denom0 = adot(planeNormal, v3sub(cap->p0, a)), denom1 = adot(planeNormal, v3sub(cap->p1, a));
  const float projPlaneN = adot(planeNormal, normal);
  const float numer = projPlaneN > 0.f ? 1.0f / projPlaneN : 0.f;
  const float t0 = denom0 * numer, t1 = denom1 * numer;
  const int con0 = radius >= t0, con1 = radius >= t1;
  if (con0 || con1) {
    const m33 rot = gjk_rotation_from_z(planeNormal);
    v3 pts[4]; v3 mn = V3(FLT_MAX, FLT_MAX, FLT_MAX), mx = v3neg(mn);
    for (int i = 0; i < 4; ++i) { pts[i] = m33mul(&rot, pb->verts[inds[i]]); mn = v3min(mn, pts[i]); mx = v3max(mx, pts[i]); }
    if (con0) {
      const v3 proj = v3negscalesub(normal, t0, cap->p0);
      const v3 point = m33mul(&rot, proj);
      if (poly_contains(pts, point, mn, mx)) { mc[*num].a = amxftransforminv(aToB, cap->p0); mc[*num].b = proj; mc[*num].n = normal; mc[*num].pen = t0; (*num)++; }
    }
    if (con1) {
      const v3 proj = v3negscalesub(normal, t1, cap->p1);
      const v3 point = m33mul(&rot, proj);
      if (poly_contains(pts, point, mn, mx)) { mc[*num].a = amxftransforminv(aToB, cap->p1); mc[*num].b = proj; mc[*num].n = normal; mc[*num].pen = t1; (*num)++; }
    }
  }
}
/* :312-359 generateEE */
PXB_D void gjk_capbox_ee(v3 p, v3 q, v3 normal, v3 a, v3 b, const mxf* aToB, MPoint* mc, int* num, float inflatedRadius) {
  const float expandedRatio = 0.005f;
  const v3 ab = v3sub(b, a);
  const v3 n = v3cross(ab, normal);
  const float d = adot(n, a), np = adot(n, p), nq = adot(n, q);
  const float signP = np - d, signQ = nq - d;
  if (signP * signQ > 0.f) return;
  const v3 pq = v3sub(q, p);
  const float npq = adot(n, pq);
  if (npq == 0.f) return;
  const float segT = (d - np) / npq;
  const v3 localPointA = v3scaleadd(pq, segT, p);
  const v3 perNormal = v3cross(normal, pq);
  const v3 ap = v3sub(localPointA, a);
  const float nom = adot(perNormal, ap), denom = adot(perNormal, ab);
  const float tValue = nom / denom;
  const float mx = 1.0f + expandedRatio, mn = 0.f - expandedRatio;
  if (tValue > mx || mn > tValue) return;
  const v3 v = v3negscalesub(ab, tValue, ap);
  const float signedDist = adot(v, normal);
  if (inflatedRadius >= signedDist) {
    mc[*num].a = amxftransforminv(aToB, localPointA); mc[*num].b = v3sub(localPointA, v); mc[*num].n = normal; mc[*num].pen = signedDist; (*num)++;
  }
}
/* :379-420 generateCapsuleBoxFullContactManifold (PCM_WITNESS_POINT_LOWER_EPS 1e-2, UPPER_EPS 5e-2: GuPCMContactGen.h) */
PXB_D int gjk_capbox_full_manifold(const GjkConvex* cap, const PolyBox* pb, v3 ext, const mxf* aToB, MPoint* mc, int* num, float contactDist,
                                           v3* normal, v3 closest, float margin, int doOverlapTest, float toleranceScale) {
  const int original = *num;
  int ref;
  if (doOverlapTest) {
    float minOverlap;
    if (!gjk_capbox_sat(cap, pb, ext, contactDist, &minOverlap, normal)) return 0;
    ref = gjk_box_polygon_index(pb, v3neg(*normal));
  } else {
    const float lowerEps = toleranceScale * 1e-2f, upperEps = toleranceScale * 5e-2f;
    const float tolerance = fmin_(fmax_(margin, lowerEps), upperEps);   /* PxClamp */
    ref = gjk_box_witness_polygon_index(pb, v3neg(*normal), closest, tolerance);
  }
  gjk_capbox_face_contacts(cap, pb, ref, aToB, mc, num, contactDist, *normal);
  if (*num - original < 2) {
    const uint8_t* inds = gjk_box_poly_refs + ref * 4;
    const float inflatedRadius = cap->margin + contactDist;
    for (int rStart = 0, rEnd = 3; rStart < 4; rEnd = rStart++)
      gjk_capbox_ee(cap->p0, cap->p1, *normal, pb->verts[inds[rStart]], pb->verts[inds[rEnd]], aToB, mc, num, inflatedRadius);
  }
  return 1;
}

/* GuPersistentContactManifold.cpp:1177-1243 reduceBatchContacts2 + :1292-1310 addBatchManifoldContacts2 */
PXB_D void gjk_add_batch2(Manifold* m, const MPoint* p, int numPoints) {
  if (numPoints <= 2) { for (int i = 0; i < numPoints; ++i) m->pts[i] = p[i]; m->n = numPoints; return; }
  int chosen[8]; for (int i = 0; i < 8; ++i) chosen[i] = 0;
  float maxDis = p[0].pen; int index = 0;
  for (int i = 1; i < numPoints; ++i) if (maxDis > p[i].pen) { maxDis = p[i].pen; index = i; }
  m->pts[0] = p[index]; chosen[index] = 1;
  v3 v = v3sub(p[0].b, m->pts[0].b);
  maxDis = adot(v, v); index = 0;
  for (int i = 1; i < numPoints; ++i) { v = v3sub(p[i].b, m->pts[0].b); const float d = adot(v, v); if (d > maxDis) { maxDis = d; index = i; } }
  m->pts[1] = p[index]; chosen[index] = 1;
  int secondIndex = index;
  const float maxDepth = p[index].pen;
  for (int i = 0; i < numPoints; ++i) {
    if (chosen[i]) continue;
    const v3 d0 = v3sub(m->pts[0].b, p[i].b), d1 = v3sub(m->pts[1].b, p[i].b);
    if (adot(d0, d0) > adot(d1, d1)) { if (maxDepth > p[i].pen) secondIndex = i; }
  }
  if (secondIndex != index) m->pts[1] = p[secondIndex];
  m->n = 2;
}

/* GuPersistentContactManifold.h:227-241, thresholds .cpp:179,187 */
PXB_D int gjk_invalidate_sphere_capsule(const Manifold* m, const xf* cur, float minMargin) {
  const float thr2[3] = {0.5f, 0.1f, 0.75f}, qthr2[3] = {0.9995f, 0.9999f, 0.9997f};
  const float thresholdP = minMargin * thr2[m->n];
  const float deltaP = max_pos_delta(*m, cur->p);
  const float deltaQ = adot4(cur->q, m->rel.q);
  return (deltaP > thresholdP) || (qthr2[m->n] > deltaQ);
}
/* GuPersistentContactManifold.cpp:783-807 (sphere / capsule variant: points stored on the core segment) */
PXB_D void gjk_manifold_to_contacts_radius(const Manifold* m, v3 normal, const xf* transf0, float radius, float contactOffset, Contacts* out) {
  out->count = 0; out->normal = normal;
  for (int i = 0; i < m->n; ++i) {
    const float dist = m->pts[i].pen - radius;
    if (contactOffset >= dist) { out->point[out->count] = v3negscalesub(normal, radius, axftransform(transf0, m->pts[i].a)); out->sep[out->count] = dist; out->count++; }
  }
}

/* pcmContactCapsuleBox: GuPCMContactCapsuleBox.cpp:75-202 (shape0 = capsule, shape1 = box).
 * Returns 0 normally; 1 when the reference would run EPA (segment inside the box), which is not restated yet -- the caller counts the pair as unsupported. */
PXB_D int gjk_pcm_capsule_box(const xf* transf0, const xf* transf1, float capsuleRadius, float capsuleHalfHeight, v3 boxExtents, float contactDist, float toleranceLength,
                                      Manifold* manifold, Contacts* out) {
  out->count = 0;
  const xf curRTrans = axfinvmul(transf1, transf0);
  const mxf aToB = amxffromxf(&curRTrans);
  const int initialContacts = manifold->n;
  const float boxMargin = box_margin(boxExtents, toleranceLength);
  const float minMargin = fmin_(boxMargin, capsuleRadius);
  const float projectBreakingThreshold = minMargin * 0.8f;
  manifold_refresh(*manifold, aToB, projectBreakingThreshold);
  const int bLostContacts = manifold->n != initialContacts;
  if (bLostContacts || gjk_invalidate_sphere_capsule(manifold, &curRTrans, minMargin)) {
    manifold->rel = curRTrans; manifold->dirty = 1;
    const GjkConvex box = gjk_cvx_box(transf1->p, boxExtents);
    const GjkConvex capsule = gjk_cvx_capsule(aToB.p, m33mul(&aToB.r, v3scale(V3(1, 0, 0), capsuleHalfHeight)), capsuleRadius);
    GjkOutput output; output.normal = output.closestA = output.closestB = output.searchDir = V3(0, 0, 0); output.penDep = 0.f;
    const v3 initialSearchDir = v3sub(capsule.center, box.center);
    int status = gjk_penetration(&capsule, &box, initialSearchDir, contactDist, 1, manifold->aInd, manifold->bInd, &manifold->nWarm, &output);
    MPoint mc[8]; int numContacts = 0; int doOverlapTest = 0;   // <= 1 GJK point + 2 face points + 4 edge points
    PolyBox pb; gjk_poly_box(&pb, boxExtents);
    v3 normal = output.normal;
    if (status == GJK_NON_INTERSECT) return 0;
    if (status == GJK_DEGENERATE) doOverlapTest = 1;
    else {
      if (status == GJK_CONTACT) {
        mc[numContacts].a = amxftransforminv(&aToB, output.closestA); mc[numContacts].b = output.closestB; mc[numContacts].n = output.normal; mc[numContacts].pen = output.penDep; numContacts++;
      } else {   /* EPA_CONTACT: the core segment overlaps the box */
        status = gjk_epa_penetration(&capsule, &box, manifold->aInd, manifold->bInd, manifold->nWarm, 1, toleranceLength, &output);
        if (status == EPA_CONTACT) {
          mc[numContacts].a = amxftransforminv(&aToB, output.closestA); mc[numContacts].b = output.closestB; mc[numContacts].n = output.normal; mc[numContacts].pen = output.penDep; numContacts++;
        } else doOverlapTest = 1;
        normal = output.normal;
      }
      if (!(initialContacts == 0 || bLostContacts || doOverlapTest)) {
        const float replaceBreakingThreshold = minMargin * 0.1f;
        add_manifold_point2(*manifold, aqrotinv(curRTrans.q, v3sub(output.closestA, curRTrans.p)), output.closestB, output.normal, output.penDep, replaceBreakingThreshold);
        const v3 n = aqrot(transf1->q, output.normal);
        gjk_manifold_to_contacts_radius(manifold, n, transf0, capsuleRadius, contactDist, out);
        return 0;
      }
    }
    /* fullContactsGenerationCapsuleBox :42-73 */
    const int origContacts = numContacts;
    if (!gjk_capbox_full_manifold(&capsule, &pb, boxExtents, &aToB, mc, &numContacts, contactDist, &normal, output.closestB, box.margin, doOverlapTest, toleranceLength)) return 0;
    const MPoint* mcp = mc;
    if (origContacts != 0 && numContacts != origContacts) { numContacts--; mcp++; }   /* new contacts replace the GJK one */
    gjk_add_batch2(manifold, mcp, numContacts);
    normal = aqrot(transf1->q, normal);
    gjk_manifold_to_contacts_radius(manifold, normal, transf0, capsuleRadius, contactDist, out);
    return 0;
  } else if (manifold->n > 0) {
    const v3 worldNormal = manifold_world_normal(*manifold, *transf1);
    gjk_manifold_to_contacts_radius(manifold, worldNormal, transf0, capsuleRadius, contactDist, out);
  }
  return 0;
}

/* ---------------- sphere vs convex hull: GuPCMContactSphereConvex.cpp:47-246, GuPCMContactGenSphereCapsule.cpp:43-152,470-497 ---------------- */
/* testPolyDataAxis :43-96 over a hull (identity scaling) */
PXB_D int gjk_hull_test_poly_axis(const GjkConvex* cap, const DevHull* h, float contactDist, float* minOverlap, v3* separatingAxis) {
  float _minOverlap = FLT_MAX; v3 tempAxis = V3(0, 1, 0);
  for (uint32_t i = 0; i < h->nPolys; ++i) {
    const v3 pn = h->plane_n(i);
    const v3 minVert = h->vert(h->poly_meta(i).z);
    const float magnitude = 1.0f / alen(pn);
    const v3 planeN = v3scale(pn, magnitude);
    const float min0 = adot(pn, minVert) * magnitude, max0 = (-h->plane_d(i)) * magnitude;
    const float tempMin = adot(cap->p0, planeN), tempMax = adot(cap->p1, planeN);
    float min1 = fmin_(tempMin, tempMax), max1 = fmax_(tempMin, tempMax);
    min1 = min1 - cap->margin; max1 = max1 + cap->margin;
    if ((min1 > max0 + contactDist) || (min0 > max1 + contactDist)) return 0;
    const float tempOverlap = max0 - min1;
    if (_minOverlap > tempOverlap) { _minOverlap = tempOverlap; tempAxis = planeN; }
  }
  *separatingAxis = tempAxis; *minOverlap = _minOverlap;
  return 1;
}
/* intersectRayPolyhedron :98-152 */
PXB_D int gjk_hull_ray(v3 a, v3 dir, const DevHull* h, float* tEnter, float* tExit) {
  float tFirst = 0.f, tLast = FLT_MAX;
  for (uint32_t k = 0; k < h->nPolys; ++k) {
    const v3 n = h->plane_n(k); const float d = h->plane_d(k);
    const float denominator = adot(n, dir), distToPlane = adot(n, a) + d;
    if (1e-7f > fabsf(denominator)) { if (distToPlane > 0.f) return 0; }
    else {
      const float tTemp = -(distToPlane / denominator);
      const int con = 0.f > denominator;
      if (con && tTemp > tFirst) tFirst = tTemp;
      if (!con && tLast > tTemp) tLast = tTemp;
    }
    if (tFirst > tLast) return 0;
  }
  *tEnter = tFirst; *tExit = tLast;
  return 1;
}
PXB_D void gjk_pcm_sphere_convex(const xf* transf0, const xf* transf1, float sphereRadius, const DevHull* hull, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out) {
  out->count = 0;
  const xf curRTrans = axfinvmul(transf1, transf0);
  const mxf aToB = amxffromxf(&curRTrans);
  const float convexMargin = gjk_hull_pcm_margin(hull, toleranceLength);
  const int initialContacts = manifold->n;
  const float minMargin = fmin_(convexMargin, sphereRadius);
  manifold_refresh(*manifold, aToB, minMargin * 0.05f);
  const int bLostContacts = manifold->n != initialContacts;
  if (bLostContacts || gjk_invalidate_sphere_capsule(manifold, &curRTrans, minMargin)) {
    manifold->rel = curRTrans; manifold->dirty = 1;
    const GjkConvex convexHull = gjk_cvx_hull(hull);
    const GjkConvex capsule = gjk_cvx_capsule(aToB.p, V3(0, 0, 0), sphereRadius);   /* CapsuleV(p, radius): p0 = p1 = p */
    GjkOutput output; output.normal = output.closestA = output.closestB = output.searchDir = V3(0, 0, 0); output.penDep = 0.f;
    const v3 initialSearchDir = v3sub(capsule.center, convexHull.center);
    int status = gjk_penetration(&capsule, &convexHull, initialSearchDir, contactDist, 1, manifold->aInd, manifold->bInd, &manifold->nWarm, &output);
    if (status == GJK_NON_INTERSECT) return;
    if (status == EPA_CONTACT) {
      status = gjk_epa_penetration(&capsule, &convexHull, manifold->aInd, manifold->bInd, manifold->nWarm, 1, toleranceLength, &output);
      if (status != EPA_CONTACT) status = GJK_DEGENERATE;   /* EPA failed: full contact generation with the overlap test, like a degenerate GJK */
      else status = GJK_CONTACT;
    }
    if (status == GJK_CONTACT) {
      manifold->pts[0].a = V3(0, 0, 0); manifold->pts[0].b = output.closestB; manifold->pts[0].n = output.normal; manifold->pts[0].pen = output.penDep; manifold->n = 1;
      const v3 worldNormal = aqrot(transf1->q, output.normal);
      out->normal = worldNormal; out->point[0] = v3negscalesub(worldNormal, sphereRadius, transf0->p); out->sep[0] = output.penDep - sphereRadius; out->count = 1;
      return;
    }
    if (status == GJK_DEGENERATE) {   /* fullContactsGenerationSphereConvex :47-82 with doOverlapTest = true */
      v3 normal = output.normal; float minOverlap;
      if (!gjk_hull_test_poly_axis(&capsule, hull, contactDist, &minOverlap, &normal)) return;
      float tEnter = 0.f, tExit = 0.f;
      const float inflatedRadius = sphereRadius + contactDist;
      const v3 dir = v3neg(normal);
      if (gjk_hull_ray(capsule.p0, dir, hull, &tEnter, &tExit) && inflatedRadius >= tEnter) {
        manifold->pts[0].a = V3(0, 0, 0); manifold->pts[0].b = v3scaleadd(dir, tEnter, capsule.p0); manifold->pts[0].n = normal; manifold->pts[0].pen = tEnter; manifold->n = 1;
        const v3 worldNormal = aqrot(transf1->q, normal);
        out->normal = worldNormal; out->point[0] = v3negscalesub(worldNormal, sphereRadius, transf0->p); out->sep[0] = tEnter - sphereRadius; out->count = 1;
      }
    }
  } else if (manifold->n > 0) {
    const v3 worldNormal = aqrot(transf1->q, manifold->pts[0].n);
    out->normal = worldNormal; out->point[0] = v3negscalesub(worldNormal, sphereRadius, transf0->p); out->sep[0] = manifold->pts[0].pen - sphereRadius; out->count = 1;
  }
}


/* ---------------- capsule vs convex hull: GuPCMContactCapsuleConvex.cpp:42-262, GuPCMContactGenSphereCapsule.cpp:154-468, GuPCMContactGenUtil.cpp:105-260 ---------------- */
/* ConvexHullNoScaleV::bruteForceSearchMinMax GuVecConvexHullNoScale.h:116-136 (SupportLocalImpl::doSupport) */
PXB_D void gjk_hull_support_minmax(const DevHull* h, v3 dir, float* mn, float* mx) {
  if (h->bigSubdiv) {   /* ConvexHullNoScaleV::supportVertexMinMax GuVecConvexHullNoScale.h:139-155: two hill climbs */
    const uint32_t maxIndex = gjk_hull_hill_climb(h, dir), minIndex = gjk_hull_hill_climb(h, v3neg(dir));
    *mn = adot(dir, h->vert(minIndex)); *mx = adot(dir, h->vert(maxIndex));
    return;
  }
  float _max = adot(h->vert(0), dir), _min = _max;
  for (uint32_t i = 1; i < h->nVerts; ++i) { const float d = adot(h->vert(i), dir); _max = d > _max ? d : _max; _min = d < _min ? d : _min; }   /* FMax / FMin */
  *mn = _min; *mx = _max;
}
/* testSATCapsulePoly :154-219 */
PXB_D int gjk_hull_sat_capsule(const GjkConvex* cap, const DevHull* h, float contactDist, float* minOverlap, v3* separatingAxis) {
  float _minOverlap = FLT_MAX; v3 tempAxis = V3(0, 1, 0);
  if (!gjk_hull_test_poly_axis(cap, h, contactDist, &_minOverlap, &tempAxis)) return 0;
  const v3 capsuleAxis = v3sub(cap->p1, cap->p0);
  for (uint32_t i = 0; i < h->nPolys; ++i) {
    const uint8_t* inds = h->vertexRefs + h->poly_meta(i).x; const uint32_t nb = h->poly_meta(i).y;
    for (uint32_t lStart = 0, lEnd = nb - 1; lStart < nb; lEnd = lStart++) {
      const v3 p10 = h->vert(inds[lStart]), p11 = h->vert(inds[lEnd]);
      const v3 dir = v3cross(capsuleAxis, v3sub(p11, p10));
      const float lenSq = adot(dir, dir);
      if (FLT_EPSILON > lenSq) continue;
      const v3 normal = gjk_v3div(dir, sqrtf(lenSq));
      float min0, max0; gjk_hull_support_minmax(h, normal, &min0, &max0);
      const float tempMin = adot(cap->p0, normal), tempMax = adot(cap->p1, normal);
      float min1 = fmin_(tempMin, tempMax), max1 = fmax_(tempMin, tempMax);
      min1 = min1 - cap->margin; max1 = max1 + cap->margin;
      if ((min1 > max0 + contactDist) || (min0 > max1 + contactDist)) return 0;
      const float tempOverlap = max0 - min1;
      if (_minOverlap > tempOverlap) { _minOverlap = tempOverlap; tempAxis = normal; }
    }
  }
  *separatingAxis = tempAxis; *minOverlap = _minOverlap;
  return 1;
}
/* generatedFaceContacts :286-310 */
PXB_D void gjk_hull_capsule_face_contacts(const GjkConvex* cap, const DevHull* h, const mxf* aToB, MPoint* mc, int* num, float contactDist, v3 normal) {
  float tEnter = 0.f, tExit = 0.f;
  const float inflatedRadius = cap->margin + contactDist;
  const v3 dir = v3neg(normal);
  if (gjk_hull_ray(cap->p0, dir, h, &tEnter, &tExit) && inflatedRadius >= tEnter) { mc[*num].a = amxftransforminv(aToB, cap->p0); mc[*num].b = v3scaleadd(dir, tEnter, cap->p0); mc[*num].n = normal; mc[*num].pen = tEnter; (*num)++; }
  if (gjk_hull_ray(cap->p1, dir, h, &tEnter, &tExit) && inflatedRadius >= tEnter) { mc[*num].a = amxftransforminv(aToB, cap->p1); mc[*num].b = v3scaleadd(dir, tEnter, cap->p1); mc[*num].n = normal; mc[*num].pen = tEnter; (*num)++; }
}
/* getPolygonIndex GuPCMContactGenUtil.cpp:105-202 (identity scaling; hulls carry their edge -> faces table) */
PXB_D int gjk_hull_polygon_index(const DevHull* h, v3 normal) {
  const v3 n = normal, nnormal = v3neg(n);
  float minProj = adot(n, h->plane_n(0));
  int closestFaceIndex = 0;
  for (uint32_t i = 1; i < h->nPolys; ++i) { const float proj = adot(n, h->plane_n(i)); if (minProj > proj) { minProj = proj; closestFaceIndex = (int)i; } }
  uint32_t closestEdge = 0xffffffffu;
  float maxDpSq = minProj * minProj;
  for (uint32_t i = 0; i < h->nEdges; ++i) {
    const uint8_t f0 = h->facesByEdges[i * 2], f1 = h->facesByEdges[i * 2 + 1];
    const v3 edgeNormal = v3add(h->plane_n(f0), h->plane_n(f1));
    const float enMagSq = adot(edgeNormal, edgeNormal), dp = adot(edgeNormal, nnormal), sqDp = dp * dp;
    if (dp >= 0.f && sqDp > maxDpSq * enMagSq) { maxDpSq = sqDp / enMagSq; closestEdge = i; }
  }
  if (closestEdge != 0xffffffffu) {
    const uint32_t f0 = h->facesByEdges[closestEdge * 2], f1 = h->facesByEdges[closestEdge * 2 + 1];
    const float dp0 = adot(h->plane_n(f0), nnormal), dp1 = adot(h->plane_n(f1), nnormal);
    closestFaceIndex = dp0 > dp1 ? (int)f0 : (int)f1;
  }
  return closestFaceIndex;
}
/* getWitnessPolygonIndex GuPCMContactGenUtil.cpp:204-260 */
PXB_D int gjk_hull_witness_polygon_index(const DevHull* h, v3 normal, v3 closest, float tolerance) {
  float pd[64];   // GPU-compatible hulls have <= 64 polygons (pxb_scene_set_convex_meshes enforces the limit)
  const float eps = -tolerance;
  float dist = v3dot(closest, h->plane_n(0)) + h->plane_d(0);
  float minDist = dist >= eps ? fabsf(dist) : FLT_MAX;
  pd[0] = minDist;
  float maxDist = dist; int maxFace = 0, closestFace = 0;
  for (uint32_t i = 1; i < h->nPolys; ++i) {
    dist = v3dot(closest, h->plane_n(i)) + h->plane_d(i);
    pd[i] = dist >= eps ? fabsf(dist) : FLT_MAX;
    if (minDist > pd[i]) { minDist = pd[i]; closestFace = (int)i; }
    if (dist > maxDist) { maxDist = dist; maxFace = (int)i; }
  }
  if (minDist == FLT_MAX) return maxFace;
  float bestProj = adot(anormalize(h->plane_n(closestFace)), normal);
  const int first = closestFace;
  for (uint32_t i = 0; i < h->nPolys; ++i) {
    if ((tolerance > (pd[i] - minDist)) && first != (int)i) {
      const float proj = adot(anormalize(h->plane_n(i)), normal);
      if (bestProj > proj) { closestFace = (int)i; bestProj = proj; }
    }
  }
  return closestFace;
}
/* generatedContactsEEContacts :361-377 */
PXB_D void gjk_hull_capsule_ee_contacts(const GjkConvex* cap, const DevHull* h, int ref, const mxf* aToB, MPoint* mc, int* num, float contactDist, v3 normal) {
  const uint8_t* inds = h->vertexRefs + h->poly_meta(ref).x; const uint32_t nb = h->poly_meta(ref).y;
  const float inflatedRadius = cap->margin + contactDist;
  for (uint32_t rStart = 0, rEnd = nb - 1; rStart < nb; rEnd = rStart++)
    gjk_capbox_ee(cap->p0, cap->p1, normal, h->vert(inds[rStart]), h->vert(inds[rEnd]), aToB, mc, num, inflatedRadius);
}
/* generateFullContactManifold(capsule, polyData, ...) :423-468 */
PXB_D int gjk_hull_capsule_full_manifold(const GjkConvex* cap, const DevHull* h, const mxf* aToB, MPoint* mc, int* num, float contactDist, v3* normal, v3 closest, float margin,
                                                 int doOverlapTest, float toleranceLength) {
  const int original = *num;
  v3 tNormal = *normal;
  if (doOverlapTest) {
    float minOverlap;
    if (!gjk_hull_sat_capsule(cap, h, contactDist, &minOverlap, &tNormal)) return 0;
    gjk_hull_capsule_face_contacts(cap, h, aToB, mc, num, contactDist, tNormal);
    if (*num - original < 2) gjk_hull_capsule_ee_contacts(cap, h, gjk_hull_polygon_index(h, v3neg(tNormal)), aToB, mc, num, contactDist, tNormal);
  } else {
    gjk_hull_capsule_face_contacts(cap, h, aToB, mc, num, contactDist, tNormal);
    if (*num - original < 2) {
      const float lowerEps = toleranceLength * 1e-2f, upperEps = toleranceLength * 5e-2f;
      const float tolerance = fmin_(fmax_(margin, lowerEps), upperEps);
      gjk_hull_capsule_ee_contacts(cap, h, gjk_hull_witness_polygon_index(h, v3neg(tNormal), closest, tolerance), aToB, mc, num, contactDist, tNormal);
    }
  }
  *normal = tNormal;
  return 1;
}
/* pcmContactCapsuleConvex :80-262 (shape0 = capsule, shape1 = convex mesh, identity mesh scale), in the three phases the kernels run separately
 * (k_gjk_refresh -> k_gjk_query -> k_gjk_manifold: every phase's warps hold only pairs that need that phase):
 *   refresh  manifold refresh + invalidation test; a manifold that stays valid emits its cached points and the pair is finished
 *   query    GJK (+ EPA) on the invalidated pairs; separated pairs and pairs that only add the GJK point are finished
 *   manifold fullContactsGenerationCapsuleConvex for the rest
 * What a later phase needs from an earlier one travels in GjkCarry (+ the manifold record itself). */
struct GjkCarry { v3 normal, closestA, closestB; int doOverlapTest; };
PXB_D int gjk_capsule_convex_refresh(const xf* transf0, const xf* transf1, float capsuleRadius, const DevHull* hull, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out, int* flags) {
  out->count = 0;
  const xf curRTrans = axfinvmul(transf1, transf0);
  const mxf aToB = amxffromxf(&curRTrans);
  const float convexMargin = gjk_hull_pcm_margin(hull, toleranceLength);
  const float minMargin = fmin_(convexMargin, capsuleRadius * 0.05f);   /* CalculateCapsuleMinMargin GuVecCapsule.h:42-47 */
  const int initialContacts = manifold->n;
  manifold_refresh(*manifold, aToB, minMargin * 1.25f);
  const int bLostContacts = manifold->n != initialContacts;
  if (bLostContacts || gjk_invalidate_sphere_capsule(manifold, &curRTrans, minMargin)) {
    manifold->rel = curRTrans; manifold->dirty = 1;
    *flags = initialContacts | (bLostContacts << 8);
    return 1;
  }
  if (manifold->n > 0) {
    const v3 worldNormal = manifold_world_normal(*manifold, *transf1);
    gjk_manifold_to_contacts_radius(manifold, worldNormal, transf0, capsuleRadius, contactDist, out);
  }
  return 0;
}
/* the two GJK shapes of the pair (capsule in the hull's frame) and the initial search direction */
PXB_D void gjk_capsule_convex_shapes(const xf* transf0, const xf* transf1, float capsuleRadius, float capsuleHalfHeight, const DevHull* hull, GjkConvex* capsule, GjkConvex* convexHull, mxf* aToB, v3* initialSearchDir) {
  const xf curRTrans = axfinvmul(transf1, transf0);
  *aToB = amxffromxf(&curRTrans);
  *convexHull = gjk_cvx_hull(hull);
  *capsule = gjk_cvx_capsule(aToB->p, m33mul(&aToB->r, v3scale(V3(1, 0, 0), capsuleHalfHeight)), capsuleRadius);
  *initialSearchDir = v3sub(capsule->center, convexHull->center);
}
/* what pcmContactCapsuleConvex does with the query's answer.  status: gjk_penetration's (never GJK_NON_INTERSECT here); epaStatus: gjk_epa_penetration's when
 * status == EPA_CONTACT.  Returns 1 when the full manifold generation has to run (carry filled), 0 when the pair is finished. */
PXB_D int gjk_capsule_convex_post(const xf* transf0, const xf* transf1, float capsuleRadius, const DevHull* hull, float contactDist, float toleranceLength, int flags, int status, int epaStatus,
                                  const GjkOutput* output, const mxf* aToB, Manifold* manifold, Contacts* out, GjkCarry* carry) {
  const int initialContacts = flags & 0xff, bLostContacts = flags >> 8;
  const float minMargin = fmin_(gjk_hull_pcm_margin(hull, toleranceLength), capsuleRadius * 0.05f);
  int doOverlapTest = 0;
  if (status == GJK_DEGENERATE) doOverlapTest = 1;
  else {
    const float replaceBreakingThreshold = minMargin * 0.05f;
    if (status == EPA_CONTACT && epaStatus != EPA_CONTACT) doOverlapTest = 1;
    if (!doOverlapTest) add_manifold_point2(*manifold, amxftransforminv(aToB, output->closestA), output->closestB, output->normal, output->penDep, replaceBreakingThreshold);
    if (!(initialContacts == 0 || bLostContacts || doOverlapTest)) {
      const v3 n = aqrot(transf1->q, output->normal);
      gjk_manifold_to_contacts_radius(manifold, n, transf0, capsuleRadius, contactDist, out);
      return 0;
    }
  }
  carry->normal = output->normal; carry->closestA = output->closestA; carry->closestB = output->closestB; carry->doOverlapTest = doOverlapTest;
  return 1;
}
PXB_D int gjk_capsule_convex_query(const xf* transf0, const xf* transf1, float capsuleRadius, float capsuleHalfHeight, const DevHull* hull, float contactDist, float toleranceLength, int flags,
                                   Manifold* manifold, Contacts* out, GjkCarry* carry) {
  out->count = 0;
  GjkConvex capsule, convexHull; mxf aToB; v3 initialSearchDir;
  gjk_capsule_convex_shapes(transf0, transf1, capsuleRadius, capsuleHalfHeight, hull, &capsule, &convexHull, &aToB, &initialSearchDir);
  GjkOutput output; output.normal = output.closestA = output.closestB = output.searchDir = V3(0, 0, 0); output.penDep = 0.f;
  const int status = gjk_penetration(&capsule, &convexHull, initialSearchDir, contactDist, 1, manifold->aInd, manifold->bInd, &manifold->nWarm, &output);
  if (status == GJK_NON_INTERSECT) return 0;
  int epaStatus = 0;
  if (status == EPA_CONTACT) epaStatus = gjk_epa_penetration(&capsule, &convexHull, manifold->aInd, manifold->bInd, manifold->nWarm, 1, toleranceLength, &output);
  return gjk_capsule_convex_post(transf0, transf1, capsuleRadius, hull, contactDist, toleranceLength, flags, status, epaStatus, &output, &aToB, manifold, out, carry);
}
/* fullContactsGenerationCapsuleConvex :42-78 */
PXB_D void gjk_capsule_convex_manifold(const xf* transf0, const xf* transf1, float capsuleRadius, float capsuleHalfHeight, const DevHull* hull, float contactDist, float toleranceLength,
                                       const GjkCarry* carry, Manifold* manifold, Contacts* out) {
  out->count = 0;
  const xf curRTrans = axfinvmul(transf1, transf0);
  const mxf aToB = amxffromxf(&curRTrans);
  const GjkConvex convexHull = gjk_cvx_hull(hull);
  const GjkConvex capsule = gjk_cvx_capsule(aToB.p, m33mul(&aToB.r, v3scale(V3(1, 0, 0), capsuleHalfHeight)), capsuleRadius);
  MPoint mc[16]; int numContacts = 0; const int doOverlapTest = carry->doOverlapTest;   // <= 2 face points + one edge-edge point per polygon edge that the segment crosses (2 for a convex polygon)
  v3 normal = carry->normal;
  if (!gjk_hull_capsule_full_manifold(&capsule, hull, &aToB, mc, &numContacts, contactDist, &normal, carry->closestB, convexHull.margin, doOverlapTest, toleranceLength)) return;
  if (numContacts > 0) {
    gjk_add_batch2(manifold, mc, numContacts);
    normal = aqrot(transf1->q, normal);
    gjk_manifold_to_contacts_radius(manifold, normal, transf0, capsuleRadius, contactDist, out);
  } else if (!doOverlapTest) {
    normal = aqrot(transf1->q, normal);
    gjk_manifold_to_contacts_radius(manifold, normal, transf0, capsuleRadius, contactDist, out);
  }
}
PXB_D void gjk_pcm_capsule_convex(const xf* transf0, const xf* transf1, float capsuleRadius, float capsuleHalfHeight, const DevHull* hull, float contactDist, float toleranceLength,
                                          Manifold* manifold, Contacts* out) {   // the three phases back to back (single-kernel variant)
  int flags; GjkCarry carry;
  if (!gjk_capsule_convex_refresh(transf0, transf1, capsuleRadius, hull, contactDist, toleranceLength, manifold, out, &flags)) return;
  if (!gjk_capsule_convex_query(transf0, transf1, capsuleRadius, capsuleHalfHeight, hull, contactDist, toleranceLength, flags, manifold, out, &carry)) return;
  gjk_capsule_convex_manifold(transf0, transf1, capsuleRadius, capsuleHalfHeight, hull, contactDist, toleranceLength, &carry, manifold, out);
}

/* ---------------- polygonal pairs: box vs hull, hull vs hull (GuPCMContactGenBoxConvex.cpp:331-720, GuPCMContactBoxConvex.cpp, GuPCMContactConvexConvex.cpp) ---------------- */
// Gu::contains GuPCMContactGenUtil.cpp:35-103 for an n-vertex polygon (poly_contains in pxb_np.cuh is the 4-vertex instance)
PXB_D bool gjk_poly_contains_n(const v3* verts, int numVerts, v3 p, v3 mn, v3 mx) {
  if ((mn.x > p.x) || (p.x > mx.x) || (mn.y > p.y) || (p.y > mx.y)) return false;
  const float tx = p.x, ty = p.y; const float eps = FLT_EPSILON;
  int inter = 0;
  for (int i = 0, j = numVerts - 1; i < numVerts; j = i++) {
    const float jy = verts[j].y, iy = verts[i].y, jx = verts[j].x, ix = verts[i].x;
    if ((tx == jx && ty == jy) || (tx == ix && ty == iy)) return true;
    const bool yflag0 = jy > ty, yflag1 = iy > ty;
    if (yflag0 != yflag1) {
      const float jix = ix - jx, jiy = iy - jy, jty = ty - jy;
      const float part1 = jty * jix, part2 = (jx + eps) * jiy, part3 = tx * jiy;
      const bool comp = jiy > 0.f;
      const float tmp = part1 + part2;
      const float comp1 = comp ? tmp : part3, comp2 = comp ? part3 : tmp;
      if (comp1 >= comp2) { if (inter == 1) return false; inter++; }
    }
  }
  return inter > 0;
}
/* ---------------- SAT branch of generateFullContactManifold: GuPCMContactGenBoxConvex.cpp:56-328, :537-603 (PCM_USE_INTERNAL_OBJECT = 1) ---------------- */
#define GJK_SAT_MAX_AXES 128   // per-thread; a GPU-compatible hull (<= 64 vertices, <= 64 polygons) has <= 126 edges (the reference's SEP_AXIS_FIXED_MEMORY is 256)
/* SupportLocalImpl::doSupport(dir, min, max) / doSupport(dir): hull = brute force (GuVecConvexHull.h:377-429), box = sign select (GuVecBox.h:165-177) */
PXB_D void gjk_poly_support_minmax(const DevHull* h, int isBox, v3 dir, float* mn, float* mx) {
  if (isBox) { const v3 e = h->internalExtents; const v3 pt = V3(dir.x > 0.f ? e.x : -e.x, dir.y > 0.f ? e.y : -e.y, dir.z > 0.f ? e.z : -e.z); *mx = adot(dir, pt); *mn = -*mx; return; }
  gjk_hull_support_minmax(h, dir, mn, mx);
}
PXB_D v3 gjk_poly_support(const DevHull* h, int isBox, v3 dir) {
  if (isBox) { const v3 e = h->internalExtents; return V3(dir.x > 0.f ? e.x : -e.x, dir.y > 0.f ? e.y : -e.y, dir.z > 0.f ? e.z : -e.z); }
  return h->vert(gjk_hull_support_index(h, dir));
}
typedef struct { const DevHull* h; int isBox; v3 center; float internalRadius; v3 internalExtents; } GjkPolyData;   /* PolygonalData: mCenter, mInternal */
PXB_D GjkPolyData gjk_poly_data(const DevHull* h, int isBox) {
  GjkPolyData p; p.h = h; p.isBox = isBox;
  p.center = isBox ? V3(0, 0, 0) : h->centerOfMass; p.internalRadius = isBox ? 0.f : h->internalRadius; p.internalExtents = h->internalExtents;
  return p;
}
/* testFaceNormal :56-156 */
PXB_D int gjk_sat_face_normal(const GjkPolyData* p0, const GjkPolyData* p1, const mxf* transform0To1, const mxf* transform1To0, float contactDist,
                                      float* minOverlap, uint32_t* feature, v3* faceNormal, int faceStatus, int* status) {
  float _minOverlap = FLT_MAX; uint32_t _feature = 0; v3 _faceNormal = *faceNormal;
  const v3 center1To0 = transform1To0->p;
  const v3 internalCenter1In0 = amxftransform(transform1To0, p1->center);
  const v3 ie1 = p1->internalExtents;
  for (uint32_t i = 0; i < p0->h->nPolys; ++i) {
    const v3 pn = p0->h->plane_n(i);
    const v3 minVert = p0->h->vert(p0->h->poly_meta(i).z);
    const float magnitude = 1.0f / alen(pn);
    const float min0 = adot(pn, minVert) * magnitude, max0 = (-p0->h->plane_d(i)) * magnitude;
    const v3 n0 = v3scale(pn, magnitude);
    const v3 n1 = m33mul(&transform0To1->r, n0);
    const v3 proj = V3(n1.x > 0.f ? ie1.x : -ie1.x, n1.y > 0.f ? ie1.y : -ie1.y, n1.z > 0.f ? ie1.z : -ie1.z);
    const float radius = fmax_(adot(n1, proj), p1->internalRadius);
    const float internalTrans = adot(internalCenter1In0, n0);
    const float _min1 = internalTrans - radius, _max1 = internalTrans + radius;
    const float _min = fmax_(min0, _min1), _max = fmin_(max0, _max1);
    if ((_max - _min) > _minOverlap) continue;
    const float translate = adot(center1To0, n0);
    float min1, max1; gjk_poly_support_minmax(p1->h, p1->isBox, n1, &min1, &max1);
    min1 = translate + min1; max1 = translate + max1;
    if ((min1 > max0 + contactDist) || (min0 > max1 + contactDist)) return 0;
    const float tempOverlap = max0 - min1;
    if (_minOverlap > tempOverlap) { _minOverlap = tempOverlap; _feature = i; _faceNormal = n0; }
  }
  if (*minOverlap > _minOverlap) { *faceNormal = _faceNormal; *minOverlap = _minOverlap; *status = faceStatus; }
  *feature = _feature;
  return 1;
}
/* buildPartialHull :159-193 + SeparatingAxes::addAxis GuSeparatingAxes.cpp:33-57 */
PXB_D void gjk_sat_partial_hull(const DevHull* h, v3* axes, uint32_t* nAxes, v3 planeP, v3 planeDir) {
  const v3 dir = anormalize(planeDir);
  for (uint32_t i = 0; i < h->nPolys; ++i) {
    const uint8_t* inds = h->vertexRefs + h->poly_meta(i).x; const uint32_t nb = h->poly_meta(i).y;
    v3 v0 = h->vert(inds[nb - 1]);
    float dist0 = adot(dir, v3sub(v0, planeP));
    for (uint32_t k = 0; k < nb; ++k) {
      const v3 v1 = h->vert(inds[k]);
      const float dist1 = adot(dir, v3sub(v1, planeP));
      if (dist0 > 0.f || dist1 > 0.f) {
        const v3 t = v3sub(v0, v1);
        const float m = t.x * t.x + t.y * t.y + t.z * t.z;                       /* PxVec3::getNormalized */
        const v3 axis = m > 0.f ? v3scale(t, 1.0f / sqrtf(m)) : V3(0, 0, 0);
        int dup = 0;
        for (uint32_t a = 0; a < *nAxes; ++a) if (fabsf(v3dot(axis, axes[a])) > 0.9999f) { dup = 1; break; }
        if (!dup && *nAxes < GJK_SAT_MAX_AXES) axes[(*nAxes)++] = axis;
      }
      v0 = v1; dist0 = dist1;
    }
  }
}
/* testEdgeNormal :195-328 */
PXB_D int gjk_sat_edge_normal(const GjkPolyData* p0, const GjkPolyData* p1, const mxf* transform0To1, const mxf* transform1To0, float contactDist,
                                      float* minOverlap, v3* edgeNormalIn0, int edgeStatus, int* status) {
  float overlap = *minOverlap;
  const v3 internalCenter1In0 = v3sub(amxftransform(transform1To0, p1->center), p0->center);
  const v3 ie1 = p1->internalExtents, ie0 = p0->internalExtents;
  const v3 center1To0 = transform1To0->p;
  const v3 dir0 = v3sub(amxftransform(transform1To0, p1->center), p0->center);
  const v3 support0 = gjk_poly_support(p0->h, p0->isBox, dir0);
  const v3 dir1 = m33mul(&transform0To1->r, v3neg(dir0));
  const v3 support1 = gjk_poly_support(p1->h, p1->isBox, dir1);
  const v3 support0In1 = amxftransform(transform0To1, support0), support1In0 = amxftransform(transform1To0, support1);
  v3 axe0[GJK_SAT_MAX_AXES], axe1[GJK_SAT_MAX_AXES]; uint32_t numAxe0 = 0, numAxe1 = 0;
  gjk_sat_partial_hull(p0->h, axe0, &numAxe0, support1In0, dir0);
  gjk_sat_partial_hull(p1->h, axe1, &numAxe1, support0In1, dir1);
  for (uint32_t i = 0; i < numAxe0; ++i) {
    const v3 v0 = axe0[i];
    for (uint32_t j = 0; j < numAxe1; ++j) {
      const v3 dir = v3cross(v0, m33mul(&transform1To0->r, axe1[j]));
      const float lenSq = adot(dir, dir);
      if (FLT_EPSILON > lenSq) continue;
      const v3 n0 = v3scale(dir, 1.0f / sqrtf(lenSq));
      const v3 n1 = m33mul(&transform0To1->r, n0);
      const v3 proj = V3(n1.x > 0.f ? ie1.x : -ie1.x, n1.y > 0.f ? ie1.y : -ie1.y, n1.z > 0.f ? ie1.z : -ie1.z);
      const float radius = fmax_(adot(n1, proj), p1->internalRadius);
      const float internalTrans = adot(internalCenter1In0, n0);
      const float _min1 = internalTrans - radius, _max1 = internalTrans + radius;
      const v3 proj0 = V3(n0.x > 0.f ? ie0.x : -ie0.x, n0.y > 0.f ? ie0.y : -ie0.y, n0.z > 0.f ? ie0.z : -ie0.z);
      const float radius0 = fmax_(adot(n0, proj0), p0->internalRadius);
      const float _min = fmax_(-radius0, _min1), _max = fmin_(radius0, _max1);
      if ((_max - _min) > overlap) continue;
      float min0, max0, min1, max1;
      gjk_poly_support_minmax(p0->h, p0->isBox, n0, &min0, &max0);
      const float translate = adot(center1To0, n0);
      gjk_poly_support_minmax(p1->h, p1->isBox, n1, &min1, &max1);
      min1 = translate + min1; max1 = translate + max1;
      if ((min1 > max0 + contactDist) || (min0 > max1 + contactDist)) return 0;
      const float tempOverlap = max0 - min1;
      if (overlap > tempOverlap) { overlap = tempOverlap; *edgeNormalIn0 = n0; *status = edgeStatus; }
    }
  }
  *minOverlap = overlap;
  return 1;
}
/* generateFullContactManifold, doOverlapTest == true (:537-603).  Returns 0 when a separating axis was found. */
enum { GJK_FS_POLYDATA0 = 0, GJK_FS_POLYDATA1 = 1, GJK_FS_EDGE = 2 };
PXB_D void gjk_poly_generated_contacts(const DevHull* poly0, const DevHull* poly1, int refIdx, int incIdx, const mxf* transform0To1, MPoint* mc, int* numContacts, float contactDist);
PXB_D int gjk_poly_full_manifold_sat(const DevHull* poly0, int isBox0, const DevHull* poly1, const xf* map0, const xf* map1, MPoint* mc, int* numContacts, float contactDist) {
  /* SupportLocal::transform is a PxTransformV (GuConvexSupportTable.h:55): quaternion transformInv, then the conversion PxMatTransformV(PxTransformV) */
  const xf t10 = axfinvmul(map0, map1), t01 = axfinvmul(map1, map0);
  const mxf transform1To0 = amxffromxf(&t10), transform0To1 = amxffromxf(&t01);
  const GjkPolyData p0 = gjk_poly_data(poly0, isBox0), p1 = gjk_poly_data(poly1, 0);
  int status = GJK_FS_POLYDATA0;
  float minOverlap = FLT_MAX; v3 minNormal = V3(0, 0, 0);
  uint32_t feature0, feature1;
  if (!gjk_sat_face_normal(&p0, &p1, &transform0To1, &transform1To0, contactDist, &minOverlap, &feature0, &minNormal, GJK_FS_POLYDATA0, &status)) return 0;
  if (!gjk_sat_face_normal(&p1, &p0, &transform1To0, &transform0To1, contactDist, &minOverlap, &feature1, &minNormal, GJK_FS_POLYDATA1, &status)) return 0;
  int doEdgeTest = 0;
  for (;;) {
    if (doEdgeTest) {
      if (!gjk_sat_edge_normal(&p0, &p1, &transform0To1, &transform1To0, contactDist, &minOverlap, &minNormal, GJK_FS_EDGE, &status)) return 0;
      if (status != GJK_FS_EDGE) return 1;
    }
    if (status == GJK_FS_POLYDATA0) {
      const v3 n = m33mul(&transform0To1.r, minNormal);
      gjk_poly_generated_contacts(poly0, poly1, (int)feature0, gjk_hull_polygon_index(poly1, n), &transform0To1, mc, numContacts, contactDist);
      if (*numContacts > 0) { const v3 nn = v3neg(n); for (int i = 0; i < *numContacts; ++i) { const v3 lb = mc[i].b; mc[i].b = mc[i].a; mc[i].a = lb; mc[i].n = nn; } }
    } else if (status == GJK_FS_POLYDATA1) {
      gjk_poly_generated_contacts(poly1, poly0, (int)feature1, gjk_hull_polygon_index(poly0, m33mul(&transform1To0.r, minNormal)), &transform1To0, mc, numContacts, contactDist);
    } else {
      const int incident = gjk_hull_polygon_index(poly0, v3neg(minNormal));
      const int reference = gjk_hull_polygon_index(poly1, m33mul(&transform0To1.r, minNormal));
      gjk_poly_generated_contacts(poly1, poly0, reference, incident, &transform1To0, mc, numContacts, contactDist);
    }
    if (*numContacts == 0 && !doEdgeTest) { doEdgeTest = 1; continue; }
    break;
  }
  return 1;
}


PXB_D float gjk_signed_2d_tri_area(v3 a, v3 b, v3 c) { const v3 ca = v3sub(a, c), cb = v3sub(b, c); return ca.x * cb.y - ca.y * cb.x; }   /* GuPCMContactGenUtil.h:56-66 */
#define GJK_POLY_MAX_CONTACTS 32   // per-thread buffer; a pair of <= 32-vertex hull polygons stays far below it (the reference's buffer holds 256)
/* generatedContacts :331-530: incident polygon (of poly1) clipped against the reference polygon (of poly0) in the reference polygon's plane */
PXB_D void gjk_poly_generated_contacts(const DevHull* poly0, const DevHull* poly1, int refIdx, int incIdx, const mxf* transform0To1, MPoint* mc, int* numContacts, float contactDist) {
  const uint4 referencePolygon = poly0->poly_meta((uint32_t)refIdx), incidentPolygon = poly1->poly_meta((uint32_t)incIdx);
  const uint8_t* inds0 = poly0->vertexRefs + referencePolygon.x; const uint8_t* inds1 = poly1->vertexRefs + incidentPolygon.x;
  const uint32_t nRef = min(referencePolygon.y, 32u), nInc = min(incidentPolygon.y, 32u);
  const v3 contactNormal = anormalize(poly0->plane_n((uint32_t)refIdx));
  const m33 rot = gjk_rotation_from_z(contactNormal);
  v3 points0In0[32], points1In0[32]; int pen1[32]; float tval1[32];
  for (uint32_t i = 0; i < nRef; ++i) points0In0[i] = poly0->vert(inds0[i]);
  for (uint32_t i = 0; i < nInc; ++i) points1In0[i] = poly1->vert(inds1[i]);
  const v3 sPoint = points1In0[0];
  const float eps = FLT_EPSILON;
  v3 rMin = V3(FLT_MAX, FLT_MAX, FLT_MAX), rMax = V3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (uint32_t i = 0; i < nRef; ++i) { points0In0[i] = m33mul(&rot, points0In0[i]); rMin = v3min(rMin, points0In0[i]); rMax = v3max(rMax, points0In0[i]); }
  rMin = v3sub(rMin, V3(eps, eps, eps)); rMax = v3add(rMax, V3(eps, eps, eps));
  const float d = points0In0[0].z, rd = d + contactDist;
  v3 iMin = V3(FLT_MAX, FLT_MAX, FLT_MAX), iMax = V3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  uint32_t inside = 0;
  for (uint32_t i = 0; i < nInc; ++i) {
    const v3 vert1 = points1In0[i];
    const v3 a = amxftransforminv(transform0To1, vert1);
    points1In0[i] = m33mul(&rot, a);
    const float z = points1In0[i].z;
    tval1[i] = z - d;
    points1In0[i].z = d;
    iMin = v3min(iMin, points1In0[i]); iMax = v3max(iMax, points1In0[i]);
    if (rd > z) {
      pen1[i] = 1;
      if (gjk_poly_contains_n(points0In0, (int)nRef, points1In0[i], rMin, rMax)) {
        inside++;
        if (*numContacts == GJK_POLY_MAX_CONTACTS) return;
        mc[*numContacts].a = vert1; mc[*numContacts].b = am33tmul(&rot, points1In0[i]); mc[*numContacts].n = contactNormal; mc[*numContacts].pen = tval1[i]; (*numContacts)++;
      }
    } else pen1[i] = 0;
  }
  if (inside == nInc) return;
  inside = 0;
  iMin = v3sub(iMin, V3(eps, eps, eps)); iMax = v3add(iMax, V3(eps, eps, eps));
  const v3 incidentNormal = anormalize(poly1->plane_n((uint32_t)incIdx));
  const v3 contactNormalIn1 = m33mul(&transform0To1->r, contactNormal);
  for (uint32_t i = 0; i < nRef; ++i) {
    if (gjk_poly_contains_n(points1In0, (int)nInc, points0In0[i], iMin, iMax)) {
      const v3 vert0 = am33tmul(&rot, points0In0[i]);
      const v3 a = amxftransform(transform0To1, vert0);
      const float nom = adot(incidentNormal, v3sub(sPoint, a)), denom = adot(incidentNormal, contactNormalIn1);
      const float t = nom / denom;
      if (t > contactDist) continue;
      inside++;
      if (*numContacts == GJK_POLY_MAX_CONTACTS) return;
      mc[*numContacts].a = v3scaleadd(contactNormalIn1, t, a); mc[*numContacts].b = vert0; mc[*numContacts].n = contactNormal; mc[*numContacts].pen = t; (*numContacts)++;
    }
  }
  if (inside == nRef) return;
  for (uint32_t iStart = 0, iEnd = nInc - 1; iStart < nInc; iEnd = iStart++) {
    if (!pen1[iStart] && !pen1[iEnd]) continue;
    const v3 ipA = points1In0[iStart], ipB = points1In0[iEnd];
    v3 ipAOri = points1In0[iStart]; ipAOri.z = tval1[iStart] + d;
    v3 ipBOri = points1In0[iEnd]; ipBOri.z = tval1[iEnd] + d;
    const v3 sMin = v3min(ipA, ipB), sMax = v3max(ipA, ipB);
    for (uint32_t rStart = 0, rEnd = nRef - 1; rStart < nRef; rEnd = rStart++) {
      const v3 rpA = points0In0[rStart], rpB = points0In0[rEnd];
      const v3 qMin = v3min(rpA, rpB), qMax = v3max(rpA, rpB);
      if ((sMin.x > qMax.x) || (qMin.x > sMax.x) || (sMin.y > qMax.y) || (qMin.y > sMax.y)) continue;
      const float a1 = gjk_signed_2d_tri_area(rpA, rpB, ipA), a2 = gjk_signed_2d_tri_area(rpA, rpB, ipB);
      if (0.f > a1 * a2) {
        const float a3 = gjk_signed_2d_tri_area(ipA, ipB, rpA), a4 = gjk_signed_2d_tri_area(ipA, ipB, rpB);
        if (0.f > a3 * a4) {
          const float t = a1 / (a2 - a1);
          const v3 pBB = v3negscalesub(v3sub(ipBOri, ipAOri), t, ipAOri);
          v3 pAA = pBB; pAA.z = d;
          const v3 pA = am33tmul(&rot, pAA);
          const v3 pB = amxftransform(transform0To1, am33tmul(&rot, pBB));
          const float pen = pBB.z - pAA.z;
          if (pen > contactDist) continue;
          if (*numContacts == GJK_POLY_MAX_CONTACTS) return;
          mc[*numContacts].a = pB; mc[*numContacts].b = pA; mc[*numContacts].n = contactNormal; mc[*numContacts].pen = pen; (*numContacts)++;
        }
      }
    }
  }
}
/* generateFullContactManifold :532-665, doOverlapTest == false (witness polygons of the GJK / EPA closest points).  map0 / map1 = world transforms of the two shapes. */
PXB_D void gjk_poly_full_manifold(const DevHull* poly0, const DevHull* poly1, const xf* map0, const xf* map1, MPoint* mc, int* numContacts, float contactDist,
                                          v3 normal, v3 closestA, v3 closestB, float marginA, float marginB, float toleranceLength) {
  const xf t10 = axfinvmul(map0, map1), t01 = axfinvmul(map1, map0);   /* PxTransformV::transformInv, then PxMatTransformV(PxTransformV) */
  const mxf transform1To0 = amxffromxf(&t10), transform0To1 = amxffromxf(&t01);
  const float lowerEps = toleranceLength * 1e-2f, upperEps = toleranceLength * 5e-2f;
  const float toleranceA = fmin_(fmax_(marginA, lowerEps), upperEps), toleranceB = fmin_(fmax_(marginB, lowerEps), upperEps);
  const v3 negNormal = v3neg(normal);
  const v3 normalIn0 = am33tmul(&transform0To1.r, normal);
  const int faceIndex1 = gjk_hull_witness_polygon_index(poly1, negNormal, closestB, toleranceB);
  const int faceIndex0 = gjk_hull_witness_polygon_index(poly0, normalIn0, amxftransforminv(&transform0To1, closestA), toleranceA);
  const v3 referenceNormal = anormalize(poly1->plane_n((uint32_t)faceIndex1)), incidentNormal = anormalize(poly0->plane_n((uint32_t)faceIndex0));
  const float referenceProject = fabsf(adot(referenceNormal, negNormal)), incidentProject = fabsf(adot(incidentNormal, normalIn0));
  if (referenceProject >= incidentProject) gjk_poly_generated_contacts(poly1, poly0, faceIndex1, faceIndex0, &transform1To0, mc, numContacts, contactDist);
  else {
    gjk_poly_generated_contacts(poly0, poly1, faceIndex0, faceIndex1, &transform0To1, mc, numContacts, contactDist);
    if (*numContacts > 0) {
      const v3 n = m33mul(&transform0To1.r, incidentNormal), nn = v3neg(n);
      for (int i = 0; i < *numContacts; ++i) { const v3 lb = mc[i].b; mc[i].b = mc[i].a; mc[i].a = lb; mc[i].n = nn; }
    }
  }
}
PXB_D v3 gjk_manifold_local_normal(const Manifold* m) { v3 n = m->pts[0].n; for (int i = 1; i < m->n; ++i) n = v3add(n, m->pts[i].n); return anormalize(n); }   /* getLocalNormal .h:710-718 */
PXB_D void gjk_manifold_to_contacts(const Manifold* m, v3 worldNormal, const xf* transf1, float contactDist, Contacts* out) {   /* .cpp:739-759 */
  out->count = 0; out->normal = worldNormal;
  for (int i = 0; i < m->n; ++i) { const float dist = m->pts[i].pen; if (contactDist >= dist) { out->point[out->count] = axftransform(transf1, m->pts[i].b); out->sep[out->count] = dist; out->count++; } }
}
/* pcmContactBoxConvex / pcmContactConvexConvex: shape A (box or hull) relative to hull B, in the same three phases as the capsule (see GjkCarry).
 * The manifold phase returns 1 when the reference would run the SAT branch (not restated), 0 otherwise. */
PXB_D int gjk_poly_convex_refresh(const xf* transf0, const xf* transf1, float marginPcmA, float radiusA, const DevHull* hullB, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out, int* flags) {
  out->count = 0;
  const xf curRTrans = axfinvmul(transf1, transf0);
  const mxf aToB = amxffromxf(&curRTrans);
  const float convexMarginB = gjk_hull_pcm_margin(hullB, toleranceLength);
  const float minMargin = fmin_(marginPcmA, convexMarginB);
  const int initialContacts = manifold->n;
  manifold_refresh(*manifold, aToB, minMargin * 0.8f);
  const int bLostContacts = manifold->n != initialContacts;
  const float radiusB = alen(hullB->internalExtents);
  if (bLostContacts || invalidate_boxconvex(*manifold, curRTrans, transf0->q, transf1->q, minMargin, radiusA, radiusB)) {
    manifold->rel = curRTrans; manifold->quatA = transf0->q; manifold->quatB = transf1->q; manifold->dirty = 1;
    *flags = initialContacts | (bLostContacts << 8);
    return 1;
  }
  if (manifold->n > 0) gjk_manifold_to_contacts(manifold, manifold_world_normal(*manifold, *transf1), transf1, contactDist, out);
  return 0;
}
/* generateOrProcessContacts* + addGJKEPAContacts: what pcmContactBoxConvex / ConvexConvex do with the query's answer (status / epaStatus as for the capsule).
 * centreA: convexA's centre in its own frame.  Returns 1 when the full manifold generation has to run. */
PXB_D int gjk_poly_convex_post(const xf* transf1, v3 centreALocal, v3 centreB, float marginPcmA, const DevHull* hullB, float contactDist, float toleranceLength, int flags, int status, int epaStatus,
                               const GjkOutput* output, const mxf* aToB, Manifold* manifold, Contacts* out, GjkCarry* carry) {
  const int initialContacts = flags & 0xff;
  const float minMargin = fmin_(marginPcmA, gjk_hull_pcm_margin(hullB, toleranceLength));
  const v3 localNor = manifold->n ? gjk_manifold_local_normal(manifold) : V3(0, 0, 0);
  const float replaceBreakingThreshold = minMargin * 0.05f;
  int doOverlapTest = 0;
  if (status == GJK_DEGENERATE) {
    const float costheta = adot(output->searchDir, output->normal);
    if (costheta > 0.9999f) {
      const v3 centreA = amxftransform(aToB, centreALocal);
      const v3 dir = anormalize(v3sub(centreA, centreB));
      if (adot(dir, output->normal) > 0.707f) gjk_add_manifold_point(manifold, amxftransforminv(aToB, output->closestA), output->closestB, output->normal, output->penDep, replaceBreakingThreshold);
      else doOverlapTest = 1;
    } else doOverlapTest = 1;
  } else if (status == GJK_CONTACT) gjk_add_manifold_point(manifold, amxftransforminv(aToB, output->closestA), output->closestB, output->normal, output->penDep, replaceBreakingThreshold);
  else {
    if (epaStatus == EPA_CONTACT) gjk_add_manifold_point(manifold, amxftransforminv(aToB, output->closestA), output->closestB, output->normal, output->penDep, replaceBreakingThreshold);
    else doOverlapTest = 1;
  }
  const int fullContactGen = (0.707106781f > adot(localNor, output->normal)) || (manifold->n < initialContacts);
  if (fullContactGen || doOverlapTest) {
    carry->normal = output->normal; carry->closestA = output->closestA; carry->closestB = output->closestB; carry->doOverlapTest = doOverlapTest;
    return 1;
  }
  const v3 newLocalNor = v3add(localNor, output->normal);
  gjk_manifold_to_contacts(manifold, anormalize(aqrot(transf1->q, newLocalNor)), transf1, contactDist, out);
  return 0;
}
PXB_D int gjk_poly_convex_query(const xf* transf0, const xf* transf1, GjkConvex* convexA, float marginPcmA, const DevHull* hullB, float contactDist, float toleranceLength, int flags,
                                Manifold* manifold, Contacts* out, GjkCarry* carry) {
  out->count = 0;
  const xf curRTrans = axfinvmul(transf1, transf0);
  const mxf aToB = amxffromxf(&curRTrans);
  gjk_cvx_make_relative(convexA, &aToB);
  const GjkConvex convexB = gjk_cvx_hull(hullB);
  GjkOutput output; output.normal = output.closestA = output.closestB = output.searchDir = V3(0, 0, 0); output.penDep = 0.f;
  const int status = gjk_penetration(convexA, &convexB, aToB.p, contactDist, 1, manifold->aInd, manifold->bInd, &manifold->nWarm, &output);
  if (status == GJK_NON_INTERSECT) return 0;
  int epaStatus = 0;
  if (status != GJK_DEGENERATE && status != GJK_CONTACT) epaStatus = gjk_epa_penetration(convexA, &convexB, manifold->aInd, manifold->bInd, manifold->nWarm, 1, toleranceLength, &output);
  return gjk_poly_convex_post(transf1, convexA->center, convexB.center, marginPcmA, hullB, contactDist, toleranceLength, flags, status, epaStatus, &output, &aToB, manifold, out, carry);
}
/* fullContactsGenerationBoxConvex / ConvexConvex */
PXB_D int gjk_poly_convex_manifold(const xf* transf0, const xf* transf1, const GjkConvex* convexA, const DevHull* polyA, const DevHull* hullB, float contactDist, float toleranceLength, const GjkCarry* carry,
                                   Manifold* manifold, Contacts* out) {
  out->count = 0;
  const GjkConvex convexB = gjk_cvx_hull(hullB);
  MPoint mc[GJK_POLY_MAX_CONTACTS]; int numContacts = 0; const int doOverlapTest = carry->doOverlapTest;
  if (doOverlapTest) { if (!gjk_poly_full_manifold_sat(polyA, convexA->type == GJK_CVX_BOX, hullB, transf0, transf1, mc, &numContacts, contactDist)) return 0; }
  else gjk_poly_full_manifold(polyA, hullB, transf0, transf1, mc, &numContacts, contactDist, carry->normal, carry->closestA, carry->closestB, convexA->margin, convexB.margin, toleranceLength);
  if (numContacts > 0) {
    if (numContacts <= PXB_MANIFOLD_CACHE) { for (int i = 0; i < numContacts; ++i) manifold->pts[i] = mc[i]; manifold->n = numContacts; }
    else { reduce_batch(*manifold, mc, numContacts, toleranceLength); manifold->n = PXB_MANIFOLD_CACHE; }
    gjk_manifold_to_contacts(manifold, manifold_world_normal(*manifold, *transf1), transf1, contactDist, out);
  } else if (!doOverlapTest) gjk_manifold_to_contacts(manifold, manifold_world_normal(*manifold, *transf1), transf1, contactDist, out);
  return 0;
}
PXB_D int gjk_pcm_poly_convex(const xf* transf0, const xf* transf1, GjkConvex* convexA, const DevHull* polyA, float marginPcmA, float radiusA, const DevHull* hullB,
                                      float contactDist, float toleranceLength, Manifold* manifold, Contacts* out) {   // the three phases back to back (single-kernel variant)
  int flags; GjkCarry carry;
  if (!gjk_poly_convex_refresh(transf0, transf1, marginPcmA, radiusA, hullB, contactDist, toleranceLength, manifold, out, &flags)) return 0;
  if (!gjk_poly_convex_query(transf0, transf1, convexA, marginPcmA, hullB, contactDist, toleranceLength, flags, manifold, out, &carry)) return 0;
  return gjk_poly_convex_manifold(transf0, transf1, convexA, polyA, hullB, contactDist, toleranceLength, &carry, manifold, out);
}
struct BoxAsHull { float4 verts[8]; float4 polys[12]; DevHull view; };
PXB_D const DevHull* gjk_box_as_hull(BoxAsHull* b, v3 ext) {   // PCMPolygonalBox as the polygonal view the hulls use (per-thread arrays)
  PolyBox pb; gjk_poly_box(&pb, ext);
  for (int i = 0; i < 8; ++i) b->verts[i] = F4(pb.verts[i], 0.f);
  for (int i = 0; i < 6; ++i) { b->polys[2 * i] = F4(pb.polys[i].n, pb.polys[i].d); b->polys[2 * i + 1] = make_float4(__uint_as_float((uint32_t)i * 4), __uint_as_float(4u), __uint_as_float((uint32_t)pb.polys[i].minIndex), 0.f); }
  b->view.nVerts = 8; b->view.nPolys = 6; b->view.nEdges = 0; b->view.internalExtents = ext; b->view.centerOfMass = V3(0, 0, 0); b->view.internalRadius = 0.f;
  b->view.verts = b->verts; b->view.polys = b->polys; b->view.vertexRefs = gjk_box_poly_refs; b->view.facesByEdges = nullptr;
  return &b->view;
}
PXB_D int gjk_pcm_box_convex(const xf* transf0, const xf* transf1, v3 boxExtents, const DevHull* hull, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out) {
  BoxAsHull bh; const DevHull* polyA = gjk_box_as_hull(&bh, boxExtents);
  GjkConvex box = gjk_cvx_box(V3(0, 0, 0), boxExtents);
  return gjk_pcm_poly_convex(transf0, transf1, &box, polyA, box_margin(boxExtents, toleranceLength), alen(boxExtents), hull, contactDist, toleranceLength, manifold, out);
}
PXB_D int gjk_pcm_convex_convex(const xf* transf0, const xf* transf1, const DevHull* hull0, const DevHull* hull1, float contactDist, float toleranceLength, Manifold* manifold, Contacts* out) {
  GjkConvex c0 = gjk_cvx_hull(hull0);
  return gjk_pcm_poly_convex(transf0, transf1, &c0, hull0, gjk_hull_pcm_margin(hull0, toleranceLength), alen(hull0->internalExtents), hull1, contactDist, toleranceLength, manifold, out);
}
