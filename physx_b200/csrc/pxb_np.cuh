// pxb_np.cuh -- narrowphase contact generation (stage 2 of the rigid-body step), device functions.
//
// Semantics follow the reference's CPU PCM path (persistent contact manifolds), which is what the
// parity target is, not the reference GPU's stateless box-box:
//   persistent manifold refresh / invalidation / reduction
//       physx/source/geomutils/src/pcm/GuPersistentContactManifold.h:160-260, :695-752
//       physx/source/geomutils/src/pcm/GuPersistentContactManifold.cpp:739-1174
//   plane vs box   physx/source/geomutils/src/pcm/GuPCMContactPlaneBox.cpp:36-209
//   box vs box     physx/source/geomutils/src/pcm/GuPCMContactBoxBox.cpp:42-971
// Known deviation: the reference's _mm_rcp_ps in the segment/AABB clip is an exact reciprocal here.
// The GJK/EPA single-point fallback of box-box (taken when the SAT passes and face clipping produces
// no point) is gjk_boxbox_gjk_fallback in pxb_gjk.cuh.
#pragma once
#include "pxb_math.cuh"

#define PXB_MANIFOLD_CACHE 4
#define PXB_MANIFOLD_F4 16   // float4 slots per persistent manifold record (14 used)

struct MPoint { v3 a, b, n; float pen; };
struct Manifold { int n; xf rel; q4 quatA, quatB; MPoint pts[PXB_MANIFOLD_CACHE]; int dirty; uint8_t aInd[4], bInd[4], nWarm; };   // aInd / bInd / nWarm: GJK warm start (mAIndice / mBIndice / mNumWarmStartPoints)   // dirty: points / frames regenerated this frame (else only pens changed)
struct Contacts { int count; v3 normal; v3 point[PXB_MANIFOLD_CACHE]; float sep[PXB_MANIFOLD_CACHE]; };

PXB_D void manifold_init(Manifold& m) {
  m.n = 0; m.rel.q = Q4(0, 0, 0, 1); m.rel.p = V3(FLT_MAX, FLT_MAX, FLT_MAX);
  m.quatA = Q4(0, 0, 0, 1); m.quatB = Q4(0, 0, 0, 1); m.dirty = 1;
  for (int i = 0; i < PXB_MANIFOLD_CACHE; ++i) { m.pts[i].a = V3(0, 0, 0); m.pts[i].b = V3(0, 0, 0); m.pts[i].n = V3(0, 0, 0); m.pts[i].pen = 0.f; m.aInd[i] = 0; m.bInd[i] = 0; }
  m.nWarm = 0;
}
// GJK pair types keep their warm-start simplex in slot 14 of the record: (aInd packed, bInd packed, count)
PXB_D void manifold_load_warm(Manifold& m, const float4* __restrict__ rec) {
  const float4 w = rec[14]; const uint32_t a = __float_as_uint(w.x), b = __float_as_uint(w.y);
  for (int i = 0; i < 4; ++i) { m.aInd[i] = (uint8_t)(a >> (8 * i)); m.bInd[i] = (uint8_t)(b >> (8 * i)); }
  m.nWarm = (uint8_t)__float_as_uint(w.z);
}
PXB_D void manifold_store_warm(const Manifold& m, float4* __restrict__ rec) {
  uint32_t a = 0, b = 0;
  for (int i = 0; i < 4; ++i) { a |= (uint32_t)m.aInd[i] << (8 * i); b |= (uint32_t)m.bInd[i] << (8 * i); }
  rec[14] = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float((uint32_t)m.nWarm), 0.f);
}

// record layout: [0]=(n, rel.p) [1]=rel.q [2]=quatA [3]=quatB [4]=pen of the 4 points [5..13]=4 points x 9 floats (a, b, n).
// The penetrations sit in one float4 because they are the only thing a persisting manifold changes per frame
// (refreshContactPoints, GuPersistentContactManifold.h:723-752): steady state writes 16-32 B per pair instead of 224 B.
PXB_D void manifold_load(Manifold& m, const float4* __restrict__ rec) {
  const float4 h = rec[0], rq = rec[1], qa = rec[2], qb = rec[3], pen = rec[4];
  m.n = __float_as_int(h.x); m.rel.p = V3(h.y, h.z, h.w); m.rel.q = Q4(rq); m.quatA = Q4(qa); m.quatB = Q4(qb); m.dirty = 0;
  float f[36];
#pragma unroll
  for (int i = 0; i < 9; ++i) { const float4 v = rec[5 + i]; f[i * 4] = v.x; f[i * 4 + 1] = v.y; f[i * 4 + 2] = v.z; f[i * 4 + 3] = v.w; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m.pts[i].a = V3(f[i * 9], f[i * 9 + 1], f[i * 9 + 2]); m.pts[i].b = V3(f[i * 9 + 3], f[i * 9 + 4], f[i * 9 + 5]);
    m.pts[i].n = V3(f[i * 9 + 6], f[i * 9 + 7], f[i * 9 + 8]);
  }
  m.pts[0].pen = pen.x; m.pts[1].pen = pen.y; m.pts[2].pen = pen.z; m.pts[3].pen = pen.w;
}
PXB_D void manifold_store(const Manifold& m, float4* __restrict__ rec) {
  rec[0] = make_float4(__int_as_float(m.n), m.rel.p.x, m.rel.p.y, m.rel.p.z); rec[1] = F4(m.rel.q); rec[2] = F4(m.quatA); rec[3] = F4(m.quatB);
  rec[4] = make_float4(m.pts[0].pen, m.pts[1].pen, m.pts[2].pen, m.pts[3].pen);
  float f[36];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[i * 9] = m.pts[i].a.x; f[i * 9 + 1] = m.pts[i].a.y; f[i * 9 + 2] = m.pts[i].a.z; f[i * 9 + 3] = m.pts[i].b.x; f[i * 9 + 4] = m.pts[i].b.y;
    f[i * 9 + 5] = m.pts[i].b.z; f[i * 9 + 6] = m.pts[i].n.x; f[i * 9 + 7] = m.pts[i].n.y; f[i * 9 + 8] = m.pts[i].n.z;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) rec[5 + i] = make_float4(f[i * 4], f[i * 4 + 1], f[i * 4 + 2], f[i * 4 + 3]);
}
PXB_D void manifold_store_pens(const Manifold& m, float4* __restrict__ rec) { rec[4] = make_float4(m.pts[0].pen, m.pts[1].pen, m.pts[2].pen, m.pts[3].pen); }

PXB_D float box_margin(v3 e, float toleranceLength) {  // GuVecBox.h:81-88
  const float mn = fmin_(e.x, fmin_(e.y, e.z));
  return fmin_(mn * 0.15f, toleranceLength * 0.15f);
}

PXB_D void manifold_refresh(Manifold& m, const mxf& aToB, float projectBreakingThreshold) {  // GuPersistentContactManifold.h:723-752
  const float sq = projectBreakingThreshold * projectBreakingThreshold;
  for (int i = m.n; i > 0; --i) {
    MPoint& mp = m.pts[i - 1];
    const v3 localAInB = amxftransform(aToB, mp.a);
    const v3 v = localAInB - mp.b;
    const float dist = adot(v, mp.n);
    const v3 projected = negscalesub(mp.n, dist, localAInB);
    const v3 diff = mp.b - projected;
    const float d2 = adot(diff, diff);
    if (d2 > sq) { m.n--; m.pts[i - 1] = m.pts[m.n]; }
    else mp.pen = dist;
  }
}
PXB_D float max_pos_delta(const Manifold& m, v3 curP) {
  const v3 d = vabs(curP - m.rel.p);
  return fmax_(d.x, fmax_(d.y, d.z));
}
PXB_D bool invalidate_plane(const Manifold& m, const xf& cur, float minMargin, float ratio) {  // :245-257
  const float thresholdP = minMargin * ratio;
  return (max_pos_delta(m, cur.p) > thresholdP) || (0.99996f > adot4(cur.q, m.rel.q));
}
PXB_D bool invalidate_boxconvex(const Manifold& m, const xf& cur, q4 quatA, q4 quatB, float minMargin, float radiusA, float radiusB) {  // :190-223
  const float thr[5] = {0.5f, 0.125f, 0.25f, 0.375f, 0.375f};
  const float qthr[5] = {0.9998f, 0.9999f, 0.9999f, 0.9999f, 0.9999f};
  const float thresholdP = minMargin * thr[m.n];
  const float deltaP = max_pos_delta(m, cur.p);
  const float thresholdQ = qthr[m.n];
  const float dqA = adot4(quatA, m.quatA), dqB = adot4(quatB, m.quatB);
  bool gen = (deltaP > thresholdP) || (thresholdQ > dqA) || (thresholdQ > dqB);
  if (!gen) {
    const float aRad = dqA < 1.0f ? acosf(dqA) : 0.f;
    const float bRad = dqB < 1.0f ? acosf(dqB) : 0.f;
    gen = (aRad * radiusA > thresholdP) || (bRad * radiusB > thresholdP);
  }
  return gen;
}

// GuPersistentContactManifold.cpp:859-1005
PXB_D void reduce_cluster(Manifold& m, const MPoint* p, int numPoints) {
  uint32_t chosen = 0u;
  float maxDist = FLT_MAX; int index = 0; int indices[4];
  for (int i = 0; i < numPoints; ++i) if (maxDist > p[i].pen) { maxDist = p[i].pen; index = i; }
  m.pts[0] = p[index]; chosen |= 1u << index; indices[0] = index;
  v3 v = p[0].b - m.pts[0].b; maxDist = adot(v, v); index = 0;
  for (int i = 1; i < numPoints; ++i) { v = p[i].b - m.pts[0].b; const float d = adot(v, v); if (d > maxDist) { maxDist = d; index = i; } }
  m.pts[1] = p[index]; chosen |= 1u << index; indices[1] = index;
  maxDist = -FLT_MAX; index = 0;
  v = m.pts[1].b - m.pts[0].b;
  const v3 cn0 = m.pts[0].n;
  v3 norm = cross(v, cn0);
  const float sqLen = adot(norm, norm);
  if (sqLen > 0.f) { const float l = sqrtf(sqLen); norm = V3(norm.x / l, norm.y / l, norm.z / l); } else norm = cn0;
  float minDist = FLT_MAX; int index1 = 0;
  for (int i = 0; i < numPoints; ++i) if (!((chosen >> i) & 1u)) {
    v = p[i].b - m.pts[0].b; const float d = adot(v, norm);
    if (d > maxDist) { maxDist = d; index = i; }
    if (minDist > d) { minDist = d; index1 = i; }
  }
  m.pts[2] = p[index]; chosen |= 1u << index; indices[2] = index;
  if (minDist * maxDist > 0.f) {
    maxDist = -FLT_MAX;
    for (int i = 0; i < numPoints; ++i) if (!((chosen >> i) & 1u)) {
      v = p[i].b - m.pts[0].b; const float d = adot(v, norm);
      if (d > maxDist) { maxDist = d; index1 = i; }
    }
  }
  m.pts[3] = p[index1]; chosen |= 1u << index1; indices[3] = index1;
  for (int i = 0; i < numPoints; ++i) if (!((chosen >> i) & 1u)) {
    maxDist = FLT_MAX; const float pen = p[i].pen; index = 0;
    for (int j = 0; j < 4; ++j) { const v3 v1 = p[i].b - m.pts[j].b; const float dist = adot(v1, v1); if (maxDist > dist) { maxDist = dist; index = j; } }
    if (p[indices[index]].pen > pen) indices[index] = i;
  }
  for (int k = 0; k < 4; ++k) m.pts[k] = p[indices[k]];
}

// GuPersistentContactManifold.cpp:1008-1174
PXB_D void reduce_batch(Manifold& m, const MPoint* p, int numPoints, float toleranceLength) {
  int chosenIdx[4]; uint8_t cand[16];
  float maxPen = p[0].pen, minPen = maxPen;
  int index = 0; cand[0] = 0; int candIndex = 0; int nbCand = numPoints;
  for (int i = 1; i < numPoints; ++i) {
    cand[i] = (uint8_t)i;
    const float pen = p[i].pen;
    minPen = fmax_(minPen, pen);
    if (maxPen > pen) { maxPen = pen; index = i; candIndex = i; }
  }
  chosenIdx[0] = index;
  nbCand--; cand[candIndex] = cand[nbCand];
  v3 v = p[cand[0]].b - p[chosenIdx[0]].b;
  float maxDist = adot(v, v); index = cand[0]; candIndex = 0;
  for (int i = 1; i < nbCand; ++i) {
    v = p[cand[i]].b - p[chosenIdx[0]].b; const float d = adot(v, v);
    if (d > maxDist) { maxDist = d; index = cand[i]; candIndex = i; }
  }
  chosenIdx[1] = index;
  nbCand--; cand[candIndex] = cand[nbCand];
  v = p[chosenIdx[1]].b - p[chosenIdx[0]].b;
  const v3 cn0 = p[chosenIdx[0]].n;
  v3 norm = cross(v, cn0);
  const float sqLen = adot(norm, norm);
  if (sqLen > 0.f) { const float l = sqrtf(sqLen); norm = V3(norm.x / l, norm.y / l, norm.z / l); } else norm = cn0;
  maxDist = -FLT_MAX; index = 0; candIndex = 0;
  float minDist = FLT_MAX; int index1 = 0, candIndex1 = 0;
  for (int i = 0; i < nbCand; ++i) {
    v = p[cand[i]].b - p[chosenIdx[0]].b; const float d = adot(v, norm);
    if (d > maxDist) { maxDist = d; index = cand[i]; candIndex = i; }
    if (minDist > d) { minDist = d; index1 = cand[i]; candIndex1 = i; }
  }
  chosenIdx[2] = index;
  nbCand--; cand[candIndex] = cand[nbCand];
  if (nbCand == candIndex1) candIndex1 = candIndex;
  if (minDist * maxDist > 0.f) {
    maxDist = -FLT_MAX;
    for (int i = 0; i < nbCand; ++i) {
      v = p[cand[i]].b - p[chosenIdx[0]].b; const float d = adot(v, norm);
      if (d > maxDist) { maxDist = d; index1 = cand[i]; candIndex1 = i; }
    }
  }
  chosenIdx[3] = index1;
  nbCand--; cand[candIndex1] = cand[nbCand];
  const float eps = toleranceLength * 0.02f;
  if ((eps > maxPen) && (minPen > eps)) {
    for (int i = 0; i < 4; ++i) {
      float pen = p[chosenIdx[i]].pen;
      if (pen > eps) {
        candIndex = 0xff;
        for (int j = 0; j < nbCand; ++j) {
          const float pen1 = p[cand[j]].pen;
          if ((pen > pen1) && (eps > pen1)) { pen = pen1; candIndex = j; }
        }
        if (candIndex < nbCand) { const int orig = chosenIdx[i]; chosenIdx[i] = cand[candIndex]; cand[candIndex] = (uint8_t)orig; }
      }
      m.pts[i] = p[chosenIdx[i]];
    }
  } else {
    for (int i = 0; i < 4; ++i) m.pts[i] = p[chosenIdx[i]];
  }
}

PXB_D v3 manifold_world_normal(const Manifold& m, const xf& trB) {  // GuPersistentContactManifold.h:695-708
  v3 n = m.pts[0].n;
  for (int i = 1; i < m.n; ++i) n = n + m.pts[i].n;
  const float sq = adot(n, n);
  const v3 nn = (sq > FLT_EPSILON) ? n : m.pts[0].n;
  return aqrot_normalize(trB.q, nn);
}

// plane (shape0) vs box (shape1): GuPCMContactPlaneBox.cpp:36-209
PXB_D void pcm_plane_box(const xf& planeTm, const xf& boxTm, v3 be, float contactDist, float toleranceLength, Manifold& man, Contacts& out) {
  const xf cur = axfinvmul(planeTm, boxTm);  // box to plane
  const v3 negPlaneNormal = anormalize(-aqbasis0(planeTm.q));
  const float margin = box_margin(be, toleranceLength);
  const int initial = man.n;
  const mxf aToB = amxffromxf(cur);
  manifold_refresh(man, aToB, margin * 0.2f);
  const bool lost = man.n != initial;
  out.count = 0; out.normal = negPlaneNormal;
  if (lost || invalidate_plane(man, cur, margin, 0.2f)) {
    const v3 ln = V3(1.f, 0.f, 0.f);
    man.n = 0; man.rel = cur; man.dirty = 1;
    const v3 t0 = aToB.r.c0 * be.x, t1 = aToB.r.c1 * be.y, t2 = aToB.r.c2 * be.z;
    const v3 nt2 = -t2;
    const float px = aToB.p.x;
    const v3 t01 = t0 + t1, t02 = t0 - t1;
    const float acceptance = contactDist - px;
    MPoint mc[8]; int num = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // corner order of the reference: (x,y,z) (x,y,-z) (x,-y,z) (x,-y,-z) (-x,y,z) (-x,y,-z) (-x,-y,z) (-x,-y,-z)
      const v3 zt = (k & 1) ? nt2 : t2;
      float s;
      switch (k >> 1) { case 0: s = (zt + t01).x; break; case 1: s = (zt + t02).x; break; case 2: s = (zt - t02).x; break; default: s = (zt - t01).x; }
      if (acceptance > s) {
        const v3 corner = V3((k & 4) ? -be.x : be.x, (k & 2) ? -be.y : be.y, (k & 1) ? -be.z : be.z);
        const float pen = s + px;
        mc[num].a = corner; mc[num].b = negscalesub(ln, pen, amxftransform(aToB, corner)); mc[num].n = ln; mc[num].pen = pen; num++;
      }
    }
    if (num <= PXB_MANIFOLD_CACHE) { for (int i = 0; i < num; ++i) man.pts[i] = mc[i]; man.n = num; }
    else { reduce_cluster(man, mc, num); man.n = PXB_MANIFOLD_CACHE; }
  }
  for (int i = 0; i < man.n; ++i) {
    const float dist = man.pts[i].pen;
    if (contactDist >= dist) { out.point[out.count] = axftransform(planeTm, man.pts[i].b); out.sep[out.count] = dist; out.count++; }
  }
}

// ---- box vs box ----
PXB_D void incident_polygon(v3* pts, v3& faceNormal, v3 axis, const mxf& t, v3 ext) {  // GuPCMContactBoxBox.cpp:42-118
  float ex = ext.x, ey = ext.y, ez = ext.z;
  const v3 u0 = t.r.c0, u1 = t.r.c1, u2 = t.r.c2;
  const float d0 = adot(u0, axis), d1 = adot(u1, axis), d2 = adot(u2, axis);
  const float a0 = fabsf(d0), a1 = fabsf(d1), a2 = fabsf(d2);
  v3 base, e1, e2;
  if (a0 >= a1 && a0 >= a2) {
    const bool con = d0 > 0.f; faceNormal = con ? -u0 : u0; ex = con ? -ex : ex;
    const v3 r0 = u0 * ex, r1 = u1 * ey, r2 = u2 * ez;
    base = t.p + r0; e1 = r1 + r2; e2 = r1 - r2;
  } else if (a1 >= a2) {
    const bool con = d1 > 0.f; faceNormal = con ? -u1 : u1; ey = con ? -ey : ey;
    const v3 r0 = u0 * ex, r1 = u1 * ey, r2 = u2 * ez;
    base = t.p + r1; e1 = r0 + r2; e2 = r0 - r2;
  } else {
    const bool con = d2 > 0.f; faceNormal = con ? -u2 : u2; ez = con ? -ez : ez;
    const v3 r0 = u0 * ex, r1 = u1 * ey, r2 = u2 * ez;
    base = t.p + r2; e1 = r0 + r1; e2 = r0 - r1;
  }
  pts[0] = base + e1; pts[1] = base + e2; pts[2] = base - e1; pts[3] = base - e2;
}

PXB_D float comp(v3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// The reference's V3RecipFast = x86 RCPSS / RCPPS, a 12-bit table approximation: where box-box edge clipping puts its points -- and so which of them sit just inside
// the friction-offset threshold -- depends on it (an exact reciprocal moves tumbling boxes by up to 3e-2 per step against the reference: friction anchors from other
// points).  The instruction is restated through the pinning host's table (pxb_rcp_table.h, tools/make_rcp_table.c): its result depends only on sign, exponent and the
// top 11 mantissa bits and scales exactly with the exponent.
#include "pxb_rcp_table.h"
PXB_D float rcp_fast(float x) {
  const uint32_t b = __float_as_uint(x); const uint32_t e = (b >> 23) & 0xffu;
  if (e == 0u || e >= 253u) return 1.0f / x;   // zero, denormals, huge, inf, nan: outside the table's domain (results the clipping code does not use)
  const uint32_t r = PXB_RCP_TABLE[(b >> 12) & 0x7ffu];
  return __uint_as_float((b & 0x80000000u) | ((((r >> 23) & 0xffu) - (e - 127u)) << 23) | (r & 0x7fffffu));
}
PXB_D bool seg_aabb(v3 p0, v3 d, v3 mx, v3 mn, float& tmin, float& tmax) {  // :121-165
  const float eps = 1e-6f;
  bool par[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    par[k] = eps > fabsf(comp(d, k));
    const bool outside = (comp(p0, k) > comp(mx, k)) || (comp(mn, k) > comp(p0, k));
    if (par[k] && outside) return false;
  }
  float ft1 = -FLT_MAX, ft2 = FLT_MAX;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float odd = rcp_fast(comp(d, k));  // reference: V3RecipFast
    const float t1 = par[k] ? 0.f : (comp(mn, k) - comp(p0, k)) * odd;
    const float t2 = par[k] ? FLT_MAX : (comp(mx, k) - comp(p0, k)) * odd;
    ft1 = fmax_(ft1, fmin_(t1, t2)); ft2 = fmin_(ft2, fmax_(t1, t2));
  }
  const float tminf = fmax_(ft1, 0.f), tmaxf = fmin_(1.f, ft2);
  tmin = tminf; tmax = tmaxf;
  return !((tminf > tmaxf) || (tminf > 1.f));
}

PXB_D bool poly_contains(const v3* verts, v3 p, v3 mn, v3 mx) {  // GuPCMContactGenUtil.cpp:35-103, 4 vertices
  if ((mn.x > p.x) || (p.x > mx.x) || (mn.y > p.y) || (p.y > mx.y)) return false;
  const float tx = p.x, ty = p.y; const float eps = FLT_EPSILON;
  int inter = 0;
  for (int i = 0, j = 3; i < 4; j = i++) {
    const float jy = verts[j].y, iy = verts[i].y, jx = verts[j].x, ix = verts[i].x;
    if ((tx == jx && ty == jy) || (tx == ix && ty == iy)) return true;
    const bool yflag0 = jy > ty, yflag1 = iy > ty;
    if (yflag0 != yflag1) {
      const float jix = ix - jx, jiy = iy - jy, jty = ty - jy;
      const float part1 = jty * jix, part2 = (jx + eps) * jiy, part3 = tx * jiy;
      const bool c = jiy > 0.f;
      const float tmp = part1 + part2;
      const float comp1 = c ? tmp : part3, comp2 = c ? part3 : tmp;
      if (comp1 >= comp2) { if (inter == 1) return false; inter++; }
    }
  }
  return inter > 0;
}

PXB_D void calc_contacts(float extentX_, float extentY_, v3* pts, v3 incN, v3 localNormal, MPoint* mc, int& numContacts, float contactDist) {  // :168-330
  const float extentX = extentX_ * 1.0001f, extentY = extentY_ * 1.0001f;
  const float nExtentX = -extentX, nExtentY = -extentY;
  bool pPen[4], pArea[4];
  v3 bmin = V3(FLT_MAX, FLT_MAX, FLT_MAX), bmax = V3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  int n = numContacts;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bmin = vmin(bmin, pts[i]); bmax = vmax(bmax, pts[i]);
    const float z = -pts[i].z;
    if (contactDist > z) {
      pPen[i] = true;
      const v3 ap = vabs(pts[i]);
      if (extentX >= ap.x && extentY >= ap.y) { pArea[i] = true; mc[n].a = V3(pts[i].x, pts[i].y, 0.f); mc[n].b = pts[i]; mc[n].n = localNormal; mc[n].pen = z; n++; }
      else pArea[i] = false;
    } else { pPen[i] = false; pArea[i] = false; }
  }
  if (n == 4) { numContacts = n; return; }
  {
    const float denom = incN.z;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const v3 q = V3((k & 2) ? nExtentX : extentX, (k & 1) ? nExtentY : extentY, 0.f);
      if (poly_contains(pts, q, bmin, bmax)) {
        const float nom = adot(incN, pts[0] - q);
        const float t = nom / denom; const float pen = -t;
        if (contactDist > pen) { mc[n].a = q; mc[n].b = V3(q.x, q.y, t); mc[n].n = localNormal; mc[n].pen = pen; n++; }
      }
    }
  }
  const v3 ext = V3(extentX, extentY, FLT_MAX);
  const v3 negExt = V3(nExtentX, nExtentY, -(contactDist + FLT_EPSILON));
  for (int rStart = 0, rEnd = 3; rStart < 4; rEnd = rStart++) {
    const v3 p0 = pts[rStart], p1 = pts[rEnd];
    if (!pPen[rStart] && !pPen[rEnd]) continue;
    const bool con0 = pPen[rStart] && pArea[rStart], con1 = pPen[rEnd] && pArea[rEnd];
    if (con0 && con1) continue;
    const v3 p0p1 = p1 - p0;
    float tmin, tmax;
    if (seg_aabb(p0, p0p1, ext, negExt, tmin, tmax)) {
      if (!con0) { const v3 ip = scaleadd(p0p1, tmin, p0); mc[n].a = V3(ip.x, ip.y, 0.f); mc[n].b = ip; mc[n].n = localNormal; mc[n].pen = -ip.z; n++; }
      if (!con1) { const v3 ip = scaleadd(p0p1, tmax, p0); mc[n].a = V3(ip.x, ip.y, 0.f); mc[n].b = ip; mc[n].n = localNormal; mc[n].pen = -ip.z; n++; }
    }
  }
  numContacts = n;
}

PXB_D float sum3(v3 v) { return v.x + (v.y + v.z); }

// GuPCMContactBoxBox.cpp:332-846; false when a separating axis exists
PXB_D bool boxbox_generate(v3 e0, v3 e1, const mxf& t0, const mxf& t1, float contactDist, MPoint* mc, int& numContacts) {
  const mxf t1To0 = amxfinvmul(t0, t1);
  const m33 r01 = mtranspose(t1To0.r);
  const float uEps = 1e-6f;
  const v3 tt = t1To0.p;
  const v3 col[3] = {t1To0.r.c0, t1To0.r.c1, t1To0.r.c2};
  const v3 rr[3] = {r01.c0, r01.c1, r01.c2};
  v3 abs1To0[3], abs0To1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    abs1To0[k] = V3(fabsf(col[k].x) + uEps, fabsf(col[k].y) + uEps, fabsf(col[k].z) + uEps);
    abs0To1[k] = V3(fabsf(rr[k].x) + uEps, fabsf(rr[k].y) + uEps, fabsf(rr[k].z) + uEps);
  }
  float sign[6], overlap[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    sign[k] = comp(tt, k);
    const float rb = sum3(vmul(abs0To1[k], e1));
    overlap[k] = ((comp(e0, k) + rb) - fabsf(sign[k])) + contactDist;
    if (0.f > overlap[k]) return false;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    sign[3 + k] = adot(tt, col[k]);
    const float ra = sum3(vmul(abs1To0[k], e0));
    overlap[3 + k] = ((ra + comp(e1, k)) - fabsf(sign[3 + k])) + contactDist;
    if (0.f > overlap[3 + k]) return false;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float absSign = fabsf(comp(col[j], i1) * comp(tt, i2) - comp(col[j], i2) * comp(tt, i1));
      float ra, rb;
      if (i == 0) ra = abs1To0[j].z * e0.y + abs1To0[j].y * e0.z;
      else if (i == 1) ra = abs1To0[j].z * e0.x + abs1To0[j].x * e0.z;
      else ra = abs1To0[j].y * e0.x + abs1To0[j].x * e0.y;
      if (j == 0) rb = abs0To1[i].z * e1.y + abs0To1[i].y * e1.z;
      else if (j == 1) rb = abs0To1[i].z * e1.x + abs0To1[i].x * e1.z;
      else rb = abs0To1[i].y * e1.x + abs0To1[i].x * e1.y;
      if (absSign > ((ra + rb) + contactDist)) return false;
    }
  }
  int feature = 0; float minOverlap = overlap[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) if (minOverlap > overlap[i]) { minOverlap = overlap[i]; feature = i; }

  float sgn = sign[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) if (feature == i) sgn = sign[i];
  const bool neg = 0.f >= sgn;
  const bool flip = feature >= 3;
  const mxf& tr = flip ? t1 : t0;          // transform owning the reference face
  const v3 er = flip ? e1 : e0;
  const int ax = flip ? feature - 3 : feature;
  const v3 A0 = tr.r.c0, A1 = tr.r.c1, A2 = tr.r.c2;
  mxf nt; v3 mtd; float exx, eyy;
  if (!flip) {
    if (ax == 0) {
      exx = er.z; eyy = er.y;
      if (neg) { mtd = A0; nt.r.c0 = -A2; nt.r.c1 = A1; nt.r.c2 = A0; nt.p = negscalesub(A0, er.x, tr.p); }
      else { mtd = -A0; nt.r.c0 = A2; nt.r.c1 = A1; nt.r.c2 = mtd; nt.p = scaleadd(A0, er.x, tr.p); }
    } else if (ax == 1) {
      exx = er.x; eyy = er.z;
      if (neg) { mtd = A1; nt.r.c0 = A0; nt.r.c1 = -A2; nt.r.c2 = A1; nt.p = negscalesub(A1, er.y, tr.p); }
      else { mtd = -A1; nt.r.c0 = A0; nt.r.c1 = A2; nt.r.c2 = mtd; nt.p = scaleadd(A1, er.y, tr.p); }
    } else {
      exx = er.x; eyy = er.y;
      if (neg) { mtd = A2; nt.r.c0 = A0; nt.r.c1 = A1; nt.r.c2 = A2; nt.p = negscalesub(A2, er.z, tr.p); }
      else { mtd = -A2; nt.r.c0 = A0; nt.r.c1 = -A1; nt.r.c2 = mtd; nt.p = scaleadd(A2, er.z, tr.p); }
    }
  } else {
    if (ax == 0) {
      exx = er.z; eyy = er.y;
      if (neg) { mtd = A0; nt.r.c0 = A2; nt.r.c1 = A1; nt.r.c2 = -A0; nt.p = scaleadd(A0, er.x, tr.p); }
      else { mtd = -A0; nt.r.c0 = -A2; nt.r.c1 = A1; nt.r.c2 = A0; nt.p = negscalesub(A0, er.x, tr.p); }
    } else if (ax == 1) {
      exx = er.x; eyy = er.z;
      if (neg) { mtd = A1; nt.r.c0 = A0; nt.r.c1 = A2; nt.r.c2 = -A1; nt.p = scaleadd(A1, er.y, tr.p); }
      else { mtd = -A1; nt.r.c0 = A0; nt.r.c1 = -A2; nt.r.c2 = A1; nt.p = negscalesub(A1, er.y, tr.p); }
    } else {
      exx = er.x; eyy = er.y;
      if (neg) { mtd = A2; nt.r.c0 = A0; nt.r.c1 = -A1; nt.r.c2 = -A2; nt.p = scaleadd(A2, er.z, tr.p); }
      else { mtd = -A2; nt.r.c0 = A0; nt.r.c1 = A1; nt.r.c2 = A2; nt.p = negscalesub(A2, er.z, tr.p); }
    }
  }
  const v3 localNormal = amtmul(nt.r, mtd);
  v3 pts[4]; v3 incN;
  if (!flip) { const mxf o = amxfinvmul(nt, t1); incident_polygon(pts, incN, -localNormal, o, e1); }
  else { const mxf o = amxfinvmul(nt, t0); incident_polygon(pts, incN, localNormal, o, e0); }
  calc_contacts(exx, eyy, pts, incN, localNormal, mc, numContacts, contactDist);
  const int n = numContacts;
  if (n != 0) {
    if (flip) for (int i = 0; i < n; ++i) { const v3 lb = mc[i].b; mc[i].b = mc[i].a; mc[i].a = lb; }
    const mxf newTo1 = amxfinvmul(t1, nt), newTo0 = amxfinvmul(t0, nt);
    const v3 nInB = mmul(newTo1.r, mc[0].n);
    for (int i = 0; i < n; ++i) { mc[i].a = amxftransform(newTo0, mc[i].a); mc[i].b = amxftransform(newTo1, mc[i].b); mc[i].n = nInB; }
  }
  return true;
}

// GuPCMContactBoxBox.cpp:848-971, in two parts so that the device-wide path can run the (expensive, divergent) regeneration over a compacted worklist:
// pcm_box_box_refresh: manifold refresh + invalidation test; returns true when the manifold has to be regenerated (frames already updated), else emits the cached points.
PXB_D bool pcm_box_box_refresh(const xf& tm0, const xf& tm1, v3 e0, v3 e1, float contactDist, float toleranceLength, Manifold& man, Contacts& out) {
  const xf cur = axfinvmul(tm1, tm0);  // A into B
  const mxf aToB = amxffromxf(cur);
  const float minMargin = fmin_(box_margin(e0, toleranceLength), box_margin(e1, toleranceLength));
  const int initial = man.n;
  manifold_refresh(man, aToB, minMargin * 0.8f);
  const bool lost = man.n != initial;
  const float radiusA = alen(e0), radiusB = alen(e1);
  out.count = 0; out.normal = V3(0, 0, 0);
  if (lost || invalidate_boxconvex(man, cur, tm0.q, tm1.q, minMargin, radiusA, radiusB)) {
    man.rel = cur; man.quatA = tm0.q; man.quatB = tm1.q; man.dirty = 1;
    return true;
  }
  if (man.n > 0) {
    out.normal = manifold_world_normal(man, tm1);
    for (int i = 0; i < man.n; ++i) {
      const float dist = man.pts[i].pen;
      if (contactDist >= dist) { out.point[out.count] = axftransform(tm1, man.pts[i].b); out.sep[out.count] = dist; out.count++; }
    }
  }
  return false;
}
// pcm_box_box_generate: SAT + clipping + reduction.  Returns true when the SAT passed but clipping found no point: the caller then runs the GJK / EPA single-point fallback (pxb_gjk.cuh)
PXB_D bool pcm_box_box_generate(const xf& tm0, const xf& tm1, v3 e0, v3 e1, float contactDist, float toleranceLength, Manifold& man, Contacts& out) {
  out.count = 0; out.normal = V3(0, 0, 0);
  mxf tv0 = amxffromxf(tm0), tv1 = amxffromxf(tm1);
  tv0.r.c0 = anormalize(tv0.r.c0); tv0.r.c1 = anormalize(tv0.r.c1); tv0.r.c2 = anormalize(tv0.r.c2);
  tv1.r.c0 = anormalize(tv1.r.c0); tv1.r.c1 = anormalize(tv1.r.c1); tv1.r.c2 = anormalize(tv1.r.c2);
  MPoint mc[16]; int num = 0;
  if (boxbox_generate(e0, e1, tv0, tv1, contactDist, mc, num)) {
    if (num > 0) {
      if (num <= PXB_MANIFOLD_CACHE) { for (int i = 0; i < num; ++i) man.pts[i] = mc[i]; man.n = num; }
      else { reduce_batch(man, mc, num, toleranceLength); man.n = PXB_MANIFOLD_CACHE; }
      out.normal = anormalize(mmul(tv1.r, man.pts[0].n));
      for (int i = 0; i < man.n; ++i) { out.point[out.count] = amxftransform(tv1, man.pts[i].b); out.sep[out.count] = man.pts[i].pen; out.count++; }
    } else return true;
  }
  return false;
}
PXB_D bool pcm_box_box(const xf& tm0, const xf& tm1, v3 e0, v3 e1, float contactDist, float toleranceLength, Manifold& man, Contacts& out) {
  if (!pcm_box_box_refresh(tm0, tm1, e0, e1, contactDist, toleranceLength, man, out)) return false;
  return pcm_box_box_generate(tm0, tm1, e0, e1, contactDist, toleranceLength, man, out);
}

// ---------------- sphere / capsule family (reference GPU kernel: sphereNphase_Kernel) ----------------
// CPU PCM semantics: GuPCMContactSphereSphere.cpp:36-69, GuPCMContactSpherePlane.cpp:36-73, GuPCMContactSphereCapsule.cpp:38-102,
// GuPCMContactSphereBox.cpp:36-131, GuPCMContactPlaneCapsule.cpp:36-123, GuPCMContactCapsuleCapsule.cpp:38-275,
// GuDistanceSegmentSegment.cpp:411-469.
PXB_D void np_sphere_sphere(v3 p0, v3 p1, float r0, float r1, float cDist, Contacts& out) {
  const v3 delta = p0 - p1;
  const float distanceSq = adot(delta, delta);
  const float radiusSum = r0 + r1, inflatedSum = radiusSum + cDist;
  if (inflatedSum * inflatedSum > distanceSq) {
    const float dist = sqrtf(distanceSq);
    const v3 normal = (0.00001f >= dist) ? V3(1.f, 0.f, 0.f) : V3(delta.x / dist, delta.y / dist, delta.z / dist);
    out.normal = normal; out.point[0] = scaleadd(normal, r1, p1); out.sep[0] = dist - radiusSum; out.count = 1;
  }
}
PXB_D void np_sphere_plane(v3 p0, float radius, const xf& planeTm, float cDist, Contacts& out) {
  const v3 c = aqrotinv(planeTm.q, p0 - planeTm.p);
  const float separation = c.x - radius;
  if (cDist >= separation) { const v3 n = aqbasis0(planeTm.q); out.normal = n; out.point[0] = negscalesub(n, radius, p0); out.sep[0] = separation; out.count = 1; }
}
PXB_D float dist_point_segment_sq(v3 a, v3 b, v3 p, float& param) {
  const v3 ap = p - a, ab = b - a;
  const float nom = adot(ap, ab), denom = adot(ab, ab);
  const float tValue = fmax_(fmin_(nom / denom, 1.f), 0.f);
  const float t = (denom == 0.f) ? 0.f : tValue;
  const v3 v = negscalesub(ab, t, ap);
  param = t;
  return adot(v, v);
}
PXB_D void np_sphere_capsule(v3 sphereCenter, float sphereRadius, const xf& capTm, float capRadius, float halfHeight, float cDist, Contacts& out) {
  const v3 tmp0 = aqbasis0(capTm.q) * halfHeight;
  const v3 s = capTm.p + tmp0, e = capTm.p - tmp0;
  const float radiusSum = sphereRadius + capRadius, inflatedSum = radiusSum + cDist;
  float t; const float squareDist = dist_point_segment_sq(s, e, sphereCenter, t);
  if (inflatedSum * inflatedSum > squareDist) {
    const v3 p = scaleadd(e - s, t, s);
    const v3 dir = sphereCenter - p;
    const float len = alen(dir);
    const v3 normal = (len > FLT_EPSILON) ? V3(dir.x / len, dir.y / len, dir.z / len) : V3(1.f, 0.f, 0.f);
    out.normal = normal; out.point[0] = negscalesub(normal, sphereRadius, sphereCenter); out.sep[0] = sqrtf(squareDist) - radiusSum; out.count = 1;
  }
}
PXB_D void np_sphere_box(v3 sphereOrigin, float radius, const xf& boxTm, v3 be, float cDist, Contacts& out) {
  const v3 c = aqrotinv(boxTm.q, sphereOrigin - boxTm.p);
  const float inflatedSum = radius + cDist;
  const v3 p = V3(fmax_(fmin_(c.x, be.x), -be.x), fmax_(fmin_(c.y, be.y), -be.y), fmax_(fmin_(c.z, be.z), -be.z));
  const v3 v = c - p;
  const float lengthSq = adot(v, v);
  if (inflatedSum * inflatedSum > lengthSq) {
    const v3 ac = vabs(c);
    if (be.x >= ac.x && be.y >= ac.y && be.z >= ac.z) {
      const v3 d = be - vabs(p);
      const bool con0 = d.x >= d.z && d.y >= d.z, con1 = d.y >= d.x && d.z >= d.x;
      const v3 sign = V3(p.x >= 0.f ? 1.f : -1.f, p.y >= 0.f ? 1.f : -1.f, p.z >= 0.f ? 1.f : -1.f);
      const v3 locNorm = con0 ? V3(0.f * sign.x, 0.f * sign.y, 1.f * sign.z) : (con1 ? V3(1.f * sign.x, 0.f * sign.y, 0.f * sign.z) : V3(0.f * sign.x, 1.f * sign.y, 0.f * sign.z));
      const float dist = -(con0 ? d.z : (con1 ? d.x : d.y));
      const v3 normal = aqrot(boxTm.q, locNorm);
      out.normal = normal; out.point[0] = sphereOrigin - normal * dist; out.sep[0] = dist - radius; out.count = 1;
    } else {
      const float recipLength = 1.0f / sqrtf(lengthSq);
      const float length = 1.0f / recipLength;
      out.normal = aqrot(boxTm.q, v * recipLength); out.point[0] = axftransform(boxTm, p); out.sep[0] = length - radius; out.count = 1;
    }
  }
}
PXB_D void add_manifold_point2(Manifold& m, v3 la, v3 lb, v3 n, float pen, float replaceBreakingThreshold) {
  const float shortest = replaceBreakingThreshold * replaceBreakingThreshold;
  for (int i = 0; i < m.n; ++i) {
    const v3 dB = m.pts[i].b - lb, dA = m.pts[i].a - la;
    if (shortest > fmin_(adot(dB, dB), adot(dA, dA))) { m.pts[i].a = la; m.pts[i].b = lb; m.pts[i].n = n; m.pts[i].pen = pen; return; }
  }
  if (m.n < 2) { m.pts[m.n].a = la; m.pts[m.n].b = lb; m.pts[m.n].n = n; m.pts[m.n].pen = pen; m.n++; return; }
  const v3 v0 = m.pts[0].b - lb, v1 = m.pts[1].b - lb;
  const int k = (adot(v0, v0) > adot(v1, v1)) ? 1 : 0;
  m.pts[k].a = la; m.pts[k].b = lb; m.pts[k].n = n; m.pts[k].pen = pen;
}
PXB_D void pcm_plane_capsule(const xf& planeTm, const xf& capTm, float radius, float halfHeight, float contactDist, Manifold& man, Contacts& out) {
  const xf aToB = axfinvmul(planeTm, capTm);
  const v3 planeNormal = anormalize(aqbasis0(planeTm.q));
  const v3 ln = V3(1.f, 0.f, 0.f);
  const v3 tmp = aqbasis0(aToB.q) * halfHeight;
  const v3 s = aToB.p + tmp, e = aToB.p - tmp;
  const float inflatedRadius = radius + contactDist;
  const int initial = man.n;
  const mxf aToBm = amxffromxf(aToB);
  manifold_refresh(man, aToBm, radius * 0.05f);
  const bool lost = man.n != initial;
  if (lost || invalidate_plane(man, aToB, radius, 0.02f)) {
    man.n = 0; man.rel = aToB; man.dirty = 1;
    if (inflatedRadius > s.x) add_manifold_point2(man, aqrotinv(aToB.q, s - aToB.p), negscalesub(ln, s.x, s), ln, s.x, radius * 0.001f);
    if (inflatedRadius > e.x) add_manifold_point2(man, aqrotinv(aToB.q, e - aToB.p), negscalesub(ln, e.x, e), ln, e.x, radius * 0.001f);
  }
  out.count = 0; out.normal = -planeNormal;
  for (int i = 0; i < man.n; ++i) {
    const float dist = man.pts[i].pen - radius;
    if (contactDist >= dist) { out.point[out.count] = negscalesub(planeNormal, radius, axftransform(capTm, man.pts[i].a)); out.sep[out.count] = dist; out.count++; }
  }
}
PXB_D float dist_seg_seg_sq(v3 p1, v3 d1, v3 p2, v3 d2, float& s, float& t) {
  const float eps = FLT_EPSILON;
  const v3 r = p1 - p2;
  const float a = dot(d1, d1), e = dot(d2, d2), b = dot(d1, d2), c = dot(d1, r);
  const float aRecip = a > eps ? 1.0f / a : 0.f, eRecip = e > eps ? 1.0f / e : 0.f;
  const float f = adot(d2, r);
  const float denom = a * e - b * b;
  const float temp = b * f - c * e;
  const float s0 = fmax_(fmin_(temp / denom, 1.f), 0.f);
  const float sTmp = (eps > denom) ? 0.5f : s0;
  const float tTmp = (b * sTmp + f) * eRecip;
  const float t2 = fmax_(fmin_(tTmp, 1.f), 0.f);
  const float comp = (b * t2 - c) * aRecip;
  const float s2 = fmax_(fmin_(comp, 1.f), 0.f);
  s = s2; t = t2;
  const v3 vv = scaleadd(d1, s2, p1) - scaleadd(d2, t2, p2);
  return adot(vv, vv);
}
// Contacts of one capsule pair share the first contact's normal (they fall into one patch when their normals
// agree within PXC_SAME_NORMAL, which holds for the near-parallel case that produces several contacts).
PXB_D void np_capsule_capsule(const xf& tm0, const xf& tm1, float r0, float hh0, float r1, float hh1, float cDist, Contacts& out) {
  const v3 positionOffset = (tm0.p + tm1.p) * 0.5f;
  const v3 p0 = tm0.p - positionOffset, p1 = tm1.p - positionOffset;
  const v3 tmp0 = aqbasis0(tm0.q) * hh0;
  const v3 s0 = p0 + tmp0, e0 = p0 - tmp0, d0 = e0 - s0;
  const v3 tmp1 = aqbasis0(tm1.q) * hh1;
  const v3 s1 = p1 + tmp1, e1 = p1 - tmp1, d1 = e1 - s1;
  const float sumRadius = r0 + r1, inflatedSum = sumRadius + cDist, inflatedSumSquared = inflatedSum * inflatedSum;
  const float a = adot(d0, d0), e = adot(d1, d1), eps = 1e-6f;
  float t0, t1;
  const float sqDist0 = dist_seg_seg_sq(s0, d0, s1, d1, t0, t1);
  if (!(inflatedSumSquared >= sqDist0)) return;
  const float sa = sqrtf(a), se = sqrtf(e);
  const v3 dir0 = (eps > a) ? V3(0, 0, 0) : V3(d0.x / sa, d0.y / sa, d0.z / sa);
  const v3 dir1 = (eps > e) ? V3(0, 0, 0) : V3(d1.x / se, d1.y / se, d1.z / se);
  if (fabsf(adot(dir0, dir1)) > 0.9998f) {
    const v3 ab0 = e0 - s0, ab1 = e1 - s1;
    const float den0 = adot(ab0, ab0), den1 = adot(ab1, ab1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float nom, den;
      if (k == 0) { nom = dot(s1 - s0, ab0); den = den0; } else if (k == 1) { nom = dot(e1 - s0, ab0); den = den0; }
      else if (k == 2) { nom = dot(s0 - s1, ab1); den = den1; } else { nom = dot(e0 - s1, ab1); den = den1; }
      const float t = (den == 0.f) ? 0.f : nom / den;
      if (!(t >= 0.f && 1.f >= t)) continue;
      v3 proj, v, base;
      if (k == 0) { proj = scaleadd(d0, t, s0); v = proj - s1; base = proj; }
      else if (k == 1) { proj = scaleadd(d0, t, s0); v = proj - e1; base = proj; }
      else if (k == 2) { proj = scaleadd(d1, t, s1); v = s0 - proj; base = s0; }
      else { proj = scaleadd(d1, t, s1); v = e0 - proj; base = e0; }
      const float sqDist = adot(v, v);
      if (sqDist > eps && inflatedSumSquared > sqDist) {
        const float dist = sqrtf(sqDist);
        const v3 normal = V3(v.x / dist, v.y / dist, v.z / dist);
        if (out.count == 0) out.normal = normal;
        out.point[out.count] = negscalesub(normal, r0, base) + positionOffset; out.sep[out.count] = dist - sumRadius; out.count++;
      }
    }
    if (out.count) return;
  }
  const v3 closestA = scaleadd(d0, t0, s0), closestB = scaleadd(d1, t1, s1);
  const bool con = eps > sqDist0;
  const v3 nrm = anormalize(con ? ((a > eps) ? d0 : V3(1.f, 0.f, 0.f)) : (closestA - closestB));
  out.normal = nrm; out.point[0] = negscalesub(nrm, r0, closestA) + positionOffset;
  out.sep[0] = (con ? 0.f : sqrtf(sqDist0)) - sumRadius; out.count = 1;
}
