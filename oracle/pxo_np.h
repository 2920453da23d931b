/* pxo_np.h -- CPU restatement of the reference's PCM narrowphase for the primitive pairs on the hot
 * path (TEST INFRASTRUCTURE).  Each function cites the reference code it follows.
 *   persistent manifold:  physx/source/geomutils/src/pcm/GuPersistentContactManifold.{h,cpp}
 *   plane-box:            physx/source/geomutils/src/pcm/GuPCMContactPlaneBox.cpp:36-209
 *   box-box:              physx/source/geomutils/src/pcm/GuPCMContactBoxBox.cpp:42-971
 * The GJK/EPA single-point fallback of pcmContactBoxBox (GuPCMContactBoxBox.cpp:918-958, taken when the SAT passes but face
 * clipping yields no point) lives in pxo_gjk.h (pxo_boxbox_gjk_fallback); pxo_pcm_box_box returns 2 to ask for it.
 * The reference's V3RecipFast (_mm_rcp_ps, 12-bit) in intersectSegmentAABB is restated as an exact
 * reciprocal; contact points produced by edge clipping therefore agree to ~4e-4 relative, not bitwise. */
#ifndef PXO_NP_H
#define PXO_NP_H
#include "pxo_math.h"
#if defined(PXO_REF_RCPPS) && defined(__SSE__)
#include <xmmintrin.h>
#endif

#define PXO_MANIFOLD_CACHE 4
#define PXO_MAX_CONTACTS 8   /* per-pair output capacity; the primitive PCM paths emit <= 4 */

typedef struct { v3 a, b; v3 n; float pen; } PxoMPoint; /* mLocalPointA, mLocalPointB, mLocalNormalPen */

typedef struct {
  int n;            /* mNumContacts */
  xf rel;           /* mRelativeTransform (p = FLT_MAX when invalid) */
  q4 quatA, quatB;  /* mQuatA / mQuatB */
  PxoMPoint pts[PXO_MANIFOLD_CACHE];
  uint8_t aInd[4], bInd[4], nWarm;   /* mAIndice / mBIndice / mNumWarmStartPoints: GJK warm start (pxo_gjk.h) */
} PxoManifold;

typedef struct {
  int count;
  v3 normal;                 /* shared by all points (single patch); points from shape1 towards shape0 */
  v3 point[PXO_MAX_CONTACTS];
  float sep[PXO_MAX_CONTACTS];
} PxoContacts;

static inline void pxo_manifold_init(PxoManifold* m) {
  memset(m, 0, sizeof(*m));
  m->rel.q = Q4(0, 0, 0, 1); m->rel.p = V3(FLT_MAX, FLT_MAX, FLT_MAX);
  m->quatA = Q4(0, 0, 0, 1); m->quatB = Q4(0, 0, 0, 1);
}

/* GuVecBox.h:81-88 */
static inline float pxo_box_margin(v3 ext, float toleranceLength) {
  const float mn = fminf_(ext.x, fminf_(ext.y, ext.z));
  return fminf_(mn * 0.15f, toleranceLength * 0.15f);
}

/* GuPersistentContactManifold.h:723-752 */
static inline void pxo_refresh(PxoManifold* m, const mxf* aToB, float projectBreakingThreshold) {
  const float sq = projectBreakingThreshold * projectBreakingThreshold;
  for (int i = m->n; i > 0; --i) {
    PxoMPoint* mp = &m->pts[i - 1];
    const v3 localAInB = amxftransform(aToB, mp->a);
    const v3 localBInB = mp->b;
    const v3 v = v3sub(localAInB, localBInB);
    const v3 ln = mp->n;
    const float dist = adot(v, ln);
    const v3 projected = v3negscalesub(ln, dist, localAInB);
    const v3 diff = v3sub(localBInB, projected);
    const float d2 = adot(diff, diff);
    if (d2 > sq) { m->n--; m->pts[i - 1] = m->pts[m->n]; }
    else mp->pen = dist;
  }
}

static inline float pxo_max_pos_delta(const PxoManifold* m, v3 curP) {
  const v3 d = v3abs(v3sub(curP, m->rel.p));
  return fmaxf_(d.x, fmaxf_(d.y, d.z));
}

/* GuPersistentContactManifold.h:245-257 */
static inline int pxo_invalidate_plane(const PxoManifold* m, const xf* cur, float minMargin, float ratio) {
  const float thresholdP = minMargin * ratio;
  const float deltaP = pxo_max_pos_delta(m, cur->p);
  const float deltaQ = adot4(cur->q, m->rel.q);
  return (deltaP > thresholdP) || (0.99996f > deltaQ);
}

/* GuPersistentContactManifold.h:190-223, thresholds .cpp:175,183 */
static inline int pxo_invalidate_boxconvex(const PxoManifold* m, const xf* cur, q4 quatA, q4 quatB, float minMargin, float radiusA, float radiusB) {
  static const float thr[5] = {0.5f, 0.125f, 0.25f, 0.375f, 0.375f};
  static const float qthr[5] = {0.9998f, 0.9999f, 0.9999f, 0.9999f, 0.9999f};
  const float thresholdP = minMargin * thr[m->n];
  const float deltaP = pxo_max_pos_delta(m, cur->p);
  const float thresholdQ = qthr[m->n];
  const float dqA = adot4(quatA, m->quatA), dqB = adot4(quatB, m->quatB);
  int gen = (deltaP > thresholdP) || (thresholdQ > dqA) || (thresholdQ > dqB);
  if (!gen) {
    const float aRad = dqA < 1.0f ? acosf(dqA) : 0.f;
    const float bRad = dqB < 1.0f ? acosf(dqB) : 0.f;
    gen = (aRad * radiusA > thresholdP) || (bRad * radiusB > thresholdP);
  }
  return gen;
}

/* GuPersistentContactManifold.cpp:859-1005 (reduceBatchContactsCluster) */
static inline void pxo_reduce_cluster(PxoManifold* m, const PxoMPoint* p, int numPoints) {
  int chosen[64]; memset(chosen, 0, sizeof(chosen));
  float maxDist = FLT_MAX; int index = 0; int indices[4];
  for (int i = 0; i < numPoints; ++i) if (maxDist > p[i].pen) { maxDist = p[i].pen; index = i; }
  m->pts[0] = p[index]; chosen[index] = 1; indices[0] = index;
  v3 v = v3sub(p[0].b, m->pts[0].b); maxDist = adot(v, v); index = 0;
  for (int i = 1; i < numPoints; ++i) { v = v3sub(p[i].b, m->pts[0].b); const float d = adot(v, v); if (d > maxDist) { maxDist = d; index = i; } }
  m->pts[1] = p[index]; chosen[index] = 1; indices[1] = index;
  maxDist = -FLT_MAX; index = -1;
  v = v3sub(m->pts[1].b, m->pts[0].b);
  const v3 cn0 = m->pts[0].n;
  v3 norm = v3cross(v, cn0);
  const float sqLen = adot(norm, norm);
  if (sqLen > 0.f) { const float l = sqrtf(sqLen); norm = V3(norm.x / l, norm.y / l, norm.z / l); } else norm = cn0;
  float minDist = FLT_MAX; int index1 = -1;
  for (int i = 0; i < numPoints; ++i) if (!chosen[i]) {
    v = v3sub(p[i].b, m->pts[0].b); const float d = adot(v, norm);
    if (d > maxDist) { maxDist = d; index = i; }
    if (minDist > d) { minDist = d; index1 = i; }
  }
  m->pts[2] = p[index]; chosen[index] = 1; indices[2] = index;
  if (minDist * maxDist > 0.f) {
    maxDist = -FLT_MAX;
    for (int i = 0; i < numPoints; ++i) if (!chosen[i]) {
      v = v3sub(p[i].b, m->pts[0].b); const float d = adot(v, norm);
      if (d > maxDist) { maxDist = d; index1 = i; }
    }
  }
  m->pts[3] = p[index1]; chosen[index1] = 1; indices[3] = index1;
  for (int i = 0; i < numPoints; ++i) if (!chosen[i]) {
    maxDist = FLT_MAX; const float pen = p[i].pen; index = 0;
    for (int j = 0; j < 4; ++j) { const v3 v1 = v3sub(p[i].b, m->pts[j].b); const float dist = adot(v1, v1); if (maxDist > dist) { maxDist = dist; index = j; } }
    if (p[indices[index]].pen > pen) indices[index] = i;
  }
  for (int k = 0; k < 4; ++k) m->pts[k] = p[indices[k]];
}

/* GuPersistentContactManifold.cpp:1008-1174 (reduceBatchContacts) */
static inline void pxo_reduce_batch(PxoManifold* m, const PxoMPoint* p, int numPoints, float toleranceLength) {
  uint8_t chosenIdx[4]; uint8_t cand[64];
  float maxPen = p[0].pen, minPen = maxPen;
  int index = 0; cand[0] = 0; int candIndex = 0; int nbCand = numPoints;
  for (int i = 1; i < numPoints; ++i) {
    cand[i] = (uint8_t)i;
    const float pen = p[i].pen;
    minPen = fmaxf_(minPen, pen);
    if (maxPen > pen) { maxPen = pen; index = i; candIndex = i; }
  }
  chosenIdx[0] = (uint8_t)index;
  nbCand--; cand[candIndex] = cand[nbCand];
  v3 v = v3sub(p[cand[0]].b, p[chosenIdx[0]].b);
  float maxDist = adot(v, v); index = cand[0]; candIndex = 0;
  for (int i = 1; i < nbCand; ++i) {
    v = v3sub(p[cand[i]].b, p[chosenIdx[0]].b); const float d = adot(v, v);
    if (d > maxDist) { maxDist = d; index = cand[i]; candIndex = i; }
  }
  chosenIdx[1] = (uint8_t)index;
  nbCand--; cand[candIndex] = cand[nbCand];
  v = v3sub(p[chosenIdx[1]].b, p[chosenIdx[0]].b);
  const v3 cn0 = p[chosenIdx[0]].n;
  v3 norm = v3cross(v, cn0);
  const float sqLen = adot(norm, norm);
  if (sqLen > 0.f) { const float l = sqrtf(sqLen); norm = V3(norm.x / l, norm.y / l, norm.z / l); } else norm = cn0;
  maxDist = -FLT_MAX; index = 0xff; candIndex = 0xff;
  float minDist = FLT_MAX; int index1 = 0xff, candIndex1 = 0xff;
  for (int i = 0; i < nbCand; ++i) {
    v = v3sub(p[cand[i]].b, p[chosenIdx[0]].b); const float d = adot(v, norm);
    if (d > maxDist) { maxDist = d; index = cand[i]; candIndex = i; }
    if (minDist > d) { minDist = d; index1 = cand[i]; candIndex1 = i; }
  }
  chosenIdx[2] = (uint8_t)index;
  nbCand--; cand[candIndex] = cand[nbCand];
  if (nbCand == candIndex1) candIndex1 = candIndex;
  if (minDist * maxDist > 0.f) {
    maxDist = -FLT_MAX;
    for (int i = 0; i < nbCand; ++i) {
      v = v3sub(p[cand[i]].b, p[chosenIdx[0]].b); const float d = adot(v, norm);
      if (d > maxDist) { maxDist = d; index1 = cand[i]; candIndex1 = i; }
    }
  }
  chosenIdx[3] = (uint8_t)index1;
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
        if (candIndex < nbCand) { const uint8_t orig = chosenIdx[i]; chosenIdx[i] = cand[candIndex]; cand[candIndex] = orig; }
      }
      m->pts[i] = p[chosenIdx[i]];
    }
  } else {
    for (int i = 0; i < 4; ++i) m->pts[i] = p[chosenIdx[i]];
  }
}

/* GuPersistentContactManifold.h:695-708 */
static inline v3 pxo_world_normal(const PxoManifold* m, const xf* trB) {
  v3 n = m->pts[0].n;
  for (int i = 1; i < m->n; ++i) n = v3add(n, m->pts[i].n);
  const float sq = adot(n, n);
  const v3 nn = (sq > FLT_EPSILON) ? n : m->pts[0].n;
  return aqrot_normalize(trB->q, nn);
}

/* ---- plane vs box (shape0 = plane, shape1 = box) : GuPCMContactPlaneBox.cpp:36-209 ---- */
static inline int pxo_pcm_plane_box(const xf* planeTm, const xf* boxTm, v3 boxExtents, float contactDist, float toleranceLength,
                                    PxoManifold* manifold, PxoContacts* out) {
  const xf* transf0 = boxTm; const xf* transf1 = planeTm;
  const xf curTransf = axfinvmul(transf1, transf0); /* box to plane */
  const v3 negPlaneNormal = anormalize(v3neg(aqbasis0(transf1->q)));
  const float boxMargin = pxo_box_margin(boxExtents, toleranceLength);
  const float projectBreakingThreshold = boxMargin * 0.2f;
  const int initialContacts = manifold->n;
  const mxf aToB = amxffromxf(&curTransf);
  pxo_refresh(manifold, &aToB, projectBreakingThreshold);
  const int bLost = manifold->n != initialContacts;
  out->count = 0; out->normal = negPlaneNormal;
  if (bLost || pxo_invalidate_plane(manifold, &curTransf, boxMargin, 0.2f)) {
    const v3 localNormal = V3(1.f, 0.f, 0.f);
    manifold->n = 0;
    manifold->rel = curTransf;
    const float bx = boxExtents.x, by = boxExtents.y, bz = boxExtents.z;
    const v3 temp0 = v3scale(aToB.r.c0, bx), temp1 = v3scale(aToB.r.c1, by), temp2 = v3scale(aToB.r.c2, bz);
    const v3 ntemp2 = v3neg(temp2);
    const float px = aToB.p.x;
    const v3 temp01 = v3add(temp0, temp1), temp02 = v3sub(temp0, temp1);
    const float s[8] = {v3add(temp2, temp01).x, v3add(ntemp2, temp01).x, v3add(temp2, temp02).x, v3add(ntemp2, temp02).x,
                        v3sub(temp2, temp02).x, v3sub(ntemp2, temp02).x, v3sub(temp2, temp01).x, v3sub(ntemp2, temp01).x};
    const v3 corner[8] = {V3(bx, by, bz), V3(bx, by, -bz), V3(bx, -by, bz), V3(bx, -by, -bz),
                          V3(-bx, by, bz), V3(-bx, by, -bz), V3(-bx, -by, bz), V3(-bx, -by, -bz)};
    const float acceptanceDist = contactDist - px;
    PxoMPoint mc[8]; int num = 0;
    for (int k = 0; k < 8; ++k) if (acceptanceDist > s[k]) {
      const float pen = s[k] + px;
      mc[num].a = corner[k];
      mc[num].b = v3negscalesub(localNormal, pen, amxftransform(&aToB, corner[k]));
      mc[num].n = localNormal; mc[num].pen = pen; num++;
    }
    if (num <= PXO_MANIFOLD_CACHE) { for (int i = 0; i < num; ++i) manifold->pts[i] = mc[i]; manifold->n = num; }
    else { pxo_reduce_cluster(manifold, mc, num); manifold->n = PXO_MANIFOLD_CACHE; }
  }
  for (int i = 0; i < manifold->n; ++i) {
    const float dist = manifold->pts[i].pen;
    if (contactDist >= dist) {
      out->point[out->count] = axftransform(transf1, manifold->pts[i].b);
      out->sep[out->count] = dist; out->count++;
    }
  }
  return manifold->n > 0;
}

/* ---- box vs box ---- */
/* GuPCMContactBoxBox.cpp:42-118 */
static inline void pxo_incident_polygon(v3* pts, v3* faceNormal, v3 axis, const mxf* t1To0, v3 extents) {
  float ex = extents.x, ey = extents.y, ez = extents.z;
  const v3 u0 = t1To0->r.c0, u1 = t1To0->r.c1, u2 = t1To0->r.c2;
  const float d0 = adot(u0, axis), d1 = adot(u1, axis), d2 = adot(u2, axis);
  const float a0 = fabsf(d0), a1 = fabsf(d1), a2 = fabsf(d2);
  if (a0 >= a1 && a0 >= a2) {
    const int con = d0 > 0.f; *faceNormal = con ? v3neg(u0) : u0; ex = con ? -ex : ex;
    const v3 r0 = v3scale(u0, ex), r1 = v3scale(u1, ey), r2 = v3scale(u2, ez);
    const v3 t0 = v3add(t1To0->p, r0), t1 = v3add(r1, r2), t2 = v3sub(r1, r2);
    pts[0] = v3add(t0, t1); pts[1] = v3add(t0, t2); pts[2] = v3sub(t0, t1); pts[3] = v3sub(t0, t2);
  } else if (a1 >= a2) {
    const int con = d1 > 0.f; *faceNormal = con ? v3neg(u1) : u1; ey = con ? -ey : ey;
    const v3 r0 = v3scale(u0, ex), r1 = v3scale(u1, ey), r2 = v3scale(u2, ez);
    const v3 t0 = v3add(t1To0->p, r1), t1 = v3add(r0, r2), t2 = v3sub(r0, r2);
    pts[0] = v3add(t0, t1); pts[1] = v3add(t0, t2); pts[2] = v3sub(t0, t1); pts[3] = v3sub(t0, t2);
  } else {
    const int con = d2 > 0.f; *faceNormal = con ? v3neg(u2) : u2; ez = con ? -ez : ez;
    const v3 r0 = v3scale(u0, ex), r1 = v3scale(u1, ey), r2 = v3scale(u2, ez);
    const v3 t0 = v3add(t1To0->p, r2), t1 = v3add(r0, r1), t2 = v3sub(r0, r1);
    pts[0] = v3add(t0, t1); pts[1] = v3add(t0, t2); pts[2] = v3sub(t0, t1); pts[3] = v3sub(t0, t2);
  }
}

/* GuPCMContactBoxBox.cpp:121-165 */
static inline int pxo_seg_aabb(v3 p0, v3 d, v3 mx, v3 mn, float* tmin, float* tmax) {
  const float eps = 1e-6f;
  const float pv[3] = {p0.x, p0.y, p0.z}, dv[3] = {d.x, d.y, d.z}, mxv[3] = {mx.x, mx.y, mx.z}, mnv[3] = {mn.x, mn.y, mn.z};
  int par[3];
  for (int k = 0; k < 3; ++k) {
    par[k] = eps > fabsf(dv[k]);
    const int outside = (pv[k] > mxv[k]) || (mnv[k] > pv[k]);
    if (par[k] && outside) return 0;
  }
  float ft1 = -FLT_MAX, ft2 = FLT_MAX;
  for (int k = 0; k < 3; ++k) {
#if defined(PXO_REF_RCPPS) && defined(__SSE__)
    const float odd = _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(dv[k]))); /* the reference's V3RecipFast on this host (pinning builds only) */
#else
    const float odd = 1.0f / dv[k]; /* reference: V3RecipFast (_mm_rcp_ps), see header note */
#endif
    const float t1 = par[k] ? 0.f : (mnv[k] - pv[k]) * odd;
    const float t2 = par[k] ? FLT_MAX : (mxv[k] - pv[k]) * odd;
    const float tt1 = fminf_(t1, t2), tt2 = fmaxf_(t1, t2);
    ft1 = fmaxf_(ft1, tt1); ft2 = fminf_(ft2, tt2);
  }
  const float tminf = fmaxf_(ft1, 0.f), tmaxf = fminf_(1.f, ft2);
  *tmin = tminf; *tmax = tmaxf;
  return !((tminf > tmaxf) || (tminf > 1.f));
}

/* GuPCMContactGenUtil.cpp:35-103 */
static inline int pxo_contains(const v3* verts, int numVerts, v3 p, v3 mn, v3 mx) {
  if ((mn.x > p.x) || (p.x > mx.x) || (mn.y > p.y) || (p.y > mx.y)) return 0;
  const float tx = p.x, ty = p.y; const float eps = FLT_EPSILON;
  int inter = 0;
  for (int i = 0, j = numVerts - 1; i < numVerts; j = i++) {
    const float jy = verts[j].y, iy = verts[i].y, jx = verts[j].x, ix = verts[i].x;
    if ((tx == jx && ty == jy) || (tx == ix && ty == iy)) return 1;
    const int yflag0 = jy > ty, yflag1 = iy > ty;
    if (yflag0 != yflag1) {
      const float jix = ix - jx, jiy = iy - jy, jty = ty - jy;
      const float part1 = jty * jix, part2 = (jx + eps) * jiy, part3 = tx * jiy;
      const int comp = jiy > 0.f;
      const float tmp = part1 + part2;
      const float comp1 = comp ? tmp : part3, comp2 = comp ? part3 : tmp;
      if (comp1 >= comp2) { if (inter == 1) return 0; inter++; }
    }
  }
  return inter > 0;
}

/* GuPCMContactBoxBox.cpp:168-330 */
static inline void pxo_calc_contacts(float extentX_, float extentY_, v3* pts, v3 incidentNormal, v3 localNormal,
                                     PxoMPoint* mc, int* numContacts, float contactDist) {
  const float extentX = extentX_ * 1.0001f, extentY = extentY_ * 1.0001f;
  const float nExtentX = -extentX, nExtentY = -extentY;
  int pPen[4], pArea[4];
  v3 bmin = V3(FLT_MAX, FLT_MAX, FLT_MAX), bmax = V3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  int n = *numContacts;
  for (int i = 0; i < 4; ++i) {
    bmin = v3min(bmin, pts[i]); bmax = v3max(bmax, pts[i]);
    const float z = -pts[i].z;
    if (contactDist > z) {
      pPen[i] = 1;
      const v3 ap = v3abs(pts[i]);
      if (extentX >= ap.x && extentY >= ap.y) { /* z bound is FLT_MAX */
        pArea[i] = 1;
        mc[n].a = V3(pts[i].x, pts[i].y, 0.f); mc[n].b = pts[i]; mc[n].n = localNormal; mc[n].pen = z; n++;
      } else pArea[i] = 0;
    } else { pPen[i] = 0; pArea[i] = 0; }
  }
  if (n == 4) { *numContacts = n; return; }
  {
    const float denom = incidentNormal.z;
    const v3 q[4] = {V3(extentX, extentY, 0.f), V3(extentX, nExtentY, 0.f), V3(nExtentX, extentY, 0.f), V3(nExtentX, nExtentY, 0.f)};
    for (int k = 0; k < 4; ++k) {
      if (pxo_contains(pts, 4, q[k], bmin, bmax)) {
        const float nom = adot(incidentNormal, v3sub(pts[0], q[k]));
        const float t = nom / denom; const float pen = -t;
        if (contactDist > pen) { mc[n].a = q[k]; mc[n].b = V3(q[k].x, q[k].y, t); mc[n].n = localNormal; mc[n].pen = pen; n++; }
      }
    }
  }
  const v3 ext = V3(extentX, extentY, FLT_MAX);
  const v3 negExt = V3(nExtentX, nExtentY, -(contactDist + FLT_EPSILON));
  for (int rStart = 0, rEnd = 3; rStart < 4; rEnd = rStart++) {
    const v3 p0 = pts[rStart], p1 = pts[rEnd];
    if (!pPen[rStart] && !pPen[rEnd]) continue;
    const int con0 = pPen[rStart] && pArea[rStart], con1 = pPen[rEnd] && pArea[rEnd];
    if (con0 && con1) continue;
    const v3 p0p1 = v3sub(p1, p0);
    float tmin, tmax;
    if (pxo_seg_aabb(p0, p0p1, ext, negExt, &tmin, &tmax)) {
      if (!con0) { const v3 ip = v3scaleadd(p0p1, tmin, p0); mc[n].a = V3(ip.x, ip.y, 0.f); mc[n].b = ip; mc[n].n = localNormal; mc[n].pen = -ip.z; n++; }
      if (!con1) { const v3 ip = v3scaleadd(p0p1, tmax, p0); mc[n].a = V3(ip.x, ip.y, 0.f); mc[n].b = ip; mc[n].n = localNormal; mc[n].pen = -ip.z; n++; }
    }
  }
  *numContacts = n;
}

static inline float pxo_sum3(v3 v) { return v.x + (v.y + v.z); }

/* GuPCMContactBoxBox.cpp:332-846.  Returns 0 when a separating axis is found. */
static inline int pxo_boxbox_generate(v3 e0, v3 e1, const mxf* t0, const mxf* t1, float contactDist, PxoMPoint* mc, int* numContacts) {
  const float ea[3] = {e0.x, e0.y, e0.z}, eb[3] = {e1.x, e1.y, e1.z};
  const mxf t1To0 = amxfinvmul(t0, t1);
  const m33 rot0To1 = m33transpose(&t1To0.r);
  const float uEps = 1e-6f;
  const float tx = t1To0.p.x, ty = t1To0.p.y, tz = t1To0.p.z;
  const v3 col[3] = {t1To0.r.c0, t1To0.r.c1, t1To0.r.c2};
  v3 abs1To0[3], abs0To1[3];
  const v3 r01[3] = {rot0To1.c0, rot0To1.c1, rot0To1.c2};
  for (int k = 0; k < 3; ++k) {
    abs1To0[k] = V3(fabsf(col[k].x) + uEps, fabsf(col[k].y) + uEps, fabsf(col[k].z) + uEps);
    abs0To1[k] = V3(fabsf(r01[k].x) + uEps, fabsf(r01[k].y) + uEps, fabsf(r01[k].z) + uEps);
  }
  float sign[6], overlap[6];
  const float tt[3] = {tx, ty, tz};
  for (int k = 0; k < 3; ++k) { /* ua0..ua2 */
    sign[k] = tt[k];
    const float rb = pxo_sum3(v3mul(abs0To1[k], e1));
    const float radiusSum = ea[k] + rb;
    overlap[k] = (radiusSum - fabsf(sign[k])) + contactDist;
    if (0.f > overlap[k]) return 0;
  }
  for (int k = 0; k < 3; ++k) { /* ub0..ub2 */
    sign[3 + k] = adot(t1To0.p, col[k]);
    const float ra = pxo_sum3(v3mul(abs1To0[k], e0));
    const float radiusSum = ra + eb[k];
    overlap[3 + k] = (radiusSum - fabsf(sign[3 + k])) + contactDist;
    if (0.f > overlap[3 + k]) return 0;
  }
  /* 9 edge-edge axes: rejection only (GuPCMContactBoxBox.cpp:443-617) */
#define C(v, i) ((i) == 0 ? (v).x : ((i) == 1 ? (v).y : (v).z))
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3; /* ua_i x ub_j */
    for (int j = 0; j < 3; ++j) {
      float absSign, ra, rb;
      /* sign: for i=0: |col_j.y*tz - col_j.z*ty| ; i=1: |col_j.z*tx - col_j.x*tz| ; i=2: |col_j.x*ty - col_j.y*tx| */
      absSign = fabsf(C(col[j], i1) * tt[i2] - C(col[j], i2) * tt[i1]);
      /* ra: i=0: abs1To0[j].z*ea1 + abs1To0[j].y*ea2 ; i=1: .z*ea0 + .x*ea2 ; i=2: .y*ea0 + .x*ea1 */
      if (i == 0) ra = abs1To0[j].z * ea[1] + abs1To0[j].y * ea[2];
      else if (i == 1) ra = abs1To0[j].z * ea[0] + abs1To0[j].x * ea[2];
      else ra = abs1To0[j].y * ea[0] + abs1To0[j].x * ea[1];
      /* rb: j=0: abs0To1[i].z*eb1 + abs0To1[i].y*eb2 ; j=1: .z*eb0 + .x*eb2 ; j=2: .y*eb0 + .x*eb1 */
      if (j == 0) rb = abs0To1[i].z * eb[1] + abs0To1[i].y * eb[2];
      else if (j == 1) rb = abs0To1[i].z * eb[0] + abs0To1[i].x * eb[2];
      else rb = abs0To1[i].y * eb[0] + abs0To1[i].x * eb[1];
      const float radiusSum = (ra + rb) + contactDist;
      if (absSign > radiusSum) return 0;
    }
  }
#undef C
  int feature = 0; float minOverlap = overlap[0];
  for (int i = 1; i < 6; ++i) if (minOverlap > overlap[i]) { minOverlap = overlap[i]; feature = i; }

  const v3 ax0[3] = {t0->r.c0, t0->r.c1, t0->r.c2}, ax1[3] = {t1->r.c0, t1->r.c1, t1->r.c2};
  mxf nt; v3 mtd; v3 incN; v3 pts[4]; int flip = 0;
  const int neg = 0.f >= sign[feature];
  switch (feature) {
    case 0:
      if (neg) { mtd = ax0[0]; nt.r.c0 = v3neg(ax0[2]); nt.r.c1 = ax0[1]; nt.r.c2 = ax0[0]; nt.p = v3negscalesub(ax0[0], ea[0], t0->p); }
      else { mtd = v3neg(ax0[0]); nt.r.c0 = ax0[2]; nt.r.c1 = ax0[1]; nt.r.c2 = mtd; nt.p = v3scaleadd(ax0[0], ea[0], t0->p); }
      break;
    case 1:
      if (neg) { mtd = ax0[1]; nt.r.c0 = ax0[0]; nt.r.c1 = v3neg(ax0[2]); nt.r.c2 = ax0[1]; nt.p = v3negscalesub(ax0[1], ea[1], t0->p); }
      else { mtd = v3neg(ax0[1]); nt.r.c0 = ax0[0]; nt.r.c1 = ax0[2]; nt.r.c2 = mtd; nt.p = v3scaleadd(ax0[1], ea[1], t0->p); }
      break;
    case 2:
      if (neg) { mtd = ax0[2]; nt.r.c0 = ax0[0]; nt.r.c1 = ax0[1]; nt.r.c2 = ax0[2]; nt.p = v3negscalesub(ax0[2], ea[2], t0->p); }
      else { mtd = v3neg(ax0[2]); nt.r.c0 = ax0[0]; nt.r.c1 = v3neg(ax0[1]); nt.r.c2 = mtd; nt.p = v3scaleadd(ax0[2], ea[2], t0->p); }
      break;
    case 3:
      flip = 1;
      if (neg) { mtd = ax1[0]; nt.r.c0 = ax1[2]; nt.r.c1 = ax1[1]; nt.r.c2 = v3neg(ax1[0]); nt.p = v3scaleadd(ax1[0], eb[0], t1->p); }
      else { mtd = v3neg(ax1[0]); nt.r.c0 = v3neg(ax1[2]); nt.r.c1 = ax1[1]; nt.r.c2 = ax1[0]; nt.p = v3negscalesub(ax1[0], eb[0], t1->p); }
      break;
    case 4:
      flip = 1;
      if (neg) { mtd = ax1[1]; nt.r.c0 = ax1[0]; nt.r.c1 = ax1[2]; nt.r.c2 = v3neg(ax1[1]); nt.p = v3scaleadd(ax1[1], eb[1], t1->p); }
      else { mtd = v3neg(ax1[1]); nt.r.c0 = ax1[0]; nt.r.c1 = v3neg(ax1[2]); nt.r.c2 = ax1[1]; nt.p = v3negscalesub(ax1[1], eb[1], t1->p); }
      break;
    default:
      flip = 1;
      if (neg) { mtd = ax1[2]; nt.r.c0 = ax1[0]; nt.r.c1 = v3neg(ax1[1]); nt.r.c2 = v3neg(ax1[2]); nt.p = v3scaleadd(ax1[2], eb[2], t1->p); }
      else { mtd = v3neg(ax1[2]); nt.r.c0 = ax1[0]; nt.r.c1 = ax1[1]; nt.r.c2 = ax1[2]; nt.p = v3negscalesub(ax1[2], eb[2], t1->p); }
      break;
  }
  const v3 localNormal = amxfrotateinv(&nt, mtd);
  if (!flip) {
    const mxf t1ToNew = amxfinvmul(&nt, t1);
    pxo_incident_polygon(pts, &incN, v3neg(localNormal), &t1ToNew, e1);
    const float exx = feature == 0 ? ea[2] : ea[0];
    const float eyy = feature == 0 ? ea[1] : (feature == 1 ? ea[2] : ea[1]);
    pxo_calc_contacts(exx, eyy, pts, incN, localNormal, mc, numContacts, contactDist);
  } else {
    const mxf t0ToNew = amxfinvmul(&nt, t0);
    pxo_incident_polygon(pts, &incN, localNormal, &t0ToNew, e0);
    const float exx = feature == 3 ? eb[2] : eb[0];
    const float eyy = feature == 3 ? eb[1] : (feature == 4 ? eb[2] : eb[1]);
    pxo_calc_contacts(exx, eyy, pts, incN, localNormal, mc, numContacts, contactDist);
  }
  const int n = *numContacts;
  if (n != 0) {
    if (flip) for (int i = 0; i < n; ++i) { const v3 lb = mc[i].b; mc[i].b = mc[i].a; mc[i].a = lb; }
    const mxf newTo1 = amxfinvmul(t1, &nt), newTo0 = amxfinvmul(t0, &nt);
    const v3 localNormalInB = mxfrotate(&newTo1, mc[0].n);
    for (int i = 0; i < n; ++i) {
      mc[i].a = amxftransform(&newTo0, mc[i].a);
      mc[i].b = amxftransform(&newTo1, mc[i].b);
      mc[i].n = localNormalInB;
    }
  }
  return 1;
}

/* GuPCMContactBoxBox.cpp:848-971 */
static inline int pxo_pcm_box_box(const xf* tm0, const xf* tm1, v3 ext0, v3 ext1, float contactDist, float toleranceLength,
                                  PxoManifold* manifold, PxoContacts* out) {
  const xf curRTrans = axfinvmul(tm1, tm0); /* A into B */
  const mxf aToB = amxffromxf(&curRTrans);
  const float minMargin = fminf_(pxo_box_margin(ext0, toleranceLength), pxo_box_margin(ext1, toleranceLength));
  const int initialContacts = manifold->n;
  const float projectBreakingThreshold = minMargin * 0.8f;
  pxo_refresh(manifold, &aToB, projectBreakingThreshold);
  const int bLost = manifold->n != initialContacts;
  const float radiusA = alen(ext0), radiusB = alen(ext1);
  out->count = 0;
  if (bLost || pxo_invalidate_boxconvex(manifold, &curRTrans, tm0->q, tm1->q, minMargin, radiusA, radiusB)) {
    manifold->rel = curRTrans; manifold->quatA = tm0->q; manifold->quatB = tm1->q;
    mxf tv0 = amxffromxf(tm0), tv1 = amxffromxf(tm1);
    tv0.r.c0 = anormalize(tv0.r.c0); tv0.r.c1 = anormalize(tv0.r.c1); tv0.r.c2 = anormalize(tv0.r.c2);
    tv1.r.c0 = anormalize(tv1.r.c0); tv1.r.c1 = anormalize(tv1.r.c1); tv1.r.c2 = anormalize(tv1.r.c2);
    PxoMPoint mc[16]; int num = 0;
    if (pxo_boxbox_generate(ext0, ext1, &tv0, &tv1, contactDist, mc, &num)) {
      if (num > 0) {
        if (num <= PXO_MANIFOLD_CACHE) { for (int i = 0; i < num; ++i) manifold->pts[i] = mc[i]; manifold->n = num; }
        else { pxo_reduce_batch(manifold, mc, num, toleranceLength); manifold->n = PXO_MANIFOLD_CACHE; }
        out->normal = anormalize(mxfrotate(&tv1, manifold->pts[0].n));
        for (int i = 0; i < manifold->n; ++i) {
          out->point[out->count] = amxftransform(&tv1, manifold->pts[i].b);
          out->sep[out->count] = manifold->pts[i].pen; out->count++;
        }
        return 1;
      }
      return 2;   /* SAT passed, clipping found no point: the caller runs the GJK / EPA single-point fallback (pxo_gjk.h: pxo_boxbox_gjk_fallback) */
    }
    return 0;
  } else if (manifold->n > 0) {
    out->normal = pxo_world_normal(manifold, tm1);
    for (int i = 0; i < manifold->n; ++i) {
      const float dist = manifold->pts[i].pen;
      if (contactDist >= dist) { out->point[out->count] = axftransform(tm1, manifold->pts[i].b); out->sep[out->count] = dist; out->count++; }
    }
    return 1;
  }
  return 0;
}

/* ---------------- sphere / capsule family ---------------- */
/* shape0 = sphere, shape1 = sphere: GuPCMContactSphereSphere.cpp:36-69 (stateless) */
static inline void pxo_sphere_sphere(v3 p0, v3 p1, float r0, float r1, float cDist, PxoContacts* out) {
  out->count = 0;
  const v3 delta = v3sub(p0, p1);
  const float distanceSq = adot(delta, delta);
  const float radiusSum = r0 + r1, inflatedSum = radiusSum + cDist;
  if (inflatedSum * inflatedSum > distanceSq) {
    const float dist = sqrtf(distanceSq);
    const v3 normal = (0.00001f >= dist) ? V3(1.f, 0.f, 0.f) : V3(delta.x / dist, delta.y / dist, delta.z / dist);
    out->normal = normal; out->point[0] = v3scaleadd(normal, r1, p1); out->sep[0] = dist - radiusSum; out->count = 1;
  }
}
/* shape0 = sphere, shape1 = plane: GuPCMContactSpherePlane.cpp:36-73 */
static inline void pxo_sphere_plane(v3 p0, float radius, const xf* planeTm, float cDist, PxoContacts* out) {
  out->count = 0;
  const v3 c = aqrotinv(planeTm->q, v3sub(p0, planeTm->p));
  const float separation = c.x - radius;
  if (cDist >= separation) {
    const v3 n = aqbasis0(planeTm->q);
    out->normal = n; out->point[0] = v3negscalesub(n, radius, p0); out->sep[0] = separation; out->count = 1;
  }
}
/* GuPCMContactSphereCapsule.cpp:38-54 */
static inline float pxo_dist_point_segment_sq(v3 a, v3 b, v3 p, float* param) {
  const v3 ap = v3sub(p, a), ab = v3sub(b, a);
  const float nom = adot(ap, ab), denom = adot(ab, ab);
  const float tValue = fmaxf_(fminf_(nom / denom, 1.f), 0.f);
  const float t = (denom == 0.f) ? 0.f : tValue;
  const v3 v = v3negscalesub(ab, t, ap);
  *param = t;
  return adot(v, v);
}
/* shape0 = sphere, shape1 = capsule: GuPCMContactSphereCapsule.cpp:56-102 */
static inline void pxo_sphere_capsule(v3 sphereCenter, float sphereRadius, const xf* capTm, float capRadius, float halfHeight, float cDist, PxoContacts* out) {
  out->count = 0;
  const v3 tmp0 = v3scale(aqbasis0(capTm->q), halfHeight);
  const v3 s = v3add(capTm->p, tmp0), e = v3sub(capTm->p, tmp0);
  const float radiusSum = sphereRadius + capRadius, inflatedSum = radiusSum + cDist;
  float t; const float squareDist = pxo_dist_point_segment_sq(s, e, sphereCenter, &t);
  if (inflatedSum * inflatedSum > squareDist) {
    const v3 p = v3scaleadd(v3sub(e, s), t, s);
    const v3 dir = v3sub(sphereCenter, p);
    const float len = alen(dir);
    const v3 normal = (len > FLT_EPSILON) ? V3(dir.x / len, dir.y / len, dir.z / len) : V3(1.f, 0.f, 0.f);
    out->normal = normal; out->point[0] = v3negscalesub(normal, sphereRadius, sphereCenter); out->sep[0] = sqrtf(squareDist) - radiusSum; out->count = 1;
  }
}
/* shape0 = sphere, shape1 = box: GuPCMContactSphereBox.cpp:36-131 */
static inline void pxo_sphere_box(v3 sphereOrigin, float radius, const xf* boxTm, v3 be, float cDist, PxoContacts* out) {
  out->count = 0;
  const v3 c = aqrotinv(boxTm->q, v3sub(sphereOrigin, boxTm->p));
  const float inflatedSum = radius + cDist;
  const v3 p = V3(fmaxf_(fminf_(c.x, be.x), -be.x), fmaxf_(fminf_(c.y, be.y), -be.y), fmaxf_(fminf_(c.z, be.z), -be.z)); /* V3Clamp = max(min(a,max),min) */
  const v3 v = v3sub(c, p);
  const float lengthSq = adot(v, v);
  if (inflatedSum * inflatedSum > lengthSq) {
    const v3 ac = v3abs(c);
    if (be.x >= ac.x && be.y >= ac.y && be.z >= ac.z) {
      const v3 d = v3sub(be, v3abs(p));
      const int con0 = d.x >= d.z && d.y >= d.z && d.z >= d.z, con1 = d.x >= d.x && d.y >= d.x && d.z >= d.x;
      const v3 sign = V3(p.x >= 0.f ? 1.f : -1.f, p.y >= 0.f ? 1.f : -1.f, p.z >= 0.f ? 1.f : -1.f);
      const v3 locNorm = con0 ? V3(0.f * sign.x, 0.f * sign.y, 1.f * sign.z) : (con1 ? V3(1.f * sign.x, 0.f * sign.y, 0.f * sign.z) : V3(0.f * sign.x, 1.f * sign.y, 0.f * sign.z));
      const float dist = -(con0 ? d.z : (con1 ? d.x : d.y));
      const v3 normal = aqrot(boxTm->q, locNorm);
      out->normal = normal; out->point[0] = v3sub(sphereOrigin, v3scale(normal, dist)); out->sep[0] = dist - radius; out->count = 1;
    } else {
      const float recipLength = 1.0f / sqrtf(lengthSq);
      const float length = 1.0f / recipLength;
      const v3 locNorm = v3scale(v, recipLength);
      out->normal = aqrot(boxTm->q, locNorm); out->point[0] = axftransform(boxTm, p); out->sep[0] = length - radius; out->count = 1;
    }
  }
}
/* GuPersistentContactManifold.cpp:552-600, :1269-1290 (addManifoldPoint2 -> replaceManifoldPoint / reduceContactSegment) */
static inline void pxo_add_manifold_point2(PxoManifold* m, v3 la, v3 lb, v3 n, float pen, float replaceBreakingThreshold) {
  const float shortest = replaceBreakingThreshold * replaceBreakingThreshold;
  for (int i = 0; i < m->n; ++i) {
    const v3 dB = v3sub(m->pts[i].b, lb), dA = v3sub(m->pts[i].a, la);
    if (shortest > fminf_(adot(dB, dB), adot(dA, dA))) { m->pts[i].a = la; m->pts[i].b = lb; m->pts[i].n = n; m->pts[i].pen = pen; return; }
  }
  if (m->n < 2) { m->pts[m->n].a = la; m->pts[m->n].b = lb; m->pts[m->n].n = n; m->pts[m->n].pen = pen; m->n++; return; }
  const v3 v0 = v3sub(m->pts[0].b, lb), v1 = v3sub(m->pts[1].b, lb);
  const int k = (adot(v0, v0) > adot(v1, v1)) ? 1 : 0;
  m->pts[k].a = la; m->pts[k].b = lb; m->pts[k].n = n; m->pts[k].pen = pen;
}
/* shape0 = plane, shape1 = capsule: GuPCMContactPlaneCapsule.cpp:36-123 (persistent manifold, <= 2 points) */
static inline void pxo_pcm_plane_capsule(const xf* planeTm, const xf* capTm, float radius, float halfHeight, float contactDist, PxoManifold* man, PxoContacts* out) {
  const xf aToB = axfinvmul(planeTm, capTm);
  const v3 planeNormal = anormalize(aqbasis0(planeTm->q));
  const v3 contactNormal = v3neg(planeNormal);
  const v3 localNormal = V3(1.f, 0.f, 0.f);
  const v3 tmp = v3scale(aqbasis0(aToB.q), halfHeight);
  const v3 s = v3add(aToB.p, tmp), e = v3sub(aToB.p, tmp);
  const float inflatedRadius = radius + contactDist;
  const float replaceBreakingThreshold = radius * 0.001f, projectBreakingThreshold = radius * 0.05f;
  const int initial = man->n;
  const mxf aToBm = amxffromxf(&aToB);
  pxo_refresh(man, &aToBm, projectBreakingThreshold);
  const int lost = man->n != initial;
  if (lost || pxo_invalidate_plane(man, &aToB, radius, 0.02f)) {
    man->n = 0; man->rel = aToB;
    const float sd0 = s.x;
    if (inflatedRadius > sd0) pxo_add_manifold_point2(man, aqrotinv(aToB.q, v3sub(s, aToB.p)), v3negscalesub(localNormal, sd0, s), localNormal, sd0, replaceBreakingThreshold);
    const float sd1 = e.x;
    if (inflatedRadius > sd1) pxo_add_manifold_point2(man, aqrotinv(aToB.q, v3sub(e, aToB.p)), v3negscalesub(localNormal, sd1, e), localNormal, sd1, replaceBreakingThreshold);
  }
  /* addManifoldContactsToContactBuffer(buffer, normal, projectionNormal, transf0, radius, contactOffset): .cpp:783-807 */
  out->count = 0; out->normal = contactNormal;
  for (int i = 0; i < man->n; ++i) {
    const float dist = man->pts[i].pen - radius;
    if (contactDist >= dist) { out->point[out->count] = v3negscalesub(planeNormal, radius, axftransform(capTm, man->pts[i].a)); out->sep[out->count] = dist; out->count++; }
  }
}
/* GuDistanceSegmentSegment.cpp:411-469 */
static inline float pxo_dist_seg_seg_sq(v3 p1, v3 d1, v3 p2, v3 d2, float* s, float* t) {
  const float eps = FLT_EPSILON;
  const v3 r = v3sub(p1, p2);
  const float a = v3dot(d1, d1), e = v3dot(d2, d2), b = v3dot(d1, d2), c = v3dot(d1, r); /* V3Dot4: (x+y)+z */
  const float aRecip = a > eps ? 1.0f / a : 0.f, eRecip = e > eps ? 1.0f / e : 0.f;
  const float f = adot(d2, r);
  const float denom = a * e - b * b;
  const float temp = b * f - c * e;
  const float s0 = fmaxf_(fminf_(temp / denom, 1.f), 0.f);
  const float sTmp = (eps > denom) ? 0.5f : s0;
  const float tTmp = (b * sTmp + f) * eRecip;
  const float t2 = fmaxf_(fminf_(tTmp, 1.f), 0.f);
  const float comp = (b * t2 - c) * aRecip;
  const float s2 = fmaxf_(fminf_(comp, 1.f), 0.f);
  *s = s2; *t = t2;
  const v3 vv = v3sub(v3scaleadd(d1, s2, p1), v3scaleadd(d2, t2, p2));
  return adot(vv, vv);
}
/* shape0 = capsule, shape1 = capsule: GuPCMContactCapsuleCapsule.cpp:78-275 (stateless, <= 4 contacts).
 * Returns nonzero in *multiNormal when the contacts' normals are not all within PXC_SAME_NORMAL of the first
 * one (the reference would then split them into several patches; we keep one patch -- see DESIGN.md). */
static inline void pxo_capsule_capsule(const xf* tm0, const xf* tm1, float r0, float hh0, float r1, float hh1, float cDist, PxoContacts* out, int* multiNormal) {
  out->count = 0; *multiNormal = 0;
  const v3 positionOffset = v3scale(v3add(tm0->p, tm1->p), 0.5f);
  const v3 p0 = v3sub(tm0->p, positionOffset), p1 = v3sub(tm1->p, positionOffset);
  const v3 tmp0 = v3scale(aqbasis0(tm0->q), hh0);
  const v3 s0 = v3add(p0, tmp0), e0 = v3sub(p0, tmp0), d0 = v3sub(e0, s0);
  const v3 tmp1 = v3scale(aqbasis0(tm1->q), hh1);
  const v3 s1 = v3add(p1, tmp1), e1 = v3sub(p1, tmp1), d1 = v3sub(e1, s1);
  const float sumRadius = r0 + r1, inflatedSum = sumRadius + cDist, inflatedSumSquared = inflatedSum * inflatedSum;
  const float a = adot(d0, d0), e = adot(d1, d1), eps = 1e-6f;
  float t0, t1;
  const float sqDist0 = pxo_dist_seg_seg_sq(s0, d0, s1, d1, &t0, &t1);
  if (!(inflatedSumSquared >= sqDist0)) return;
  v3 normals[4];
  const float sa = sqrtf(a), se = sqrtf(e);
  const v3 dir0 = (eps > a) ? V3(0, 0, 0) : V3(d0.x / sa, d0.y / sa, d0.z / sa);
  const v3 dir1 = (eps > e) ? V3(0, 0, 0) : V3(d1.x / se, d1.y / se, d1.z / se);
  const float cosv = fabsf(adot(dir0, dir1));
  if (cosv > 0.9998f) {
    /* pcmDistancePointSegmentTValue22(s0,e0, s1,e1, s1,e1,s0,e0): t = dot(p - a, ab)/dot(ab,ab) */
    const v3 ab0 = v3sub(e0, s0), ab1 = v3sub(e1, s1);
    const float den0 = adot(ab0, ab0), den1 = adot(ab1, ab1);
    const float nom[4] = {v3dot(v3sub(s1, s0), ab0), v3dot(v3sub(e1, s0), ab0), v3dot(v3sub(s0, s1), ab1), v3dot(v3sub(e0, s1), ab1)};
    const float den[4] = {den0, den0, den1, den1};
    float t[4];
    for (int k = 0; k < 4; ++k) t[k] = (den[k] == 0.f) ? 0.f : nom[k] / den[k];
    for (int k = 0; k < 4; ++k) {
      if (!(t[k] >= 0.f && 1.f >= t[k])) continue;
      v3 proj, v, base;
      if (k == 0) { proj = v3scaleadd(d0, t[0], s0); v = v3sub(proj, s1); base = proj; }
      else if (k == 1) { proj = v3scaleadd(d0, t[1], s0); v = v3sub(proj, e1); base = proj; }
      else if (k == 2) { proj = v3scaleadd(d1, t[2], s1); v = v3sub(s0, proj); base = s0; }
      else { proj = v3scaleadd(d1, t[3], s1); v = v3sub(e0, proj); base = e0; }
      const float sqDist = adot(v, v);
      if (sqDist > eps && inflatedSumSquared > sqDist) {
        const float dist = sqrtf(sqDist);
        const v3 normal = V3(v.x / dist, v.y / dist, v.z / dist);
        normals[out->count] = normal;
        out->point[out->count] = v3add(v3negscalesub(normal, r0, base), positionOffset); out->sep[out->count] = dist - sumRadius; out->count++;
      }
    }
    if (out->count) {
      out->normal = normals[0];
      for (int k = 1; k < out->count; ++k) if (!(v3dot(normals[k], normals[0]) >= 0.999f)) *multiNormal = 1;
      return;
    }
  }
  const v3 closestA = v3scaleadd(d0, t0, s0), closestB = v3scaleadd(d1, t1, s1);
  const int con = eps > sqDist0;
  const v3 _normal = con ? ((a > eps) ? d0 : V3(1.f, 0.f, 0.f)) : v3sub(closestA, closestB);
  const v3 normal = anormalize(_normal);
  out->normal = normal; out->point[0] = v3add(v3negscalesub(normal, r0, closestA), positionOffset);
  out->sep[0] = (con ? 0.f : sqrtf(sqDist0)) - sumRadius; out->count = 1;
}
#endif
