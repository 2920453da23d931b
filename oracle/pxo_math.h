/* pxo_math.h -- scalar float math for the CPU oracle (TEST INFRASTRUCTURE, never linked by physx_b200/).
 * Follows the reference's scalar formulas so rounding stays close:
 *   quaternion rotate:      physx/include/foundation/PxQuat.h:286-295
 *   quaternion -> matrix:   physx/include/foundation/PxMat33.h:136-163
 *   transform ops:          physx/include/foundation/PxTransform.h
 * Compile with -ffp-contract=off (the reference x86-64 build has no FMA contraction). */
#ifndef PXO_MATH_H
#define PXO_MATH_H
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdint.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } q4;
typedef struct { q4 q; v3 p; } xf;           /* PxTransform */
typedef struct { v3 c0, c1, c2; } m33;       /* column major like PxMat33 */
typedef struct { m33 r; v3 p; } mxf;         /* PxMatTransformV */

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3scale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 v3neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float v3dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 v3cross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline v3 v3abs(v3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline float v3lensq(v3 a) { return v3dot(a, a); }
static inline float v3len(v3 a) { return sqrtf(v3dot(a, a)); }
static inline v3 v3scaleadd(v3 a, float s, v3 b) { return V3(a.x * s + b.x, a.y * s + b.y, a.z * s + b.z); }       /* a*s+b */
static inline v3 v3negscalesub(v3 a, float s, v3 b) { return V3(b.x - a.x * s, b.y - a.y * s, b.z - a.z * s); }   /* b-a*s */
static inline v3 v3normalize(v3 a) { float l = v3len(a); return V3(a.x / l, a.y / l, a.z / l); }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline v3 v3min(v3 a, v3 b) { return V3(fminf_(a.x, b.x), fminf_(a.y, b.y), fminf_(a.z, b.z)); }
static inline v3 v3max(v3 a, v3 b) { return V3(fmaxf_(a.x, b.x), fmaxf_(a.y, b.y), fmaxf_(a.z, b.z)); }

static inline q4 Q4(float x, float y, float z, float w) { q4 r = {x, y, z, w}; return r; }
static inline float q4dot(q4 a, q4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static inline q4 q4conj(q4 a) { return Q4(-a.x, -a.y, -a.z, a.w); }
static inline q4 q4mul(q4 a, q4 b) { /* PxQuat::operator* */
  return Q4(a.w * b.x + b.w * a.x + a.y * b.z - b.y * a.z,
            a.w * b.y + b.w * a.y + a.z * b.x - b.z * a.x,
            a.w * b.z + b.w * a.z + a.x * b.y - b.x * a.y,
            a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
static inline q4 q4normalized(q4 a) {
  float s = 1.0f / sqrtf(q4dot(a, a));
  return Q4(a.x * s, a.y * s, a.z * s, a.w * s);
}
static inline v3 q4rot(q4 q, v3 v) {
  const float vx = 2.0f * v.x, vy = 2.0f * v.y, vz = 2.0f * v.z;
  const float w2 = q.w * q.w - 0.5f;
  const float dot2 = (q.x * vx + q.y * vy + q.z * vz);
  return V3((vx * w2 + (q.y * vz - q.z * vy) * q.w + q.x * dot2), (vy * w2 + (q.z * vx - q.x * vz) * q.w + q.y * dot2),
            (vz * w2 + (q.x * vy - q.y * vx) * q.w + q.z * dot2));
}
static inline v3 q4rotinv(q4 q, v3 v) {
  const float vx = 2.0f * v.x, vy = 2.0f * v.y, vz = 2.0f * v.z;
  const float w2 = q.w * q.w - 0.5f;
  const float dot2 = (q.x * vx + q.y * vy + q.z * vz);
  return V3((vx * w2 - (q.y * vz - q.z * vy) * q.w + q.x * dot2), (vy * w2 - (q.z * vx - q.x * vz) * q.w + q.y * dot2),
            (vz * w2 - (q.x * vy - q.y * vx) * q.w + q.z * dot2));
}
static inline v3 q4basis0(q4 q) { /* PxQuat::getBasisVector0 */
  const float x2 = q.x * 2.0f, w2 = q.w * 2.0f;
  return V3((q.w * w2) - 1.0f + q.x * x2, (q.z * w2) + q.y * x2, (-q.y * w2) + q.z * x2);
}
static inline m33 m33fromq(q4 q) {
  const float x = q.x, y = q.y, z = q.z, w = q.w;
  const float x2 = x + x, y2 = y + y, z2 = z + z;
  const float xx = x2 * x, yy = y2 * y, zz = z2 * z;
  const float xy = x2 * y, xz = x2 * z, xw = x2 * w;
  const float yz = y2 * z, yw = y2 * w, zw = z2 * w;
  m33 m;
  m.c0 = V3(1.0f - yy - zz, xy + zw, xz - yw);
  m.c1 = V3(xy - zw, 1.0f - xx - zz, yz + xw);
  m.c2 = V3(xz + yw, yz - xw, 1.0f - xx - yy);
  return m;
}
static inline v3 m33mul(const m33* m, v3 v) { /* M*v */
  return V3(m->c0.x * v.x + m->c1.x * v.y + m->c2.x * v.z, m->c0.y * v.x + m->c1.y * v.y + m->c2.y * v.z,
            m->c0.z * v.x + m->c1.z * v.y + m->c2.z * v.z);
}
static inline v3 m33tmul(const m33* m, v3 v) { /* M^T*v */
  return V3(v3dot(m->c0, v), v3dot(m->c1, v), v3dot(m->c2, v));
}
static inline m33 m33transpose(const m33* m) {
  m33 r; r.c0 = V3(m->c0.x, m->c1.x, m->c2.x); r.c1 = V3(m->c0.y, m->c1.y, m->c2.y); r.c2 = V3(m->c0.z, m->c1.z, m->c2.z); return r;
}

static inline v3 xftransform(const xf* t, v3 v) { return v3add(q4rot(t->q, v), t->p); }
static inline v3 xftransforminv(const xf* t, v3 v) { return q4rotinv(t->q, v3sub(v, t->p)); }
/* a.transformInv(b): b expressed in a's frame  (PxTransform::transformInv(const PxTransform&)) */
static inline xf xfinvmul(const xf* a, const xf* b) {
  xf r; q4 qinv = q4conj(a->q);
  r.p = q4rot(qinv, v3sub(b->p, a->p));
  r.q = q4mul(qinv, b->q);
  return r;
}
static inline mxf mxffromxf(const xf* t) { mxf m; m.r = m33fromq(t->q); m.p = t->p; return m; }
static inline v3 mxftransform(const mxf* t, v3 v) { return v3add(m33mul(&t->r, v), t->p); }
static inline v3 mxftransforminv(const mxf* t, v3 v) { return m33tmul(&t->r, v3sub(v, t->p)); }
static inline v3 mxfrotate(const mxf* t, v3 v) { return m33mul(&t->r, v); }
static inline v3 mxfrotateinv(const mxf* t, v3 v) { return m33tmul(&t->r, v); }
/* a.transformInv(b) for matrix transforms: rot = a.rot^T * b.rot, p = a.rot^T*(b.p-a.p)  (PxMatTransformV) */
static inline mxf mxfinvmul(const mxf* a, const mxf* b) {
  mxf r;
  r.r.c0 = m33tmul(&a->r, b->r.c0); r.r.c1 = m33tmul(&a->r, b->r.c1); r.r.c2 = m33tmul(&a->r, b->r.c2);
  r.p = m33tmul(&a->r, v3sub(b->p, a->p));
  return r;
}

/* ---- "aos" variants: same operation ORDER as the reference's SSE2 vector layer (no SSE4.2), used wherever
 * the reference code is written against physx/include/foundation/PxVecMath*.h / PxVecQuat.h / PxVecTransform.h.
 *   V3Dot  = (x*x' + z*z') + y*y'          PxVecMathSSE.h:965-981
 *   V4Dot  = (x*x' + z*z') + (y*y' + w*w') PxVecMathSSE.h:1714-1740
 *   M33TrnspsMulV3 = V3Dot per column      unix/sse2/PxUnixSse2InlineAoS.h:360-366
 *   QuatRotate/QuatRotateInv/QuatTransform PxVecQuat.h:185-267, QuatMul :269-283, QuatGetBasisVector0 :114-135 */
static inline float adot(v3 a, v3 b) { return (a.x * b.x + a.z * b.z) + (a.y * b.y); }
static inline float adot4(q4 a, q4 b) { return (a.x * b.x + a.z * b.z) + (a.y * b.y + a.w * b.w); }
static inline float alensq(v3 a) { return adot(a, a); }
static inline float alen(v3 a) { return sqrtf(adot(a, a)); }
static inline v3 anormalize(v3 a) { const float l = sqrtf(adot(a, a)); return V3(a.x / l, a.y / l, a.z / l); }
static inline v3 aqrot_noscale(q4 q, v3 v) {
  const v3 u = V3(q.x, q.y, q.z);
  const float w2 = q.w * q.w + (-0.5f);
  const v3 a = v3scale(v, w2);
  const v3 temp = v3scaleadd(v3cross(u, v), q.w, a);
  return v3scaleadd(u, adot(u, v), temp);
}
static inline v3 aqrot(q4 q, v3 v) { return v3scale(aqrot_noscale(q, v), 2.0f); }
static inline v3 aqrotinv(q4 q, v3 v) {
  const v3 u = V3(q.x, q.y, q.z);
  const float w2 = q.w * q.w + (-0.5f);
  const v3 a = v3scale(v, w2);
  const v3 temp = v3negscalesub(v3cross(u, v), q.w, a);
  return v3scale(v3scaleadd(u, adot(u, v), temp), 2.0f);
}
static inline v3 aqrot_normalize(q4 q, v3 v) { return anormalize(aqrot_noscale(q, v)); }
static inline q4 aqmul(q4 a, q4 b) {
  const v3 ia = V3(a.x, a.y, a.z), ib = V3(b.x, b.y, b.z);
  const float real = a.w * b.w - v3dot(ia, ib); /* V4Dot3 = (x+y)+z */
  const v3 imag = v3add(v3add(v3scale(ia, b.w), v3scale(ib, a.w)), v3cross(ia, ib));
  return Q4(imag.x, imag.y, imag.z, real);
}
static inline v3 aqbasis0(q4 q) {
  const float x2 = q.x * 2.0f, w2 = q.w * 2.0f;
  const v3 a = v3scale(V3(q.x, q.y, q.z), x2);
  const v3 ab = v3scaleadd(V3(q.w, q.z, -q.y), w2, a);
  return V3(ab.x - 1.0f, ab.y, ab.z);
}
static inline v3 axftransform(const xf* t, v3 v) { return v3scaleadd(aqrot_noscale(t->q, v), 2.0f, t->p); }
static inline xf axfinvmul(const xf* a, const xf* b) { /* PxTransformV::transformInv(PxTransformV) */
  xf r; const q4 qinv = q4conj(a->q);
  r.p = aqrot(qinv, v3sub(b->p, a->p));
  r.q = aqmul(qinv, b->q);
  return r;
}
/* Cm::getStaticGlobalPoseAligned / getDynamicGlobalPoseAligned (common/src/CmTransformUtils.h:40-133): the shape's world pose as the transform cache holds it.
 * atransform_fast(a, b) = a * b, atransform_inv_fast(a, b) = a^-1 * b, in the aos operation order of transformFast / transformInvFast. */
static inline xf atransform_fast(const xf* a, const xf* b) {
  const float wa = a->q.w, wb = b->q.w; const v3 va = V3(a->q.x, a->q.y, a->q.z), vb = V3(b->q.x, b->q.y, b->q.z);
  const float wo = wa * wb - adot(va, vb);
  const v3 vo = v3scaleadd(va, wb, v3scaleadd(vb, wa, v3cross(va, vb)));
  const v3 t1 = v3scale(b->p, wa * wa + (-0.5f));
  const v3 t2 = v3scaleadd(v3cross(va, b->p), wa, t1);
  const v3 t3 = v3scaleadd(va, adot(va, b->p), t2);
  xf o; o.p = v3scaleadd(t3, 2.f, a->p); o.q = Q4(vo.x, vo.y, vo.z, wo); return o;
}
static inline xf atransform_inv_fast(const xf* a, const xf* b) {
  const float wa = a->q.w, wb = b->q.w; const v3 va = V3(a->q.x, a->q.y, a->q.z), vb = V3(b->q.x, b->q.y, b->q.z);
  const float wo = wa * wb + adot(va, vb);
  const v3 vo = v3negscalesub(va, wb, v3scaleadd(vb, wa, v3cross(vb, va)));
  const v3 pt = v3sub(b->p, a->p);
  const v3 t1 = v3scale(pt, wa * wa + (-0.5f));
  const v3 t2 = v3scaleadd(v3cross(pt, va), wa, t1);
  const v3 t3 = v3scaleadd(va, adot(va, pt), t2);
  xf o; o.p = v3add(t3, t3); o.q = Q4(vo.x, vo.y, vo.z, wo); return o;
}
/* scalar PxTransform algebra of the API layer (foundation/PxTransform.h): a * b = (a.q.rotate(b.p) + a.p, a.q * b.q), getInverse = (q.rotateInv(-p), q*), getNormalized */
static inline xf xfmul(const xf* a, const xf* b) { xf o; o.p = v3add(q4rot(a->q, b->p), a->p); o.q = q4mul(a->q, b->q); return o; }
static inline xf xfinverse(const xf* a) { xf o; o.p = q4rotinv(a->q, v3neg(a->p)); o.q = q4conj(a->q); return o; }
static inline xf xfnormalized(const xf* a) { const float s = 1.0f / sqrtf(a->q.x * a->q.x + a->q.y * a->q.y + a->q.z * a->q.z + a->q.w * a->q.w); xf o; o.p = a->p; o.q = Q4(a->q.x * s, a->q.y * s, a->q.z * s, a->q.w * s); return o; }
static inline v3 am33tmul(const m33* m, v3 v) { return V3(adot(m->c0, v), adot(m->c1, v), adot(m->c2, v)); }
static inline v3 amxftransform(const mxf* t, v3 v) { return v3add(t->p, m33mul(&t->r, v)); }
static inline v3 amxftransforminv(const mxf* t, v3 v) { return am33tmul(&t->r, v3sub(v, t->p)); }
static inline v3 amxfrotateinv(const mxf* t, v3 v) { return am33tmul(&t->r, v); }
static inline mxf amxfinvmul(const mxf* a, const mxf* b) { /* PxMatTransformV::transformInv: M33MulM33(M33Trnsps(rot), src.rot) */
  mxf r; const m33 at = m33transpose(&a->r);
  r.r.c0 = m33mul(&at, b->r.c0); r.r.c1 = m33mul(&at, b->r.c1); r.r.c2 = m33mul(&at, b->r.c2);
  r.p = am33tmul(&a->r, v3sub(b->p, a->p));
  return r;
}
/* PxMatTransformV(const PxTransformV&): 3-output QuatGetMat33V, PxVecMathSSE.h:54-69 */
static inline m33 am33fromq(q4 q);
static inline mxf amxffromxf(const xf* t) { mxf m; m.p = t->p; m.r = am33fromq(t->q); return m; }
/* also PxMat33Padded(const PxQuat&), physx/include/foundation/PxSIMDHelpers.h:45-60 */
static inline m33 am33fromq(q4 q) {
  struct { m33 r; } m_; 
#define m m_
  const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z, w2 = q.w + q.w;
  const float wx = x2 * q.w, wy = y2 * q.w, wz = z2 * q.w, ww1 = w2 * q.w + (-1.0f);
  m.r.c0 = V3(q.x * x2 + ww1, q.y * x2 + wz, q.z * x2 + (-wy));
  m.r.c1 = V3(q.x * y2 + (-wz), q.y * y2 + ww1, q.z * y2 + wx);
  m.r.c2 = V3(q.x * z2 + wy, q.y * z2 + (-wx), q.z * z2 + ww1);
  return m.r;
#undef m
}
#endif
