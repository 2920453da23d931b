// pxb_math.cuh -- device-side float math for the rigid-body step kernels.
//
// Parity note: the reference's CPU path is compiled for x86-64 without FMA contraction, and its
// narrowphase / solver prep are written against an SSE2 vector layer whose reductions associate as
// (x+z)+y (physx/include/foundation/PxVecMathSSE.h:965-981).  Results of thresholded decisions
// (isSeparated, manifold invalidation, clipping) depend on those last bits, so this file keeps the
// same operation order and the translation unit is compiled with -fmad=false.  The "a*" helpers
// mirror the vector layer (V3Dot, QuatRotate, QuatMul, 3-output QuatGetMat33V ...), the plain helpers
// mirror the scalar PxVec3/PxQuat/PxTransform classes (physx/include/foundation/PxQuat.h:286-295,
// PxMat33.h:136-163).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define PXB_HD __host__ __device__ __forceinline__
#define PXB_D __device__ __forceinline__

struct v3 { float x, y, z; };
struct q4 { float x, y, z, w; };
struct xf { q4 q; v3 p; };
struct m33 { v3 c0, c1, c2; };
struct mxf { m33 r; v3 p; };

PXB_D v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
PXB_D v3 V3(const float4& f) { return V3(f.x, f.y, f.z); }
PXB_D q4 Q4(float x, float y, float z, float w) { q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
PXB_D q4 Q4(const float4& f) { return Q4(f.x, f.y, f.z, f.w); }
PXB_D float4 F4(v3 a, float w) { return make_float4(a.x, a.y, a.z, w); }
PXB_D float4 F4(q4 a) { return make_float4(a.x, a.y, a.z, a.w); }
PXB_D v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
PXB_D v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
PXB_D v3 operator-(v3 a) { return V3(-a.x, -a.y, -a.z); }
PXB_D v3 operator*(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
PXB_D v3 vmul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
PXB_D float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          // PxVec3::dot
PXB_D v3 cross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PXB_D v3 vabs(v3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
PXB_D float fmin_(float a, float b) { return a < b ? a : b; }  // SSE min/max semantics for ordered inputs
PXB_D float fmax_(float a, float b) { return a > b ? a : b; }
PXB_D v3 vmin(v3 a, v3 b) { return V3(fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)); }
PXB_D v3 vmax(v3 a, v3 b) { return V3(fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)); }
PXB_D v3 scaleadd(v3 a, float s, v3 b) { return V3(a.x * s + b.x, a.y * s + b.y, a.z * s + b.z); }      // a*s+b
PXB_D v3 negscalesub(v3 a, float s, v3 b) { return V3(b.x - a.x * s, b.y - a.y * s, b.z - a.z * s); }  // b-a*s
PXB_D float lensq(v3 a) { return dot(a, a); }

PXB_D q4 conj(q4 a) { return Q4(-a.x, -a.y, -a.z, a.w); }
PXB_D float qdot(q4 a, q4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
PXB_D q4 qmul(q4 a, q4 b) {  // PxQuat::operator*
  return Q4(a.w * b.x + b.w * a.x + a.y * b.z - b.y * a.z, a.w * b.y + b.w * a.y + a.z * b.x - b.z * a.x,
            a.w * b.z + b.w * a.z + a.x * b.y - b.x * a.y, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
PXB_D q4 qnormalized(q4 a) { const float s = 1.0f / sqrtf(qdot(a, a)); return Q4(a.x * s, a.y * s, a.z * s, a.w * s); }
PXB_D v3 qrot(q4 q, v3 v) {  // PxQuat::rotate
  const float vx = 2.0f * v.x, vy = 2.0f * v.y, vz = 2.0f * v.z;
  const float w2 = q.w * q.w - 0.5f;
  const float dot2 = (q.x * vx + q.y * vy + q.z * vz);
  return V3((vx * w2 + (q.y * vz - q.z * vy) * q.w + q.x * dot2), (vy * w2 + (q.z * vx - q.x * vz) * q.w + q.y * dot2),
            (vz * w2 + (q.x * vy - q.y * vx) * q.w + q.z * dot2));
}
PXB_D v3 qrotinv(q4 q, v3 v) {  // PxQuat::rotateInv
  const float vx = 2.0f * v.x, vy = 2.0f * v.y, vz = 2.0f * v.z;
  const float w2 = q.w * q.w - 0.5f;
  const float dot2 = (q.x * vx + q.y * vy + q.z * vz);
  return V3((vx * w2 - (q.y * vz - q.z * vy) * q.w + q.x * dot2), (vy * w2 - (q.z * vx - q.x * vz) * q.w + q.y * dot2),
            (vz * w2 - (q.x * vy - q.y * vx) * q.w + q.z * dot2));
}
PXB_D v3 qbasis0(q4 q) {  // PxQuat::getBasisVector0
  const float x2 = q.x * 2.0f, w2 = q.w * 2.0f;
  return V3((q.w * w2) - 1.0f + q.x * x2, (q.z * w2) + q.y * x2, (-q.y * w2) + q.z * x2);
}
PXB_D m33 mfromq(q4 q) {  // PxMat33(const PxQuat&)
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
PXB_D v3 mmul(const m33& m, v3 v) {  // M*v, left-associated sums (also M33MulV3)
  return V3(m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z, m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z,
            m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z);
}
PXB_D m33 mtranspose(const m33& m) {
  m33 r; r.c0 = V3(m.c0.x, m.c1.x, m.c2.x); r.c1 = V3(m.c0.y, m.c1.y, m.c2.y); r.c2 = V3(m.c0.z, m.c1.z, m.c2.z); return r;
}
PXB_D v3 xftransform(const xf& t, v3 v) { return qrot(t.q, v) + t.p; }
PXB_D v3 xftransforminv(const xf& t, v3 v) { return qrotinv(t.q, v - t.p); }
PXB_D xf xfinvmul(const xf& a, const xf& b) {  // a.transformInv(b)
  xf r; const q4 qi = conj(a.q); r.p = qrot(qi, b.p - a.p); r.q = qmul(qi, b.q); return r;
}

// ---- vector-layer ("aos") op order ----
PXB_D float adot(v3 a, v3 b) { return (a.x * b.x + a.z * b.z) + (a.y * b.y); }                  // V3Dot
PXB_D float adot4(q4 a, q4 b) { return (a.x * b.x + a.z * b.z) + (a.y * b.y + a.w * b.w); }     // V4Dot
PXB_D float alen(v3 a) { return sqrtf(adot(a, a)); }
PXB_D v3 anormalize(v3 a) { const float l = sqrtf(adot(a, a)); return V3(a.x / l, a.y / l, a.z / l); }
PXB_D v3 aqrot_noscale(q4 q, v3 v) {
  const v3 u = V3(q.x, q.y, q.z);
  const float w2 = q.w * q.w + (-0.5f);
  const v3 a = v * w2;
  const v3 t = scaleadd(cross(u, v), q.w, a);
  return scaleadd(u, adot(u, v), t);
}
PXB_D v3 aqrot(q4 q, v3 v) { return aqrot_noscale(q, v) * 2.0f; }                                // QuatRotate
PXB_D v3 aqrotinv(q4 q, v3 v) {                                                                 // QuatRotateInv
  const v3 u = V3(q.x, q.y, q.z);
  const float w2 = q.w * q.w + (-0.5f);
  const v3 a = v * w2;
  const v3 t = negscalesub(cross(u, v), q.w, a);
  return scaleadd(u, adot(u, v), t) * 2.0f;
}
PXB_D v3 aqrot_normalize(q4 q, v3 v) { return anormalize(aqrot_noscale(q, v)); }                // QuatRotateAndNormalize
PXB_D q4 aqmul(q4 a, q4 b) {                                                                    // QuatMul
  const v3 ia = V3(a.x, a.y, a.z), ib = V3(b.x, b.y, b.z);
  const float real = a.w * b.w - dot(ia, ib);
  const v3 im = ((ia * b.w) + (ib * a.w)) + cross(ia, ib);
  return Q4(im.x, im.y, im.z, real);
}
PXB_D v3 aqbasis0(q4 q) {                                                                       // QuatGetBasisVector0
  const float x2 = q.x * 2.0f, w2 = q.w * 2.0f;
  const v3 a = V3(q.x, q.y, q.z) * x2;
  const v3 ab = scaleadd(V3(q.w, q.z, -q.y), w2, a);
  return V3(ab.x - 1.0f, ab.y, ab.z);
}
PXB_D v3 axftransform(const xf& t, v3 v) { return scaleadd(aqrot_noscale(t.q, v), 2.0f, t.p); } // QuatTransform
PXB_D xf axfinvmul(const xf& a, const xf& b) {                                                  // PxTransformV::transformInv
  xf r; const q4 qi = conj(a.q); r.p = aqrot(qi, b.p - a.p); r.q = aqmul(qi, b.q); return r;
}
PXB_D v3 amtmul(const m33& m, v3 v) { return V3(adot(m.c0, v), adot(m.c1, v), adot(m.c2, v)); } // M33TrnspsMulV3
PXB_D v3 amxftransform(const mxf& t, v3 v) { return t.p + mmul(t.r, v); }
PXB_D mxf amxfinvmul(const mxf& a, const mxf& b) {                                              // PxMatTransformV::transformInv
  mxf r; const m33 at = mtranspose(a.r);
  r.r.c0 = mmul(at, b.r.c0); r.r.c1 = mmul(at, b.r.c1); r.r.c2 = mmul(at, b.r.c2);
  r.p = amtmul(a.r, b.p - a.p);
  return r;
}
PXB_D m33 amfromq(q4 q) {                        // 3-output QuatGetMat33V (PxVecMathSSE.h:54-69) = PxMat33Padded(const PxQuat&)
  const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z, w2 = q.w + q.w;
  const float wx = x2 * q.w, wy = y2 * q.w, wz = z2 * q.w, ww1 = w2 * q.w + (-1.0f);
  m33 r;
  r.c0 = V3(q.x * x2 + ww1, q.y * x2 + wz, q.z * x2 + (-wy));
  r.c1 = V3(q.x * y2 + (-wz), q.y * y2 + ww1, q.z * y2 + wx);
  r.c2 = V3(q.x * z2 + wy, q.y * z2 + (-wx), q.z * z2 + ww1);
  return r;
}
PXB_D mxf amxffromxf(const xf& t) { mxf m; m.p = t.p; m.r = amfromq(t.q); return m; }
