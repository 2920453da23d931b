// pxb_solver.cuh -- TGS contact-constraint prep / solve / integration math (stages 3-4), device functions.
//
// Matches the reference CPU TGS path (the parity target):
//   unconstrained velocity  physx/source/lowleveldynamics/src/DyBodyCoreIntegrator.h:39-81
//   solver body setup       DyTGSDynamics.cpp:154-243 (copyToSolverBodyDataStep), common/src/CmUtils.h:57-70
//   friction correlation    DyFrictionCorrelation.cpp:56-330, DyContactPrepShared.h:52-131
//   contact prep            DyTGSContactPrep.cpp:322-823
//   solve                   DyTGSContactPrep.cpp:1492-1873
//   integration             DyTGSDynamics.cpp:1403-1476, :1549-1580
// Scope: rigid dynamic vs rigid dynamic/static, one contact patch per pair (all primitive PCM pairs),
// rigid (restitution >= 0) contacts, no dominance / kinematics.
#pragma once
#include "pxb_math.cuh"
#include "pxb_np.cuh"

#define PXB_FRICTION_F4 8   // float4 slots per persistent friction-patch record

struct FrictionPatch {      // Dy::FrictionPatch (lowleveldynamics/src/DyFrictionPatch.h)
  v3 body0Normal, body1Normal; v3 body0Anchors[2], body1Anchors[2]; q4 relativeQuat;
  int anchorCount, broken, valid; float sf, df, rest;
};
PXB_D void friction_load(FrictionPatch& f, const float4* __restrict__ r) {
  const float4 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], g = r[5], h = r[6];
  f.body0Normal = V3(a.x, a.y, a.z); f.anchorCount = __float_as_int(a.w);
  f.body1Normal = V3(b.x, b.y, b.z); f.broken = __float_as_int(b.w);
  f.body0Anchors[0] = V3(c.x, c.y, c.z); f.valid = __float_as_int(c.w);
  f.body0Anchors[1] = V3(d.x, d.y, d.z); f.sf = d.w;
  f.body1Anchors[0] = V3(e.x, e.y, e.z); f.df = e.w;
  f.body1Anchors[1] = V3(g.x, g.y, g.z); f.rest = g.w;
  f.relativeQuat = Q4(h);
}
PXB_D void friction_store(const FrictionPatch& f, float4* __restrict__ r) {
  r[0] = F4(f.body0Normal, __int_as_float(f.anchorCount)); r[1] = F4(f.body1Normal, __int_as_float(f.broken));
  r[2] = F4(f.body0Anchors[0], __int_as_float(f.valid)); r[3] = F4(f.body0Anchors[1], f.sf);
  r[4] = F4(f.body1Anchors[0], f.df); r[5] = F4(f.body1Anchors[1], f.rest); r[6] = F4(f.relativeQuat);
}

// PxFrictionPatch write-back for PxDirectGPUAPI::copyContactData (the reference's writeBackContactBlockFriction, gpusolver/src/CUDA/solverBlockCommon.cuh:33-66):
// per pair 4 float4 = impulse at anchor 0 (w: anchor count), impulse at anchor 1, world anchor 0, world anchor 1.  impulse[k] = t0 * applied[2k] + t1 * applied[2k+1];
// the anchors are body 0's local anchors of the pair's friction patch in body 0's start-of-step frame (what contact prep solved with).
PXB_D void friction_report_store(float4* __restrict__ out, uint32_t i, int numFriction, float4 t0, float4 t1, float4 fap, const float4* __restrict__ frRec, float4 p0, float4 q0) {
  float4* o = out + (size_t)i * 4;
  const int anchors = numFriction >> 1;
  const v3 T0 = V3(t0.x, t0.y, t0.z), T1 = V3(t1.x, t1.y, t1.z);
  const v3 i0 = anchors >= 1 ? T0 * fap.x + T1 * fap.y : V3(0, 0, 0), i1 = anchors >= 2 ? T0 * fap.z + T1 * fap.w : V3(0, 0, 0);
  xf tm; tm.p = V3(p0.x, p0.y, p0.z); tm.q = Q4(q0);
  const float4 a0 = frRec[2], a1 = frRec[3];
  const v3 w0 = anchors >= 1 ? axftransform(tm, V3(a0.x, a0.y, a0.z)) : V3(0, 0, 0), w1 = anchors >= 2 ? axftransform(tm, V3(a1.x, a1.y, a1.z)) : V3(0, 0, 0);
  o[0] = F4(i0, __int_as_float(anchors)); o[1] = F4(i1, 0.f); o[2] = F4(w0, 0.f); o[3] = F4(w1, 0.f);
}

PXB_D void transform_inertia(v3 d, const m33& M, m33& out) {  // Cm::transformInertiaTensor, M(r,c) = column c row r
  const float axx = d.x * M.c0.x, axy = d.x * M.c0.y, axz = d.x * M.c0.z;
  const float byx = d.y * M.c1.x, byy = d.y * M.c1.y, byz = d.y * M.c1.z;
  const float czx = d.z * M.c2.x, czy = d.z * M.c2.y, czz = d.z * M.c2.z;
  const float m00 = axx * M.c0.x + byx * M.c1.x + czx * M.c2.x;
  const float m11 = axy * M.c0.y + byy * M.c1.y + czy * M.c2.y;
  const float m22 = axz * M.c0.z + byz * M.c1.z + czz * M.c2.z;
  const float m01 = axx * M.c0.y + byx * M.c1.y + czx * M.c2.y;
  const float m02 = axx * M.c0.z + byx * M.c1.z + czx * M.c2.z;
  const float m12 = axy * M.c0.z + byy * M.c1.z + czy * M.c2.z;
  out.c0 = V3(m00, m01, m02); out.c1 = V3(m01, m11, m12); out.c2 = V3(m02, m12, m22);
}

PXB_D void unconstrained_velocity(v3 gravity, float dt, float linDamping, float angDamping, float maxLinVelSq, float maxAngVelSq, v3& lv, v3& av) {
  v3 l = lv, a = av;
  const float oml = 1.0f - linDamping * dt, oma = 1.0f - angDamping * dt;
  l = l + (gravity * dt) * 1.0f;
  const float lm = oml >= 0.f ? oml : 0.f, am = oma >= 0.f ? oma : 0.f;
  l = l * lm; a = a * am;
  const float lsq = lensq(l); if (lsq > maxLinVelSq) l = l * sqrtf(maxLinVelSq / lsq);
  const float asq = lensq(a); if (asq > maxAngVelSq) a = a * sqrtf(maxAngVelSq / asq);
  lv = l; av = a;
}

// PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES: the torque-free gyroscopic term applied to the angular velocity when the solver body is built
// (copyToSolverBodyDataStep DyTGSDynamics.cpp:177-193 = copyToSolverBodyData DyRigidBodyToSolverBody.cpp:53-70; scalar PxVec3 / PxQuat arithmetic)
PXB_D v3 gyroscopic(v3 av, v3 invInertia, q4 q, float dt) {
  const v3 localInertia = V3(invInertia.x == 0.f ? 0.f : 1.f / invInertia.x, invInertia.y == 0.f ? 0.f : 1.f / invInertia.y, invInertia.z == 0.f ? 0.f : 1.f / invInertia.z);
  const v3 localAngVel = qrotinv(q, av);
  const v3 origMom = vmul(localInertia, localAngVel);
  const v3 c = cross(localAngVel, origMom); const v3 torque = V3(-c.x, -c.y, -c.z);
  v3 newMom = origMom + torque * dt;
  const float denom = sqrtf(newMom.x * newMom.x + newMom.y * newMom.y + newMom.z * newMom.z);
  const float ratio = denom > 0.f ? sqrtf(origMom.x * origMom.x + origMom.y * origMom.y + origMom.z * origMom.z) / denom : 0.f;
  newMom = newMom * ratio;
  return av + qrot(q, vmul(invInertia, newMom) - localAngVel);
}

// PxRigidDynamicLockFlag bits (linear x,y,z = 1,2,4; angular x,y,z = 8,16,32) travel in the unused .z lane of the body's second inertia float4
PXB_D v3 lock3(v3 v, uint32_t bits) { if (bits & 1u) v.x = 0.f; if (bits & 2u) v.y = 0.f; if (bits & 4u) v.z = 0.f; return v; }
// External force / torque for this step (PxDirectGPUAPI eFORCE / eTORQUE = PxRigidBody::addForce / addTorque(eFORCE)):
// NpRigidBodyTemplate.h:507-528 (linAcc = F * invMass, angAcc = world inverse inertia * T, inertia as in :315-320) and
// Sc::BodySim::updateForces ScBodySim.cpp:656-720 (v += acc * dt before the unconstrained-velocity pass).
static __device__ __noinline__ void apply_external_force(v3 F, v3 T, float invMass, float4 invI, q4 q, float dt, v3& lv, v3& av) {   // rare path: out of line, the solve kernels are register bound
  if (F.x != 0.f || F.y != 0.f || F.z != 0.f) { const v3 linAcc = F * invMass; lv = lv + (V3(0, 0, 0) + linAcc * dt); }
  if (T.x != 0.f || T.y != 0.f || T.z != 0.f) {
    const m33 rot = amfromq(q);
    m33 invIW; transform_inertia(V3(invI.x, invI.y, invI.z), rot, invIW);
    const v3 angAcc = mmul(invIW, T);
    av = av + (V3(0, 0, 0) + angAcc * dt);
  }
}
// integrateCoreStep: returns updated (p, deltaQ, deltaLinDt, deltaAngDt); lock flags zero the locked velocity components first (DyTGSDynamics.cpp:1405-1422)
PXB_D void integrate_core_step(v3& linVel, v3& angState, const m33& sqrtInvInertia, float dt, v3& p, q4& deltaQ, v3& dLin, v3& dAng, uint32_t lock) {
  if (lock) { linVel = lock3(linVel, lock & 7u); angState = lock3(angState, (lock >> 3) & 7u); }
  const v3 delta = linVel * dt;
  const v3 w3 = mmul(sqrtInvInertia, angState);
  const float w2 = lensq(w3);
  p = p + delta;
  if (w2 != 0.0f) {
    const float w = sqrtf(w2);
    const float v = dt * w * 0.5f;
    float s, q; sincosf(v, &s, &q);   // PxSinCos: one shared range reduction; same values as sinf / cosf
    s /= w;
    const v3 pqr = w3 * s;
    q4 r = qmul(Q4(pqr.x, pqr.y, pqr.z, 0.f), deltaQ);
    r.x += deltaQ.x * q; r.y += deltaQ.y * q; r.z += deltaQ.z * q; r.w += deltaQ.w * q;
    deltaQ = qnormalized(r);
  }
  dAng = dAng + angState * dt;
  dLin = dLin + delta;
}

// Friction patch correlation for one contact patch (see header of oracle/pxo_solver.h for the derivation
// from getFrictionPatches / correlatePatches / growPatches).
PXB_D void friction_correlate(FrictionPatch& fp, const Contacts& c, const xf& f0, const xf& f1, float sf, float df, float rest,
                              float correlationDistance, float frictionOffsetThreshold) {
  const float SAME_NORMAL = 0.999f;
  bool keepOld = false; v3 oldWorldNormal = V3(0, 0, 0);
  if (fp.valid && !fp.broken && fp.anchorCount != 0) {
    const xf b1To0 = xfinvmul(f0, f1);
    if (dot(fp.body0Normal, qrot(b1To0.q, fp.body1Normal)) > SAME_NORMAL) {
      bool separated = false;
      for (int a = 0; a < fp.anchorCount; ++a) {
        const v3 p1 = xftransform(b1To0, fp.body1Anchors[a]);
        if (!(fabsf(dot(fp.body0Anchors[a] - p1, fp.body0Normal)) < correlationDistance)) { separated = true; break; }
      }
      if (!separated) { keepOld = true; oldWorldNormal = qrot(f0.q, fp.body0Normal); }
    }
  }
  v3 bmin = c.point[0], bmax = c.point[0];
  for (int i = 1; i < c.count; ++i) { bmin = vmin(bmin, c.point[i]); bmax = vmax(bmax, c.point[i]); }
  const v3 pn = c.normal;
  const bool correlated = keepOld && !((dot(pn, oldWorldNormal) < SAME_NORMAL) || fp.rest != rest || fp.sf != sf || fp.df != df);
  if (!correlated) {
    fp.body0Normal = qrotinv(f0.q, pn); fp.body1Normal = qrotinv(f1.q, pn);
    fp.relativeQuat = qmul(conj(f0.q), f1.q);
    fp.anchorCount = 0; fp.broken = 0; fp.sf = sf; fp.df = df; fp.rest = rest;
  }
  fp.valid = 1;
  if (fp.anchorCount == 2) {
    const float diagSq = lensq(bmax - bmin);
    const float anchorSq = lensq(fp.body0Anchors[0] - fp.body0Anchors[1]);
    if ((anchorSq * 4.f) >= diagSq) return;
    fp.anchorCount = 0;
  }
  v3 wa0 = V3(0, 0, 0), wa1 = V3(0, 0, 0); int anchorCount = 0; float pointDistSq = 0.f;
  if (fp.anchorCount == 1) { wa0 = xftransform(f0, fp.body0Anchors[0]); anchorCount = 1; }
  const float eps = 1e-8f;
  for (int j = 0; j < c.count; ++j) {
    const v3 wp = c.point[j];
    if (c.sep[j] < frictionOffsetThreshold) {
      if (anchorCount == 0) { wa0 = wp; anchorCount = 1; }
      else if (anchorCount == 1) { pointDistSq = lensq(wp - wa0); if (pointDistSq > eps) { wa1 = wp; anchorCount = 2; } }
      else {
        const float d0 = lensq(wp - wa0), d1 = lensq(wp - wa1);
        if (d0 > d1) { if (d0 > pointDistSq) { wa1 = wp; pointDistSq = d0; } }
        else if (d1 > pointDistSq) { wa0 = wp; pointDistSq = d1; }
      }
    }
  }
  if (fp.anchorCount < 1 && anchorCount >= 1) { fp.body0Anchors[0] = xftransforminv(f0, wa0); fp.body1Anchors[0] = xftransforminv(f1, wa0); }
  if (fp.anchorCount < 2 && anchorCount >= 2) { fp.body0Anchors[1] = xftransforminv(f0, wa1); fp.body1Anchors[1] = xftransforminv(f1, wa1); }
  if (anchorCount == 0) { fp.body0Anchors[0] = V3(0, 0, 0); fp.body1Anchors[0] = V3(0, 0, 0); }
  fp.anchorCount = anchorCount;
}

struct SolverParams {
  float dt, stepDt, invStepDt, invTotalDt, biasCoefficient, bounceThreshold, frictionOffsetThreshold, correlationDistance;
  float restDistance, staticFriction, dynamicFriction, restitution;
};

// One prepared contact point / friction row, as stored in the SoA row arrays.
struct SPoint { v3 raXnI, rbXnI; float velMultiplier, separation, biasCoefficient, targetVelocity, recipResponse, appliedForce; };
struct SFriction { v3 normal; float error; v3 raXnI; float targetVel; v3 rbXnI; float velMultiplier; float appliedForce, frictionScale, biasScale; };

PXB_D void prep_point(SPoint& s, v3 point, float separation, v3 normal, v3 p0, v3 p1, const m33& sI0, const m33& sI1,
                      v3 angVel0, v3 angVel1, float norVel0, float norVel1, float invMassNorLenSq0, float invMassNorLenSq1,
                      const SolverParams& P, float invDtp8, const bool kin1 = false) {
  const v3 ra = point - p0, rb = point - p1;
  const v3 raXn = cross(ra, normal), rbXn = cross(rb, normal);
  const float angV0 = adot(raXn, angVel0), angV1 = adot(rbXn, angVel1);
  const float vrel1 = norVel0 + angV0, vrel2 = norVel1 + angV1;
  const float vrel = vrel1 - vrel2;
  const v3 raXnI = mmul(sI0, raXn), rbXnI = mmul(sI1, rbXn);
  const float i0 = adot(raXnI, raXnI) * 1.f, i1 = adot(rbXnI, rbXnI) * 1.f;
  const float resp0 = invMassNorLenSq0 + i0, resp1 = i1 - invMassNorLenSq1;
  const float unitResponse = resp0 + resp1;
  const float penetration = separation - P.restDistance;
  const bool isSeparated = penetration > 0.f;
  const float penetrationInvDt = penetration * P.invTotalDt;
  const bool isGreater2 = (P.restitution > 0.f) && (P.bounceThreshold > vrel) && ((-vrel) > penetrationInvDt);
  const float ratio = P.dt + (isGreater2 ? (penetration / vrel) : (-P.dt));
  const float recipResponse = (unitResponse > 0.f) ? (1.f / unitResponse) : 0.f;
  const float biasCoeff = -(isSeparated ? P.invStepDt : invDtp8);
  float totalError = penetration;
  float targetVelocity = 0.f + (isGreater2 ? ((-vrel) * P.restitution) : 0.f);
  totalError = targetVelocity * ratio + totalError;
  if (kin1) targetVelocity = targetVelocity + vrel2;   // kinematic body B: its velocity along the normal is the row's target velocity (DyTGSContactPrep.cpp:406-409)
  s.raXnI = raXnI; s.rbXnI = rbXnI; s.velMultiplier = recipResponse; s.separation = totalError;
  s.biasCoefficient = biasCoeff; s.targetVelocity = targetVelocity; s.recipResponse = recipResponse; s.appliedForce = 0.f;
}

PXB_D void prep_friction_row(SFriction& f, v3 ra, v3 rb, v3 error, v3 tdir, const m33& sI0, const m33& sI1,
                             float invMassNorLenSq0, float invMassNorLenSq1, float scale, float frictionScale, float frictionBiasScale) {
  const v3 raXn = cross(ra, tdir), rbXn = cross(rb, tdir);
  const v3 raXnI = mmul(sI0, raXn), rbXnI = mmul(sI1, rbXn);
  const float resp0 = invMassNorLenSq0 + adot(raXnI, raXnI) * 1.f;
  const float resp1 = adot(rbXnI, rbXnI) * 1.f - invMassNorLenSq1;
  const float unitResponse = resp0 + resp1;
  f.normal = tdir; f.error = adot(error, tdir); f.raXnI = raXnI; f.targetVel = 0.f; f.rbXnI = rbXnI;
  f.velMultiplier = (unitResponse > 0.f) ? (scale / unitResponse) : 0.f;
  f.appliedForce = 0.f; f.frictionScale = frictionScale; f.biasScale = frictionBiasScale;
}
