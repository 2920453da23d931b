/* pxo_pgs.h -- CPU restatement of the reference's PGS rigid-body contact solver, scalar path (TEST INFRASTRUCTURE).
 * Follows:
 *   solver body data        physx/source/lowleveldynamics/src/DyRigidBodyToSolverBody.cpp:38-112 (copyToSolverBodyData)
 *   contact prep            DyContactPrep.cpp:60-365 (setupFinalizeSolverConstraints), DyContactPrepShared.h:298-391 (constructContactConstraint)
 *   solve                   DySolverConstraintsShared.h:49-109 (solveDynamicContacts), DySolverConstraints.cpp:221-371 (solveContact;
 *                           solveContact_BStatic :373-506 is the same arithmetic with the static body's zero terms dropped)
 *   conclude / write-back   DySolverConstraints.cpp:508-551, :553-640
 *   iteration loop          DySolverControl.cpp:163-405 (solveV_Blocks: position iterations, friction only in the last three,
 *                           the last one concludes; saveMotionVelocities; velocity iterations, at least one, the last writes back)
 *   integration             DyBodyCoreIntegrator.h:83-185 (integrateCore), DyDynamics.cpp:1398-1423
 * PGS solves for velocity DELTAS: PxSolverBody starts at zero and the pre-solver velocity is folded into the rows' target
 * velocity (constructContactConstraint: targetVelocity -= vrel).  Restrictions as pxo_solver.h. */
#ifndef PXO_PGS_H
#define PXO_PGS_H
#include "pxo_solver.h"

typedef struct { v3 linVel, angVel; m33 sqrtInvInertia; float invMass, penBiasClamp; xf body2World; } PxoPgsBodyData;   /* PxSolverBodyData */
typedef struct { v3 linVel, angState; } PxoPgsBody;                                                                      /* PxSolverBody (deltas) */

typedef struct { v3 raXn, rbXn; float velMultiplier, biasedErr, unbiasedErr, impulseMultiplier, maxImpulse, appliedForce; } PxoPgsPoint;
typedef struct { v3 normal, raXn, rbXn; float appliedForce, velMultiplier, bias, targetVel; } PxoPgsFriction;
typedef struct {
  int body0, body1;
  v3 normal; float invMass0, invMass1, angDom0, angDom1, staticFriction, dynamicFriction;
  int numNormal, numFriction, broken;
  PxoPgsPoint pts[PXO_MAX_CONTACTS];
  PxoPgsFriction fr[4];
} PxoPgsConstraint;

static inline void pxo_pgs_body_data_init(PxoPgsBodyData* d, v3 lv, v3 av, float invMass, v3 invInertia, const xf* pose, float maxDepenVel, uint32_t lockFlags) {
  /* copyToSolverBodyData :72-98: only the ANGULAR locks reach the data (the linear ones are written to data.linearVelocity before it is overwritten with `lin`) */
  av = pxo_lock3(av, (lockFlags >> 3) & 7u);
  const m33 rot = am33fromq(pose->q);
  const v3 sqrtInvI = V3(invInertia.x == 0.f ? 0.f : sqrtf(invInertia.x), invInertia.y == 0.f ? 0.f : sqrtf(invInertia.y), invInertia.z == 0.f ? 0.f : sqrtf(invInertia.z));
  pxo_transform_inertia(sqrtInvI, &rot, &d->sqrtInvInertia);
  d->linVel = lv; d->angVel = av; d->invMass = invMass; d->penBiasClamp = -maxDepenVel; d->body2World = *pose;
}
static inline void pxo_pgs_static_data_init(PxoPgsBodyData* d) { memset(d, 0, sizeof(*d)); d->penBiasClamp = -FLT_MAX; d->body2World.q = Q4(0, 0, 0, 1); }

/* setupFinalizeSolverConstraints for one friction patch / one contact patch */
static inline void pxo_pgs_prep_contact(PxoPgsConstraint* k, const PxoContacts* c, const PxoFrictionPatch* fp, const PxoPgsBodyData* d0, const PxoPgsBodyData* d1,
                                        const xf* bodyFrame0, const xf* bodyFrame1, float sf, float df, float restitution, float restDistance,
                                        float invDt, float dt, float bounceThreshold) {
  (void)dt;
  const float angD0 = 1.f, angD1 = 1.f;
  const float invMass0_dom0 = 1.f * d0->invMass, invMass1_dom1 = (-1.f) * d1->invMass;
  const float maxPenBias = fmaxf_(d0->penBiasClamp, d1->penBiasClamp);
  const v3 linVel0 = d0->linVel, linVel1 = d1->linVel, angVel0 = d0->angVel, angVel1 = d1->angVel;
  const float invDtp8 = invDt * 0.8f;
  k->invMass0 = invMass0_dom0; k->invMass1 = -invMass1_dom1; k->angDom0 = angD0; k->angDom1 = angD1; k->broken = 0;
  const v3 normal = c->normal;
  const float normalLenSq = alensq(normal);
  const v3 nv = v3sub(v3mul(normal, linVel0), v3mul(normal, linVel1));   /* V3NegMulSub(normal, linVel1, V3Mul(normal, linVel0)) */
  const float norVel = (nv.x + nv.y) + nv.z;
  const float invMassNorLenSq0 = invMass0_dom0 * normalLenSq, invMassNorLenSq1 = invMass1_dom1 * normalLenSq;
  k->normal = normal; k->numNormal = c->count;
  for (int j = 0; j < c->count; ++j) {   /* constructContactConstraint (solverOffsetSlop = 0, ccdMaxSeparation = PX_MAX_F32, target velocity 0) */
    PxoPgsPoint* s = &k->pts[j];
    const v3 point = c->point[j]; const float separation = c->sep[j];
    const float cTargetVel = 0.f;
    const v3 ra = v3sub(point, bodyFrame0->p), rb = v3sub(point, bodyFrame1->p);
    const v3 raXn = v3cross(ra, normal), rbXn = v3cross(rb, normal);
    const float vRelAng = adot(raXn, angVel0) - adot(rbXn, angVel1);
    const float vrel = norVel + vRelAng;
    const v3 raXnI = m33mul(&d0->sqrtInvInertia, raXn), rbXnI = m33mul(&d1->sqrtInvInertia, rbXn);
    const float resp0 = invMassNorLenSq0 + adot(raXnI, raXnI) * angD0;
    const float resp1 = adot(rbXnI, rbXnI) * angD1 - invMassNorLenSq1;
    const float unitResponse = resp0 + resp1;
    const float penetration = separation - restDistance;
    const float penetrationInvDt = penetration * invDt;
    const int isSeparated = penetration >= 0.f;
    const int collidingWithVrel = (-vrel) > penetrationInvDt;
    const int isGreater2 = (restitution > 0.f) && (bounceThreshold > vrel) && collidingWithVrel;
    float targetVelocity = cTargetVel + (isGreater2 ? ((-vrel) * restitution) : 0.f);
    targetVelocity = targetVelocity - vrel;
    const float recipResponse = (unitResponse > 0.f) ? (1.0f / unitResponse) : 0.f;
    const float velMultiplier = recipResponse;
    const float penetrationInvDtScaled = isSeparated ? penetrationInvDt : (penetration * invDtp8);
    float scaledBias = velMultiplier * fmaxf_(maxPenBias, penetrationInvDtScaled);
    if (isGreater2) scaledBias = 0.f;   /* ccdSeparationCondition is always true without speculative CCD */
    s->biasedErr = targetVelocity * velMultiplier + (-scaledBias);
    s->unbiasedErr = targetVelocity * velMultiplier + (isGreater2 ? 0.f : (-fmaxf_(scaledBias, 0.f)));
    s->impulseMultiplier = 1.0f; s->raXn = raXnI; s->rbXn = rbXnI; s->velMultiplier = velMultiplier; s->maxImpulse = FLT_MAX; s->appliedForce = 0.f;
  }
  const float frictionCoefficient = (fp->anchorCount == 2) ? 0.5f : 1.f;
  k->staticFriction = sf * frictionCoefficient; k->dynamicFriction = df * frictionCoefficient;
  const int haveFriction = fp->anchorCount != 0;
  k->numFriction = haveFriction ? fp->anchorCount * 2 : 0;
  if (haveFriction) {
    const v3 linVrel = v3sub(linVel0, linVel1);
    const v3 t0Fallback1 = V3(0.f, -normal.z, normal.y), t0Fallback2 = V3(-normal.y, normal.x, 0.f);
    const v3 t0Fallback = (0.70710678f > fabsf(normal.x)) ? t0Fallback1 : t0Fallback2;
    v3 t0 = v3sub(linVrel, v3scale(normal, adot(normal, linVrel)));
    t0 = (alensq(t0) > 0.0001f) ? t0 : t0Fallback;
    t0 = anormalize(t0);
    const v3 t1 = v3cross(normal, t0);   /* not normalised in the PGS path */
    for (int j = 0; j < fp->anchorCount; ++j) {
      const v3 ra = aqrot(bodyFrame0->q, fp->body0Anchors[j]), rb = aqrot(bodyFrame1->q, fp->body1Anchors[j]);
      const v3 error = v3sub(v3add(ra, bodyFrame0->p), v3add(rb, bodyFrame1->p));
      for (int t = 0; t < 2; ++t) {
        const v3 tdir = t == 0 ? t0 : t1;
        PxoPgsFriction* f = &k->fr[j * 2 + t];
        const v3 raXn = v3cross(ra, tdir), rbXn = v3cross(rb, tdir);
        const v3 raXnI = m33mul(&d0->sqrtInvInertia, raXn), rbXnI = m33mul(&d1->sqrtInvInertia, rbXn);
        const float resp0 = invMass0_dom0 + angD0 * adot(raXnI, raXnI);
        const float resp1 = angD1 * adot(rbXnI, rbXnI) - invMass1_dom1;
        const float resp = resp0 + resp1;
        const float velMultiplier = (resp > 0.f) ? (0.8f / resp) : 0.f;
        float targetVel = 0.f;   /* V3Dot(tvel, t) with a zero contact target velocity */
        const float vrel1 = adot(tdir, linVel0) + adot(raXn, angVel0);
        const float vrel2 = adot(tdir, linVel1) + adot(rbXn, angVel1);
        const float vrel = vrel1 - vrel2;
        targetVel = targetVel - vrel;
        f->normal = tdir; f->appliedForce = 0.f; f->raXn = raXnI; f->velMultiplier = velMultiplier; f->rbXn = rbXnI; f->bias = adot(tdir, error) * invDt; f->targetVel = targetVel;
      }
    }
  }
}

/* solveContact (b1 may be the zero static body: its terms are exact zeros) */
static inline void pxo_pgs_solve_contact(PxoPgsConstraint* k, PxoPgsBody* b0, PxoPgsBody* b1, int doFriction) {
  v3 linVel0 = b0->linVel, linVel1 = b1->linVel, angState0 = b0->angState, angState1 = b1->angState;
  const float invMassA = k->invMass0, invMassB = k->invMass1, angDom0 = k->angDom0, angDom1 = k->angDom1;
  const v3 n = k->normal;
  float accum = 0.f;
  {
    const v3 delLinVel0 = v3scale(n, invMassA), delLinVel1 = v3scale(n, invMassB);
    for (int i = 0; i < k->numNormal; ++i) {
      PxoPgsPoint* c = &k->pts[i];
      const v3 raXn = c->raXn, rbXn = c->rbXn;
      const float appliedForce = c->appliedForce, velMultiplier = c->velMultiplier;
      const v3 v0 = v3add(v3mul(linVel0, n), v3mul(angState0, raXn));
      const v3 v1 = v3add(v3mul(linVel1, n), v3mul(angState1, rbXn));
      const v3 dv = v3sub(v0, v1);
      const float normalVel = (dv.x + dv.y) + dv.z;
      const float _deltaF = fmaxf_(c->biasedErr - normalVel * velMultiplier, -appliedForce);
      const float _newForce = c->impulseMultiplier * appliedForce + _deltaF;
      const float newForce = fminf_(_newForce, c->maxImpulse);
      const float deltaF = newForce - appliedForce;
      linVel0 = v3scaleadd(delLinVel0, deltaF, linVel0);
      linVel1 = v3negscalesub(delLinVel1, deltaF, linVel1);
      angState0 = v3scaleadd(raXn, deltaF * angDom0, angState0);
      angState1 = v3negscalesub(rbXn, deltaF * angDom1, angState1);
      c->appliedForce = newForce;
      accum = accum + newForce;
    }
  }
  if (doFriction && k->numFriction) {
    const float maxFrictionImpulse = k->staticFriction * accum, maxDynFrictionImpulse = k->dynamicFriction * accum;
    const float negMaxDynFrictionImpulse = -maxDynFrictionImpulse;
    int broken = 0;
    for (int i = 0; i < k->numFriction; ++i) {
      PxoPgsFriction* f = &k->fr[i];
      const v3 normal = f->normal, raXn = f->raXn, rbXn = f->rbXn;
      const float appliedForce = f->appliedForce, bias = f->bias, velMultiplier = f->velMultiplier, targetVel = f->targetVel;
      const v3 delLinVel0 = v3scale(normal, invMassA), delLinVel1 = v3scale(normal, invMassB);
      const v3 v0 = v3add(v3mul(linVel0, normal), v3mul(angState0, raXn));
      const v3 v1 = v3add(v3mul(linVel1, normal), v3mul(angState1, rbXn));
      const v3 dv = v3sub(v0, v1);
      const float normalVel = (dv.x + dv.y) + dv.z;
      const float tmp1 = appliedForce - (bias - targetVel) * velMultiplier;
      const float totalImpulse = tmp1 - normalVel * velMultiplier;
      const int clamp = fabsf(totalImpulse) > maxFrictionImpulse;
      const float totalClamped = fminf_(maxDynFrictionImpulse, fmaxf_(negMaxDynFrictionImpulse, totalImpulse));
      const float newAppliedForce = clamp ? totalClamped : totalImpulse;
      broken = broken || clamp;
      const float deltaF = newAppliedForce - appliedForce;
      linVel0 = v3scaleadd(delLinVel0, deltaF, linVel0);
      linVel1 = v3negscalesub(delLinVel1, deltaF, linVel1);
      angState0 = v3scaleadd(raXn, deltaF * angDom0, angState0);
      angState1 = v3negscalesub(rbXn, deltaF * angDom1, angState1);
      f->appliedForce = newAppliedForce;
    }
    k->broken = broken;
  }
  b0->linVel = linVel0; b0->angState = angState0;
  if (k->body1 >= 0) { b1->linVel = linVel1; b1->angState = angState1; }
}

static inline void pxo_pgs_conclude(PxoPgsConstraint* k) {   /* concludeContact */
  for (int i = 0; i < k->numNormal; ++i) k->pts[i].biasedErr = k->pts[i].unbiasedErr;
  for (int i = 0; i < k->numFriction; ++i) k->fr[i].bias = 0.f;
}

/* integrateCore: pose from the motion velocity (state after the position iterations), velocity from the final state */
static inline void pxo_pgs_integrate(PxoPgsBodyData* d, PxoPgsBody* b, v3 motionLin, v3 motionAng, float dt, uint32_t lockFlags, v3* outMotionLin, v3* outMotionAng) {
  if (lockFlags) {   /* integrateCore :86-124 */
    const uint32_t l = lockFlags & 7u, a = (lockFlags >> 3) & 7u;
    motionLin = pxo_lock3(motionLin, l); b->linVel = pxo_lock3(b->linVel, l); d->linVel = pxo_lock3(d->linVel, l);
    motionAng = pxo_lock3(motionAng, a); b->angState = pxo_lock3(b->angState, a);
  }
  const v3 linearMotionVel = v3add(d->linVel, motionLin);
  d->body2World.p = v3add(d->body2World.p, v3scale(linearMotionVel, dt));
  const v3 angularMotionVel = v3add(d->angVel, m33mul(&d->sqrtInvInertia, motionAng));
  float w = v3lensq(angularMotionVel);
  if (w != 0.0f) {
    w = sqrtf(w);   /* (the 1e7 rad/s clamp of the reference is out of reach: maxAngularVelocity is 100) */
    const float v = dt * w * 0.5f;
    float s = sinf(v), q = cosf(v);
    s /= w;
    const v3 pqr = v3scale(angularMotionVel, s);
    const q4 quatVel = Q4(pqr.x, pqr.y, pqr.z, 0);
    q4 result = q4mul(quatVel, d->body2World.q);
    result.x += d->body2World.q.x * q; result.y += d->body2World.q.y * q; result.z += d->body2World.q.z * q; result.w += d->body2World.q.w * q;
    d->body2World.q = q4normalized(result);
  }
  *outMotionLin = linearMotionVel; *outMotionAng = angularMotionVel;   /* motionVelocityArray[i] after integrateCore (sleepCheck input) */
  d->linVel = v3add(d->linVel, b->linVel);
  d->angVel = v3add(d->angVel, m33mul(&d->sqrtInvInertia, b->angState));
}
#endif
