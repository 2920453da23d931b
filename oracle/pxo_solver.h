/* pxo_solver.h -- CPU restatement of the reference's TGS rigid-body contact solver (scalar path)
 * (TEST INFRASTRUCTURE).  Follows:
 *   unconstrained velocity  physx/source/lowleveldynamics/src/DyBodyCoreIntegrator.h:39-81
 *   solver body setup       DyTGSDynamics.cpp:154-243 (copyToSolverBodyDataStep), CmUtils.h:57-70
 *   friction correlation    DyFrictionCorrelation.cpp:56-330, DyContactPrepShared.h:52-131
 *   contact prep            DyTGSContactPrep.cpp:322-823 (constructContactConstraintStep, setupFinalizeSolverConstraints)
 *   solve                   DyTGSContactPrep.cpp:1492-1873 (solveDynamicContactsStep, solveContact)
 *   colouring               DyConstraintPartition.cpp:460-568 (first fit), :203-262 (static placement)
 *   iteration loop          DyTGSDynamics.cpp:2515-2793 (iterativeSolveIsland)
 *   integration             DyTGSDynamics.cpp:1403-1476 (integrateCoreStep), :1549-1580 (copyBackBodies)
 * Restrictions: rigid dynamic vs rigid dynamic/static contacts only, one contact patch per pair,
 * no kinematics/articulations/joints, compliant contacts (restitution<0) not restated. */
#ifndef PXO_SOLVER_H
#define PXO_SOLVER_H
#include "pxo_math.h"
#include "pxo_np.h"

typedef struct {            /* Dy::FrictionPatch (DyFrictionPatch.h) -- persistent per pair */
  v3 body0Normal, body1Normal;
  v3 body0Anchors[2], body1Anchors[2];
  q4 relativeQuat;
  int anchorCount, broken;
  float staticFriction, dynamicFriction, restitution;
} PxoFrictionPatch;

typedef struct {
  v3 linVel, angState;      /* PxTGSSolverBodyVel: linearVelocity, angularVelocity (sqrt-inertia space) */
  v3 deltaLinDt, deltaAngDt;
  m33 sqrtInvInertia;       /* PxTGSSolverBodyTxInertia */
  v3 body2WorldP; q4 deltaQ;
  float invMass, penBiasClamp, maxContactImpulse;
  v3 origLinVel, origAngVel;
  int hasConstraints;
  int isKinematic;          /* PxTGSSolverBodyVel::isKinematic: zero solver velocity, the body's velocity enters the rows' target velocities */
  uint32_t lockFlags;       /* PxRigidDynamicLockFlag bits: linear x,y,z = 1,2,4; angular x,y,z = 8,16,32 */
} PxoSolverBody;

typedef struct { v3 raXnI, rbXnI; float velMultiplier, separation, biasCoefficient, targetVelocity, recipResponse, maxImpulse, appliedForce; } PxoSPoint;
typedef struct { v3 normal; float error; v3 raXnI; float targetVel; v3 rbXnI; float velMultiplier; float appliedForce, frictionScale, biasScale; } PxoSFriction;

typedef struct {
  int body0, body1;         /* solver body indices; -1 = static world body */
  v3 normal; float invMass0, invMass1, angDom0, angDom1, maxPenBias, staticFriction, dynamicFriction;
  int numNormal, numFriction, broken;
  PxoSPoint pts[PXO_MAX_CONTACTS];
  PxoSFriction fr[4];
} PxoConstraint;

typedef struct {
  float dt, stepDt, invStepDt, invTotalDt, biasCoefficient, bounceThreshold, frictionOffsetThreshold, correlationDistance;
} PxoSolverParams;

static inline void pxo_transform_inertia(v3 invD, const m33* M, m33* out) { /* CmUtils.h:57-70; M(r,c) = column c, row r */
#define MM(r, c) ((c) == 0 ? ((r) == 0 ? M->c0.x : (r) == 1 ? M->c0.y : M->c0.z) : (c) == 1 ? ((r) == 0 ? M->c1.x : (r) == 1 ? M->c1.y : M->c1.z) : ((r) == 0 ? M->c2.x : (r) == 1 ? M->c2.y : M->c2.z))
  const float axx = invD.x * MM(0, 0), axy = invD.x * MM(1, 0), axz = invD.x * MM(2, 0);
  const float byx = invD.y * MM(0, 1), byy = invD.y * MM(1, 1), byz = invD.y * MM(2, 1);
  const float czx = invD.z * MM(0, 2), czy = invD.z * MM(1, 2), czz = invD.z * MM(2, 2);
  const float m00 = axx * MM(0, 0) + byx * MM(0, 1) + czx * MM(0, 2);
  const float m11 = axy * MM(1, 0) + byy * MM(1, 1) + czy * MM(1, 2);
  const float m22 = axz * MM(2, 0) + byz * MM(2, 1) + czz * MM(2, 2);
  const float m01 = axx * MM(1, 0) + byx * MM(1, 1) + czx * MM(1, 2);
  const float m02 = axx * MM(2, 0) + byx * MM(2, 1) + czx * MM(2, 2);
  const float m12 = axy * MM(2, 0) + byy * MM(2, 1) + czy * MM(2, 2);
#undef MM
  out->c0 = V3(m00, m01, m02); out->c1 = V3(m01, m11, m12); out->c2 = V3(m02, m12, m22);
}

/* DyBodyCoreIntegrator.h:39-81 */
static inline void pxo_unconstrained_velocity(v3 gravity, float dt, float linDamping, float angDamping, float maxLinVelSq, float maxAngVelSq, v3* lv, v3* av, int disableGravity) {
  v3 l = *lv, a = *av;
  const float oml = 1.0f - linDamping * dt, oma = 1.0f - angDamping * dt;
  if (!disableGravity) l = v3add(l, v3scale(v3scale(gravity, dt), 1.0f)); /* gravity*dt*accelScale, accelScale = 1; PxActorFlag::eDISABLE_GRAVITY skips the term */
  const float lm = oml >= 0.f ? oml : 0.f, am = oma >= 0.f ? oma : 0.f;
  l = v3scale(l, lm); a = v3scale(a, am);
  const float lsq = v3lensq(l); if (lsq > maxLinVelSq) l = v3scale(l, sqrtf(maxLinVelSq / lsq));
  const float asq = v3lensq(a); if (asq > maxAngVelSq) a = v3scale(a, sqrtf(maxAngVelSq / asq));
  *lv = l; *av = a;
}

/* PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES: DyTGSDynamics.cpp:177-193 = DyRigidBodyToSolverBody.cpp:53-70 (scalar PxVec3 / PxQuat arithmetic) */
static inline v3 pxo_gyroscopic(v3 av, v3 invInertia, q4 q, float dt) {
  const v3 localInertia = V3(invInertia.x == 0.f ? 0.f : 1.f / invInertia.x, invInertia.y == 0.f ? 0.f : 1.f / invInertia.y, invInertia.z == 0.f ? 0.f : 1.f / invInertia.z);
  const v3 localAngVel = q4rotinv(q, av);
  const v3 origMom = v3mul(localInertia, localAngVel);
  const v3 c = v3cross(localAngVel, origMom); const v3 torque = V3(-c.x, -c.y, -c.z);
  v3 newMom = v3add(origMom, v3scale(torque, dt));
  const float denom = sqrtf(newMom.x * newMom.x + newMom.y * newMom.y + newMom.z * newMom.z);
  const float ratio = denom > 0.f ? sqrtf(origMom.x * origMom.x + origMom.y * origMom.y + origMom.z * origMom.z) / denom : 0.f;
  newMom = v3scale(newMom, ratio);
  const v3 newDeltaAngVel = q4rot(q, v3sub(v3mul(invInertia, newMom), localAngVel));
  return v3add(av, newDeltaAngVel);
}

/* DyTGSDynamics.cpp:154-243 (gyroscopic forces: the caller applies pxo_gyroscopic to `av` first) */
static inline v3 pxo_lock3(v3 v, uint32_t bits) { if (bits & 1u) v.x = 0.f; if (bits & 2u) v.y = 0.f; if (bits & 4u) v.z = 0.f; return v; }
static inline void pxo_solver_body_init(PxoSolverBody* b, v3 lv, v3 av, float invMass, v3 invInertia, const xf* pose, float maxDepenVel, uint32_t lockFlags) {
  lv = pxo_lock3(lv, lockFlags & 7u); av = pxo_lock3(av, (lockFlags >> 3) & 7u);   /* DyTGSDynamics.cpp:195-222 */
  b->lockFlags = lockFlags;
  const m33 rot = am33fromq(pose->q); /* PxMat33Padded rotation(globalPose.q) */
  const v3 sqrtInvI = V3(invInertia.x == 0.f ? 0.f : sqrtf(invInertia.x), invInertia.y == 0.f ? 0.f : sqrtf(invInertia.y), invInertia.z == 0.f ? 0.f : sqrtf(invInertia.z));
  const v3 sqrtI = V3(sqrtInvI.x == 0.f ? 0.f : 1.0f / sqrtInvI.x, sqrtInvI.y == 0.f ? 0.f : 1.0f / sqrtInvI.y, sqrtInvI.z == 0.f ? 0.f : 1.0f / sqrtInvI.z);
  pxo_transform_inertia(sqrtInvI, &rot, &b->sqrtInvInertia);
  b->body2WorldP = pose->p; b->deltaQ = Q4(0, 0, 0, 1);
  m33 sqrtInertia; pxo_transform_inertia(sqrtI, &rot, &sqrtInertia);
  b->linVel = lv; b->angState = m33mul(&sqrtInertia, av);
  b->deltaLinDt = V3(0, 0, 0); b->deltaAngDt = V3(0, 0, 0);
  b->invMass = invMass; b->penBiasClamp = -maxDepenVel; b->maxContactImpulse = FLT_MAX;
  b->origLinVel = lv; b->origAngVel = av; b->isKinematic = 0;
}
/* copyToSolverBodyDataStepKinematic, DyTGSDynamics.cpp:245-274 */
static inline void pxo_kinematic_body_init(PxoSolverBody* b, v3 lv, v3 av, const xf* pose, float maxDepenVel) {
  memset(b, 0, sizeof(*b));
  b->body2WorldP = pose->p; b->deltaQ = Q4(0, 0, 0, 1);
  b->invMass = 0.f; b->penBiasClamp = -maxDepenVel; b->maxContactImpulse = FLT_MAX;
  b->origLinVel = lv; b->origAngVel = av; b->isKinematic = 1;
}

static inline void pxo_static_body_init(PxoSolverBody* b) {
  memset(b, 0, sizeof(*b)); b->deltaQ = Q4(0, 0, 0, 1); b->penBiasClamp = -FLT_MAX; b->maxContactImpulse = FLT_MAX;
}

/* DyTGSDynamics.cpp:1403-1476 */
static inline void pxo_integrate_core_step(PxoSolverBody* b, float dt) {
  if (b->lockFlags) { b->linVel = pxo_lock3(b->linVel, b->lockFlags & 7u); b->angState = pxo_lock3(b->angState, (b->lockFlags >> 3) & 7u); }   /* :1405-1422 (the angular lock acts on the sqrt-inertia-space state) */
  const v3 delta = v3scale(b->linVel, dt);
  const v3 unmolested = b->angState;
  const v3 angMotionVel = m33mul(&b->sqrtInvInertia, b->angState);
  const float w2 = v3lensq(angMotionVel);
  b->body2WorldP = v3add(b->body2WorldP, delta);
  if (w2 != 0.0f) {
    const float w = sqrtf(w2);
    const float v = dt * w * 0.5f;
    float s = sinf(v), q = cosf(v);
    s /= w;
    const v3 pqr = v3scale(angMotionVel, s);
    const q4 quatVel = Q4(pqr.x, pqr.y, pqr.z, 0);
    q4 result = q4mul(quatVel, b->deltaQ);
    result.x += b->deltaQ.x * q; result.y += b->deltaQ.y * q; result.z += b->deltaQ.z * q; result.w += b->deltaQ.w * q;
    b->deltaQ = q4normalized(result);
  }
  b->deltaAngDt = v3add(b->deltaAngDt, v3scale(unmolested, dt));
  b->deltaLinDt = v3add(b->deltaLinDt, delta);
}

/* Friction patch correlation for the single-contact-patch case.
 * DyContactPrepShared.h:73-131 (getFrictionPatches) + DyFrictionCorrelation.cpp:133-330 (correlatePatches, growPatches).
 * Returns the index (0/1) of the friction patch that owns the contacts; fp is updated in place to the
 * patch that persists to the next frame. contactID[2] receives the anchor->contact indices (0xffff = none). */
static inline void pxo_friction_correlate(PxoFrictionPatch* fp, int hadPatch, const PxoContacts* c, const xf* bodyFrame0, const xf* bodyFrame1,
                                          float sf, float df, float rest, float correlationDistance, float frictionOffsetThreshold, int contactID[2]) {
  const float SAME_NORMAL = 0.999f;
  int keepOld = 0;
  v3 oldWorldNormal = V3(0, 0, 0);
  if (hadPatch && !fp->broken && fp->anchorCount != 0) {
    const xf body1ToBody0 = xfinvmul(bodyFrame0, bodyFrame1);
    if (v3dot(fp->body0Normal, q4rot(body1ToBody0.q, fp->body1Normal)) > SAME_NORMAL) {
      int separated = 0;
      for (int a = 0; a < fp->anchorCount; ++a) {
        const v3 p1 = xftransform(&body1ToBody0, fp->body1Anchors[a]);
        if (!(fabsf(v3dot(v3sub(fp->body0Anchors[a], p1), fp->body0Normal)) < correlationDistance)) { separated = 1; break; }
      }
      if (!separated) { keepOld = 1; oldWorldNormal = q4rot(bodyFrame0->q, fp->body0Normal); }
    }
  }
  /* contact patch bounds (createContactPatches) */
  v3 bmin = c->point[0], bmax = c->point[0];
  for (int i = 1; i < c->count; ++i) { bmin = v3min(bmin, c->point[i]); bmax = v3max(bmax, c->point[i]); }
  const v3 patchNormal = c->normal;
  int correlated = keepOld && !((v3dot(patchNormal, oldWorldNormal) < SAME_NORMAL) || fp->restitution != rest || fp->staticFriction != sf || fp->dynamicFriction != df);
  contactID[0] = 0xffff; contactID[1] = 0xffff;
  if (!correlated) { /* initFrictionPatch */
    fp->body0Normal = q4rotinv(bodyFrame0->q, patchNormal);
    fp->body1Normal = q4rotinv(bodyFrame1->q, patchNormal);
    fp->relativeQuat = q4mul(q4conj(bodyFrame0->q), bodyFrame1->q);
    fp->anchorCount = 0; fp->broken = 0; fp->staticFriction = sf; fp->dynamicFriction = df; fp->restitution = rest;
  }
  /* growPatches */
  if (fp->anchorCount == 2) {
    const v3 dim = v3sub(bmax, bmin);
    const float diagSq = v3lensq(dim);
    const float anchorSq = v3lensq(v3sub(fp->body0Anchors[0], fp->body0Anchors[1]));
    if ((anchorSq * 4.f) >= diagSq) return; /* keep both anchors */
    fp->anchorCount = 0;
  }
  v3 worldAnchors[2]; int anchorCount = 0; float pointDistSq = 0.f;
  if (fp->anchorCount == 1) worldAnchors[anchorCount++] = xftransform(bodyFrame0, fp->body0Anchors[0]);
  const float eps = 1e-8f;
  for (int j = 0; j < c->count; ++j) {
    const v3 wp = c->point[j];
    if (c->sep[j] < frictionOffsetThreshold) {
      switch (anchorCount) {
        case 0: contactID[0] = j; worldAnchors[0] = wp; anchorCount++; break;
        case 1:
          pointDistSq = v3lensq(v3sub(wp, worldAnchors[0]));
          if (pointDistSq > eps) { contactID[1] = j; worldAnchors[1] = wp; anchorCount++; }
          break;
        default: {
          const float dist0 = v3lensq(v3sub(wp, worldAnchors[0])), dist1 = v3lensq(v3sub(wp, worldAnchors[1]));
          if (dist0 > dist1) { if (dist0 > pointDistSq) { contactID[1] = j; worldAnchors[1] = wp; pointDistSq = dist0; } }
          else if (dist1 > pointDistSq) { contactID[0] = j; worldAnchors[0] = wp; pointDistSq = dist1; }
        }
      }
    }
  }
  for (int j = fp->anchorCount; j < anchorCount; ++j) {
    fp->body0Anchors[j] = xftransforminv(bodyFrame0, worldAnchors[j]);
    fp->body1Anchors[j] = xftransforminv(bodyFrame1, worldAnchors[j]);
  }
  if (anchorCount == 0) { fp->body0Anchors[0] = V3(0, 0, 0); fp->body1Anchors[0] = V3(0, 0, 0); }
  fp->anchorCount = anchorCount;
}

/* DyTGSContactPrep.cpp:426-823 + :322-424 */
static inline void pxo_prep_contact(PxoConstraint* k, const PxoContacts* c, const PxoFrictionPatch* fp, const int contactID[2],
                                    const PxoSolverBody* b0, const PxoSolverBody* b1, const xf* bodyFrame0, const xf* bodyFrame1,
                                    float sf, float df, float restitution, float restDistance, const PxoSolverParams* P) {
  const float d0 = 1.f, d1 = 1.f, angD0 = 1.f, angD1 = 1.f;
  const float invMass0_dom0 = d0 * b0->invMass, invMass1_dom1 = (-d1) * b1->invMass;
  const float maxPenBias = fmaxf_(b0->penBiasClamp, b1->penBiasClamp);
  const v3 linVel0 = b0->origLinVel, linVel1 = b1->origLinVel, angVel0 = b0->origAngVel, angVel1 = b1->origAngVel;
  const float invDt = P->invStepDt, dt = P->stepDt, totalDt = P->dt, invTotalDt = P->invTotalDt;
  const float scale = fminf_(0.8f, P->biasCoefficient);
  const float invDtp8 = invDt * scale;
  const float frictionBiasScale = invDt * scale;
  (void)dt;
  k->invMass0 = invMass0_dom0; k->invMass1 = -invMass1_dom1;
  const v3 normal = c->normal;
  const float normalLenSq = alensq(normal);
  const float norVel0 = adot(linVel0, normal), norVel1 = adot(linVel1, normal);
  const float invMassNorLenSq0 = invMass0_dom0 * normalLenSq, invMassNorLenSq1 = invMass1_dom1 * normalLenSq;
  k->normal = normal; k->maxPenBias = maxPenBias; k->angDom0 = angD0; k->angDom1 = angD1;
  k->staticFriction = sf; k->dynamicFriction = df; k->broken = 0;
  k->numNormal = c->count;
  for (int j = 0; j < c->count; ++j) { /* constructContactConstraintStep */
    PxoSPoint* s = &k->pts[j];
    const v3 point = c->point[j]; const float separation = c->sep[j];
    const float cTargetVel = 0.f;
    const v3 ra = v3sub(point, bodyFrame0->p), rb = v3sub(point, bodyFrame1->p);
    v3 raXn = v3cross(ra, normal), rbXn = v3cross(rb, normal);
    const float angV0 = adot(raXn, angVel0), angV1 = adot(rbXn, angVel1);
    const float vrel1 = norVel0 + angV0, vrel2 = norVel1 + angV1;
    const float vrel = vrel1 - vrel2;
    /* solverOffsetSlop = 0: raXn/rbXn unchanged */
    const v3 raXnI = m33mul(&b0->sqrtInvInertia, raXn), rbXnI = m33mul(&b1->sqrtInvInertia, rbXn);
    const float i0 = adot(raXnI, raXnI) * angD0, i1 = adot(rbXnI, rbXnI) * angD1;
    const float resp0 = invMassNorLenSq0 + i0, resp1 = i1 - invMassNorLenSq1;
    const float unitResponse = resp0 + resp1;
    const float penetration = separation - restDistance;
    const int isSeparated = penetration > 0.f;
    const float penetrationInvDt = penetration * invTotalDt;
    const int isGreater2 = (restitution > 0.f) && (P->bounceThreshold > vrel) && ((-vrel) > penetrationInvDt);
    const float ratio = totalDt + (isGreater2 ? (penetration / vrel) : (-totalDt));
    const float recipResponse = (unitResponse > 0.f) ? (1.f / unitResponse) : 0.f;
    const float biasCoeff = -(isSeparated ? invDt : invDtp8);
    const float velMultiplier = recipResponse;
    float totalError = penetration;
    float targetVelocity = cTargetVel + (isGreater2 ? ((-vrel) * restitution) : 0.f);
    totalError = targetVelocity * ratio + totalError;
    if (b0->isKinematic) targetVelocity = targetVelocity - vrel1;   /* DyTGSContactPrep.cpp:406-409 */
    if (b1->isKinematic) targetVelocity = targetVelocity + vrel2;
    s->raXnI = raXnI; s->rbXnI = rbXnI; s->velMultiplier = velMultiplier; s->separation = totalError;
    s->biasCoefficient = biasCoeff; s->targetVelocity = targetVelocity; s->recipResponse = recipResponse;
    s->maxImpulse = FLT_MAX; s->appliedForce = 0.f;
  }
  const int haveFriction = fp->anchorCount != 0;
  k->numFriction = haveFriction ? fp->anchorCount * 2 : 0;
  if (haveFriction) {
    const v3 linVrel = v3sub(linVel0, linVel1);
    const v3 t0Fallback1 = V3(0.f, -normal.z, normal.y), t0Fallback2 = V3(-normal.y, normal.x, 0.f);
    const v3 t0Fallback = (0.70710678f > fabsf(normal.x)) ? t0Fallback1 : t0Fallback2;
    v3 t0 = v3sub(linVrel, v3scale(normal, adot(normal, linVrel)));
    t0 = (alensq(t0) > 0.0001f) ? t0 : t0Fallback;
    t0 = anormalize(t0);
    const v3 t1 = anormalize(v3cross(normal, t0));
    const v3 relTr = v3sub(bodyFrame0->p, bodyFrame1->p);
    const float norVelT[2][2] = {{adot(linVel0, t0), adot(linVel1, t0)}, {adot(linVel0, t1), adot(linVel1, t1)}};   /* norVel00 / 01 / 10 / 11, DyTGSContactPrep.cpp:670-673 */
    const float frictionScale = (fp->anchorCount == 2) ? 0.5f : 1.f;
    for (int j = 0; j < fp->anchorCount; ++j) {
      const v3 ra = aqrot(bodyFrame0->q, fp->body0Anchors[j]), rb = aqrot(bodyFrame1->q, fp->body1Anchors[j]);
      const v3 error = v3add(v3sub(ra, rb), relTr);
      (void)contactID; /* target velocity of the anchor's contact is zero here */
      for (int t = 0; t < 2; ++t) {
        const v3 tdir = t == 0 ? t0 : t1;
        PxoSFriction* f = &k->fr[j * 2 + t];
        const v3 raXn = v3cross(ra, tdir), rbXn = v3cross(rb, tdir);
        const v3 raXnI = m33mul(&b0->sqrtInvInertia, raXn), rbXnI = m33mul(&b1->sqrtInvInertia, rbXn);
        const float resp0 = invMassNorLenSq0 + adot(raXnI, raXnI) * angD0;
        const float resp1 = adot(rbXnI, rbXnI) * angD1 - invMassNorLenSq1;
        const float unitResponse = resp0 + resp1;
        const float velMultiplier = (unitResponse > 0.f) ? (scale / unitResponse) : 0.f;
        float targetVel = 0.f;   /* kinematic bodies: their velocity along the tangent becomes the row's target velocity (:724-727, :761-764) */
        if (b0->isKinematic) targetVel = targetVel - (norVelT[t][0] + adot(raXn, angVel0));
        if (b1->isKinematic) targetVel = targetVel + (norVelT[t][1] + adot(rbXn, angVel1));
        f->normal = tdir; f->error = adot(error, tdir); f->raXnI = raXnI; f->targetVel = targetVel; f->rbXnI = rbXnI;
        f->velMultiplier = velMultiplier; f->appliedForce = 0.f; f->frictionScale = frictionScale; f->biasScale = frictionBiasScale;
      }
    }
  }
}

/* DyTGSContactPrep.cpp:1492-1873 */
static inline void pxo_solve_contact(PxoConstraint* k, PxoSolverBody* b0, PxoSolverBody* b1, float minPen, float elapsedTime) {
  v3 linVel0 = b0->linVel, linVel1 = b1->linVel, angState0 = b0->angState, angState1 = b1->angState;
  const v3 angMotion0 = b0->deltaAngDt, angMotion1 = b1->deltaAngDt;
  const v3 relMotion = v3sub(b0->deltaLinDt, b1->deltaLinDt);
  const float invMassA = k->invMass0, invMassB = k->invMass1, angDom0 = k->angDom0, angDom1 = k->angDom1;
  const v3 n = k->normal;
  const float maxPenBias = k->maxPenBias;
  float accum = 0.f;
  {
    const v3 nim0 = v3scale(n, invMassA), nim1 = v3scale(n, invMassB);
    const float deltaV = adot(relMotion, n);
    for (int i = 0; i < k->numNormal; ++i) {
      PxoSPoint* c = &k->pts[i];
      const v3 raXnI = c->raXnI, rbXnI = c->rbXnI;
      const float deltaAng = adot(angMotion0, raXnI) - adot(angMotion1, rbXnI);
      const float targetVel = c->targetVelocity;
      const float deltaBias = (deltaV + deltaAng) - targetVel * elapsedTime;
      const float sep = fmaxf_(minPen, c->separation + deltaBias);
      const float bias = fminf_(-maxPenBias, c->biasCoefficient * sep);
      const v3 v0 = v3add(v3mul(linVel0, n), v3mul(angState0, raXnI));
      const v3 v1 = v3add(v3mul(linVel1, n), v3mul(angState1, rbXnI));
      const v3 dv = v3sub(v0, v1);
      const float normalVel = dv.x + dv.y + dv.z;
      const float biasNV = bias * c->recipResponse;
      const float lambda = biasNV - (normalVel - targetVel) * c->velMultiplier;
      const float appliedForce = c->appliedForce;
      const float _deltaF = fmaxf_(lambda, -appliedForce);
      const float _newForce = appliedForce + _deltaF;
      const float newForce = fminf_(_newForce, c->maxImpulse);
      const float deltaF = newForce - appliedForce;
      linVel0 = v3scaleadd(nim0, deltaF, linVel0);
      linVel1 = v3negscalesub(nim1, deltaF, linVel1);
      angState0 = v3scaleadd(raXnI, deltaF * angDom0, angState0);
      angState1 = v3negscalesub(rbXnI, deltaF * angDom1, angState1);
      c->appliedForce = newForce;
      accum = accum + newForce;
    }
  }
  if (k->numFriction) {
    const float maxFrictionImpulse = k->staticFriction * accum;
    const float maxDynFrictionImpulse = k->dynamicFriction * accum;
    int broken = 0;
    for (int i = 0; i < k->numFriction; i += 2) {
      PxoSFriction *f0 = &k->fr[i], *f1 = &k->fr[i + 1];
      const float frictionScale = f0->frictionScale, biasScale = f0->biasScale;
      const v3 normal0 = f0->normal, normal1 = f1->normal;
      const v3 raXnI0 = f0->raXnI, rbXnI0 = f0->rbXnI, raXnI1 = f1->raXnI, rbXnI1 = f1->rbXnI;
      const float appliedForce0 = f0->appliedForce, appliedForce1 = f1->appliedForce;
      const float targetVel0 = f0->targetVel, targetVel1 = f1->targetVel;
      float deltaV0 = (adot(raXnI0, angMotion0) - adot(rbXnI0, angMotion1)) + adot(normal0, relMotion);
      float deltaV1 = (adot(raXnI1, angMotion0) - adot(rbXnI1, angMotion1)) + adot(normal1, relMotion);
      deltaV0 = deltaV0 - targetVel0 * elapsedTime; deltaV1 = deltaV1 - targetVel1 * elapsedTime;
      const float error0 = f0->error + deltaV0, error1 = f1->error + deltaV1;
      const float bias0 = error0 * biasScale, bias1 = error1 * biasScale;
      const float velMultiplier0 = f0->velMultiplier, velMultiplier1 = f1->velMultiplier;
      const v3 delLinVel00 = v3scale(normal0, invMassA), delLinVel10 = v3scale(normal0, invMassB);
      const v3 delLinVel01 = v3scale(normal1, invMassA), delLinVel11 = v3scale(normal1, invMassB);
      const v3 v00 = v3add(v3mul(linVel0, normal0), v3mul(angState0, raXnI0));
      const v3 v10 = v3add(v3mul(linVel1, normal0), v3mul(angState1, rbXnI0));
      const v3 d0 = v3sub(v00, v10); const float normalVel0 = d0.x + d0.y + d0.z;
      const v3 v01 = v3add(v3mul(linVel0, normal1), v3mul(angState0, raXnI1));
      const v3 v11 = v3add(v3mul(linVel1, normal1), v3mul(angState1, rbXnI1));
      const v3 d1 = v3sub(v01, v11); const float normalVel1 = d1.x + d1.y + d1.z;
      const float tmp10 = appliedForce0 - (bias0 - targetVel0) * velMultiplier0;
      const float tmp11 = appliedForce1 - (bias1 - targetVel1) * velMultiplier1;
      const float totalImpulse0 = tmp10 - normalVel0 * velMultiplier0;
      const float totalImpulse1 = tmp11 - normalVel1 * velMultiplier1;
      const float totalImpulse = sqrtf(totalImpulse0 * totalImpulse0 + totalImpulse1 * totalImpulse1);
#ifdef PXO_BLOCK_FRICTION /* DyTGSContactPrepBlock.cpp:2538-2650: the 4-wide path's clamp (1e-5 slack, sticky broken flag) */
      const int clamp = totalImpulse > ((frictionScale * k->staticFriction) * accum + 1e-5f);
      broken = broken || clamp;
      const float totalClamped = broken ? fminf_((frictionScale * k->dynamicFriction) * accum, totalImpulse) : totalImpulse;
      const float ratio = (totalImpulse > 0.f) ? (totalClamped / totalImpulse) : 0.f;
      const float newAppliedForce0 = totalImpulse0 * ratio, newAppliedForce1 = totalImpulse1 * ratio;
#else
      const int clamp = totalImpulse > (frictionScale * maxFrictionImpulse);
      const float totalClamped = clamp ? fminf_(frictionScale * maxDynFrictionImpulse, totalImpulse) : totalImpulse;
      const float ratio = (totalImpulse > 0.f) ? (totalClamped / totalImpulse) : 0.f;
      const float newAppliedForce0 = totalImpulse0 * ratio, newAppliedForce1 = totalImpulse1 * ratio;
      broken = broken || clamp;
#endif
      const float deltaF0 = newAppliedForce0 - appliedForce0, deltaF1 = newAppliedForce1 - appliedForce1;
      linVel0 = v3scaleadd(delLinVel00, deltaF0, v3scaleadd(delLinVel01, deltaF1, linVel0));
      linVel1 = v3negscalesub(delLinVel10, deltaF0, v3negscalesub(delLinVel11, deltaF1, linVel1));
      angState0 = v3scaleadd(raXnI0, deltaF0 * angDom0, v3scaleadd(raXnI1, deltaF1 * angDom0, angState0));
      angState1 = v3negscalesub(rbXnI0, deltaF0 * angDom1, v3negscalesub(rbXnI1, deltaF1 * angDom1, angState1));
      f0->appliedForce = newAppliedForce0; f1->appliedForce = newAppliedForce1;
    }
    k->broken = broken;
  }
  b0->linVel = linVel0; b0->angState = angState0;
  if (k->body1 >= 0) { b1->linVel = linVel1; b1->angState = angState1; }
}

/* First-fit colouring in input order: DyConstraintPartition.cpp:460-568 for dynamic-dynamic constraints;
 * static constraints go to partition maxNormalProgress(body)+k (:203-262).  colour[] receives the
 * partition index of every constraint; returns the number of partitions.  bodyMask/bodyMaxDyn/bodyStatic
 * are scratch arrays of nBodies entries.  (64 colours; overflow returns -1.) */
static inline int pxo_colour(const int* body0, const int* body1, int n, int nBodies, int* colour, uint64_t* bodyMask, int* bodyMaxDyn, int* bodyStatic) {
  memset(bodyMask, 0, sizeof(uint64_t) * nBodies); memset(bodyMaxDyn, 0, sizeof(int) * nBodies); memset(bodyStatic, 0, sizeof(int) * nBodies);
  int nPart = 0;
  for (int i = 0; i < n; ++i) {
    const int a = body0[i], b = body1[i];
    if (a >= 0 && b >= 0) {
      const uint64_t comb = ~bodyMask[a] & ~bodyMask[b];   /* 64 colours = two of the reference's 32-colour rounds (:520-552): overflow constraints see only each other's colours */
      if (comb == 0) return -1;
      int p = 0; while (!((comb >> p) & 1u)) p++;
      bodyMask[a] |= 1ull << p; bodyMask[b] |= 1ull << p;
      if (p + 1 > bodyMaxDyn[a]) bodyMaxDyn[a] = p + 1;
      if (p + 1 > bodyMaxDyn[b]) bodyMaxDyn[b] = p + 1;
      colour[i] = p;
    } else colour[i] = -1;
  }
  for (int i = 0; i < n; ++i) {
    const int a = body0[i], b = body1[i];
    if (!(a >= 0 && b >= 0)) { const int d = a >= 0 ? a : b; colour[i] = bodyMaxDyn[d] + bodyStatic[d]++; }
    if (colour[i] + 1 > nPart) nPart = colour[i] + 1;
  }
  return nPart;
}
/* PXB_FLAG_RELAXED_PARTITIONING (include/physx_b200.h): the same Jones-Plassmann rounds as k_colour_partition's relaxed branch
 * (physx_b200/csrc/pxb_engine.cu), run sequentially.  Not a reference algorithm: it exists so the relaxed GPU mode stays testable
 * bit for bit.  Static contacts are placed exactly as in pxo_colour. */
static inline int pxo_colour_relaxed(const int* body0, const int* body1, int n, int nBodies, int* colour, uint64_t* bodyMask, int* bodyMaxDyn, int* bodyStatic) {
  memset(bodyMask, 0, sizeof(uint64_t) * nBodies); memset(bodyMaxDyn, 0, sizeof(int) * nBodies); memset(bodyStatic, 0, sizeof(int) * nBodies);
  uint64_t* best = (uint64_t*)calloc(nBodies ? nBodies : 1, sizeof(uint64_t));
  int remaining = 0;
  for (int i = 0; i < n; ++i) { colour[i] = -1; if (body0[i] >= 0 && body1[i] >= 0) remaining++; }
  for (uint32_t round = 1; remaining > 0; ++round) {
    for (int i = 0; i < n; ++i) {
      if (colour[i] >= 0 || body0[i] < 0 || body1[i] < 0) continue;
      const uint64_t bid = ((uint64_t)round << 32) | ((uint32_t)i * 2654435761u + 1u);
      if (best[body0[i]] < bid) best[body0[i]] = bid;
      if (best[body1[i]] < bid) best[body1[i]] = bid;
    }
    for (int i = 0; i < n; ++i) {
      if (colour[i] >= 0 || body0[i] < 0 || body1[i] < 0) continue;
      const uint64_t bid = ((uint64_t)round << 32) | ((uint32_t)i * 2654435761u + 1u);
      const int a = body0[i], b = body1[i];
      if (best[a] != bid || best[b] != bid) continue;
      const uint64_t comb = ~bodyMask[a] & ~bodyMask[b];
      if (comb == 0) { free(best); return -1; }
      int p = 0; while (!((comb >> p) & 1u)) p++;
      bodyMask[a] |= 1ull << p; bodyMask[b] |= 1ull << p;
      colour[i] = p; remaining--;
    }
  }
  free(best);
  int nPart = 0;
  for (int i = 0; i < n; ++i) {
    const int a = body0[i], b = body1[i];
    if (!(a >= 0 && b >= 0)) {
      const int d = a >= 0 ? a : b; const uint64_t m = bodyMask[d]; int mx = 0; while (mx < 64 && (m >> mx)) mx++;
      colour[i] = mx + bodyStatic[d]++;
    }
    if (colour[i] + 1 > nPart) nPart = colour[i] + 1;
  }
  (void)bodyMaxDyn;
  return nPart;
}
#endif
