// pxb_pgs.cuh -- PGS variant of the contact solver (PxSolverType::ePGS), device functions + the device-wide kernels.
// Included from pxb_env.cuh after RegRows (both solver types share the 25-float4 row record and its memory image).
//
// Reference CPU path matched (scalar variant; the oracle restatement is oracle/pxo_pgs.h):
//   prep        DyContactPrep.cpp:60-365 setupFinalizeSolverConstraints, DyContactPrepShared.h:298-391 constructContactConstraint
//   solve       DySolverConstraintsShared.h:49-109, DySolverConstraints.cpp:221-371 (solveContact; _BStatic = same arithmetic, zero terms dropped)
//   conclude    DySolverConstraints.cpp:508-551
//   loop        DySolverControl.cpp:163-405 solveV_Blocks: friction only in the last three position iterations, the last position
//               iteration concludes (biased -> unbiased error), motion velocities saved, >= 1 velocity iteration
//   integrate   DyBodyCoreIntegrator.h:83-185 integrateCore
// Reference GPU kernels replaced: solveContactParallel / concludeBlocks / writebackBlocks (gpusolver/src/CUDA/solverMultiBlock.cu,
// solverBlock.cuh) and integrateCoreParallelLaunch (integration.cu).
// PGS solves for velocity DELTAS: solver bodies start at zero, the pre-solver velocity is folded into the rows' target velocity.
// RegRows field use for PGS: pa = (raXn, velMultiplier), pb = (rbXn, biasedErr), pc0 = unbiasedErr[4], pc1 = friction targetVel[4],
// t0/t1 = friction directions, fa = (raXn, velMultiplier), fb = (rbXn, bias), h1.zw = friction coefficients x anchor scale.
#pragma once

__device__ __forceinline__ void prep_constraint_pgs(RegRows& r, uint32_t i, uint32_t b0, uint32_t b1, const PrepBodies& B, const float4* __restrict__ cHdr,
                                                    const float4* __restrict__ cPts, float4* __restrict__ frec, const SolverParams& P, const bool noFriction = false) {
  Contacts con; const float4 h = cHdr[i]; con.normal = V3(h.x, h.y, h.z); con.count = __float_as_int(h.w);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float4 p = cPts[(size_t)i * 4 + j]; con.point[j] = V3(p.x, p.y, p.z); con.sep[j] = p.w; }
  const xf& f0 = B.f0; const xf& f1 = B.f1;
  FrictionPatch fp; friction_load(fp, frec);
  friction_correlate(fp, con, f0, f1, P.staticFriction, P.dynamicFriction, P.restitution, P.correlationDistance, P.frictionOffsetThreshold + P.restDistance);
  if (noFriction) fp.anchorCount = 0;   // PxMaterialFlag::eDISABLE_FRICTION
  friction_store(fp, frec);
  const float maxPenBias = fmax_(B.pen0, B.pen1);
  const v3 linVel0 = B.linVel0, linVel1 = B.linVel1, angVel0 = B.angVel0, angVel1 = B.angVel1;
  const m33& sI0 = B.sI0; const m33& sI1 = B.sI1;
  const float invMass0_dom0 = 1.f * B.invMass0, invMass1_dom1 = (-1.f) * B.invMass1;
  const float invDt = P.invTotalDt, invDtp8 = invDt * 0.8f;
  const v3 normal = con.normal;
  const float normalLenSq = adot(normal, normal);
  const v3 nv = vmul(normal, linVel0) - vmul(normal, linVel1);
  const float norVel = (nv.x + nv.y) + nv.z;
  const float imn0 = invMass0_dom0 * normalLenSq, imn1 = invMass1_dom1 * normalLenSq;
  const float frictionCoefficient = (fp.anchorCount == 2) ? 0.5f : 1.f;
  const bool haveFriction = fp.anchorCount != 0;
  const uint32_t numFriction = haveFriction ? (uint32_t)fp.anchorCount * 2u : 0u;
  r.h0 = F4(normal, 0.f);
  r.h1 = make_float4(invMass0_dom0, -invMass1_dom1, P.staticFriction * frictionCoefficient, P.dynamicFriction * frictionCoefficient);
  r.h2 = make_uint4(b0, b1, (uint32_t)con.count | (numFriction << 8), i);
  r.pc0 = r.pc1 = r.ap = r.fap = make_float4(0, 0, 0, 0); r.t0 = r.t1 = make_float4(0, 0, 0, 0); r.broken = 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.pa[j] = r.pb[j] = make_float4(0, 0, 0, 0);
    if (j < con.count) {   // constructContactConstraint (solverOffsetSlop 0, no speculative CCD, zero contact target velocity)
      const v3 ra = con.point[j] - f0.p, rb = con.point[j] - f1.p;
      const v3 raXn = cross(ra, normal), rbXn = cross(rb, normal);
      const float vRelAng = adot(raXn, angVel0) - adot(rbXn, angVel1);
      const float vrel = norVel + vRelAng;
      const v3 raXnI = mmul(sI0, raXn), rbXnI = mmul(sI1, rbXn);
      const float resp0 = imn0 + adot(raXnI, raXnI) * 1.f, resp1 = adot(rbXnI, rbXnI) * 1.f - imn1;
      const float unitResponse = resp0 + resp1;
      const float penetration = con.sep[j] - P.restDistance;
      const float penetrationInvDt = penetration * invDt;
      const bool isSeparated = penetration >= 0.f;
      const bool isGreater2 = (P.restitution > 0.f) && (P.bounceThreshold > vrel) && ((-vrel) > penetrationInvDt);
      float targetVelocity = 0.f + (isGreater2 ? ((-vrel) * P.restitution) : 0.f);
      targetVelocity = targetVelocity - vrel;
      const float velMultiplier = (unitResponse > 0.f) ? (1.0f / unitResponse) : 0.f;
      const float penetrationInvDtScaled = isSeparated ? penetrationInvDt : (penetration * invDtp8);
      float scaledBias = velMultiplier * fmax_(maxPenBias, penetrationInvDtScaled);
      if (isGreater2) scaledBias = 0.f;
      const float biasedErr = targetVelocity * velMultiplier + (-scaledBias);
      const float unbiasedErr = targetVelocity * velMultiplier + (isGreater2 ? 0.f : (-fmax_(scaledBias, 0.f)));
      r.pa[j] = F4(raXnI, velMultiplier); r.pb[j] = F4(rbXnI, biasedErr); f4set(r.pc0, j, unbiasedErr);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) r.fa[j] = r.fb[j] = make_float4(0, 0, 0, 0);
  if (haveFriction) {
    const v3 linVrel = linVel0 - linVel1;
    const v3 fb1 = V3(0.f, -normal.z, normal.y), fb2 = V3(-normal.y, normal.x, 0.f);
    const v3 t0Fallback = (0.70710678f > fabsf(normal.x)) ? fb1 : fb2;
    v3 t0 = linVrel - normal * adot(normal, linVrel);
    t0 = (adot(t0, t0) > 0.0001f) ? t0 : t0Fallback;
    t0 = anormalize(t0);
    const v3 t1 = cross(normal, t0);   // not normalised on the PGS path
    r.t0 = F4(t0, 0.f); r.t1 = F4(t1, 0.f);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j < fp.anchorCount) {
        const v3 ra = aqrot(f0.q, fp.body0Anchors[j]), rb = aqrot(f1.q, fp.body1Anchors[j]);
        const v3 error = (ra + f0.p) - (rb + f1.p);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const v3 tdir = t == 0 ? t0 : t1;
          const v3 raXn = cross(ra, tdir), rbXn = cross(rb, tdir);
          const v3 raXnI = mmul(sI0, raXn), rbXnI = mmul(sI1, rbXn);
          const float resp0 = invMass0_dom0 + 1.f * adot(raXnI, raXnI), resp1 = 1.f * adot(rbXnI, rbXnI) - invMass1_dom1;
          const float resp = resp0 + resp1;
          const float velMultiplier = (resp > 0.f) ? (0.8f / resp) : 0.f;
          const float vrel1 = adot(tdir, linVel0) + adot(raXn, angVel0), vrel2 = adot(tdir, linVel1) + adot(rbXn, angVel1);
          const float targetVel = 0.f - (vrel1 - vrel2);
          r.fa[j * 2 + t] = F4(raXnI, velMultiplier); r.fb[j * 2 + t] = F4(rbXnI, adot(tdir, error) * invDt); f4set(r.pc1, j * 2 + t, targetVel);
        }
      }
    }
  }
}

template <bool FR>
__device__ __forceinline__ void solve_constraint_pgs(RegRows& r, const bool doFriction, float4* bLin, float4* bAng, const FrView fr = FrView()) {
  const uint32_t b0 = r.h2.x, b1 = r.h2.y;
  const int numNormal = (int)(r.h2.z & 0xff), numFriction = (int)((r.h2.z >> 8) & 0xff);
  const v3 n = V3(r.h0.x, r.h0.y, r.h0.z);
  const float invMassA = r.h1.x, invMassB = r.h1.y;
  v3 linVel0 = V3(bLin[b0]), angState0 = V3(bAng[b0]);
  v3 linVel1 = V3(0, 0, 0), angState1 = V3(0, 0, 0);
  if (b1 != NONE32) { linVel1 = V3(bLin[b1]); angState1 = V3(bAng[b1]); }
  float accum = 0.f;
  {
    const v3 del0 = n * invMassA, del1 = n * invMassB;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < numNormal) {
        const float4 A = r.pa[j], B = r.pb[j];
        const v3 raXn = V3(A.x, A.y, A.z), rbXn = V3(B.x, B.y, B.z);
        const float applied = f4get(r.ap, j), velMultiplier = A.w;
        const v3 dv = (vmul(linVel0, n) + vmul(angState0, raXn)) - (vmul(linVel1, n) + vmul(angState1, rbXn));
        const float normalVel = (dv.x + dv.y) + dv.z;
        const float dF_ = fmax_(B.w - normalVel * velMultiplier, -applied);
        const float newForce = fmin_(1.0f * applied + dF_, FLT_MAX);
        const float deltaF = newForce - applied;
        linVel0 = scaleadd(del0, deltaF, linVel0); linVel1 = negscalesub(del1, deltaF, linVel1);
        angState0 = scaleadd(raXn, deltaF * 1.f, angState0); angState1 = negscalesub(rbXn, deltaF * 1.f, angState1);
        f4set(r.ap, j, newForce);
        accum = accum + newForce;
      }
    }
  }
  if (doFriction && numFriction) {
    const float maxFrictionImpulse = r.h1.z * accum, maxDynFrictionImpulse = r.h1.w * accum;
    const float negMaxDyn = -maxDynFrictionImpulse;
    const float4 T0 = FR ? fr.p[0] : r.t0, T1 = FR ? fr.p[fr.stride] : r.t1, TV = FR ? fr.p[11 * fr.stride] : r.pc1; float4 fap = FR ? fr.p[10 * fr.stride] : r.fap;
    bool broken = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < numFriction) {
        const float4 T = (j & 1) ? T1 : T0;
        const v3 normal = V3(T.x, T.y, T.z);
        const float4 A = FR ? fr.p[(2 + j) * fr.stride] : r.fa[j], B = FR ? fr.p[(6 + j) * fr.stride] : r.fb[j];
        const v3 raXn = V3(A.x, A.y, A.z), rbXn = V3(B.x, B.y, B.z);
        const float applied = f4get(fap, j), bias = B.w, velMultiplier = A.w, targetVel = f4get(TV, j);
        const v3 del0 = normal * invMassA, del1 = normal * invMassB;
        const v3 dv = (vmul(linVel0, normal) + vmul(angState0, raXn)) - (vmul(linVel1, normal) + vmul(angState1, rbXn));
        const float normalVel = (dv.x + dv.y) + dv.z;
        const float tmp1 = applied - (bias - targetVel) * velMultiplier;
        const float totalImpulse = tmp1 - normalVel * velMultiplier;
        const bool clamp = fabsf(totalImpulse) > maxFrictionImpulse;
        const float totalClamped = fmin_(maxDynFrictionImpulse, fmax_(negMaxDyn, totalImpulse));
        const float newApplied = clamp ? totalClamped : totalImpulse;
        broken = broken || clamp;
        const float deltaF = newApplied - applied;
        linVel0 = scaleadd(del0, deltaF, linVel0); linVel1 = negscalesub(del1, deltaF, linVel1);
        angState0 = scaleadd(raXn, deltaF * 1.f, angState0); angState1 = negscalesub(rbXn, deltaF * 1.f, angState1);
        f4set(fap, j, newApplied);
      }
    }
    if (FR) fr.p[10 * fr.stride] = fap; else r.fap = fap;
    r.broken = broken ? 1u : 0u;
  }
  bLin[b0] = F4(linVel0, 0.f); bAng[b0] = F4(angState0, 0.f);
  if (b1 != NONE32) { bLin[b1] = F4(linVel1, 0.f); bAng[b1] = F4(angState1, 0.f); }
}

// concludeContact: biased error -> unbiased error, friction bias -> 0
template <bool FR>
__device__ __forceinline__ void conclude_constraint_pgs(RegRows& r, const FrView fr = FrView()) {
#pragma unroll
  for (int j = 0; j < 4; ++j) { r.pb[j].w = f4get(r.pc0, j); if (FR) fr.p[(6 + j) * fr.stride].w = 0.f; else r.fb[j].w = 0.f; }
}

// integrateCore: pose from the motion velocity (deltas after the position iterations), velocity from the final deltas
// (lock flags: DyBodyCoreIntegrator.h:86-124); outMotionLin / outMotionAng = motionVelocityArray after integrateCore (sleepCheck input)
__device__ __forceinline__ void integrate_core_pgs(v3& p, q4& q, v3& linVel, v3& angVel, const m33& sqrtInvInertia, v3 motionLin, v3 motionAng, v3 deltaLin, v3 deltaAng, float dt,
                                                   uint32_t lock, v3& outMotionLin, v3& outMotionAng) {
  if (lock) {
    const uint32_t l = lock & 7u, a = (lock >> 3) & 7u;
    motionLin = lock3(motionLin, l); deltaLin = lock3(deltaLin, l); linVel = lock3(linVel, l);
    motionAng = lock3(motionAng, a); deltaAng = lock3(deltaAng, a);
  }
  const v3 linearMotionVel = linVel + motionLin;
  p = p + linearMotionVel * dt;
  const v3 angularMotionVel = angVel + mmul(sqrtInvInertia, motionAng);
  float w = lensq(angularMotionVel);
  if (w != 0.0f) {
    w = sqrtf(w);
    const float v = dt * w * 0.5f;
    float s, c; sincosf(v, &s, &c);
    s /= w;
    const v3 pqr = angularMotionVel * s;
    q4 res = qmul(Q4(pqr.x, pqr.y, pqr.z, 0.f), q);
    res.x += q.x * c; res.y += q.y * c; res.z += q.z * c; res.w += q.w * c;
    q = qnormalized(res);
  }
  outMotionLin = linearMotionVel; outMotionAng = angularMotionVel;
  linVel = linVel + deltaLin;
  angVel = angVel + mmul(sqrtInvInertia, deltaAng);
}

#ifdef PXB_SOLVE_KERNELS   // compiled by pxb_solve.cu only
// ---------------------------------------------------------------------------------------------
// device-wide solver kernels, both solver types (scenes without environment ids); rows live in the RegRows memory image (25 x cap float4)
template <bool PGS>
__global__ void __launch_bounds__(128) k_prep_rows(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ ordered, const uint32_t* __restrict__ conPair, const uint32_t* __restrict__ pairSlots,
                       const uint2* __restrict__ pairBodies, const uint32_t* __restrict__ geomFlags, const float4* __restrict__ cHdr, const float4* __restrict__ cPts,
                       const float4* __restrict__ pos, const float4* __restrict__ quat, const float4* __restrict__ linVel, const float4* __restrict__ sbOrigAng,
                       const float4* __restrict__ invInertia, const float4* __restrict__ sbIA, const float4* __restrict__ sbIB, float4* __restrict__ frictions, SolverParams P, Rows R, const MaterialArgs M,
                       const float4* __restrict__ angVel, float4* __restrict__ kinFtv) {   // kinFtv: non-null in scenes with kinematic bodies (TGS only)
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= counters[C_NCON]) return;
  const uint32_t c = ordered[k]; const uint32_t i = conPair[c];
  const uint2 bb = pairBodies[i]; const uint32_t b0 = bb.x, b1 = bb.y;
  const uint32_t gf1 = geomFlags[b1];
  const bool dyn1 = gf_dynamic(gf1);
  const bool kin1 = (gf1 & 0x800u) != 0;   // kinematic body B: a static solver body whose velocity enters the rows (TGS: as target velocity, copyToSolverBodyDataStepKinematic DyTGSDynamics.cpp:245-274; PGS: through the
                                           // pre-solver velocities the rows are built from, KinematicCopyTask DyDynamics.cpp:1855-1870)
  RegRows r;
  if (__float_as_int(cHdr[i].w) == 0) {  // empty constraint kept only for the colouring (see k_flag_ordered)
    r.h0 = r.h1 = make_float4(0, 0, 0, 0); r.h2 = make_uint4(b0, dyn1 ? b1 : NONE32, 0u, i); r.pc0 = r.pc1 = r.ap = r.t0 = r.t1 = r.fap = make_float4(0, 0, 0, 0); r.broken = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) r.pa[j] = r.pb[j] = r.fa[j] = r.fb[j] = make_float4(0, 0, 0, 0);
    rows_store(R, k, r);
    return;
  }
  PrepBodies B;
  { const float4 p = pos[b0]; B.f0.p = V3(p.x, p.y, p.z); B.f0.q = Q4(quat[b0]); const float4 q = pos[b1]; B.f1.p = V3(q.x, q.y, q.z); B.f1.q = Q4(quat[b1]); }
  B.invMass0 = pos[b0].w; B.invMass1 = dyn1 ? pos[b1].w : 0.f;
  B.pen0 = -invInertia[b0].w; B.pen1 = dyn1 ? -invInertia[b1].w : -FLT_MAX;
  B.linVel0 = V3(linVel[b0]); B.linVel1 = dyn1 ? V3(linVel[b1]) : V3(0, 0, 0);
  B.angVel0 = V3(sbOrigAng[b0]); B.angVel1 = dyn1 ? V3(sbOrigAng[b1]) : V3(0, 0, 0);
  B.sI0 = load_sym(sbIA[b0], sbIB[b0]);
  if (dyn1) B.sI1 = load_sym(sbIA[b1], sbIB[b1]); else { B.sI1.c0 = B.sI1.c1 = B.sI1.c2 = V3(0, 0, 0); }
  if (kin1) { B.pen1 = -invInertia[b1].w; B.linVel1 = V3(linVel[b1]); B.angVel1 = V3(angVel[b1]); }
  bool noFriction = false;
  if (M.matTab) noFriction = pair_material(M, b0, b1, P);   // material table: this pair's combined coefficients (P is this thread's copy)
  if (M.shapeOff) P.restDistance = M.shapeOff[b0].y + M.shapeOff[b1].y;   // per-shape rest offsets: the pair's rest distance is their sum (PxcNpWorkUnit::restDistance)
  if (PGS) prep_constraint_pgs(r, i, b0, dyn1 ? b1 : NONE32, B, cHdr, cPts, frictions + (size_t)pairSlots[i] * PXB_FRICTION_F4, P, noFriction);
  else prep_constraint_regs(r, i, b0, dyn1 ? b1 : NONE32, B, cHdr, cPts, frictions + (size_t)pairSlots[i] * PXB_FRICTION_F4, P, noFriction, kin1, (kin1 && kinFtv) ? kinFtv + i : nullptr);
  rows_store(R, k, r);
}

// a15/a16: the whole TGS iteration loop in ONE cooperative launch (iterativeSolveIsland, DyTGSDynamics.cpp:2515-2793):
// position iterations = {solve every partition in order; integrate the sub-step}, then velocity iterations.
// Replaces solveBlockUnified x partitions x iterations + propagateAverageSolverBodyVelocityTGS (~65 launches in the reference).
template <bool KIN>
__global__ void __launch_bounds__(256, PXB_SOLVE_CTAS_PER_SM) k_solve_tgs(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ partStart, uint32_t posIters, uint32_t velIters, float stepDt, Rows R,
                            float4* __restrict__ sbLin, float4* __restrict__ sbAng, float4* __restrict__ sbDLin, float4* __restrict__ sbDAng, const float4* __restrict__ sbIA, const float4* __restrict__ sbIB,
                            float4* __restrict__ sbP, float4* __restrict__ sbQ, const uint32_t* __restrict__ bodyHasCon, uint32_t nDyn, const uint32_t* __restrict__ dynActor, const float4* __restrict__ kinFtv) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t nPart = counters[C_NPART];
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  if (nPart == 0) return;
  float elapsed = 0.f;
  for (uint32_t it = 0; it < posIters + velIters; ++it) {
    const bool vel = it >= posIters;
    const float minPen = vel ? 0.f : -FLT_MAX;
    for (uint32_t p = 0; p < nPart; ++p) {
      const uint32_t b = partStart[p], e = partStart[p + 1];
      for (uint32_t k = b + gtid; k < e; k += gsize) {
        RegRows r; rows_load(R, k, r);
        if ((r.h2.z & 0xff) == 0) continue;   // empty constraint kept only for the colouring
        solve_constraint_regs<false, KIN>(r, minPen, elapsed, sbLin, sbAng, sbDLin, sbDAng, FrView(), kinFtv);
        rows_store_state(R, k, r);
      }
      grid.sync();
    }
    if (!vel) {
      for (uint32_t d = gtid; d < nDyn; d += gsize) {
        const uint32_t a = dynActor[d];
        if (!bodyHasCon[a]) continue;
        v3 p = V3(sbP[a]); q4 dq = Q4(sbQ[a]); v3 dl = V3(sbDLin[a]), da = V3(sbDAng[a]);
        const float4 ib = sbIB[a]; const uint32_t lock = __float_as_uint(ib.z);
        v3 lv = V3(sbLin[a]), as = V3(sbAng[a]);
        integrate_core_step(lv, as, load_sym(sbIA[a], ib), stepDt, p, dq, dl, da, lock);
        if (lock) { sbLin[a] = F4(lv, 0.f); sbAng[a] = F4(as, 0.f); }
        sbP[a] = F4(p, 0.f); sbQ[a] = F4(dq); sbDLin[a] = F4(dl, 0.f); sbDAng[a] = F4(da, 0.f);
      }
      elapsed += stepDt;
      grid.sync();
    }
  }
}

// the whole PGS iteration loop in ONE cooperative launch (solveV_Blocks, DySolverControl.cpp:163-405)
__global__ void __launch_bounds__(256, PXB_SOLVE_CTAS_PER_SM) k_solve_pgs(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ partStart, uint32_t posIters, uint32_t velItersIn, Rows R,
                            float4* __restrict__ sbLin, float4* __restrict__ sbAng, float4* __restrict__ sbDLin, float4* __restrict__ sbDAng, uint32_t nDyn, const uint32_t* __restrict__ dynActor) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t nPart = counters[C_NPART];
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  for (uint32_t it = posIters; it > 0; --it) {
    const bool doFriction = it <= 3;
    for (uint32_t p = 0; p < nPart; ++p) {
      const uint32_t b = partStart[p], e = partStart[p + 1];
      for (uint32_t k = b + gtid; k < e; k += gsize) {
        RegRows r; rows_load(R, k, r);
        if ((r.h2.z & 0xff) == 0) continue;
        solve_constraint_pgs<false>(r, doFriction, sbLin, sbAng);
        if (it == 1) { conclude_constraint_pgs<false>(r); rows_store(R, k, r); } else rows_store_state(R, k, r);
      }
      grid.sync();
    }
  }
  for (uint32_t d = gtid; d < nDyn; d += gsize) { const uint32_t a = dynActor[d]; sbDLin[a] = sbLin[a]; sbDAng[a] = sbAng[a]; }   // saveMotionVelocities
  grid.sync();
  const uint32_t velIters = velItersIn ? velItersIn : 1u;
  for (uint32_t it = 0; it < velIters; ++it)
    for (uint32_t p = 0; p < nPart; ++p) {
      const uint32_t b = partStart[p], e = partStart[p + 1];
      for (uint32_t k = b + gtid; k < e; k += gsize) {
        RegRows r; rows_load(R, k, r);
        if ((r.h2.z & 0xff) == 0) continue;
        solve_constraint_pgs<false>(r, true, sbLin, sbAng);
        rows_store_state(R, k, r);
      }
      grid.sync();
    }
}

// a17: writeBackContact (DyTGSContactPrep.cpp:1875-1937 / DySolverConstraints.cpp:553-640): applied forces -> contact force stream, broken flag -> friction patch
__global__ void k_writeback_rows(const uint32_t* __restrict__ counters, Rows R, const uint32_t* __restrict__ pairSlots, float* __restrict__ cForce, float4* __restrict__ frictions,
                                 float4* __restrict__ frReport, const uint2* __restrict__ pairBodies, const float4* __restrict__ pos, const float4* __restrict__ quat) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= counters[C_NCON]) return;
  const size_t s = R.stride;
  const float4 c = R.f[2 * s + k]; const uint32_t z = __float_as_uint(c.z), i = __float_as_uint(c.w);
  const int numNormal = (int)(z & 0xff), numFriction = (int)((z >> 8) & 0xff);
  const float4 ap = R.f[13 * s + k];
  for (int j = 0; j < numNormal; ++j) cForce[(size_t)i * 4 + j] = f4get(ap, j);
  if (numFriction && R.broken[k]) frictions[(size_t)pairSlots[i] * PXB_FRICTION_F4 + 1].w = __int_as_float(1);
  if (frReport) { const uint32_t a0 = pairBodies[i].x; friction_report_store(frReport, i, numFriction, R.f[14 * s + k], R.f[15 * s + k], R.f[24 * s + k], frictions + (size_t)pairSlots[i] * PXB_FRICTION_F4, pos[a0], quat[a0]); }
}

__global__ void k_finalize_bodies_pgs(uint32_t nDyn, const uint32_t* __restrict__ dynActor, float dt, float4* __restrict__ pos, float4* __restrict__ quat, float4* __restrict__ linVel,
                                      float4* __restrict__ angVel, const float4* __restrict__ sbLin, const float4* __restrict__ sbAng, const float4* __restrict__ sbDLin,
                                      const float4* __restrict__ sbDAng, const float4* __restrict__ sbIA, const float4* __restrict__ sbIB, const float4* __restrict__ invInertia, SleepArgs S,
                                      const uint32_t* __restrict__ geomFlags) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nDyn) return;
  const uint32_t a = dynActor[d];
  if (body_asleep(S, a) || !gf_dynamic(geomFlags[a])) return;   // (removed actors lose their dynamic bit; kinematic bodies move in k_kin_finalize)
  const float4 p4 = pos[a]; v3 p = V3(p4.x, p4.y, p4.z); q4 q = Q4(quat[a]); v3 lv = V3(linVel[a]), av = V3(angVel[a]);
  const m33 sI = load_sym(sbIA[a], sbIB[a]);
  v3 motionLin, motionAng;   // motionVelocityArray after integrateCore
  integrate_core_pgs(p, q, lv, av, sI, V3(sbDLin[a]), V3(sbDAng[a]), V3(sbLin[a]), V3(sbAng[a]), dt, __float_as_uint(sbIB[a].z), motionLin, motionAng);
  pos[a] = make_float4(p.x, p.y, p.z, p4.w); quat[a] = F4(q); linVel[a] = F4(lv, 0.f); angVel[a] = F4(av, 0.f);
  if (S.threshold > 0.f) sleep_check_dev(S, a, q, invInertia[a], p4.w, motionLin, motionAng);
}
#endif
