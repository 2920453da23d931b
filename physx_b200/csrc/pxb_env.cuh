// pxb_env.cuh -- the environment-partitioned fast path of the rigid-body step (BASELINE configs 2 and 5:
// thousands of independent RL environments in one scene, `PxActor::setEnvironmentID`).
//
// Environments never interact (env-ID filter, gpubroadphase/src/CUDA/broadphase.cu:62-80), so every island is
// contained in one environment.  That turns the two device-wide problems of the step into per-environment
// problems small enough for one SM's shared memory:
//   k_env_bp     a1-a7   one WARP per environment: bounds -> all-pairs AABB test in shared memory, emitted in
//                        sorted key order -> found/lost diff against the environment's last-frame segment.
//                        Replaces bounds + 2 x 5-pass radix sort + grid sweep + lifecycle (~40 launches).
//   k_env_solve  a12-a18 one CTA per environment: pre-integration, order-preserving first-fit colouring,
//                        contact prep straight into SHARED MEMORY rows, all TGS iterations with CTA barriers
//                        (no grid sync), write-back and integration.  Rows never touch HBM; iterations 2..n
//                        of the reference's "re-stream every row per iteration" traffic are served on chip.
// Both produce bit-identical results to the device-wide path (same arithmetic, same Gauss-Seidel order inside
// every island); tests/test_gpu_parity.py checks env path == global path == oracle.
// Included by pxb_engine.cu after the shared helpers (bp_test, lower_bound_u64, solve_constraint, ...).
#pragma once

#define ENV_BP_WARPS 4
#define ENV_MAX_LIST 288        // actors per environment incl. the shared env-less statics (eligibility limit)
#define ENV_MAX_GLOBALS 32

struct EnvBpArgs {
  uint32_t nEnv, maxList, bitsA, cap, ringMask; int externalTight; float contactOffset;
  const uint32_t *envStart, *envList;
  const float4 *pos, *quat, *dims; const uint32_t *geomFlags, *envId; float* tight;
  const uint64_t* oldKeys; const uint32_t* oldSlots; const uint2* oldSeg;
  uint64_t* newKeys; uint32_t* newSlots; uint2* newSeg;
  uint32_t *counters, *freeRing; uint64_t *createdKeys, *deletedKeys; float4 *manifolds, *frictions;
};

__device__ __forceinline__ void env_pair_norm(uint32_t n, uint32_t& i, uint32_t& j) {
  while (i + 1 < n && j >= n) { j = j - n + i + 2; ++i; }
}

__global__ void __launch_bounds__(32 * ENV_BP_WARPS) k_env_bp(const EnvBpArgs A) {
  extern __shared__ float4 envBpSmem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t e = blockIdx.x * ENV_BP_WARPS + warp;
  if (e >= A.nEnv) return;   // warps are independent: no CTA-wide barrier below
  float4* sMin = envBpSmem + (size_t)warp * 2 * A.maxList; float4* sMax = sMin + A.maxList;
  uint32_t* sAct = reinterpret_cast<uint32_t*>(envBpSmem + (size_t)ENV_BP_WARPS * 2 * A.maxList) + (size_t)warp * A.maxList;
  const uint32_t ls = A.envStart[e], n = A.envStart[e + 1] - ls;
  // a1/a2: bounds of this environment's actors (+ the shared env-less statics), inflated, into shared memory
  for (uint32_t k = lane; k < n; k += 32) {
    const uint32_t a = A.envList[ls + k]; const uint32_t gf = A.geomFlags[a], env = A.envId[a];
    float mn[3], mx[3];
    if (A.externalTight) { for (int c = 0; c < 3; ++c) { mn[c] = A.tight[a * 6 + c]; mx[c] = A.tight[a * 6 + 3 + c]; } }
    else {
      const float4 p4 = A.pos[a];
      tight_bounds(gf & 0xff, V3(p4.x, p4.y, p4.z), Q4(A.quat[a]), A.dims[a], mn, mx);
      if (env == e || e == 0) for (int c = 0; c < 3; ++c) { A.tight[a * 6 + c] = mn[c]; A.tight[a * 6 + 3 + c] = mx[c]; }
    }
    const float co = A.contactOffset;
    sMin[k] = make_float4(mn[0] - co, mn[1] - co, mn[2] - co, __uint_as_float(env));
    sMax[k] = make_float4(mx[0] + co, mx[1] + co, mx[2] + co, __uint_as_float(gf));
    sAct[k] = a;
  }
  __syncwarp();
  // a4/a5: all pairs (i<j) of the list in row-major order = ascending (lo,hi) key order because the list is sorted by
  // actor index.  Pass 1 counts, one atomic reserves the segment, pass 2 emits.
  const uint32_t total = n * (n - 1) / 2;
  uint32_t cnt = 0;
  {
    uint32_t i = 0, j = 1 + lane;
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
      env_pair_norm(n, i, j);
      const bool hit = (i + 1 < n) && bp_test(sMin[i], sMax[i], sMin[j], sMax[j]);
      cnt += __popc(__ballot_sync(0xffffffffu, hit));
      j += 32;
    }
  }
  uint32_t base = 0;
  if (lane == 0 && cnt) base = atomicAdd(&A.counters[C_NPAIRS_NEW], cnt);
  base = __shfl_sync(0xffffffffu, base, 0);
  if (base + cnt > A.cap) {   // capacity exceeded: report, keep the flat list crash-free (sentinel keys), drop the segment
    if (lane == 0) atomicOr(&A.counters[C_ERROR], (uint32_t)E_PAIR_OVERFLOW);
    for (uint32_t t = base + lane; t < min(base + cnt, A.cap); t += 32) { A.newKeys[t] = ~0ull; A.newSlots[t] = 0; }
    cnt = 0;
  }
  {
    uint32_t i = 0, j = 1 + lane, w = 0;
    for (uint32_t t0 = 0; t0 < total && cnt; t0 += 32) {
      env_pair_norm(n, i, j);
      const bool hit = (i + 1 < n) && bp_test(sMin[i], sMax[i], sMin[j], sMax[j]);
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (hit) A.newKeys[base + w + __popc(m & ((1u << lane) - 1u))] = ((uint64_t)sAct[i] << A.bitsA) | sAct[j];
      w += __popc(m);
      j += 32;
    }
  }
  __syncwarp();
  // a7: pair lifecycle against last frame's segment of the same environment (both sorted)
  const uint2 os = A.oldSeg[e]; const uint32_t ob = os.x, oc = os.y;
  for (uint32_t t = lane; t < cnt; t += 32) {
    const uint64_t k = A.newKeys[base + t];
    uint32_t slot = NONE32;
    if (t < oc && A.oldKeys[ob + t] == k) slot = A.oldSlots[ob + t];
    else { const uint32_t p = lower_bound_u64(A.oldKeys + ob, oc, k); if (p < oc && A.oldKeys[ob + p] == k) slot = A.oldSlots[ob + p]; }
    if (slot == NONE32) {
      // pops only consume ring entries that existed when the step began (C_FREE_SNAP), pushes of this step land behind them
      const uint32_t h = atomicAdd(&A.counters[C_FREE_HEAD], 1u);
      if ((int32_t)(A.counters[C_FREE_SNAP] - h) <= 0) { atomicOr(&A.counters[C_ERROR], (uint32_t)E_PAIR_OVERFLOW); slot = 0; }
      else slot = A.freeRing[h & A.ringMask];
      A.createdKeys[atomicAdd(&A.counters[C_NCREATED], 1u)] = k;
      float4* m = A.manifolds + (size_t)slot * PXB_MANIFOLD_F4;
      m[0] = make_float4(__int_as_float(0), FLT_MAX, FLT_MAX, FLT_MAX); m[1] = make_float4(0, 0, 0, 1); m[2] = make_float4(0, 0, 0, 1); m[3] = make_float4(0, 0, 0, 1);
      float4* f = A.frictions + (size_t)slot * PXB_FRICTION_F4;
      f[0] = make_float4(0, 0, 0, __int_as_float(0)); f[1] = make_float4(0, 0, 0, __int_as_float(0)); f[2] = make_float4(0, 0, 0, __int_as_float(0));
    }
    A.newSlots[base + t] = slot;
  }
  for (uint32_t t = lane; t < oc; t += 32) {
    const uint64_t k = A.oldKeys[ob + t];
    if (t < cnt && A.newKeys[base + t] == k) continue;
    const uint32_t p = lower_bound_u64(A.newKeys + base, cnt, k);
    if (p < cnt && A.newKeys[base + p] == k) continue;
    A.freeRing[atomicAdd(&A.counters[C_FREE_TAIL], 1u) & A.ringMask] = A.oldSlots[ob + t];
    A.deletedKeys[atomicAdd(&A.counters[C_NDELETED], 1u)] = k;
  }
  if (lane == 0) A.newSeg[e] = make_uint2(base, cnt);
}

// ---------------------------------------------------------------------------------------------
struct EnvSolveArgs {
  uint32_t nEnv, maxList, conCap, cap, posIters, velIters; float dt, gx, gy, gz; SolverParams P;
  const uint32_t *envStart, *envList, *actorLocal; const uint2* seg;
  float4 *pos, *quat, *linVel, *angVel; const float4 *invInertia, *damp; const uint32_t* geomFlags;
  const uint32_t* pairSlots; const uint2* pairBodies; const float4 *cHdr, *cPts; float* cForce; float4* frictions;
  uint32_t *conPair, *conB0, *conB1, *conColour, *ordered, *broken;   // per-pair-index scratch (global, L2 resident)
  float4 *rowA, *rowB; uint4* rowC; float4 *ptA, *ptB, *ptC, *frA, *frB, *frC, *frD;  // global rows: only for environments that do not fit conCap
  uint32_t* counters;
};

struct Rows { float4 *rowA, *rowB; uint4* rowC; float4 *ptA, *ptB, *ptC, *frA, *frB, *frC, *frD; uint32_t stride; };

template <int T>
__device__ __forceinline__ void env_solve_body(const EnvSolveArgs& A, const Rows R, const uint32_t e, const uint32_t base, const uint32_t nCon, const uint32_t n, const uint32_t* __restrict__ list,
                                               float4* bLin, float4* bAng, float4* bDLin, float4* bDAng, float4* bIA, float4* bIB, float4* bP, float4* bQ,
                                               const uint32_t* sPartStart, const uint32_t nPart) {
  const uint32_t tid = threadIdx.x;
  // a14: contact prep, one thread per constraint in partition-major order
  for (uint32_t pos = tid; pos < nCon; pos += T) {
    const uint32_t k = A.ordered[base + pos]; const uint32_t i = base + A.conPair[base + k];
    const uint32_t l0 = A.conB0[base + k], l1 = A.conB1[base + k];
    const uint2 bb = A.pairBodies[i];
    PrepBodies B;
    { const float4 p = A.pos[bb.x]; B.f0.p = V3(p.x, p.y, p.z); B.f0.q = Q4(A.quat[bb.x]); B.invMass0 = p.w; const float4 q = A.pos[bb.y]; B.f1.p = V3(q.x, q.y, q.z); B.f1.q = Q4(A.quat[bb.y]); }
    const bool dyn1 = l1 != NONE32;
    B.invMass1 = dyn1 ? A.pos[bb.y].w : 0.f;
    B.pen0 = -A.invInertia[bb.x].w; B.pen1 = dyn1 ? -A.invInertia[bb.y].w : -FLT_MAX;
    B.linVel0 = V3(bLin[l0]); B.angVel0 = V3(bQ[l0]); B.sI0 = load_sym(bIA[l0], bIB[l0]);
    if (dyn1) { B.linVel1 = V3(bLin[l1]); B.angVel1 = V3(bQ[l1]); B.sI1 = load_sym(bIA[l1], bIB[l1]); }
    else { B.linVel1 = V3(0, 0, 0); B.angVel1 = V3(0, 0, 0); B.sI1.c0 = B.sI1.c1 = B.sI1.c2 = V3(0, 0, 0); }
    prep_constraint(pos, R.stride, i, l0, l1, B, A.cHdr, A.cPts, A.frictions + (size_t)A.pairSlots[i] * PXB_FRICTION_F4, A.P,
                    R.rowA, R.rowB, R.rowC, R.ptA, R.ptB, R.ptC, R.frA, R.frB, R.frC, R.frD);
    A.broken[base + pos] = 0u;
  }
  __syncthreads();
  for (uint32_t b = tid; b < n; b += T) bQ[b] = make_float4(0, 0, 0, 1);   // bQ carried the unconstrained angular velocity during prep
  __syncthreads();
  // a15/a16: iterativeSolveIsland (DyTGSDynamics.cpp:2515-2793) with CTA barriers between partitions
  const float stepDt = A.P.stepDt;
  float elapsed = 0.f;
  for (uint32_t it = 0; it < A.posIters + A.velIters; ++it) {
    const bool vel = it >= A.posIters;
    const float minPen = vel ? 0.f : -FLT_MAX;
    for (uint32_t p = 0; p < nPart; ++p) {
      const uint32_t pb = sPartStart[p], pe = sPartStart[p + 1];
      for (uint32_t k = pb + tid; k < pe; k += T)
        solve_constraint(k, R.stride, minPen, elapsed, R.rowA, R.rowB, R.rowC, R.ptA, R.ptB, R.ptC, R.frA, R.frB, R.frC, R.frD, bLin, bAng, bDLin, bDAng, A.broken + base);
      __syncthreads();
    }
    if (!vel) {
      for (uint32_t b = tid; b < n; b += T) {
        if (!__float_as_uint(bP[b].w)) continue;
        v3 p = V3(bP[b]); q4 dq = Q4(bQ[b]); v3 dl = V3(bDLin[b]), da = V3(bDAng[b]);
        integrate_core_step(V3(bLin[b]), V3(bAng[b]), load_sym(bIA[b], bIB[b]), stepDt, p, dq, dl, da);
        bP[b] = F4(p, __uint_as_float(1u)); bQ[b] = F4(dq); bDLin[b] = F4(dl, 0.f); bDAng[b] = F4(da, 0.f);
      }
      elapsed += stepDt;
      __syncthreads();
    }
  }
  // a17: writeBackContact
  for (uint32_t pos = tid; pos < nCon; pos += T) {
    const uint4 rc = R.rowC[pos]; const uint32_t i = rc.w; const uint32_t numNormal = rc.z & 0xff, numFriction = (rc.z >> 8) & 0xff;
    for (uint32_t j = 0; j < numNormal; ++j) A.cForce[(size_t)i * 4 + j] = R.ptC[(size_t)j * R.stride + pos].w;
    if (numFriction && A.broken[base + pos]) A.frictions[(size_t)A.pairSlots[i] * PXB_FRICTION_F4 + 1].w = __int_as_float(1);
  }
}

template <int T>
__global__ void __launch_bounds__(T) k_env_solve(const EnvSolveArgs A) {
  extern __shared__ float4 envSmem[];
  __shared__ uint32_t sPartCnt[MAX_PARTITIONS + 1], sPartStart[MAX_PARTITIONS + 1], sWarp[T / 32 + 1], sMisc[4];
  const uint32_t e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ls = A.envStart[e], n = A.envStart[e + 1] - ls; const uint32_t* list = A.envList + ls;
  const uint2 sg = A.seg[e]; const uint32_t base = sg.x, m = sg.y;
  const uint32_t nb = A.maxList;
  float4 *bLin = envSmem, *bAng = bLin + nb, *bDLin = bAng + nb, *bDAng = bDLin + nb, *bIA = bDAng + nb, *bIB = bIA + nb, *bP = bIB + nb, *bQ = bP + nb;
  float4* rowsSmem = bQ + nb;
  uint32_t* bMask = reinterpret_cast<uint32_t*>(rowsSmem + (size_t)31 * A.conCap); uint32_t* bFirst = bMask + nb; uint32_t* bStat = bFirst + nb;
  // a12: preIntegrateBodies (DyTGSDynamics.cpp:992-1021) into shared memory
  for (uint32_t b = tid; b < n; b += T) {
    const uint32_t a = list[b];
    bMask[b] = 0; bStat[b] = 0;
    if (!(A.geomFlags[a] & 0x100u)) { bP[b] = make_float4(0, 0, 0, 0); continue; }
    const float4 dm = A.damp[a]; const float4 ii = A.invInertia[a]; const float4 p4 = A.pos[a];
    v3 lv = V3(A.linVel[a]), av = V3(A.angVel[a]);
    unconstrained_velocity(V3(A.gx, A.gy, A.gz), A.dt, dm.x, dm.y, dm.z, dm.w, lv, av);
    const m33 rot = amfromq(Q4(A.quat[a]));
    const v3 sqrtInvI = V3(ii.x == 0.f ? 0.f : sqrtf(ii.x), ii.y == 0.f ? 0.f : sqrtf(ii.y), ii.z == 0.f ? 0.f : sqrtf(ii.z));
    const v3 sqrtI = V3(sqrtInvI.x == 0.f ? 0.f : 1.0f / sqrtInvI.x, sqrtInvI.y == 0.f ? 0.f : 1.0f / sqrtInvI.y, sqrtInvI.z == 0.f ? 0.f : 1.0f / sqrtInvI.z);
    m33 sI, sInertia; transform_inertia(sqrtInvI, rot, sI); transform_inertia(sqrtI, rot, sInertia);
    bLin[b] = F4(lv, 0.f); bAng[b] = F4(mmul(sInertia, av), 0.f); bDLin[b] = make_float4(0, 0, 0, 0); bDAng[b] = make_float4(0, 0, 0, 0);
    bIA[b] = make_float4(sI.c0.x, sI.c0.y, sI.c0.z, sI.c1.y); bIB[b] = make_float4(sI.c1.z, sI.c2.z, 0.f, 0.f);
    bP[b] = make_float4(p4.x, p4.y, p4.z, __uint_as_float(0u)); bQ[b] = F4(av, 0.f);
  }
  for (uint32_t p = tid; p < MAX_PARTITIONS + 1; p += T) sPartCnt[p] = 0;
  if (tid == 0) { sMisc[0] = 0; sMisc[1] = 0; }
  __syncthreads();
  // constraint list = this environment's pairs that produced contacts, in ascending key order (ordered compaction)
  uint32_t nCon = 0;
  for (uint32_t t0 = 0; t0 < m; t0 += T) {
    const uint32_t t = t0 + tid; bool f = false;
    if (t < m) {
      f = __float_as_int(A.cHdr[base + t].w) > 0;
      if (!f) A.frictions[(size_t)A.pairSlots[base + t] * PXB_FRICTION_F4 + 2].w = __int_as_float(0);   // no contacts: the friction patch is dropped
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) sWarp[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = nCon, tot = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) { const uint32_t c = sWarp[w]; if (w < (int)warp) off += c; tot += c; }
    if (f) {
      const uint32_t k = off + __popc(bal & ((1u << lane) - 1u));
      const uint2 bb = A.pairBodies[base + t];
      const uint32_t l0 = A.actorLocal[bb.x]; const uint32_t l1 = (A.geomFlags[bb.y] & 0x100u) ? A.actorLocal[bb.y] : NONE32;
      A.conPair[base + k] = t; A.conB0[base + k] = l0; A.conB1[base + k] = l1; A.conColour[base + k] = NONE32;
      bP[l0].w = __uint_as_float(1u); if (l1 != NONE32) bP[l1].w = __uint_as_float(1u);   // hasConstraints (benign same-value races)
      if (l1 == NONE32) atomicAdd(&bStat[l0], 1u);
    }
    nCon += tot;
    __syncthreads();
  }
  // a13: first-fit colouring in constraint order (classifyConstraintDesc, DyConstraintPartition.cpp:475-568).  A constraint
  // is ready once it is the lowest-numbered uncoloured constraint on both of its bodies; ready constraints are body-disjoint.
  for (;;) {
    for (uint32_t b = tid; b < n; b += T) bFirst[b] = NONE32;
    __syncthreads();
    int undone = 0;
    for (uint32_t k = tid; k < nCon; k += T) {
      const uint32_t l1 = A.conB1[base + k];
      if (l1 == NONE32 || A.conColour[base + k] != NONE32) continue;
      atomicMin(&bFirst[A.conB0[base + k]], k); atomicMin(&bFirst[l1], k); undone = 1;
    }
    if (!__syncthreads_or(undone)) break;
    for (uint32_t k = tid; k < nCon; k += T) {
      const uint32_t l1 = A.conB1[base + k];
      if (l1 == NONE32 || A.conColour[base + k] != NONE32) continue;
      const uint32_t l0 = A.conB0[base + k];
      if (bFirst[l0] != k || bFirst[l1] != k) continue;
      const uint32_t ma = bMask[l0], mb = bMask[l1]; const uint32_t comb = ~ma & ~mb;
      uint32_t col = 31;
      if (comb) col = __ffs(comb) - 1; else atomicOr(&A.counters[C_ERROR], (uint32_t)E_COLOUR_OVERFLOW);
      A.conColour[base + k] = col; bMask[l0] = ma | (1u << col); bMask[l1] = mb | (1u << col);
    }
    __syncthreads();
  }
  // static contacts of a body go to partitions maxDynamicColour(body) + rank among the body's static contacts (:203-262)
  for (uint32_t k = tid; k < nCon; k += T) {
    uint32_t col;
    if (A.conB1[base + k] == NONE32) {
      const uint32_t l0 = A.conB0[base + k]; uint32_t rank = 0;
      if (bStat[l0] > 1) for (uint32_t kk = 0; kk < k; ++kk) if (A.conB1[base + kk] == NONE32 && A.conB0[base + kk] == l0) ++rank;
      const uint32_t mk = bMask[l0]; col = (mk ? 32u - __clz(mk) : 0u) + rank;
      if (col >= MAX_PARTITIONS) { col = MAX_PARTITIONS - 1; atomicOr(&A.counters[C_ERROR], (uint32_t)E_PARTITION_OVERFLOW); }
      A.conColour[base + k] = col;
    } else col = A.conColour[base + k];
    atomicAdd(&sPartCnt[col], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t s = 0, np = 0;
    for (uint32_t p = 0; p < MAX_PARTITIONS; ++p) { const uint32_t c = sPartCnt[p]; sPartStart[p] = s; sPartCnt[p] = s; s += c; if (c) np = p + 1; }
    sPartStart[MAX_PARTITIONS] = s; sMisc[0] = np;
    if (nCon) { atomicAdd(&A.counters[C_NCON], nCon); atomicMax(&A.counters[C_NPART], np); atomicMax(&A.counters[C_MAXCONENV], nCon); }
  }
  __syncthreads();
  const uint32_t nPart = sMisc[0];
  for (uint32_t k = tid; k < nCon; k += T) A.ordered[base + atomicAdd(&sPartCnt[A.conColour[base + k]], 1u)] = k;
  __syncthreads();
  if (nCon) {
    Rows R;
    if (nCon <= A.conCap) {
      const uint32_t c = A.conCap; float4* r = rowsSmem;
      R.rowA = r; R.rowB = r + c; R.rowC = reinterpret_cast<uint4*>(r + 2 * c); R.ptA = r + 3 * c; R.ptB = r + 7 * c; R.ptC = r + 11 * c;
      R.frA = r + 15 * c; R.frB = r + 19 * c; R.frC = r + 23 * c; R.frD = r + 27 * c; R.stride = c;
      env_solve_body<T>(A, R, e, base, nCon, n, list, bLin, bAng, bDLin, bDAng, bIA, bIB, bP, bQ, sPartStart, nPart);
    } else {   // oversize environment: rows stream through global memory (L2), same arithmetic
      R.rowA = A.rowA + base; R.rowB = A.rowB + base; R.rowC = A.rowC + base; R.ptA = A.ptA + base; R.ptB = A.ptB + base; R.ptC = A.ptC + base;
      R.frA = A.frA + base; R.frB = A.frB + base; R.frC = A.frC + base; R.frD = A.frD + base; R.stride = A.cap;
      env_solve_body<T>(A, R, e, base, nCon, n, list, bLin, bAng, bDLin, bDAng, bIA, bIB, bP, bQ, sPartStart, nPart);
    }
  } else {
    for (uint32_t b = tid; b < n; b += T) bQ[b] = make_float4(0, 0, 0, 1);
  }
  __syncthreads();
  // a18: copyBackBodies (DyTGSDynamics.cpp:1549-1580); bodies without constraints take one full-dt step (:2573-2577)
  for (uint32_t b = tid; b < n; b += T) {
    const uint32_t a = list[b];
    if (!(A.geomFlags[a] & 0x100u)) continue;
    const m33 sI = load_sym(bIA[b], bIB[b]);
    v3 p = V3(bP[b]); q4 dq = Q4(bQ[b]);
    const v3 lv = V3(bLin[b]), as = V3(bAng[b]);
    if (!__float_as_uint(bP[b].w)) { v3 dl = V3(0, 0, 0), da = V3(0, 0, 0); integrate_core_step(lv, as, sI, A.dt, p, dq, dl, da); }
    const float invMass = A.pos[a].w;
    const q4 q = qnormalized(qmul(dq, Q4(A.quat[a])));
    A.pos[a] = make_float4(p.x, p.y, p.z, invMass); A.quat[a] = F4(q);
    A.linVel[a] = F4(lv, 0.f); A.angVel[a] = F4(mmul(sI, as), 0.f);
  }
}
