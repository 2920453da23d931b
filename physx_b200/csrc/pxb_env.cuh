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
// Instantiated by pxb_env.cu; shared helpers (tight_bounds, lower_bound_u64, sleep_check_dev, ...) come from pxb_common.cuh.
#pragma once
#include "pxb_common.cuh"

#define ENV_BP_WARPS 4
#define ENV_MAX_LIST 288        // actors per environment incl. the shared env-less statics (eligibility limit)
#define ENV_MAX_GLOBALS 32

struct EnvBpArgs {
  uint32_t nEnv, maxList, bitsA, cap, ringMask; int externalTight; float contactOffset;
  const uint32_t *envStart, *envList;
  const float4 *pos, *quat, *dims; const uint32_t *geomFlags, *envId; float* tight; HullArrays hulls;
  const uint64_t* oldKeys; const uint32_t* oldSlots; const uint2* oldSeg;
  uint64_t* newKeys; uint32_t* newSlots; uint2* newSeg;
  uint32_t *counters, *freeRing, *slotColour; uint64_t *createdKeys, *deletedKeys; float4 *manifolds, *frictions; TouchLists touch; LocalPoses L; const uint32_t* aggId; const float2* shapeOff;
  // temporal coherence of k_env_bp (NULL candKeys = off): per environment a list of candidate pairs (list positions i << 16 | j, ascending) whose bounds overlapped when expanded
  // by candMargin, and the bounds that list was built from
  uint32_t* candKeys; uint32_t* candCount; float4 *refMin, *refMax; uint32_t candCap; float candMargin;   // shapeOff: per-actor (contactOffset, restOffset), LOCAL instantiation only
};

#define ENV_BP_STAGE 256   // pair keys staged per warp in shared memory before the segment base is known

// branch-free AABB test of the environment path: closed-interval overlap on all axes (PxgIntegerAABB::intersects / ABP
// intersect2D semantics) and at least one dynamic actor (BpFiltering.h:99-114).  Every list member is in the warp's own
// environment or env-less, so the environment filter (broadphase.cu:62-80) always passes here.
// AGG (the LOCAL instantiations): the .w of the minima carries an aggregate key -- the aggregate id for members of an aggregate without self collisions, a value
// unique in the list otherwise -- and equal keys never pair.
template <bool AGG = false>
__device__ __forceinline__ bool env_bp_test(const float4& amin, const float4& amax, const float4& bmin, const float4& bmax) {
  const bool sep = (amin.x > bmax.x) | (bmin.x > amax.x) | (amin.y > bmax.y) | (bmin.y > amax.y) | (amin.z > bmax.z) | (bmin.z > amax.z);
  const bool ok = !sep & (((__float_as_uint(amax.w) | __float_as_uint(bmax.w)) & 0x100u) != 0);
  return AGG ? (ok & (__float_as_uint(amin.w) != __float_as_uint(bmin.w))) : ok;
}
__device__ __forceinline__ uint32_t env_agg_key(const uint32_t* __restrict__ aggId, uint32_t a, uint32_t k) {
  const uint32_t g = aggId ? aggId[a] : 0u;
  return (g && !(g & 0x80000000u)) ? g : (0x40000000u | k);
}

// a7: pair lifecycle of one environment against last frame's segment of the same environment (both sorted); `tid` of `stride` cooperating threads.
// Returns false when this thread saw a key that was not at the same position last frame (segment changed).
__device__ __forceinline__ bool env_bp_lifecycle(const EnvBpArgs& A, uint32_t e, uint32_t base, uint32_t cnt, uint32_t tid, uint32_t stride) {
  const uint2 os = A.oldSeg[e]; const uint32_t ob = os.x, oc = os.y & 0x7fffffffu;
  bool same = cnt == oc;   // segment identical to last frame's (steady state): k_env_solve may reuse last frame's colouring
  for (uint32_t t = tid; t < cnt; t += stride) {
    const uint64_t k = A.newKeys[base + t];
    uint32_t slot = NONE32;
    if (t < oc && A.oldKeys[ob + t] == k) slot = A.oldSlots[ob + t];
    else {
      same = false; const uint32_t p = lower_bound_u64(A.oldKeys + ob, oc, k); if (p < oc && A.oldKeys[ob + p] == k) slot = A.oldSlots[ob + p]; }
    if (slot == NONE32) {
      // pops only consume ring entries that existed when the step began (C_FREE_SNAP), pushes of this step land behind them
      const uint32_t h = atomicAdd(&A.counters[C_FREE_HEAD], 1u);
      if ((int32_t)(A.counters[C_FREE_SNAP] - h) <= 0) { atomicOr(&A.counters[C_ERROR], (uint32_t)E_PAIR_OVERFLOW); slot = 0; }
      else slot = A.freeRing[h & A.ringMask];
      A.createdKeys[atomicAdd(&A.counters[C_NCREATED], 1u)] = k;
      float4* m = A.manifolds + (size_t)slot * PXB_MANIFOLD_F4;
      m[0] = make_float4(__int_as_float(0), FLT_MAX, FLT_MAX, FLT_MAX); m[1] = make_float4(0, 0, 0, 1); m[2] = make_float4(0, 0, 0, 1); m[3] = make_float4(0, 0, 0, 1); m[14] = make_float4(0, 0, 0, 0);
      float4* f = A.frictions + (size_t)slot * PXB_FRICTION_F4;
      f[0] = make_float4(0, 0, 0, __int_as_float(0)); f[1] = make_float4(0, 0, 0, __int_as_float(0)); f[2] = make_float4(0, 0, 0, __int_as_float(0));
      A.slotColour[slot] = NONE32; A.touch.state[slot] = 0u;
    }
    A.newSlots[base + t] = slot;
  }
  for (uint32_t t = tid; t < oc; t += stride) {
    const uint64_t k = A.oldKeys[ob + t];
    if (t < cnt && A.newKeys[base + t] == k) continue;
    const uint32_t p = lower_bound_u64(A.newKeys + base, cnt, k);
    if (p < cnt && A.newKeys[base + p] == k) continue;
    A.freeRing[atomicAdd(&A.counters[C_FREE_TAIL], 1u) & A.ringMask] = A.oldSlots[ob + t];
    A.deletedKeys[atomicAdd(&A.counters[C_NDELETED], 1u)] = k;
    touch_event(A.touch, A.counters, A.oldSlots[ob + t], k, false);
  }
  return same;
}

template <bool HULLS, bool LOCAL>   // HULLS: the scene holds convex meshes; LOCAL: it has local poses (the plain instantiation carries neither code)
__global__ void __launch_bounds__(32 * ENV_BP_WARPS) k_env_bp(const EnvBpArgs A) {
  extern __shared__ float4 envBpSmem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t e = blockIdx.x * ENV_BP_WARPS + warp;
  if (e >= A.nEnv) return;   // warps are independent: no CTA-wide barrier below
  float4* sMin = envBpSmem + (size_t)warp * 2 * A.maxList; float4* sMax = sMin + A.maxList;
  uint64_t* sStage = reinterpret_cast<uint64_t*>(envBpSmem + (size_t)ENV_BP_WARPS * 2 * A.maxList) + (size_t)warp * ENV_BP_STAGE;
  uint32_t* sAct = reinterpret_cast<uint32_t*>(envBpSmem + (size_t)ENV_BP_WARPS * 2 * A.maxList) + ENV_BP_WARPS * ENV_BP_STAGE * 2 + (size_t)warp * A.maxList;
  const uint32_t ls = A.envStart[e], n = A.envStart[e + 1] - ls;
  const uint32_t nCand = A.candKeys ? A.candCount[e] : NONE32;   // NONE32: no valid candidate list
  bool moved = false;
  // a1/a2: bounds of this environment's actors (+ the shared env-less statics), inflated, into shared memory
  for (uint32_t k = lane; k < n; k += 32) {
    const uint32_t a = A.envList[ls + k]; const uint32_t gf = A.geomFlags[a], env = A.envId[a];
    float mn[3], mx[3];
    const bool own = env == e || e == 0;   // the shared env-less statics are in every environment's list: environment 0 writes their bounds and their transform-cache entry
    xf shape; shape.p = V3(0, 0, 0); shape.q = Q4(0, 0, 0, 1);
    if (LOCAL) shape = shape_world_pose(A.L, a, A.pos[a], A.quat[a], own);
    else if (!A.externalTight) { const float4 p4 = A.pos[a]; shape.p = V3(p4.x, p4.y, p4.z); shape.q = Q4(A.quat[a]); }
    if (A.externalTight) { for (int c = 0; c < 3; ++c) { mn[c] = A.tight[a * 6 + c]; mx[c] = A.tight[a * 6 + 3 + c]; } }
    else {
      tight_bounds(gf & 0xff, shape.p, shape.q, A.dims[a], mn, mx, HULLS ? &A.hulls : nullptr);
      if (own) for (int c = 0; c < 3; ++c) { A.tight[a * 6 + c] = mn[c]; A.tight[a * 6 + 3 + c] = mx[c]; }
    }
    const float co = (LOCAL && A.shapeOff) ? A.shapeOff[a].x : A.contactOffset;   // every bound is inflated by its own shape's contact offset
    sMin[k] = make_float4(mn[0] - co, mn[1] - co, mn[2] - co, __uint_as_float(LOCAL ? env_agg_key(A.aggId, a, k) : env));
    sMax[k] = make_float4(mx[0] + co, mx[1] + co, mx[2] + co, __uint_as_float(gf_dynamic(gf) ? gf : (gf & ~0x100u)));   // kinematic bodies pair with dynamic ones only, like statics (the test below reads the dynamic bit)
    sAct[k] = a;
    if (nCand != NONE32) {   // has a face of this bound moved further than the candidate list allows?
      const float4 rm = A.refMin[ls + k], rx = A.refMax[ls + k]; const float4 cm = sMin[k], cx = sMax[k]; const float lim = 0.9f * A.candMargin;
      moved |= (fabsf(cm.x - rm.x) > lim) | (fabsf(cm.y - rm.y) > lim) | (fabsf(cm.z - rm.z) > lim) | (fabsf(cx.x - rx.x) > lim) | (fabsf(cx.y - rx.y) > lim) | (fabsf(cx.z - rx.z) > lim);
    }
  }
  __syncwarp();
  // a4/a5: all pairs (i<j) of the list, row by row = ascending (lo,hi) key order because the list is sorted by actor
  // index.  Keys are staged in shared memory; one atomic reserves the environment's segment of the flat pair list.
  // Temporal coherence (the role of the reference's INCREMENTAL sweep): while no bound has moved by more than the margin since the candidate list was built, every
  // overlapping pair is a candidate (a.min <= b.max now implies a.min_ref - m <= b.max_ref + m), so only the candidates are tested -- with the same exact test, in the
  // same ascending order: the pair set is identical.  Otherwise the all-pairs enumeration runs and rebuilds the list.
  const bool useCand = nCand != NONE32 && !__any_sync(0xffffffffu, moved);
  const uint32_t* candList = A.candKeys ? A.candKeys + (size_t)e * A.candCap : nullptr;
  uint32_t cnt = 0;
  if (useCand) {
    for (uint32_t c0 = 0; c0 < nCand; c0 += 32) {
      const uint32_t c = c0 + lane; uint32_t ij = 0; bool hit = false;
      if (c < nCand) { ij = candList[c]; const uint32_t i = ij >> 16, j = ij & 0xffffu; hit = env_bp_test<LOCAL>(sMin[i], sMax[i], sMin[j], sMax[j]); }
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        const uint32_t w = cnt + __popc(m & ((1u << lane) - 1u));
        if (hit && w < ENV_BP_STAGE) sStage[w] = ((uint64_t)sAct[ij >> 16] << A.bitsA) | sAct[ij & 0xffffu];
        cnt += __popc(m);
      }
    }
  } else {
    uint32_t ccnt = 0; uint32_t* cw = A.candKeys ? A.candKeys + (size_t)e * A.candCap : nullptr; const float m2 = cw ? 2.f * A.candMargin : 0.f;
    for (uint32_t i = 0; i + 1 < n; ++i) {
      const float4 amin = sMin[i], amax = sMax[i]; const uint64_t hiKey = (uint64_t)sAct[i] << A.bitsA;
      const float4 emin = make_float4(amin.x - m2, amin.y - m2, amin.z - m2, amin.w), emax = make_float4(amax.x + m2, amax.y + m2, amax.z + m2, amax.w);   // both bounds expanded by the margin
      for (uint32_t j0 = i + 1; j0 < n; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool ch = j < n && env_bp_test<LOCAL>(emin, emax, sMin[j], sMax[j]);   // the expanded test first: a miss (the common case) is a miss of the exact test too
        const uint32_t cmk = __ballot_sync(0xffffffffu, ch);
        if (cmk) {
          const bool hit = cw ? (ch && env_bp_test<LOCAL>(amin, amax, sMin[j], sMax[j])) : ch;   // (without candidate lists the margin is zero and the first test was the exact one)
          const uint32_t m = __ballot_sync(0xffffffffu, hit);
          if (m) {
            const uint32_t w = cnt + __popc(m & ((1u << lane) - 1u));
            if (hit && w < ENV_BP_STAGE) sStage[w] = hiKey | sAct[j];
            cnt += __popc(m);
          }
          if (cw) { const uint32_t w = ccnt + __popc(cmk & ((1u << lane) - 1u)); if (ch && w < A.candCap) cw[w] = (i << 16) | j; ccnt += __popc(cmk); }
        }
      }
    }
    if (cw) {
      if (lane == 0) A.candCount[e] = ccnt <= A.candCap ? ccnt : NONE32;   // too many candidates: this environment keeps enumerating all pairs
      for (uint32_t k = lane; k < n; k += 32) { A.refMin[ls + k] = sMin[k]; A.refMax[ls + k] = sMax[k]; }
    }
  }
  uint32_t base = 0;
  if (lane == 0 && cnt) base = atomicAdd(&A.counters[C_NPAIRS_NEW], cnt);
  base = __shfl_sync(0xffffffffu, base, 0);
  __syncwarp();
  if (base + cnt > A.cap) {   // capacity exceeded: report, keep the flat list crash-free (sentinel keys), drop the segment
    if (lane == 0) atomicOr(&A.counters[C_ERROR], (uint32_t)E_PAIR_OVERFLOW);
    for (uint32_t t = base + lane; t < min(base + cnt, A.cap); t += 32) { A.newKeys[t] = ~0ull; A.newSlots[t] = 0; }
    cnt = 0;
  }
  if (cnt <= ENV_BP_STAGE) { for (uint32_t t = lane; t < cnt; t += 32) A.newKeys[base + t] = sStage[t]; }
  else if (useCand) {   // more pairs than the staging area holds: enumerate again, straight into the segment
    uint32_t w = 0;
    for (uint32_t c0 = 0; c0 < nCand; c0 += 32) {
      const uint32_t c = c0 + lane; uint32_t ij = 0; bool hit = false;
      if (c < nCand) { ij = candList[c]; const uint32_t i = ij >> 16, j = ij & 0xffffu; hit = env_bp_test<LOCAL>(sMin[i], sMax[i], sMin[j], sMax[j]); }
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (hit) A.newKeys[base + w + __popc(m & ((1u << lane) - 1u))] = ((uint64_t)sAct[ij >> 16] << A.bitsA) | sAct[ij & 0xffffu];
      w += __popc(m);
    }
  } else {
    uint32_t w = 0;
    for (uint32_t i = 0; i + 1 < n; ++i) {
      const float4 amin = sMin[i], amax = sMax[i]; const uint64_t hiKey = (uint64_t)sAct[i] << A.bitsA;
      for (uint32_t j0 = i + 1; j0 < n; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool hit = j < n && env_bp_test<LOCAL>(amin, amax, sMin[j], sMax[j]);
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (hit) A.newKeys[base + w + __popc(m & ((1u << lane) - 1u))] = hiKey | sAct[j];
        w += __popc(m);
      }
    }
  }
  __syncwarp();
  bool same = env_bp_lifecycle(A, e, base, cnt, lane, 32u);
  same = __all_sync(0xffffffffu, same);
  if (lane == 0) A.newSeg[e] = make_uint2(base, cnt | (same ? 0x80000000u : 0u));
}

// The same stage with a whole CTA per environment, for scenes of FEW environments (BASELINE config 1 is one environment of 101 actors): with one warp the
// step is that warp's latency (46 us for 100 boxes); here the rows of the all-pairs enumeration are dealt to the CTA's warps.  Two passes keep the key order:
// pass 1 counts the hits of every row, a scan gives each row its offset in the environment's segment, pass 2 enumerates again and writes the keys in place.
#define ENV_BP_CTA_THREADS 256
template <bool HULLS, bool LOCAL>
__global__ void __launch_bounds__(ENV_BP_CTA_THREADS) k_env_bp_cta(const EnvBpArgs A) {
  extern __shared__ float4 envBpSmem[];
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, W = ENV_BP_CTA_THREADS / 32;
  const uint32_t e = blockIdx.x;
  float4* sMin = envBpSmem; float4* sMax = sMin + A.maxList;
  uint32_t* sAct = reinterpret_cast<uint32_t*>(sMax + A.maxList); uint32_t* sRow = sAct + A.maxList;   // sRow: hits per row, then the row's offset
  __shared__ uint32_t sBase, sCnt;
  const uint32_t ls = A.envStart[e], n = A.envStart[e + 1] - ls;
  for (uint32_t k = tid; k < n; k += ENV_BP_CTA_THREADS) {   // a1 / a2 as in k_env_bp
    const uint32_t a = A.envList[ls + k]; const uint32_t gf = A.geomFlags[a], env = A.envId[a];
    float mn[3], mx[3];
    const bool own = env == e || e == 0;
    xf shape; shape.p = V3(0, 0, 0); shape.q = Q4(0, 0, 0, 1);
    if (LOCAL) shape = shape_world_pose(A.L, a, A.pos[a], A.quat[a], own);
    else if (!A.externalTight) { const float4 p4 = A.pos[a]; shape.p = V3(p4.x, p4.y, p4.z); shape.q = Q4(A.quat[a]); }
    if (A.externalTight) { for (int c = 0; c < 3; ++c) { mn[c] = A.tight[a * 6 + c]; mx[c] = A.tight[a * 6 + 3 + c]; } }
    else {
      tight_bounds(gf & 0xff, shape.p, shape.q, A.dims[a], mn, mx, HULLS ? &A.hulls : nullptr);
      if (own) for (int c = 0; c < 3; ++c) { A.tight[a * 6 + c] = mn[c]; A.tight[a * 6 + 3 + c] = mx[c]; }
    }
    const float co = (LOCAL && A.shapeOff) ? A.shapeOff[a].x : A.contactOffset;   // every bound is inflated by its own shape's contact offset
    sMin[k] = make_float4(mn[0] - co, mn[1] - co, mn[2] - co, __uint_as_float(LOCAL ? env_agg_key(A.aggId, a, k) : env));
    sMax[k] = make_float4(mx[0] + co, mx[1] + co, mx[2] + co, __uint_as_float(gf_dynamic(gf) ? gf : (gf & ~0x100u)));   // kinematic bodies pair with dynamic ones only, like statics (the test below reads the dynamic bit)
    sAct[k] = a;
  }
  __syncthreads();
  for (uint32_t i = warp; i < n; i += W) {   // pass 1: hits per row (the last row has none)
    uint32_t c = 0;
    if (i + 1 < n) {
      const float4 amin = sMin[i], amax = sMax[i];
      for (uint32_t j0 = i + 1; j0 < n; j0 += 32) { const uint32_t j = j0 + lane; c += __popc(__ballot_sync(0xffffffffu, j < n && env_bp_test<LOCAL>(amin, amax, sMin[j], sMax[j]))); }
    }
    if (lane == 0) sRow[i] = c;
  }
  __syncthreads();
  if (warp == 0) {   // exclusive scan of the row counts: every lane takes a run of consecutive rows
    const uint32_t per = (n + 31) / 32, r0 = lane * per, r1 = min(n, r0 + per);
    uint32_t sum = 0; for (uint32_t i = r0; i < r1; ++i) sum += sRow[i];
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
    uint32_t run = incl - sum;
    for (uint32_t i = r0; i < r1; ++i) { const uint32_t c = sRow[i]; sRow[i] = run; run += c; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) { sCnt = total; sBase = total ? atomicAdd(&A.counters[C_NPAIRS_NEW], total) : 0u; }
  }
  __syncthreads();
  const uint32_t base = sBase; uint32_t cnt = sCnt;
  if (base + cnt > A.cap) {   // capacity exceeded: report, keep the flat list crash-free (sentinel keys), drop the segment
    if (tid == 0) atomicOr(&A.counters[C_ERROR], (uint32_t)E_PAIR_OVERFLOW);
    for (uint32_t t = base + tid; t < min(base + cnt, A.cap); t += ENV_BP_CTA_THREADS) { A.newKeys[t] = ~0ull; A.newSlots[t] = 0; }
    cnt = 0;
  } else {
    for (uint32_t i = warp; i + 1 < n; i += W) {   // pass 2: the same enumeration, keys written at their final position
      const float4 amin = sMin[i], amax = sMax[i]; const uint64_t hiKey = (uint64_t)sAct[i] << A.bitsA;
      uint32_t w = sRow[i];
      for (uint32_t j0 = i + 1; j0 < n; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool hit = j < n && env_bp_test<LOCAL>(amin, amax, sMin[j], sMax[j]);
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (hit) A.newKeys[base + w + __popc(m & ((1u << lane) - 1u))] = hiKey | sAct[j];
        w += __popc(m);
      }
    }
  }
  __syncthreads();   // the segment's keys are read back by other threads of the CTA below
  const bool same = env_bp_lifecycle(A, e, base, cnt, tid, ENV_BP_CTA_THREADS);
  const int allSame = __syncthreads_and(same ? 1 : 0);
  if (tid == 0) A.newSeg[e] = make_uint2(base, cnt | (allSame ? 0x80000000u : 0u));
}

// ---------------------------------------------------------------------------------------------
struct EnvSolveArgs {
  uint32_t nEnv, maxList, conCap, cap, posIters, velIters; float dt, gx, gy, gz; SolverParams P;
  const uint32_t *envStart, *envList, *actorLocal; const uint2* seg;
  float4 *pos, *quat, *linVel, *angVel; const float4 *invInertia, *damp; const uint32_t* geomFlags;
  const uint32_t* pairSlots; const uint2* pairBodies; const float4 *cHdr, *cPts; float* cForce; float4* frictions; float4* frReport;
  uint32_t *conPair, *conB0, *conB1, *conColour, *ordered, *broken;   // per-pair-index scratch (global, L2 resident)
  uint32_t* slotColour;   // per persistent pair slot: partition of the pair's constraint last frame (NONE32 = no contacts)
  float4* rowScratch;   // 25 x cap float4, field-major: memory image of RegRows for environments with more constraints than threads
  uint32_t* counters; unsigned long long* timing; SleepArgs S;
  uint32_t anyLocks;   // some actor carries PxRigidDynamicLockFlags (uniform fast path otherwise)
  float4 *extForce, *extTorque;   // pending eFORCE / eTORQUE writes (NULL until the application uses them)
  MaterialArgs M;   // material table (matTab NULL: the scene's single material in P)
  float4* kinFtv; uint32_t kinOn;   // scenes with kinematic bodies (kinOn; EXT instantiation): TGS friction target velocities per pair index (NULL with PGS)
  const ExportTable* exportTab; const uint2* envDyn; const uint32_t* dynActor;   // fused state export: targets, per environment {first dynamic-body index, count}
};
#ifdef PXB_ENV_TIMING
#define ENV_T(i) do { __syncthreads(); if (threadIdx.x == 0) { const long long c_ = clock64(); A.timing[(size_t)blockIdx.x * 16 + (i)] = (unsigned long long)(c_ - t_prev); t_prev = c_; } } while (0)
#else
#define ENV_T(i) do { } while (0)
#endif

// Solver rows of ONE constraint in the environment path's compact layout (25 float4 = 400 B; the device-wide path's rows
// are 31 float4): contact header, per point (raXnI, velMultiplier | rbXnI, separation), the per-point scalars packed by
// field, the two friction directions shared by both anchors, per friction row (raXnI, error | rbXnI, velMultiplier) and the
// accumulated impulses.  A thread that owns a constraint for the whole solve keeps this record in REGISTERS.
struct RegRows {
  float4 h0;            // normal.xyz, maxPenBias
  float4 h1;            // invMass0*dom0, invMass1*dom1, staticFriction, dynamicFriction
  uint4 h2;             // body0, body1 (NONE32 = static), numNormal | numFriction << 8, pair index
  float4 pa[4], pb[4];  // raXnI.xyz, velMultiplier (= recipResponse) | rbXnI.xyz, separation
  float4 pc0, pc1, ap;  // biasCoefficient[4], targetVelocity[4], appliedForce[4]
  float4 t0, t1;        // friction direction 0 .xyz, frictionScale | direction 1 .xyz, biasScale
  float4 fa[4], fb[4];  // raXnI.xyz, error | rbXnI.xyz, velMultiplier   (row = anchor*2 + direction)
  float4 fap;           // applied friction impulses[4]
  uint32_t broken;
};
PXB_D float f4get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
PXB_D void f4set(float4& v, int j, float x) { if (j == 0) v.x = x; else if (j == 1) v.y = x; else if (j == 2) v.z = x; else v.w = x; }

// a14 into registers: same arithmetic as prep_constraint (createFinalizeSolverContactsStep, DyTGSContactPrep.cpp:1297-1490;
// friction correlation DyFrictionCorrelation.cpp:56-330).
__device__ __forceinline__ void prep_constraint_regs(RegRows& r, uint32_t i, uint32_t b0, uint32_t b1, const PrepBodies& B, const float4* __restrict__ cHdr,
                                                     const float4* __restrict__ cPts, float4* __restrict__ frec, const SolverParams& P, const bool noFriction = false,
                                                     const bool kin1 = false, float4* __restrict__ kinFtvOut = nullptr) {   // kin1: body B is kinematic (device-wide path only)
  Contacts con; const float4 h = cHdr[i]; con.normal = V3(h.x, h.y, h.z); con.count = __float_as_int(h.w);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float4 p = cPts[(size_t)i * 4 + j]; con.point[j] = V3(p.x, p.y, p.z); con.sep[j] = p.w; }
  const xf& f0 = B.f0; const xf& f1 = B.f1;
  FrictionPatch fp; friction_load(fp, frec);
  friction_correlate(fp, con, f0, f1, P.staticFriction, P.dynamicFriction, P.restitution, P.correlationDistance, P.frictionOffsetThreshold + P.restDistance);
  if (noFriction) fp.anchorCount = 0;   // PxMaterialFlag::eDISABLE_FRICTION: no friction rows (haveFriction = !disableStrongFriction && anchorCount != 0)
  friction_store(fp, frec);
  const float maxPenBias = fmax_(B.pen0, B.pen1);
  const v3 linVel0 = B.linVel0, linVel1 = B.linVel1, angVel0 = B.angVel0, angVel1 = B.angVel1;
  const m33& sI0 = B.sI0; const m33& sI1 = B.sI1;
  const float invMass0_dom0 = 1.f * B.invMass0, invMass1_dom1 = (-1.f) * B.invMass1;
  const float scale = fmin_(0.8f, P.biasCoefficient);
  const float invDtp8 = P.invStepDt * scale, frictionBiasScale = P.invStepDt * scale;
  const v3 normal = con.normal;
  const float normalLenSq = adot(normal, normal);
  const float norVel0 = adot(linVel0, normal), norVel1 = adot(linVel1, normal);
  const float imn0 = invMass0_dom0 * normalLenSq, imn1 = invMass1_dom1 * normalLenSq;
  const bool haveFriction = fp.anchorCount != 0;
  const uint32_t numFriction = haveFriction ? (uint32_t)fp.anchorCount * 2u : 0u;
  r.h0 = F4(normal, maxPenBias);
  r.h1 = make_float4(invMass0_dom0, -invMass1_dom1, P.staticFriction, P.dynamicFriction);
  r.h2 = make_uint4(b0, b1, (uint32_t)con.count | (numFriction << 8) | (kin1 ? 0x10000u : 0u), i);   // bit 16: friction target velocities in the side table (kinematic body B)
  r.pc0 = r.pc1 = r.ap = r.fap = make_float4(0, 0, 0, 0); r.t0 = r.t1 = make_float4(0, 0, 0, 0); r.broken = 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.pa[j] = r.pb[j] = make_float4(0, 0, 0, 0);
    if (j < con.count) {
      SPoint s; prep_point(s, con.point[j], con.sep[j], normal, f0.p, f1.p, sI0, sI1, angVel0, angVel1, norVel0, norVel1, imn0, imn1, P, invDtp8, kin1);
      r.pa[j] = F4(s.raXnI, s.velMultiplier); r.pb[j] = F4(s.rbXnI, s.separation);
      f4set(r.pc0, j, s.biasCoefficient); f4set(r.pc1, j, s.targetVelocity);
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
    const v3 t1 = anormalize(cross(normal, t0));
    const v3 relTr = f0.p - f1.p;
    const float frictionScale = (fp.anchorCount == 2) ? 0.5f : 1.f;
    r.t0 = F4(t0, frictionScale); r.t1 = F4(t1, frictionBiasScale);
    float4 ftv = make_float4(0, 0, 0, 0);   // friction target velocities (row = anchor * 2 + direction): the kinematic body's velocity along the tangent at the anchor (DyTGSContactPrep.cpp:724-727, :761-764)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j < fp.anchorCount) {
        const v3 ra = aqrot(f0.q, fp.body0Anchors[j]), rb = aqrot(f1.q, fp.body1Anchors[j]);
        const v3 error = (ra - rb) + relTr;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          SFriction f; prep_friction_row(f, ra, rb, error, t == 0 ? t0 : t1, sI0, sI1, imn0, imn1, scale, frictionScale, frictionBiasScale);
          r.fa[j * 2 + t] = F4(f.raXnI, f.error); r.fb[j * 2 + t] = F4(f.rbXnI, f.velMultiplier);
          if (kin1) { const v3 td = t == 0 ? t0 : t1; f4set(ftv, j * 2 + t, 0.f + (adot(linVel1, td) + adot(cross(rb, td), angVel1))); }
        }
      }
    }
    if (kin1 && kinFtvOut) *kinFtvOut = ftv;
  } else if (kin1 && kinFtvOut) *kinFtvOut = make_float4(0, 0, 0, 0);
}

// a15 on a register-resident record: same arithmetic and operation order as solve_constraint (solveContact,
// DyTGSContactPrep.cpp:1581-1873).  Friction rows have targetVel == 0 (prep_friction_row), so `x - 0*t` and `bias - 0` of
// the reference are the identity and are dropped.
// FR: the friction half of the record (t0, t1, fa, fb, fap [, pc1 for PGS]) lives in shared memory as fr[field * frStride] (this
// thread's column) instead of registers: it is only touched in the friction section, which keeps the register count at 128
// and lets 8 instead of 6 environments be resident per SM.
struct FrView { float4* p; uint32_t stride; };
PXB_D void fr_spill(const FrView& v, const RegRows& r) {
  v.p[0] = r.t0; v.p[v.stride] = r.t1; v.p[10 * v.stride] = r.fap; v.p[11 * v.stride] = r.pc1;
#pragma unroll
  for (int j = 0; j < 4; ++j) { v.p[(2 + j) * v.stride] = r.fa[j]; v.p[(6 + j) * v.stride] = r.fb[j]; }
}
// KIN (device-wide path, scenes with kinematic bodies): friction rows against a kinematic body carry target velocities, kept in a side table indexed by the pair
template <bool FR, bool KIN = false>
__device__ __forceinline__ void solve_constraint_regs(RegRows& r, const float minPen, const float elapsedTime, float4* bLin, float4* bAng, const float4* bDLin, const float4* bDAng, const FrView fr = FrView(),
                                                      const float4* __restrict__ kinFtv = nullptr) {
  const uint32_t b0 = r.h2.x, b1 = r.h2.y;
  const int numNormal = (int)(r.h2.z & 0xff), numFriction = (int)((r.h2.z >> 8) & 0xff);
  const v3 n = V3(r.h0.x, r.h0.y, r.h0.z); const float maxPenBias = r.h0.w;
  const float invMassA = r.h1.x, invMassB = r.h1.y;
  v3 linVel0 = V3(bLin[b0]), angState0 = V3(bAng[b0]);
  const v3 angMotion0 = V3(bDAng[b0]); v3 relMotion = V3(bDLin[b0]);
  v3 linVel1 = V3(0, 0, 0), angState1 = V3(0, 0, 0), angMotion1 = V3(0, 0, 0);
  if (b1 != NONE32) { linVel1 = V3(bLin[b1]); angState1 = V3(bAng[b1]); angMotion1 = V3(bDAng[b1]); relMotion = relMotion - V3(bDLin[b1]); }
  float accum = 0.f;
  {
    const v3 nim0 = n * invMassA, nim1 = n * invMassB;
    const float deltaV = adot(relMotion, n);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < numNormal) {
        const float4 A = r.pa[j], B = r.pb[j];
        const v3 raXnI = V3(A.x, A.y, A.z), rbXnI = V3(B.x, B.y, B.z);
        const float deltaAng = adot(angMotion0, raXnI) - adot(angMotion1, rbXnI);
        const float targetVel = f4get(r.pc1, j);
        const float deltaBias = (deltaV + deltaAng) - targetVel * elapsedTime;
        const float sep = fmax_(minPen, B.w + deltaBias);
        const float bias = fmin_(-maxPenBias, f4get(r.pc0, j) * sep);
        const v3 dv = (vmul(linVel0, n) + vmul(angState0, raXnI)) - (vmul(linVel1, n) + vmul(angState1, rbXnI));
        const float normalVel = (dv.x + dv.y) + dv.z;
        const float biasNV = bias * A.w;
        const float lambda = biasNV - (normalVel - targetVel) * A.w;
        const float applied = f4get(r.ap, j);
        const float dF_ = fmax_(lambda, -applied);
        const float newForce = fmin_(applied + dF_, FLT_MAX);
        const float deltaF = newForce - applied;
        linVel0 = scaleadd(nim0, deltaF, linVel0); linVel1 = negscalesub(nim1, deltaF, linVel1);
        angState0 = scaleadd(raXnI, deltaF * 1.f, angState0); angState1 = negscalesub(rbXnI, deltaF * 1.f, angState1);
        f4set(r.ap, j, newForce);
        accum = accum + newForce;
      }
    }
  }
  if (numFriction) {
    const float maxFrictionImpulse = r.h1.z * accum, maxDynFrictionImpulse = r.h1.w * accum;
    const float4 T0 = FR ? fr.p[0] : r.t0, T1 = FR ? fr.p[fr.stride] : r.t1; float4 fap = FR ? fr.p[10 * fr.stride] : r.fap;
    const float frictionScale = T0.w, biasScale = T1.w;
    const v3 normal0 = V3(T0.x, T0.y, T0.z), normal1 = V3(T1.x, T1.y, T1.z);
    bool broken = false;
    float4 ftv = make_float4(0, 0, 0, 0);
    if (KIN && (r.h2.z & 0x10000u)) ftv = kinFtv[r.h2.w];
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
      if (j < numFriction) {
        const float4 A0 = FR ? fr.p[(2 + j) * fr.stride] : r.fa[j], B0 = FR ? fr.p[(6 + j) * fr.stride] : r.fb[j];
        const float4 A1 = FR ? fr.p[(3 + j) * fr.stride] : r.fa[j + 1], B1 = FR ? fr.p[(7 + j) * fr.stride] : r.fb[j + 1];
        const v3 raXnI0 = V3(A0.x, A0.y, A0.z), rbXnI0 = V3(B0.x, B0.y, B0.z), raXnI1 = V3(A1.x, A1.y, A1.z), rbXnI1 = V3(B1.x, B1.y, B1.z);
        const float applied0 = f4get(fap, j), applied1 = f4get(fap, j + 1);
        float deltaV0 = (adot(raXnI0, angMotion0) - adot(rbXnI0, angMotion1)) + adot(normal0, relMotion);
        float deltaV1 = (adot(raXnI1, angMotion0) - adot(rbXnI1, angMotion1)) + adot(normal1, relMotion);
        const float tv0 = KIN ? f4get(ftv, j) : 0.f, tv1 = KIN ? f4get(ftv, j + 1) : 0.f;
        if (KIN) { deltaV0 = deltaV0 - tv0 * elapsedTime; deltaV1 = deltaV1 - tv1 * elapsedTime; }
        float bias0 = (A0.w + deltaV0) * biasScale, bias1 = (A1.w + deltaV1) * biasScale;
        const float vm0 = B0.w, vm1 = B1.w;
        const v3 d0 = (vmul(linVel0, normal0) + vmul(angState0, raXnI0)) - (vmul(linVel1, normal0) + vmul(angState1, rbXnI0));
        const v3 d1 = (vmul(linVel0, normal1) + vmul(angState0, raXnI1)) - (vmul(linVel1, normal1) + vmul(angState1, rbXnI1));
        const float normalVel0 = (d0.x + d0.y) + d0.z, normalVel1 = (d1.x + d1.y) + d1.z;
        if (KIN) { bias0 = bias0 - tv0; bias1 = bias1 - tv1; }
        const float tmp10 = applied0 - bias0 * vm0, tmp11 = applied1 - bias1 * vm1;
        const float total0 = tmp10 - normalVel0 * vm0, total1 = tmp11 - normalVel1 * vm1;
        const float total = sqrtf(total0 * total0 + total1 * total1);
        const bool clamp = total > (frictionScale * maxFrictionImpulse);
        const float totalClamped = clamp ? fmin_(frictionScale * maxDynFrictionImpulse, total) : total;
        const float ratio = (total > 0.f) ? (totalClamped / total) : 0.f;
        const float new0 = total0 * ratio, new1 = total1 * ratio;
        broken = broken || clamp;
        const float dF0 = new0 - applied0, dF1 = new1 - applied1;
        linVel0 = scaleadd(normal0 * invMassA, dF0, scaleadd(normal1 * invMassA, dF1, linVel0));
        linVel1 = negscalesub(normal0 * invMassB, dF0, negscalesub(normal1 * invMassB, dF1, linVel1));
        angState0 = scaleadd(raXnI0, dF0 * 1.f, scaleadd(raXnI1, dF1 * 1.f, angState0));
        angState1 = negscalesub(rbXnI0, dF0 * 1.f, negscalesub(rbXnI1, dF1 * 1.f, angState1));
        f4set(fap, j, new0); f4set(fap, j + 1, new1);
      }
    }
    if (FR) fr.p[10 * fr.stride] = fap; else r.fap = fap;
    r.broken = broken ? 1u : 0u;  // hdr->broken is overwritten by every solve call (Store_From_BoolV)
  }
  bLin[b0] = F4(linVel0, 0.f); bAng[b0] = F4(angState0, 0.f);
  if (b1 != NONE32) { bLin[b1] = F4(linVel1, 0.f); bAng[b1] = F4(angState1, 0.f); }
}

// memory image of RegRows for environments with more constraints than threads: 25 float4 + 1 u32 per constraint, field-major
struct Rows { float4* f; uint32_t* broken; uint32_t stride; };
PXB_D void rows_store(const Rows& R, uint32_t k, const RegRows& r) {
  const size_t s = R.stride; float4* f = R.f + k;
  f[0] = r.h0; f[s] = r.h1; f[2 * s] = make_float4(__uint_as_float(r.h2.x), __uint_as_float(r.h2.y), __uint_as_float(r.h2.z), __uint_as_float(r.h2.w));
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[(3 + j) * s] = r.pa[j]; f[(7 + j) * s] = r.pb[j]; f[(16 + j) * s] = r.fa[j]; f[(20 + j) * s] = r.fb[j]; }
  f[11 * s] = r.pc0; f[12 * s] = r.pc1; f[13 * s] = r.ap; f[14 * s] = r.t0; f[15 * s] = r.t1; f[24 * s] = r.fap; R.broken[k] = r.broken;
}
PXB_D void rows_load(const Rows& R, uint32_t k, RegRows& r) {
  const size_t s = R.stride; const float4* f = R.f + k;
  r.h0 = f[0]; r.h1 = f[s]; { const float4 c = f[2 * s]; r.h2 = make_uint4(__float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), __float_as_uint(c.w)); }
#pragma unroll
  for (int j = 0; j < 4; ++j) { r.pa[j] = f[(3 + j) * s]; r.pb[j] = f[(7 + j) * s]; r.fa[j] = f[(16 + j) * s]; r.fb[j] = f[(20 + j) * s]; }
  r.pc0 = f[11 * s]; r.pc1 = f[12 * s]; r.ap = f[13 * s]; r.t0 = f[14 * s]; r.t1 = f[15 * s]; r.fap = f[24 * s]; r.broken = R.broken[k];
}
PXB_D void rows_store_state(const Rows& R, uint32_t k, const RegRows& r) { const size_t s = R.stride; R.f[13 * s + k] = r.ap; R.f[24 * s + k] = r.fap; R.broken[k] = r.broken; }

#include "pxb_pgs.cuh"

struct ConLists { uint32_t *conPair, *b0, *b1, *colour, *ordered; };   // per-constraint scratch of one environment (shared memory, or global when it does not fit)

// vLin / vAng: per-body pre-solver (unconstrained) linear and world angular velocity
template <bool PGS, bool EXT>
__device__ __forceinline__ void env_prep_one(const EnvSolveArgs& A, const ConLists& L, uint32_t base, uint32_t pos, const float4* vLin, const float4* bIA, const float4* bIB, const float4* vAng, RegRows& r) {
  const uint32_t k = L.ordered[pos]; const uint32_t i = base + L.conPair[k];
  const uint32_t l0 = L.b0[k], l1 = L.b1[k];
  const uint2 bb = A.pairBodies[i];
  PrepBodies B;
  { const float4 p = A.pos[bb.x]; B.f0.p = V3(p.x, p.y, p.z); B.f0.q = Q4(A.quat[bb.x]); B.invMass0 = p.w; const float4 q = A.pos[bb.y]; B.f1.p = V3(q.x, q.y, q.z); B.f1.q = Q4(A.quat[bb.y]); }
  const bool dyn1 = l1 != NONE32;
  B.invMass1 = dyn1 ? A.pos[bb.y].w : 0.f;
  B.pen0 = -A.invInertia[bb.x].w; B.pen1 = dyn1 ? -A.invInertia[bb.y].w : -FLT_MAX;
  B.linVel0 = V3(vLin[l0]); B.angVel0 = V3(vAng[l0]); B.sI0 = load_sym(bIA[l0], bIB[l0]);
  if (dyn1) { B.linVel1 = V3(vLin[l1]); B.angVel1 = V3(vAng[l1]); B.sI1 = load_sym(bIA[l1], bIB[l1]); }
  else { B.linVel1 = V3(0, 0, 0); B.angVel1 = V3(0, 0, 0); B.sI1.c0 = B.sI1.c1 = B.sI1.c2 = V3(0, 0, 0); }
  if (EXT && (A.M.matTab || A.M.shapeOff || A.kinOn)) {   // material table / per-shape rest offsets / kinematic bodies: this pair's own parameters (the plain instantiation carries none of this)
    SolverParams Pm = A.P; const bool noFriction = A.M.matTab ? pair_material(A.M, bb.x, bb.y, Pm) : false;
    if (A.M.shapeOff) Pm.restDistance = A.M.shapeOff[bb.x].y + A.M.shapeOff[bb.y].y;
    const bool kin1 = A.kinOn && (A.geomFlags[bb.y] & 0x800u);   // kinematic body B (see k_prep_rows)
    if (kin1) { B.pen1 = -A.invInertia[bb.y].w; B.linVel1 = V3(A.linVel[bb.y]); B.angVel1 = V3(A.angVel[bb.y]); }
    if (PGS) prep_constraint_pgs(r, i, l0, l1, B, A.cHdr, A.cPts, A.frictions + (size_t)A.pairSlots[i] * PXB_FRICTION_F4, Pm, noFriction);
    else prep_constraint_regs(r, i, l0, l1, B, A.cHdr, A.cPts, A.frictions + (size_t)A.pairSlots[i] * PXB_FRICTION_F4, Pm, noFriction, kin1, (kin1 && A.kinFtv) ? A.kinFtv + i : nullptr);
    return;
  }
  if (PGS) prep_constraint_pgs(r, i, l0, l1, B, A.cHdr, A.cPts, A.frictions + (size_t)A.pairSlots[i] * PXB_FRICTION_F4, A.P);
  else prep_constraint_regs(r, i, l0, l1, B, A.cHdr, A.cPts, A.frictions + (size_t)A.pairSlots[i] * PXB_FRICTION_F4, A.P);
}
// a17: writeBackContact (DyTGSContactPrep.cpp:1875-1937)
__device__ __forceinline__ void env_writeback_one(const EnvSolveArgs& A, const RegRows& r, const FrView fr = FrView()) {
  const uint32_t i = r.h2.w; const int numNormal = (int)(r.h2.z & 0xff), numFriction = (int)((r.h2.z >> 8) & 0xff);
#pragma unroll
  for (int j = 0; j < 4; ++j) if (j < numNormal) A.cForce[(size_t)i * 4 + j] = f4get(r.ap, j);
  if (numFriction && r.broken) A.frictions[(size_t)A.pairSlots[i] * PXB_FRICTION_F4 + 1].w = __int_as_float(1);
  if (A.frReport) {   // contact reports on: friction impulses + world anchors (the body poses in global memory are still the start-of-step ones here)
    const uint32_t a0 = A.pairBodies[i].x;
    friction_report_store(A.frReport, i, numFriction, fr.p ? fr.p[0] : r.t0, fr.p ? fr.p[fr.stride] : r.t1, fr.p ? fr.p[10 * fr.stride] : r.fap,
                          A.frictions + (size_t)A.pairSlots[i] * PXB_FRICTION_F4, A.pos[a0], A.quat[a0]);
  }
}
template <int T, bool EXT>
__device__ __forceinline__ void env_integrate_substep(uint32_t n, float stepDt, float4* bLin, float4* bAng, float4* bDLin, float4* bDAng, const float4* bIA, const float4* bIB, float4* bP, float4* bQ) {
  for (uint32_t b = threadIdx.x; b < n; b += T) {
    if (!__float_as_uint(bP[b].w)) continue;
    v3 p = V3(bP[b]); q4 dq = Q4(bQ[b]); v3 dl = V3(bDLin[b]), da = V3(bDAng[b]);
    const float4 ib = bIB[b]; const uint32_t lock = EXT ? __float_as_uint(ib.z) : 0u;   // PxRigidDynamicLockFlags only in the EXT instantiation
    v3 lv = V3(bLin[b]), as = V3(bAng[b]);
    integrate_core_step(lv, as, load_sym(bIA[b], ib), stepDt, p, dq, dl, da, lock);
    if (EXT && lock) { bLin[b] = F4(lv, 0.f); bAng[b] = F4(as, 0.f); }
    bP[b] = F4(p, __uint_as_float(1u)); bQ[b] = F4(dq); bDLin[b] = F4(dl, 0.f); bDAng[b] = F4(da, 0.f);
  }
}

// prep + all TGS iterations + write-back of one environment (iterativeSolveIsland, DyTGSDynamics.cpp:2515-2793, with CTA
// barriers between partitions).  REG: every thread owns exactly one constraint (nCon <= T) and keeps its rows in registers
// for the whole solve; otherwise rows stream through the memory image R (global scratch, L2 resident).
template <int T, bool REG, bool PGS, bool EXT>
__device__ __forceinline__ void env_solve_body(const EnvSolveArgs& A, const Rows R, const ConLists L, const uint32_t base, const uint32_t nCon, const uint32_t n,
                                               float4* bLin, float4* bAng, float4* bDLin, float4* bDAng, float4* bIA, float4* bIB, float4* bP, float4* bQ,
                                               const uint32_t* sPartStart, const uint32_t nPart, float4* sFr, long long& t_prev) {
  const uint32_t tid = threadIdx.x;
  RegRows mine;
  const float4* vLin = PGS ? bDLin : bLin; const float4* vAng = PGS ? bDAng : bQ;
  const FrView fr = {sFr + tid, T};
  if (REG) { if (tid < nCon) { env_prep_one<PGS, EXT>(A, L, base, tid, vLin, bIA, bIB, vAng, mine); fr_spill(fr, mine); } }
  else for (uint32_t pos = tid; pos < nCon; pos += T) { RegRows r; env_prep_one<PGS, EXT>(A, L, base, pos, vLin, bIA, bIB, vAng, r); rows_store(R, pos, r); }
  __syncthreads();
  ENV_T(4);
  if (PGS) {
    // solveV_Blocks (DySolverControl.cpp:163-405): position iterations (friction in the last three, the last one concludes),
    // saveMotionVelocities, then max(velIters, 1) velocity iterations; bP / bQ keep the motion velocities for integrateCore
    for (uint32_t it = A.posIters; it > 0; --it) {
      const bool doFriction = it <= 3;
      for (uint32_t p = 0; p < nPart; ++p) {
        const uint32_t pb = sPartStart[p], pe = sPartStart[p + 1];
        if (REG) { if (tid >= pb && tid < pe) { solve_constraint_pgs<true>(mine, doFriction, bLin, bAng, fr); if (it == 1) conclude_constraint_pgs<true>(mine, fr); } }
        else for (uint32_t k = pb + tid; k < pe; k += T) { RegRows r; rows_load(R, k, r); solve_constraint_pgs<false>(r, doFriction, bLin, bAng); if (it == 1) { conclude_constraint_pgs<false>(r); rows_store(R, k, r); } else rows_store_state(R, k, r); }
        __syncthreads();
      }
    }
    for (uint32_t b = tid; b < n; b += T) { bP[b] = bLin[b]; bQ[b] = bAng[b]; }
    __syncthreads();
    const uint32_t velIters = A.velIters ? A.velIters : 1u;
    for (uint32_t it = 0; it < velIters; ++it)
      for (uint32_t p = 0; p < nPart; ++p) {
        const uint32_t pb = sPartStart[p], pe = sPartStart[p + 1];
        if (REG) { if (tid >= pb && tid < pe) solve_constraint_pgs<true>(mine, true, bLin, bAng, fr); }
        else for (uint32_t k = pb + tid; k < pe; k += T) { RegRows r; rows_load(R, k, r); solve_constraint_pgs<false>(r, true, bLin, bAng); rows_store_state(R, k, r); }
        __syncthreads();
      }
  } else {
    for (uint32_t b = tid; b < n; b += T) bQ[b] = make_float4(0, 0, 0, 1);   // bQ carried the unconstrained angular velocity during prep
    __syncthreads();
    const float stepDt = A.P.stepDt;
    float elapsed = 0.f;
    for (uint32_t it = 0; it < A.posIters + A.velIters; ++it) {
      const bool vel = it >= A.posIters;
      const float minPen = vel ? 0.f : -FLT_MAX;
      for (uint32_t p = 0; p < nPart; ++p) {
        const uint32_t pb = sPartStart[p], pe = sPartStart[p + 1];
        if (REG) { if (tid >= pb && tid < pe) solve_constraint_regs<true, EXT>(mine, minPen, elapsed, bLin, bAng, bDLin, bDAng, fr, A.kinFtv); }   // EXT: friction target velocities of rows against kinematic bodies
        else for (uint32_t k = pb + tid; k < pe; k += T) { RegRows r; rows_load(R, k, r); solve_constraint_regs<false, EXT>(r, minPen, elapsed, bLin, bAng, bDLin, bDAng, FrView(), A.kinFtv); rows_store_state(R, k, r); }
        __syncthreads();
      }
      if (!vel) {
        env_integrate_substep<T, EXT>(n, stepDt, bLin, bAng, bDLin, bDAng, bIA, bIB, bP, bQ);
        elapsed += stepDt;
        __syncthreads();
      }
    }
  }
  ENV_T(5);
  if (REG) { if (tid < nCon) env_writeback_one(A, mine, fr); }
  else for (uint32_t pos = tid; pos < nCon; pos += T) { RegRows r; rows_load(R, pos, r); env_writeback_one(A, r); }
}

// dynamic shared memory of k_env_solve (host mirror: env_solve_smem in pxb_engine.cu):
//   8 x maxList float4 body state | u32: 4 x maxList (64-bit colour masks, first-in-line, static counts), 5 x conCap constraint lists (padded to 16 B) | 12 x T float4 friction rows
// conCap = list capacity (pairs of the environment); environments with more pairs keep their lists in global scratch.
#ifndef PXB_ENV_CTAS64
#define PXB_ENV_CTAS64 8
#endif
template <int T, bool PGS, bool EXT>   // EXT: scenes that use PxRigidDynamicLockFlags or eFORCE / eTORQUE writes (the plain instantiation carries none of that code)
// (measured and not kept: a register cap of 144 = 7 CTAs/SM instead of the residency target -- solve 0.2006 vs 0.1937 ms on config 2, tools/gpu_run52.sh; the 320 B of stack are the
//  dynamically indexed row arrays, not spills, so more registers buy nothing)
__global__ void __launch_bounds__(T, (T <= 64 ? PXB_ENV_CTAS64 : (T <= 128 ? PXB_ENV_CTAS64 / 2 : 1))) k_env_solve(const EnvSolveArgs A) {
  extern __shared__ float4 envSmem[];
  __shared__ uint32_t sPartCnt[MAX_PARTITIONS + 1], sPartStart[MAX_PARTITIONS + 1], sWarp[T / 32 + 1], sMisc[4];
  const uint32_t e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ls = A.envStart[e], n = A.envStart[e + 1] - ls; const uint32_t* list = A.envList + ls;
  const uint2 sg = A.seg[e]; const uint32_t base = sg.x, m = sg.y & 0x7fffffffu; const bool sameSeg = (sg.y >> 31) != 0;
  const uint32_t nb = A.maxList, lc = A.conCap;
  float4 *bLin = envSmem, *bAng = bLin + nb, *bDLin = bAng + nb, *bDAng = bDLin + nb, *bIA = bDAng + nb, *bIB = bIA + nb, *bP = bIB + nb, *bQ = bP + nb;
  unsigned long long* bMask = reinterpret_cast<unsigned long long*>(bQ + nb);   // the 64 dynamic colours a body's constraints hold (as on the device-wide path)
  uint32_t* bFirst = reinterpret_cast<uint32_t*>(bMask + nb); uint32_t* bStat = bFirst + nb;
  uint32_t* sLists = bStat + nb;
  float4* sFr = reinterpret_cast<float4*>(sLists + 5 * lc + ((4 - ((4 * nb + 5 * lc) & 3)) & 3));   // 12 x T float4: friction half of the register rows
  ConLists L;
  if (m <= lc) { L.conPair = sLists; L.b0 = sLists + lc; L.b1 = sLists + 2 * lc; L.colour = sLists + 3 * lc; L.ordered = sLists + 4 * lc; }
  else { L.conPair = A.conPair + base; L.b0 = A.conB0 + base; L.b1 = A.conB1 + base; L.colour = A.conColour + base; L.ordered = A.ordered + base; }
  long long t_prev = clock64();
  // a12: preIntegrateBodies (DyTGSDynamics.cpp:992-1021) into shared memory
  for (uint32_t b = tid; b < n; b += T) {
    const uint32_t a = list[b];
    bMask[b] = 0; bStat[b] = 0;
    const uint32_t gf = A.geomFlags[a];
    if (!gf_dynamic(gf) || body_asleep(A.S, a)) { bP[b] = make_float4(0, 0, 0, 0); continue; }   // statics, kinematic and sleeping bodies take no part
    const float4 dm = A.damp[a]; const float4 ii = A.invInertia[a]; const float4 p4 = A.pos[a];
    v3 lv = V3(A.linVel[a]), av = V3(A.angVel[a]);
    if (EXT && A.extForce) {   // pending eFORCE / eTORQUE writes: consumed by this step
      const float4 F = A.extForce[a], Tq = A.extTorque[a];
      if (F.x != 0.f || F.y != 0.f || F.z != 0.f || Tq.x != 0.f || Tq.y != 0.f || Tq.z != 0.f) {
        apply_external_force(V3(F), V3(Tq), p4.w, ii, Q4(A.quat[a]), A.dt, lv, av);
        A.extForce[a] = make_float4(0, 0, 0, 0); A.extTorque[a] = make_float4(0, 0, 0, 0);
      }
    }
    unconstrained_velocity((EXT && (gf & 0x1000u)) ? V3(0, 0, 0) : V3(A.gx, A.gy, A.gz), A.dt, dm.x, dm.y, dm.z, dm.w, lv, av);   // eDISABLE_GRAVITY
    if (EXT && (gf & 0x2000u)) av = gyroscopic(av, V3(ii.x, ii.y, ii.z), Q4(A.quat[a]), A.dt);   // eENABLE_GYROSCOPIC_FORCES
    const uint32_t lock = EXT ? (gf >> 16) & 0x3fu : 0u;   // PxRigidDynamicLockFlags: TGS locks both velocities, PGS only the angular one (see k_preintegrate)
    if (EXT && lock) { if (!PGS) lv = lock3(lv, lock & 7u); av = lock3(av, (lock >> 3) & 7u); }
    const m33 rot = amfromq(Q4(A.quat[a]));
    const v3 sqrtInvI = V3(ii.x == 0.f ? 0.f : sqrtf(ii.x), ii.y == 0.f ? 0.f : sqrtf(ii.y), ii.z == 0.f ? 0.f : sqrtf(ii.z));
    const v3 sqrtI = V3(sqrtInvI.x == 0.f ? 0.f : 1.0f / sqrtInvI.x, sqrtInvI.y == 0.f ? 0.f : 1.0f / sqrtInvI.y, sqrtInvI.z == 0.f ? 0.f : 1.0f / sqrtInvI.z);
    m33 sI, sInertia; transform_inertia(sqrtInvI, rot, sI); transform_inertia(sqrtI, rot, sInertia);
    if (PGS) {   // solver bodies hold velocity deltas (start at zero); bDLin / bDAng hold the pre-solver velocities (PxSolverBodyData)
      bLin[b] = make_float4(0, 0, 0, 0); bAng[b] = make_float4(0, 0, 0, 0); bDLin[b] = F4(lv, 0.f); bDAng[b] = F4(av, 0.f);
    } else {
      bLin[b] = F4(lv, 0.f); bAng[b] = F4(mmul(sInertia, av), 0.f); bDLin[b] = make_float4(0, 0, 0, 0); bDAng[b] = make_float4(0, 0, 0, 0);
    }
    bIA[b] = make_float4(sI.c0.x, sI.c0.y, sI.c0.z, sI.c1.y); bIB[b] = make_float4(sI.c1.z, sI.c2.z, __uint_as_float(lock), 0.f);
    bP[b] = make_float4(p4.x, p4.y, p4.z, __uint_as_float(0u)); bQ[b] = F4(av, 0.f);
  }
  for (uint32_t p = tid; p < MAX_PARTITIONS + 1; p += T) sPartCnt[p] = 0;
  if (tid == 0) { sMisc[0] = 0; sMisc[1] = 0; }
  __syncthreads();
  ENV_T(0);
  // constraint list = this environment's pairs that produced contacts, in ascending key order (ordered compaction)
  // Steady state: when the pair segment and the contact/no-contact state of every pair are unchanged, the ordered constraint
  // list is last frame's, so last frame's partitions (kept per persistent pair slot) are the first-fit result again.
  uint32_t nCon = 0; int stale = sameSeg ? 0 : 1;
  for (uint32_t t0 = 0; t0 < m; t0 += T) {
    const uint32_t t = t0 + tid; bool f = false; uint32_t prevCol = NONE32;
    if (t < m) {
      const bool touching = __float_as_int(A.cHdr[base + t].w) > 0;
      f = touching;
      if (f && A.S.threshold > 0.f) { const uint2 bb = A.pairBodies[base + t]; if (A.S.asleep[bb.x] && (!(A.geomFlags[bb.y] & 0x100u) || A.S.asleep[bb.y])) f = false; }   // sleeping islands are not solved
      const uint32_t slot = A.pairSlots[base + t];
#ifndef PXB_NO_PREFETCH
      if (touching) { prefetch_l2(A.frictions + (size_t)slot * PXB_FRICTION_F4); prefetch_l2(A.cPts + (size_t)(base + t) * 4); }   // prep reads them after the colouring
#endif
      prevCol = A.slotColour[slot];
      if (f != (prevCol != NONE32)) stale = 1;
      if (!touching) A.frictions[(size_t)slot * PXB_FRICTION_F4 + 2].w = __int_as_float(0);   // no contacts: the friction patch is dropped (a sleeping pair keeps its patch)
      if (!f && prevCol != NONE32) A.slotColour[slot] = NONE32;
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
      const uint32_t l0 = A.actorLocal[bb.x]; const uint32_t l1 = gf_dynamic(A.geomFlags[bb.y]) ? A.actorLocal[bb.y] : NONE32;
      L.conPair[k] = t; L.b0[k] = l0; L.b1[k] = l1; L.colour[k] = prevCol;
      bP[l0].w = __uint_as_float(1u); if (l1 != NONE32) bP[l1].w = __uint_as_float(1u);   // hasConstraints (benign same-value races)
      if (l1 == NONE32) atomicAdd(&bStat[l0], 1u);
    }
    nCon += tot;
    __syncthreads();
  }
  const bool recolour = __syncthreads_or(stale) != 0;
  ENV_T(1);
  if (recolour) {
  for (uint32_t k = tid; k < nCon; k += T) L.colour[k] = NONE32;
  __syncthreads();
  // a13: first-fit colouring in constraint order (classifyConstraintDesc, DyConstraintPartition.cpp:475-568).  A constraint
  // is ready once it is the lowest-numbered uncoloured constraint on both of its bodies; ready constraints are body-disjoint.
  for (;;) {
    for (uint32_t b = tid; b < n; b += T) bFirst[b] = NONE32;
    __syncthreads();
    int undone = 0;
    for (uint32_t k = tid; k < nCon; k += T) {
      const uint32_t l1 = L.b1[k];
      if (l1 == NONE32 || L.colour[k] != NONE32) continue;
      atomicMin(&bFirst[L.b0[k]], k); atomicMin(&bFirst[l1], k); undone = 1;
    }
    if (!__syncthreads_or(undone)) break;
    for (uint32_t k = tid; k < nCon; k += T) {
      const uint32_t l1 = L.b1[k];
      if (l1 == NONE32 || L.colour[k] != NONE32) continue;
      const uint32_t l0 = L.b0[k];
      if (bFirst[l0] != k || bFirst[l1] != k) continue;
      const unsigned long long ma = bMask[l0], mb = bMask[l1]; const unsigned long long comb = ~ma & ~mb;
      uint32_t col = 63;
      if (comb) col = __ffsll((long long)comb) - 1; else atomicOr(&A.counters[C_ERROR], (uint32_t)E_COLOUR_OVERFLOW);
      L.colour[k] = col; bMask[l0] = ma | (1ull << col); bMask[l1] = mb | (1ull << col);
    }
    __syncthreads();
  }
  ENV_T(2);
  // static contacts of a body go to partitions maxDynamicColour(body) + rank among the body's static contacts (:203-262)
  for (uint32_t k = tid; k < nCon; k += T) {
    uint32_t col;
    if (L.b1[k] == NONE32) {
      const uint32_t l0 = L.b0[k]; uint32_t rank = 0;
      if (bStat[l0] > 1) for (uint32_t kk = 0; kk < k; ++kk) if (L.b1[kk] == NONE32 && L.b0[kk] == l0) ++rank;
      const unsigned long long mk = bMask[l0]; col = (mk ? 64u - __clzll((long long)mk) : 0u) + rank;
      if (col >= MAX_PARTITIONS) { col = MAX_PARTITIONS - 1; atomicOr(&A.counters[C_ERROR], (uint32_t)E_PARTITION_OVERFLOW); }
      L.colour[k] = col;
    } else col = L.colour[k];
    A.slotColour[A.pairSlots[base + L.conPair[k]]] = col;
  }
  }   // recolour
  for (uint32_t k = tid; k < nCon; k += T) atomicAdd(&sPartCnt[L.colour[k]], 1u);
  __syncthreads();
  if (tid == 0) {
    uint32_t s = 0, np = 0;
    for (uint32_t p = 0; p < MAX_PARTITIONS; ++p) { const uint32_t c = sPartCnt[p]; sPartStart[p] = s; sPartCnt[p] = s; s += c; if (c) np = p + 1; if (s == nCon) { for (uint32_t q = p + 1; q <= MAX_PARTITIONS; ++q) sPartStart[q] = s; break; } }
    sMisc[0] = np;
    if (nCon) { atomicAdd(&A.counters[C_NCON], nCon); atomicMax(&A.counters[C_NPART], np); atomicMax(&A.counters[C_MAXCONENV], nCon); }
    if (m) atomicMax(&A.counters[C_MAXPAIRENV], m);
  }
  __syncthreads();
  const uint32_t nPart = sMisc[0];
  for (uint32_t k = tid; k < nCon; k += T) L.ordered[atomicAdd(&sPartCnt[L.colour[k]], 1u)] = k;
  __syncthreads();
  ENV_T(3);
  if (nCon) {
    Rows R; R.f = A.rowScratch + base; R.broken = A.broken + base; R.stride = A.cap;
    if (nCon <= T) env_solve_body<T, true, PGS, EXT>(A, R, L, base, nCon, n, bLin, bAng, bDLin, bDAng, bIA, bIB, bP, bQ, sPartStart, nPart, sFr, t_prev);
    else env_solve_body<T, false, PGS, EXT>(A, R, L, base, nCon, n, bLin, bAng, bDLin, bDAng, bIA, bIB, bP, bQ, sPartStart, nPart, sFr, t_prev);
  } else {
    for (uint32_t b = tid; b < n; b += T) { if (PGS) { bP[b] = make_float4(0, 0, 0, 0); bQ[b] = make_float4(0, 0, 0, 0); } else bQ[b] = make_float4(0, 0, 0, 1); }
  }
  __syncthreads();
  ENV_T(6);
  // a18: copyBackBodies (DyTGSDynamics.cpp:1549-1580); bodies without constraints take one full-dt step (:2573-2577)
  for (uint32_t b = tid; b < n; b += T) {
    const uint32_t a = list[b];
    if (!gf_dynamic(A.geomFlags[a]) || body_asleep(A.S, a)) continue;
    const float4 ib = bIB[b]; const uint32_t lock = EXT ? __float_as_uint(ib.z) : 0u;
    const m33 sI = load_sym(bIA[b], ib);
    if (PGS) {   // integrate (DyDynamics.cpp:1398-1423): every body, with or without constraints
      const float4 p4 = A.pos[a]; v3 p = V3(p4.x, p4.y, p4.z); q4 q = Q4(A.quat[a]); v3 lv = V3(bDLin[b]), av = V3(bDAng[b]);
      v3 motionLin, motionAng;
      integrate_core_pgs(p, q, lv, av, sI, V3(bP[b]), V3(bQ[b]), V3(bLin[b]), V3(bAng[b]), A.dt, lock, motionLin, motionAng);
      A.pos[a] = make_float4(p.x, p.y, p.z, p4.w); A.quat[a] = F4(q); A.linVel[a] = F4(lv, 0.f); A.angVel[a] = F4(av, 0.f);
      if (A.S.threshold > 0.f) sleep_check_dev(A.S, a, q, A.invInertia[a], p4.w, motionLin, motionAng);
      continue;
    }
    v3 p = V3(bP[b]); q4 dq = Q4(bQ[b]);
    v3 lv = V3(bLin[b]), as = V3(bAng[b]);
    v3 dl = V3(bDLin[b]), da = V3(bDAng[b]);
    if (!__float_as_uint(bP[b].w)) { dl = V3(0, 0, 0); da = V3(0, 0, 0); integrate_core_step(lv, as, sI, A.dt, p, dq, dl, da, lock); }
    const float invMass = A.pos[a].w;
    const q4 q = qnormalized(qmul(dq, Q4(A.quat[a])));
    A.pos[a] = make_float4(p.x, p.y, p.z, invMass); A.quat[a] = F4(q);
    A.linVel[a] = F4(lv, 0.f); A.angVel[a] = F4(mmul(sI, as), 0.f);
    if (A.S.threshold > 0.f) { const float invDt = 1.0f / A.dt; sleep_check_dev(A.S, a, q, A.invInertia[a], invMass, dl * invDt, mmul(sI, da * invDt)); }
  }
  // a19 fused into the step: this environment's packed state block goes straight into every registered export target (the learner's tensor
  // on every peer GPU over NVLink, or mapped pinned host memory), so no pack kernel and no copy follow the step.
  if (A.exportTab) {
    const uint32_t nT = A.exportTab->n;
    if (nT) {
      __syncthreads();   // the final states above were written by other threads of this CTA
      const uint2 dr = A.envDyn[e];
      export_packed_range(A.exportTab, nT, A.dynActor, dr.x, dr.y, A.pos, A.quat, A.linVel, A.angVel, tid, T);
    }
  }
  ENV_T(7);
}
