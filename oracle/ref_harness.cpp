// ref_harness: drives the UNMODIFIED reference CPU SDK (PhysX 5.6.1, built by oracle/ref_build.mk)
// through its public API on a scene file (oracle/scene_format.h).  Test infrastructure only:
// it is the parity oracle ("whole-scene oracle") and the CPU baseline
// (eABP broadphase, CPU TGS/PGS solver, PxDefaultCpuDispatcher with T threads).
//
//   ref_harness run <scene.bin> --steps N [--threads T] [--states out.bin] [--bp out.bin]
//                   [--contacts out.bin] [--warmup W] [--quiet]
//
// Outputs
//   --states   float32 [N+1][nDyn][13]  (pos3 quat4 linVel3 angVel3; row 0 = initial state)
//   --bp       per step: u32 nActors, float bounds[nActors][6] (tight world AABBs the step's
//              broadphase sees), u32 nCreated, u32 nDeleted, then (u32 a,u32 b) pairs (a<b), both
//              lists sorted.  Pairs come from a stage-level PxBroadPhase(eABP)+PxAABBManager fed
//              those bounds with distance = contactOffset  (physx/include/PxBroadPhase.h:492-796).
//   --contacts per step: u32 nPairs, then per pair: u32 actor0, u32 actor1, u32 nContacts and
//              nContacts x {pos3, normal3, separation, impulse3}  (PxContactPair::extractContacts)
//   --order    per step: u32 nEdges, then (u32 actor0, u32 actor1) in the order the island manager
//              feeds contact managers to the solver (IG::Island edge lists walked exactly like
//              DynamicsTGSContext::prepareBodiesAndConstraints, DyTGSDynamics.cpp:822-905), dumped
//              at the moment the solver task (ScScene.updateDynamics) is submitted, i.e. the lists the step's
//              solve really uses (a custom inline PxCpuDispatcher intercepts the task; island splits caused
//              by lost edges only happen after the solve, so a post-step dump would be wrong).
//   stdout     one JSON line with timing.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include <algorithm>
#include <chrono>
#include <string>
#include "PxPhysicsAPI.h"
// Internal headers are included (with access opened up) ONLY so that --order can dump the order in which
// the reference's island manager hands contact managers to the solver; nothing internal is modified.
#define private public
#define protected public
#include "NpScene.h"
#include "ScScene.h"
#include "PxsSimpleIslandManager.h"
#include "PxsIslandSim.h"
#include "PxsContactManager.h"
#undef private
#undef protected
#include "scene_format.h"
#include "GuConvexMesh.h"
#include "GuBigConvexData.h"

using namespace physx;

static PxDefaultAllocator gAllocator;
static PxDefaultErrorCallback gErrorCallback;

struct ContactDump {
  struct Pair { uint32_t a0, a1; std::vector<float> data; uint32_t n; };
  std::vector<Pair> pairs;
};
static ContactDump gContacts;

class EventCb : public PxSimulationEventCallback {
public:
  void onConstraintBreak(PxConstraintInfo*, PxU32) override {}
  void onWake(PxActor**, PxU32) override {}
  void onSleep(PxActor**, PxU32) override {}
  void onTrigger(PxTriggerPair*, PxU32) override {}
  void onAdvance(const PxRigidBody* const*, const PxTransform*, const PxU32) override {}
  void onContact(const PxContactPairHeader& hdr, const PxContactPair* pairs, PxU32 nbPairs) override {
    for (PxU32 i = 0; i < nbPairs; i++) {
      const PxContactPair& cp = pairs[i];
      if (cp.flags & (PxContactPairFlag::eREMOVED_SHAPE_0 | PxContactPairFlag::eREMOVED_SHAPE_1)) continue;
      PxContactPairPoint pts[64];
      PxU32 n = cp.extractContacts(pts, 64);
      ContactDump::Pair p;
      p.a0 = uint32_t(size_t(hdr.actors[0]->userData));
      p.a1 = uint32_t(size_t(hdr.actors[1]->userData));
      p.n = n;
      for (PxU32 k = 0; k < n; k++) {
        const float rec[10] = {pts[k].position.x, pts[k].position.y, pts[k].position.z,
                               pts[k].normal.x, pts[k].normal.y, pts[k].normal.z, pts[k].separation,
                               pts[k].impulse.x, pts[k].impulse.y, pts[k].impulse.z};
        p.data.insert(p.data.end(), rec, rec + 10);
      }
      gContacts.pairs.push_back(p);
    }
  }
};

// Inline dispatcher (same behaviour as PxDefaultCpuDispatcherCreate(0)) that calls a hook right before the
// solver task runs.
struct InlineDispatcher : public PxCpuDispatcher {
  void (*hook)(void*) = nullptr; void* user = nullptr;
  void submitTask(PxBaseTask& task) override {
    if (hook && task.getName() && !strcmp(task.getName(), "ScScene.updateDynamics")) hook(user);
    task.run(); task.release();
  }
  uint32_t getWorkerCount() const override { return 0; }
};

static bool gWantContacts = false;
static PxMaterial* gDefaultMat = nullptr;
static bool gUseDefaultFilter = false;   // scene file carries a filter section: the pair goes through the extension's PxDefaultSimulationFilterShader first
static PxFilterFlags filterShader(PxFilterObjectAttributes a0, PxFilterData fd0, PxFilterObjectAttributes a1, PxFilterData fd1,
                                  PxPairFlags& pairFlags, const void* cb, PxU32) {
  if (gUseDefaultFilter) {
    const PxFilterFlags f = PxDefaultSimulationFilterShader(a0, fd0, a1, fd1, pairFlags, nullptr, 0);
    if (f & PxFilterFlag::eSUPPRESS) return f;
  }
  pairFlags = PxPairFlag::eCONTACT_DEFAULT;
  if (cb && *reinterpret_cast<const int*>(cb))
    pairFlags |= PxPairFlag::eNOTIFY_TOUCH_FOUND | PxPairFlag::eNOTIFY_TOUCH_PERSISTS | PxPairFlag::eNOTIFY_CONTACT_POINTS;
  return PxFilterFlag::eDEFAULT;
}

static std::vector<uint8_t> readFile(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> b(n);
  if (fread(b.data(), 1, n, f) != size_t(n)) { fprintf(stderr, "short read\n"); exit(2); }
  fclose(f);
  return b;
}

struct Hull { std::vector<PxVec3> verts; PxConvexMesh* mesh = nullptr; };

int main(int argc, char** argv) {
  const char* cookOut = nullptr;
  if (argc == 4 && strcmp(argv[1], "cook") == 0) { cookOut = argv[3]; argc = 3; }
  else if (argc < 3 || strcmp(argv[1], "run") != 0) {
    fprintf(stderr, "usage: ref_harness run <scene.bin> --steps N [--threads T] [--states f] [--bp f] [--contacts f] [--warmup W] [--hulls f]\n");
    return 2;
  }
  const char* scenePath = argv[2];
  int steps = 100, threads = 1, warmup = 0;
  const char* gpuPlugin = nullptr; bool gpuBp = false, gpuDynamics = false, directGpuApi = false; int gpuBpShift = -1;
  const char *statesPath = nullptr, *bpPath = nullptr, *contactsPath = nullptr, *hullsPath = nullptr, *orderPath = nullptr, *sleepPath = nullptr, *forcesPath = nullptr, *kinPath = nullptr;
  for (int i = 3; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--steps") steps = atoi(argv[++i]);
    else if (a == "--threads") threads = atoi(argv[++i]);
    else if (a == "--warmup") warmup = atoi(argv[++i]);
    else if (a == "--states") statesPath = argv[++i];
    else if (a == "--bp") bpPath = argv[++i];
    else if (a == "--contacts") contactsPath = argv[++i];
    else if (a == "--hulls") hullsPath = argv[++i];
    else if (a == "--order") orderPath = argv[++i];
    else if (a == "--forces") forcesPath = argv[++i];   // f32[blocks][nDyn][6] = force xyz, torque xyz: block s (mod blocks) is applied with addForce / addTorque(eFORCE) before step s
    else if (a == "--kin-targets") kinPath = argv[++i];   // f32[blocks][nKinematic][7] = PxTransform (q.xyzw, p.xyz): block s is handed to setKinematicTarget before step s (no target once the blocks run out)
    else if (a == "--gpu-plugin") gpuPlugin = argv[++i];   // (GPU-enabled host build only) path of a libPhysXGpu_64.so to load through PxSetPhysXGpuLoadHook
    else if (a == "--gpu-bp") gpuBp = true;                // PxBroadPhaseType::eGPU (CPU dynamics)
    else if (a == "--gpu-dynamics") gpuDynamics = true;    // + PxSceneFlag::eENABLE_GPU_DYNAMICS
    else if (a == "--direct-gpu-api") directGpuApi = true; // + PxSceneFlag::eENABLE_DIRECT_GPU_API (no state read-back to the host objects: timing runs only)
    else if (a == "--gpu-bp-shift") gpuBpShift = atoi(argv[++i]);   // PxGpuBroadPhaseDesc::gpuBroadPhaseNbBitsShiftX/Y/Z (default 4; SURVEY 8d asks for the shift-0 variant too)
    else if (a == "--sleep") sleepPath = argv[++i];   // per step, per dynamic actor: f32 wakeCounter, u32 isSleeping
  }
  gWantContacts = contactsPath != nullptr;
  static int wantContactsFlag; wantContactsFlag = gWantContacts ? 1 : 0;

  std::vector<uint8_t> buf = readFile(scenePath);
  const PxbSceneHeader& H = *reinterpret_cast<const PxbSceneHeader*>(buf.data());
  if (H.magic != PXB_SCENE_MAGIC) { fprintf(stderr, "bad magic\n"); return 2; }
  const PxbActorRec* recs = reinterpret_cast<const PxbActorRec*>(buf.data() + sizeof(PxbSceneHeader));
  const PxbMaterialRec* matRecs = reinterpret_cast<const PxbMaterialRec*>(recs + H.nActors);   // material table (may be empty)
  const PxbLocalPoseRec* localPoses = H.reserved[3] == PXB_LOCAL_POSE_MAGIC ? reinterpret_cast<const PxbLocalPoseRec*>(matRecs + H.reserved[2]) : nullptr;   // PxShape::setLocalPose / setCMassLocalPose per actor
  const uint8_t* afterLocal = reinterpret_cast<const uint8_t*>(matRecs + H.reserved[2]) + (localPoses ? sizeof(PxbLocalPoseRec) * size_t(H.nActors) : 0);
  const PxbFilterShaderConfig* filterCfg = (H.reserved[0] & PXB_FLAG_FILTER_SECTION) ? reinterpret_cast<const PxbFilterShaderConfig*>(afterLocal) : nullptr;
  const uint32_t* filterData = filterCfg ? reinterpret_cast<const uint32_t*>(filterCfg + 1) : nullptr;
  const uint8_t* afterFilter = afterLocal + (filterCfg ? sizeof(PxbFilterShaderConfig) + 16 * size_t(H.nActors) : 0);
  const float* shapeOffsets = (H.reserved[0] & PXB_FLAG_SHAPE_OFFSETS) ? reinterpret_cast<const float*>(afterFilter) : nullptr;   // per shape: contactOffset, restOffset
  const uint8_t* hp = afterFilter + (shapeOffsets ? 8 * size_t(H.nActors) : 0);
  if (filterCfg) {   // the extension's global filter state, through its public setters
    gUseDefaultFilter = true;
    for (PxU16 g = 0; g < 32; g++) for (PxU16 h = 0; h < 32; h++) PxSetGroupCollisionFlag(g, h, ((filterCfg->collisionTable[g] >> h) & 1u) != 0);
    PxSetFilterOps(PxFilterOp::Enum(filterCfg->ops[0]), PxFilterOp::Enum(filterCfg->ops[1]), PxFilterOp::Enum(filterCfg->ops[2]));
    PxSetFilterBool(filterCfg->filterBool != 0);
    PxGroupsMask k0, k1;
    k0.bits0 = PxU16(filterCfg->constants[0] & 0xffff); k0.bits1 = PxU16(filterCfg->constants[0] >> 16); k0.bits2 = PxU16(filterCfg->constants[1] & 0xffff); k0.bits3 = PxU16(filterCfg->constants[1] >> 16);
    k1.bits0 = PxU16(filterCfg->constants[2] & 0xffff); k1.bits1 = PxU16(filterCfg->constants[2] >> 16); k1.bits2 = PxU16(filterCfg->constants[3] & 0xffff); k1.bits3 = PxU16(filterCfg->constants[3] >> 16);
    PxSetFilterConstants(k0, k1);
  }

  PxFoundation* foundation = PxCreateFoundation(PX_PHYSICS_VERSION, gAllocator, gErrorCallback);
  PxTolerancesScale scale; scale.length = H.toleranceLength; scale.speed = 10.0f * H.toleranceLength;
  PxPhysics* physics = PxCreatePhysics(PX_PHYSICS_VERSION, *foundation, scale, false, nullptr);

  std::vector<Hull> hulls(H.nHulls);
  for (uint32_t h = 0; h < H.nHulls; h++) {
    uint32_t nv; memcpy(&nv, hp, 4); hp += 4;
    hulls[h].verts.resize(nv);
    memcpy(hulls[h].verts.data(), hp, nv * 12); hp += nv * 12;
    PxConvexMeshDesc d;
    d.points.count = nv; d.points.stride = sizeof(PxVec3); d.points.data = hulls[h].verts.data();
    d.flags = PxConvexFlag::eCOMPUTE_CONVEX;
    d.vertexLimit = 64;
    PxCookingParams cp(scale); cp.buildGPUData = true;
    hulls[h].mesh = PxCreateConvexMesh(cp, d, physics->getPhysicsInsertionCallback());
    if (!hulls[h].mesh) { fprintf(stderr, "hull cook failed\n"); return 3; }
  }
  if (cookOut) {   // `ref_harness cook scene.bin out.bin`: the cooked section of the scene format (scene_format.h) from Gu::ConvexHullData
    FILE* f = fopen(cookOut, "wb");
    for (auto& h : hulls) {
      const Gu::ConvexHullData& hd = static_cast<Gu::ConvexMesh*>(h.mesh)->getHull();
      PxbCookedHullHeader ch; memset(&ch, 0, sizeof(ch));
      ch.nVerts = hd.mNbHullVertices; ch.nPolys = hd.mNbPolygons; ch.nEdges = PxU32(PxU16(hd.mNbEdges));
      uint32_t nIdx = 0; for (uint32_t p = 0; p < ch.nPolys; p++) nIdx = PxMax(nIdx, uint32_t(hd.mPolygons[p].mVRef8) + hd.mPolygons[p].mNbVerts);
      ch.nIdx = nIdx;
      memcpy(ch.centerOfMass, &hd.mCenterOfMass, 12); memcpy(ch.boundsCenter, &hd.mAABB.mCenter, 12); memcpy(ch.boundsExtents, &hd.mAABB.mExtents, 12);
      ch.internalRadius = hd.mInternal.mInternalRadius; memcpy(ch.internalExtents, &hd.mInternal.mInternalExtents, 12);
      const Gu::BigConvexRawData* big = hd.mBigConvexRawData;   // hill-climbing data: only hulls of more than 32 vertices carry it
      if (big) ch.reserved[0] = PxU32(big->mSubdiv) | (big->mNbAdjVerts << 16);
      PxReal mass; PxMat33 inertia; PxVec3 com; h.mesh->getMassInformation(mass, inertia, com);
      ch.unitMass = mass; ch.unitInertiaDiag[0] = inertia.column0.x; ch.unitInertiaDiag[1] = inertia.column1.y; ch.unitInertiaDiag[2] = inertia.column2.z; memcpy(ch.unitCom, &com, 12);
      fwrite(&ch, sizeof(ch), 1, f);
      fwrite(hd.getHullVertices(), 12, ch.nVerts, f);
      for (uint32_t p = 0; p < ch.nPolys; p++) {
        const Gu::HullPolygonData& pd = hd.mPolygons[p];
        PxbCookedPoly cp_; cp_.plane[0] = pd.mPlane.n.x; cp_.plane[1] = pd.mPlane.n.y; cp_.plane[2] = pd.mPlane.n.z; cp_.plane[3] = pd.mPlane.d;
        cp_.vref = pd.mVRef8; cp_.nbVerts = pd.mNbVerts; cp_.minIndex = pd.mMinIndex; cp_.pad = 0;
        fwrite(&cp_, sizeof(cp_), 1, f);
      }
      const uint8_t zero[4] = {0, 0, 0, 0};
      fwrite(hd.getVertexData8(), 1, nIdx, f); fwrite(zero, 1, (4 - nIdx % 4) % 4, f);
      fwrite(hd.getFacesByEdges8(), 1, 2 * ch.nEdges, f); fwrite(zero, 1, (4 - (2 * ch.nEdges) % 4) % 4, f);
      if (big) {
        const uint32_t ns = 6u * big->mSubdiv * big->mSubdiv;   // mSamples: one vertex index per cube-map texel
        fwrite(big->mSamples, 1, ns, f); fwrite(zero, 1, (4 - ns % 4) % 4, f);
        fwrite(big->mValencies, 4, ch.nVerts, f);               // Gu::Valency {u16 mCount, u16 mOffset}
        fwrite(big->mAdjacentVerts, 1, big->mNbAdjVerts, f); fwrite(zero, 1, (4 - big->mNbAdjVerts % 4) % 4, f);
      }
    }
    fclose(f);
    return 0;
  }
  if (hullsPath) {
    // cooked hull dump: per hull: u32 nVerts, u32 nPolys, verts xyz, per poly: plane(nx ny nz d), u32 nIdx, u32 idx[nIdx]
    FILE* f = fopen(hullsPath, "wb");
    for (auto& h : hulls) {
      uint32_t nv = h.mesh->getNbVertices(), np = h.mesh->getNbPolygons();
      fwrite(&nv, 4, 1, f); fwrite(&np, 4, 1, f);
      fwrite(h.mesh->getVertices(), 12, nv, f);
      const PxU8* ib = h.mesh->getIndexBuffer();
      for (uint32_t p = 0; p < np; p++) {
        PxHullPolygon poly; h.mesh->getPolygonData(p, poly);
        fwrite(poly.mPlane, 4, 4, f);
        uint32_t n = poly.mNbVerts; fwrite(&n, 4, 1, f);
        for (uint32_t k = 0; k < n; k++) { uint32_t v = ib[poly.mIndexBase + k]; fwrite(&v, 4, 1, f); }
      }
    }
    fclose(f);
  }

  PxSceneDesc sd(scale);
  sd.gravity = PxVec3(H.gravity[0], H.gravity[1], H.gravity[2]);
  PxDefaultCpuDispatcher* dispatcher = PxDefaultCpuDispatcherCreate(threads);
  static InlineDispatcher inlineDispatcher;
  sd.cpuDispatcher = orderPath ? static_cast<PxCpuDispatcher*>(&inlineDispatcher) : static_cast<PxCpuDispatcher*>(dispatcher);
  sd.filterShader = filterShader;
  sd.filterShaderData = &wantContactsFlag;
  sd.filterShaderDataSize = sizeof(int);
  sd.broadPhaseType = PxBroadPhaseType::eABP;
#if PX_SUPPORT_GPU_PHYSX
  // The UNMODIFIED host SDK with a PhysXGpu plugin loaded through its own loader (PxPhysXGpuModuleLoader.cpp): used to run the repo's
  // libPhysXGpu_64.so (plugin/) inside the reference -- the scene is created exactly as an application would.
  struct LoadHook : public PxGpuLoadHook { const char* name; const char* getPhysXGpuDllName() const override { return name; } };
  static LoadHook hook;
  PxCudaContextManager* cudaMgr = nullptr;
  if (gpuBp || gpuDynamics) {
    if (gpuPlugin) { hook.name = gpuPlugin; PxSetPhysXGpuLoadHook(&hook); }
    PxCudaContextManagerDesc cd;
    cudaMgr = PxCreateCudaContextManager(*foundation, cd, nullptr);
    if (!cudaMgr || !cudaMgr->contextIsValid()) { fprintf(stderr, "no CUDA context manager (plugin not loaded or no GPU)\n"); return 4; }
    sd.cudaContextManager = cudaMgr;
    sd.broadPhaseType = PxBroadPhaseType::eGPU;
    if (gpuDynamics) sd.flags |= PxSceneFlag::eENABLE_GPU_DYNAMICS;
    if (gpuDynamics && directGpuApi) sd.flags |= PxSceneFlag::eENABLE_DIRECT_GPU_API;
    static PxGpuBroadPhaseDesc bpDesc;
    if (gpuBpShift >= 0) { bpDesc.gpuBroadPhaseNbBitsShiftX = bpDesc.gpuBroadPhaseNbBitsShiftY = bpDesc.gpuBroadPhaseNbBitsShiftZ = PxU8(gpuBpShift); sd.gpuBroadPhaseDesc = &bpDesc; }
    sd.gpuDynamicsConfig.foundLostPairsCapacity = PxMax(1u << 20, 8u * H.nActors);
    if (gpuDynamics) {   // PxGpuDynamicsMemoryConfig sized for the scene (contacts >= 8 M, patches >= 2 M at BASELINE sizes)
      sd.gpuDynamicsConfig.maxRigidContactCount = PxMax(1u << 20, 16u * H.nActors);
      sd.gpuDynamicsConfig.maxRigidPatchCount = PxMax(1u << 18, 8u * H.nActors);
      sd.gpuDynamicsConfig.tempBufferCapacity = PxMax(16u << 20, 256u * H.nActors);
      sd.gpuDynamicsConfig.heapCapacity = PxMax(64u << 20, 1024u * H.nActors);
      sd.gpuMaxNumPartitions = 8;
    }
    fprintf(stderr, "GPU plugin: %s on %s\n", gpuPlugin ? gpuPlugin : "(default libPhysXGpu_64.so)", cudaMgr->getDeviceName());
  }
#else
  if (gpuBp || gpuDynamics || gpuPlugin) { fprintf(stderr, "this ref_harness was built without PX_SUPPORT_GPU_PHYSX: use oracle/_ref_gpu/ref_harness\n"); return 2; }
#endif
  sd.solverType = H.solverType == PXB_SOLVER_TGS ? PxSolverType::eTGS : PxSolverType::ePGS;
  sd.flags |= PxSceneFlag::eENABLE_PCM;
  sd.bounceThresholdVelocity = H.bounceThreshold;
  sd.frictionOffsetThreshold = H.frictionOffsetThreshold;
  sd.frictionCorrelationDistance = H.frictionCorrelationDistance;
  EventCb cb;
  if (gWantContacts) sd.simulationEventCallback = &cb;
  PxScene* scene = physics->createScene(sd);
  PxMaterial* mat = physics->createMaterial(H.staticFriction, H.dynamicFriction, H.restitution);
  gDefaultMat = mat;
  std::vector<PxMaterial*> mats(H.reserved[2]);
  for (uint32_t m = 0; m < H.reserved[2]; m++) {   // PxMaterial per table entry: coefficients, combine modes, eDISABLE_FRICTION
    mats[m] = physics->createMaterial(matRecs[m].staticFriction, matRecs[m].dynamicFriction, matRecs[m].restitution);
    mats[m]->setFrictionCombineMode(PxCombineMode::Enum(matRecs[m].bits & 15u)); mats[m]->setRestitutionCombineMode(PxCombineMode::Enum((matRecs[m].bits >> 4) & 15u));
    if ((matRecs[m].bits >> 8) & 1u) mats[m]->setFlag(PxMaterialFlag::eDISABLE_FRICTION, true);
  }

  std::vector<PxRigidActor*> actors(H.nActors);
  std::vector<PxRigidDynamic*> dyn, kin;   // kin: the kinematic ones, in dynamic-body order
  std::vector<PxShape*> shapes(H.nActors);
  // PxAggregate per aggregate id (bit 31 of the id = self collisions): created and added to the scene up front, members join in record order
  std::map<uint32_t, std::vector<uint32_t>> aggMembers; std::map<uint32_t, PxAggregate*> aggs;
  for (uint32_t i = 0; i < H.nActors; i++) if (recs[i].aggregate) aggMembers[recs[i].aggregate].push_back(i);
  for (auto& kv : aggMembers) {
    const PxU32 n = PxU32(kv.second.size());
    aggs[kv.first] = physics->createAggregate(n, n, PxGetAggregateFilterHint(PxAggregateType::eGENERIC, (kv.first & 0x80000000u) != 0));
    scene->addAggregate(*aggs[kv.first]);
  }
  for (uint32_t i = 0; i < H.nActors; i++) {
    const PxbActorRec& r = recs[i];
    PxTransform pose(PxVec3(r.pos[0], r.pos[1], r.pos[2]), PxQuat(r.quat[0], r.quat[1], r.quat[2], r.quat[3]));
    PxRigidActor* a;
    PxMaterial* mat = mats.empty() ? ::gDefaultMat : mats[r.materialIndex < mats.size() ? r.materialIndex : 0];
    if (r.flags & PXB_ACTOR_DYNAMIC) a = physics->createRigidDynamic(pose); else a = physics->createRigidStatic(pose);
    PxShape* s = nullptr;
    switch (r.geomType) {
      case PXB_GEOM_SPHERE: s = PxRigidActorExt::createExclusiveShape(*a, PxSphereGeometry(r.dims[0]), *mat); break;
      case PXB_GEOM_PLANE: s = PxRigidActorExt::createExclusiveShape(*a, PxPlaneGeometry(), *mat); break;
      case PXB_GEOM_CAPSULE: s = PxRigidActorExt::createExclusiveShape(*a, PxCapsuleGeometry(r.dims[0], r.dims[1]), *mat); break;
      case PXB_GEOM_BOX: s = PxRigidActorExt::createExclusiveShape(*a, PxBoxGeometry(r.dims[0], r.dims[1], r.dims[2]), *mat); break;
      case PXB_GEOM_CONVEX: s = PxRigidActorExt::createExclusiveShape(*a, PxConvexMeshGeometry(hulls[r.hullIdx].mesh), *mat); break;
      default: fprintf(stderr, "bad geom %u\n", r.geomType); return 2;
    }
    s->setContactOffset(shapeOffsets ? shapeOffsets[2 * i] : H.contactOffset);
    s->setRestOffset(shapeOffsets ? shapeOffsets[2 * i + 1] : H.restOffset);
    if (filterData) s->setSimulationFilterData(PxFilterData(filterData[4 * i], filterData[4 * i + 1], filterData[4 * i + 2], filterData[4 * i + 3]));
    if (localPoses) { const PxbLocalPoseRec& l = localPoses[i]; s->setLocalPose(PxTransform(PxVec3(l.shapeP[0], l.shapeP[1], l.shapeP[2]), PxQuat(l.shapeQ[0], l.shapeQ[1], l.shapeQ[2], l.shapeQ[3]))); }
    shapes[i] = s;
    a->userData = reinterpret_cast<void*>(size_t(i));
    if (r.flags & PXB_ACTOR_DYNAMIC) {
      PxRigidDynamic* d = static_cast<PxRigidDynamic*>(a);
      d->setMass(r.mass);
      d->setMassSpaceInertiaTensor(PxVec3(r.inertia[0], r.inertia[1], r.inertia[2]));
      if (localPoses) { const PxbLocalPoseRec& l = localPoses[i]; d->setCMassLocalPose(PxTransform(PxVec3(l.bodyP[0], l.bodyP[1], l.bodyP[2]), PxQuat(l.bodyQ[0], l.bodyQ[1], l.bodyQ[2], l.bodyQ[3]))); }
      else d->setCMassLocalPose(PxTransform(PxIdentity));
      d->setLinearVelocity(PxVec3(r.linVel[0], r.linVel[1], r.linVel[2]));
      d->setAngularVelocity(PxVec3(r.angVel[0], r.angVel[1], r.angVel[2]));
      d->setLinearDamping(r.linDamping);
      d->setAngularDamping(r.angDamping);
      d->setMaxLinearVelocity(r.maxLinVel);
      d->setMaxAngularVelocity(r.maxAngVel);
      d->setMaxDepenetrationVelocity(r.maxDepenetrationVel);
      d->setSolverIterationCounts(H.posIters, H.velIters);
      d->setSleepThreshold(H.sleepThreshold);
      if (H.sleepThreshold == 0.0f) d->setWakeCounter(1e9f);
      d->setRigidDynamicLockFlags(PxRigidDynamicLockFlags(PxU8((r.flags >> 8) & 0x3f)));
      if (r.flags & PXB_ACTOR_DISABLE_GRAVITY) d->setActorFlag(PxActorFlag::eDISABLE_GRAVITY, true);
      if (r.flags & PXB_ACTOR_GYROSCOPIC) d->setRigidBodyFlag(PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES, true);
      if (r.flags & PXB_ACTOR_KINEMATIC) { d->setRigidBodyFlag(PxRigidBodyFlag::eKINEMATIC, true); kin.push_back(d); }
      dyn.push_back(d);
    }
    actors[i] = a;
    if (r.aggregate) aggs[r.aggregate]->addActor(*a); else scene->addActor(*a);   // (an aggregate that is in the scene inserts the actor right away: actor ids stay in record order)
  }

  FILE* fs = statesPath ? fopen(statesPath, "wb") : nullptr;
  FILE* fb = bpPath ? fopen(bpPath, "wb") : nullptr;
  FILE* fc = contactsPath ? fopen(contactsPath, "wb") : nullptr;
  FILE* fo = orderPath ? fopen(orderPath, "wb") : nullptr;
  FILE* fsl = sleepPath ? fopen(sleepPath, "wb") : nullptr;
  static FILE* sFo; static PxScene* sScene; sFo = fo; sScene = scene;
  static void (*sDump)();
  static uint32_t sIdShift; sIdShift = uint32_t(aggs.size());   // element ids: every aggregate took one from the pool before the first shape (they are created up front), shapes follow in record order
  auto dumpOrder = []() {
    FILE* fo = sFo; PxScene* scene = sScene;
    if (!fo) return;
    NpScene* np = static_cast<NpScene*>(scene);
    IG::SimpleIslandManager* im = np->getScScene().getSimpleIslandManager();
    const IG::IslandSim& is = im->getAccurateIslandSim();
    std::vector<uint32_t> edges;
    for (PxU32 i = 0; i < is.getNbActiveIslands(); i++) {
      const IG::Island& isl = is.getIsland(is.getActiveIslands()[i]);
      IG::EdgeIndex e = isl.mFirstEdge[IG::Edge::eCONTACT_MANAGER];
      while (e != IG_INVALID_EDGE) {
        PxsContactManager* cm = im->getContactManager(e);
        if (cm) { edges.push_back(cm->getWorkUnit().mTransformCache0 - sIdShift); edges.push_back(cm->getWorkUnit().mTransformCache1 - sIdShift); }
        e = is.getEdge(e).mNextIslandEdge;
      }
    }
    uint32_t n = uint32_t(edges.size() / 2);
    fwrite(&n, 4, 1, fo); fwrite(edges.data(), 4, edges.size(), fo);
  };
  sDump = dumpOrder;

  auto dumpStates = [&]() {
    if (!fs) return;
    std::vector<float> row(dyn.size() * PXB_STATE_FLOATS);
    for (size_t i = 0; i < dyn.size(); i++) {
      PxTransform t = dyn[i]->getGlobalPose();
      PxVec3 lv = dyn[i]->getLinearVelocity(), av = dyn[i]->getAngularVelocity();
      float* o = &row[i * PXB_STATE_FLOATS];
      o[0] = t.p.x; o[1] = t.p.y; o[2] = t.p.z; o[3] = t.q.x; o[4] = t.q.y; o[5] = t.q.z; o[6] = t.q.w;
      o[7] = lv.x; o[8] = lv.y; o[9] = lv.z; o[10] = av.x; o[11] = av.y; o[12] = av.z;
    }
    fwrite(row.data(), 4, row.size(), fs);
  };

  // stage-level broadphase oracle
  PxBroadPhase* bp = nullptr; PxAABBManager* aabb = nullptr;
  if (fb) {
    PxBroadPhaseDesc bpd(PxBroadPhaseType::eABP);
    bpd.mDiscardKinematicVsKinematic = true; bpd.mDiscardStaticVsKinematic = true;   // what a scene's broadphase is created with under PxPairFilteringMode::eDEFAULT (ScScene.cpp: kineKine / staticKine filtering != eKEEP)
    bp = PxCreateBroadPhase(bpd);
    aabb = PxCreateAABBManager(*bp);
  }
  auto bpStep = [&](bool first) {
    if (!fb) return;
    std::vector<float> bounds(H.nActors * 6);
    for (uint32_t i = 0; i < H.nActors; i++) {
      PxBounds3 b;
      PxGeometryQuery::computeGeomBounds(b, shapes[i]->getGeometry(), actors[i]->getGlobalPose() * shapes[i]->getLocalPose(), 0.0f, 1.0f);
      memcpy(&bounds[i * 6], &b.minimum.x, 24);
      const bool isDyn = recs[i].flags & PXB_ACTOR_DYNAMIC;
      if (first) {
        // (standalone broadphase: objects of one filter group do not collide -- the members of an aggregate without self collisions share the group of its first member)
        const uint32_t gid = (recs[i].aggregate && !(recs[i].aggregate & 0x80000000u)) ? aggMembers[recs[i].aggregate][0] : i;
        PxBpFilterGroup g = (recs[i].flags & PXB_ACTOR_KINEMATIC) ? PxGetBroadPhaseKinematicFilterGroup(gid) : isDyn ? PxGetBroadPhaseDynamicFilterGroup(gid) : PxGetBroadPhaseStaticFilterGroup();
        aabb->addObject(i, b, g, shapeOffsets ? shapeOffsets[2 * i] : H.contactOffset);   // contact distance of the object = its shape's contact offset
      } else if (isDyn) {
        aabb->updateObject(i, &b, nullptr);
      }
    }
    aabb->update();
    PxBroadPhaseResults res; aabb->fetchResults(res);
    auto norm = [](const PxBroadPhasePair* p, PxU32 n) {
      std::vector<std::pair<uint32_t, uint32_t>> v(n);
      for (PxU32 i = 0; i < n; i++) v[i] = std::make_pair(std::min(p[i].mID0, p[i].mID1), std::max(p[i].mID0, p[i].mID1));
      std::sort(v.begin(), v.end());
      return v;
    };
    auto c = norm(res.mCreatedPairs, res.mNbCreatedPairs), d = norm(res.mDeletedPairs, res.mNbDeletedPairs);
    uint32_t n = H.nActors, nc = uint32_t(c.size()), nd = uint32_t(d.size());
    fwrite(&n, 4, 1, fb); fwrite(bounds.data(), 4, bounds.size(), fb);
    fwrite(&nc, 4, 1, fb); fwrite(&nd, 4, 1, fb);
    for (auto& p : c) { fwrite(&p.first, 4, 1, fb); fwrite(&p.second, 4, 1, fb); }
    for (auto& p : d) { fwrite(&p.first, 4, 1, fb); fwrite(&p.second, 4, 1, fb); }
  };

  std::vector<float> forces;
  if (forcesPath) { std::vector<uint8_t> fb_ = readFile(forcesPath); forces.resize(fb_.size() / 4); memcpy(forces.data(), fb_.data(), forces.size() * 4); }
  std::vector<float> kinTargets;
  if (kinPath) { std::vector<uint8_t> kb_ = readFile(kinPath); kinTargets.resize(kb_.size() / 4); memcpy(kinTargets.data(), kb_.data(), kinTargets.size() * 4); }
  inlineDispatcher.hook = [](void*) { sDump(); };
  dumpStates();
  double totalMs = 0; std::vector<double> stepMs;
  for (int s = 0; s < steps + warmup; s++) {
    bpStep(s == 0);
    gContacts.pairs.clear();
    if (!forces.empty()) {   // PxRigidBody::addForce / addTorque (PxForceMode::eFORCE): the CPU-side equivalent of PxDirectGPUAPI eFORCE / eTORQUE writes
      const size_t blocks = forces.size() / (dyn.size() * 6); const float* f = forces.data() + (size_t(s) % blocks) * dyn.size() * 6;
      for (size_t i = 0; i < dyn.size(); i++) {
        const PxVec3 F(f[i * 6], f[i * 6 + 1], f[i * 6 + 2]), T(f[i * 6 + 3], f[i * 6 + 4], f[i * 6 + 5]);
        if (!F.isZero()) dyn[i]->addForce(F, PxForceMode::eFORCE);
        if (!T.isZero()) dyn[i]->addTorque(T, PxForceMode::eFORCE);
      }
    }
    if (!kin.empty() && size_t(s + 1) * kin.size() * 7 <= kinTargets.size()) {   // PxRigidDynamic::setKinematicTarget
      const float* k = kinTargets.data() + size_t(s) * kin.size() * 7;
      for (size_t i = 0; i < kin.size(); i++) if (k[i * 7] == k[i * 7]) /* NaN row: no target this step */ kin[i]->setKinematicTarget(PxTransform(PxVec3(k[i * 7 + 4], k[i * 7 + 5], k[i * 7 + 6]), PxQuat(k[i * 7], k[i * 7 + 1], k[i * 7 + 2], k[i * 7 + 3])));
    }
    auto t0 = std::chrono::steady_clock::now();
    scene->simulate(H.dt);
    scene->fetchResults(true);
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (s >= warmup) { totalMs += ms; stepMs.push_back(ms); }
    dumpStates();
    if (fsl) for (size_t i = 0; i < dyn.size(); i++) { const float wc = dyn[i]->getWakeCounter(); const uint32_t sl = dyn[i]->isSleeping() ? 1u : 0u; fwrite(&wc, 4, 1, fsl); fwrite(&sl, 4, 1, fsl); }
    if (fc) {
      std::sort(gContacts.pairs.begin(), gContacts.pairs.end(), [](const ContactDump::Pair& x, const ContactDump::Pair& y) {
        return std::make_pair(x.a0, x.a1) < std::make_pair(y.a0, y.a1); });
      uint32_t np = uint32_t(gContacts.pairs.size());
      fwrite(&np, 4, 1, fc);
      for (auto& p : gContacts.pairs) {
        fwrite(&p.a0, 4, 1, fc); fwrite(&p.a1, 4, 1, fc); fwrite(&p.n, 4, 1, fc);
        fwrite(p.data.data(), 4, p.data.size(), fc);
      }
    }
  }
  if (fs) fclose(fs);
  if (fb) fclose(fb);
  if (fc) fclose(fc);
  if (fo) fclose(fo);
  if (fsl) fclose(fsl);
  std::sort(stepMs.begin(), stepMs.end());
  double med = stepMs.empty() ? 0 : stepMs[stepMs.size() / 2];
  double p95 = stepMs.empty() ? 0 : stepMs[std::min(stepMs.size() - 1, size_t(stepMs.size() * 0.95))];
  PxSimulationStatistics st; scene->getSimulationStatistics(st);
  printf("{\"impl\": \"reference-cpu\", \"version\": \"5.6.1\", \"bodies\": %zu, \"steps\": %d, \"threads\": %d, \"solver\": \"%s\", "
         "\"ms_per_step\": %.6f, \"ms_median\": %.6f, \"ms_p95\": %.6f, \"body_steps_per_s\": %.1f, \"nb_discrete_contact_pairs\": %u}\n",
         dyn.size(), steps, threads, H.solverType == PXB_SOLVER_TGS ? "tgs" : "pgs",
         totalMs / std::max(1, steps), med, p95, dyn.size() * double(steps) / (totalMs / 1000.0),
         st.nbDiscreteContactPairsTotal);
  scene->release();
  dispatcher->release();
  physics->release();
#if PX_SUPPORT_GPU_PHYSX
  if (cudaMgr) cudaMgr->release();
#endif
  foundation->release();
  return 0;
}
