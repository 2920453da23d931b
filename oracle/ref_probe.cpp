// ref_probe: investigation tool (test infrastructure, not shipped) that steps a scene file with the
// UNMODIFIED reference CPU SDK and prints, per requested step, the order in which the reference's
// island manager hands contact managers to the solver (IG::Island edge lists walked exactly like
// DynamicsTGSContext::prepareBodiesAndConstraints, DyTGSDynamics.cpp:822-905).  Used to pin
// the constraint-input order that our engine reproduces (DESIGN.md "constraint order").
//   ref_probe <scene.bin> <steps> [printEvery]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "PxPhysicsAPI.h"
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
using namespace physx;
static PxDefaultAllocator gAllocator;
static PxDefaultErrorCallback gErrorCallback;
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf(n); fread(buf.data(), 1, n, f); fclose(f);
  int steps = atoi(argv[2]); int every = argc > 3 ? atoi(argv[3]) : 1;
  const PxbSceneHeader& H = *reinterpret_cast<const PxbSceneHeader*>(buf.data());
  const PxbActorRec* recs = reinterpret_cast<const PxbActorRec*>(buf.data() + sizeof(PxbSceneHeader));
  PxFoundation* foundation = PxCreateFoundation(PX_PHYSICS_VERSION, gAllocator, gErrorCallback);
  PxTolerancesScale scale;
  PxPhysics* physics = PxCreatePhysics(PX_PHYSICS_VERSION, *foundation, scale, false, nullptr);
  PxSceneDesc sd(scale);
  sd.gravity = PxVec3(H.gravity[0], H.gravity[1], H.gravity[2]);
  sd.cpuDispatcher = PxDefaultCpuDispatcherCreate(1);
  sd.filterShader = PxDefaultSimulationFilterShader;
  sd.broadPhaseType = PxBroadPhaseType::eABP;
  sd.solverType = H.solverType == PXB_SOLVER_TGS ? PxSolverType::eTGS : PxSolverType::ePGS;
  PxScene* scene = physics->createScene(sd);
  PxMaterial* mat = physics->createMaterial(H.staticFriction, H.dynamicFriction, H.restitution);
  for (uint32_t i = 0; i < H.nActors; i++) {
    const PxbActorRec& r = recs[i];
    PxTransform pose(PxVec3(r.pos[0], r.pos[1], r.pos[2]), PxQuat(r.quat[0], r.quat[1], r.quat[2], r.quat[3]));
    PxRigidActor* a = (r.flags & 1) ? (PxRigidActor*)physics->createRigidDynamic(pose) : (PxRigidActor*)physics->createRigidStatic(pose);
    PxShape* s = nullptr;
    if (r.geomType == PXB_GEOM_PLANE) s = PxRigidActorExt::createExclusiveShape(*a, PxPlaneGeometry(), *mat);
    else if (r.geomType == PXB_GEOM_BOX) s = PxRigidActorExt::createExclusiveShape(*a, PxBoxGeometry(r.dims[0], r.dims[1], r.dims[2]), *mat);
    else if (r.geomType == PXB_GEOM_SPHERE) s = PxRigidActorExt::createExclusiveShape(*a, PxSphereGeometry(r.dims[0]), *mat);
    else if (r.geomType == PXB_GEOM_CAPSULE) s = PxRigidActorExt::createExclusiveShape(*a, PxCapsuleGeometry(r.dims[0], r.dims[1]), *mat);
    s->setContactOffset(H.contactOffset);
    if (r.flags & 1) {
      PxRigidDynamic* d = static_cast<PxRigidDynamic*>(a);
      d->setMass(r.mass); d->setMassSpaceInertiaTensor(PxVec3(r.inertia[0], r.inertia[1], r.inertia[2]));
      d->setSolverIterationCounts(H.posIters, H.velIters);
      d->setSleepThreshold(H.sleepThreshold); if (H.sleepThreshold == 0.f) d->setWakeCounter(1e9f);
    }
    scene->addActor(*a);
  }
  NpScene* np = static_cast<NpScene*>(scene);
  for (int s = 0; s < steps; s++) {
    scene->simulate(H.dt); scene->fetchResults(true);
    if (s % every) continue;
    Sc::Scene& sc = np->getScScene();
    IG::SimpleIslandManager* im = sc.getSimpleIslandManager();
    const IG::IslandSim& is = im->getAccurateIslandSim();
    printf("step %d: %u active islands\n", s, is.getNbActiveIslands());
    for (PxU32 i = 0; i < is.getNbActiveIslands(); i++) {
      const IG::Island& isl = is.getIsland(is.getActiveIslands()[i]);
      printf("  island %u nodes:", is.getActiveIslands()[i]);
      PxNodeIndex ni = isl.mRootNode;
      while (ni.isValid()) { printf(" %u", ni.index()); ni = is.getNode(ni).mNextNode; }
      printf("\n   edges (tc0,tc1):");
      IG::EdgeIndex e = isl.mFirstEdge[IG::Edge::eCONTACT_MANAGER];
      while (e != IG_INVALID_EDGE) {
        PxsContactManager* cm = im->getContactManager(e);
        if (cm) printf(" (%u,%u)", cm->getWorkUnit().mTransformCache0, cm->getWorkUnit().mTransformCache1);
        else printf(" (null)");
        e = is.getEdge(e).mNextIslandEdge;
      }
      printf("\n");
    }
  }
  return 0;
}
