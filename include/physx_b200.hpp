// physx_b200.hpp -- header-only C++ host-side mirror of the reference scene interface over the C ABI (physx_b200.h).
// Names follow the reference: PxScene::simulate / fetchResults (physx/include/PxScene.h), PxDirectGPUAPI::getRigidDynamicData /
// setRigidDynamicData (physx/include/PxDirectGPUAPI.h:311-463), PxRigidDynamic::getWakeCounter / isSleeping.  Errors throw
// pxb::Error carrying pxb_last_error(); there is no CPU fallback (scene creation fails without a CUDA device).
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "physx_b200.h"

namespace pxb {

struct Error : std::runtime_error { int code; Error(int c, const std::string& m) : std::runtime_error(m), code(c) {} };
inline void check(int rc) { if (rc < 0) throw Error(rc, std::string("physx_b200 error ") + std::to_string(rc) + ": " + pxb_last_error()); }

// PxSceneDesc defaults of the hot path (SnippetHelloWorld.cpp:97-110: gravity -9.81, material 0.5 / 0.5 / 0.6)
inline PxbSceneDesc defaultSceneDesc(uint32_t maxActors, uint32_t solverType = PXB_SOLVER_TGS) {
  PxbSceneDesc d; std::memset(&d, 0, sizeof(d));
  d.gravity[1] = -9.81f; d.solverType = solverType;
  d.bounceThresholdVelocity = 2.0f; d.frictionOffsetThreshold = 0.04f; d.frictionCorrelationDistance = 0.025f; d.toleranceLength = 1.0f;
  d.staticFriction = 0.5f; d.dynamicFriction = 0.5f; d.restitution = 0.6f; d.contactOffset = 0.02f; d.restOffset = 0.f;
  d.posIters = 4; d.velIters = 1; d.maxActors = maxActors; d.maxPairs = 0; d.device = 0;
  return d;
}
inline void setSleepThreshold(PxbSceneDesc& d, float t) { std::memcpy(&d.reserved[4], &t, 4); }   // PxRigidDynamic::setSleepThreshold (uniform)

// actor builders: PxCreateDynamic(box, density) + PxRigidBodyExt::updateMassAndInertia closed forms; PxCreatePlane
inline PxbActorRec dynamicBox(float x, float y, float z, float hx, float hy, float hz, float density = 10.f, uint32_t envId = 0xffffffffu) {
  PxbActorRec r; std::memset(&r, 0, sizeof(r));
  r.flags = PXB_ACTOR_DYNAMIC; r.geomType = PXB_GEOM_BOX; r.envId = envId;
  r.pos[0] = x; r.pos[1] = y; r.pos[2] = z; r.quat[3] = 1.f; r.dims[0] = hx; r.dims[1] = hy; r.dims[2] = hz;
  const float m = density * 8.f * hx * hy * hz, t = 1.f / 3.f;
  r.mass = m; r.inertia[0] = m * t * (hy * hy + hz * hz); r.inertia[1] = m * t * (hx * hx + hz * hz); r.inertia[2] = m * t * (hx * hx + hy * hy);
  r.angDamping = 0.05f; r.maxLinVel = 1.0e16f; r.maxAngVel = 100.f; r.maxDepenetrationVel = 1.0e32f;
  return r;
}
// PxQuat::getNormalized iterated to its fixed point: the SDK normalises every pose handed to createRigidStatic/Dynamic
// (physx/source/physx/src/NpPhysics.cpp), so actor records carry quaternions that normalisation leaves unchanged.
inline void normalizeQuat(float q[4]) {
  for (int it = 0; it < 8; ++it) {
    const float m = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), s = 1.0f / m;
    const float n[4] = {q[0] * s, q[1] * s, q[2] * s, q[3] * s};
    if (n[0] == q[0] && n[1] == q[1] && n[2] == q[2] && n[3] == q[3]) break;
    std::memcpy(q, n, sizeof(n));
  }
}
inline PxbActorRec groundPlane() {   // normal +Y: PxPlane(0,1,0,0) = local +X rotated 90 degrees about Z
  PxbActorRec r; std::memset(&r, 0, sizeof(r));
  r.geomType = PXB_GEOM_PLANE; r.envId = 0xffffffffu; r.quat[2] = std::sqrt(0.5f); r.quat[3] = std::sqrt(0.5f); normalizeQuat(r.quat);
  r.maxLinVel = 1.0e16f; r.maxAngVel = 100.f; r.maxDepenetrationVel = 1.0e32f; r.angDamping = 0.05f;
  return r;
}

struct State { float p[3], q[4], linVel[3], angVel[3]; };   // 13 floats per dynamic body, dynamic-body order
static_assert(sizeof(State) == 52, "packed state record");

class Scene {
 public:
  explicit Scene(const PxbSceneDesc& desc) { check(pxb_scene_create(&desc, &h_)); }
  ~Scene() { if (h_) pxb_scene_release(h_); }
  Scene(const Scene&) = delete; Scene& operator=(const Scene&) = delete;

  // cooked convex hulls (what the host's PxCreateConvexMesh produced; layout in physx_b200.h): once, before the actors that refer to them by hullIdx
  void setConvexMeshes(const void* cooked, size_t bytes, uint32_t nHulls) { check(pxb_scene_set_convex_meshes(h_, cooked, bytes, nHulls)); }
  void addActors(const std::vector<PxbActorRec>& recs) { if (!recs.empty()) check(pxb_scene_add_actors(h_, recs.data(), (uint32_t)recs.size())); }
  uint32_t getNbActors() const { return pxb_scene_num_actors(h_); }
  uint32_t getNbDynamics() const { return pxb_scene_num_dynamic(h_); }
  // PxScene::simulate / fetchResults
  void simulate(float dt) { check(pxb_scene_simulate(h_, dt)); }
  bool fetchResults(bool block = true) { const int rc = pxb_scene_fetch_results(h_, block ? 1 : 0); check(rc); return rc == 0; }
  // PxDirectGPUAPI (host buffers; *_device / *_async variants are in the C header)
  void getRigidDynamicData(void* data, int dataType, uint32_t nb, const uint32_t* indices = nullptr) { check(pxb_get_rigid_dynamic_data(h_, data, indices, dataType, nb)); }
  void setMassProperties(const uint32_t* indices, const float* massInertia4, uint32_t nb) { check(pxb_scene_set_mass_properties(h_, indices, massInertia4, nb)); }   // PxRigidBody::setMass / setMassSpaceInertiaTensor
  void setGravity(const float g[3]) { check(pxb_scene_set_gravity(h_, g)); }   // PxScene::setGravity
  // PxRigidDynamic::setKinematicTarget for a batch of kinematic bodies (dynamic-body indices, PxTransform rows q.xyzw p.xyz)
  void setKinematicTargets(const uint32_t* indices, const float* poses, uint32_t nb) { check(pxb_scene_set_kinematic_targets(h_, indices, poses, nb)); }
  // PxDirectGPUAPI::getRigidDynamicData(data, gpuIndices, dataType, nbElements, startEvent, finishEvent) on device memory (PxDirectGPUAPI.h:311-340)
  void getRigidDynamicDataDevice(void* devData, const uint32_t* devIndices, int dataType, uint32_t nb, void* startEvent = nullptr, void* finishEvent = nullptr) { check(pxb_get_rigid_dynamic_data_device_ev(h_, devData, devIndices, dataType, nb, startEvent, finishEvent)); }
  void setRigidDynamicDataDevice(const void* devData, const uint32_t* devIndices, int dataType, uint32_t nb, void* startEvent = nullptr, void* finishEvent = nullptr) { check(pxb_set_rigid_dynamic_data_device_ev(h_, devData, devIndices, dataType, nb, startEvent, finishEvent)); }
  void setRigidDynamicData(const void* data, int dataType, uint32_t nb, const uint32_t* indices = nullptr) { check(pxb_set_rigid_dynamic_data(h_, data, indices, dataType, nb)); }
  std::vector<State> getStates() { std::vector<State> s(getNbDynamics()); if (!s.empty()) check(pxb_scene_get_states(h_, &s[0].p[0])); return s; }
  void getSleepData(std::vector<float>& wakeCounters, std::vector<uint32_t>& asleep) {
    wakeCounters.resize(getNbDynamics()); asleep.resize(getNbDynamics());
    if (!asleep.empty()) check(pxb_scene_get_sleep_data(h_, wakeCounters.data(), asleep.data()));
  }
  // PxMaterial table (PxMaterial::setFrictionCombineMode / setRestitutionCombineMode / eDISABLE_FRICTION per entry; PxbActorRec.materialIndex refers to it): before the actors
  void setMaterials(const std::vector<PxbMaterial>& table) { if (!table.empty()) check(pxb_scene_set_materials(h_, table.data(), (uint32_t)table.size())); }
  // PxShape::setLocalPose / PxRigidBody::setCMassLocalPose for actors [first, first + n): 7 floats each (p.xyz, q.xyzw); poses in and out of the API stay actor poses
  void setLocalPoses(uint32_t first, uint32_t n, const float* shape2Actor, const float* body2Actor) { check(pxb_scene_set_local_poses(h_, first, n, shape2Actor, body2Actor)); }
  // PxDefaultSimulationFilterShader on the device: the extension's global state (PxSetGroupCollisionFlag / PxSetFilterOps / PxSetFilterBool / PxSetFilterConstants)
  // and the shapes' PxFilterData (4 words per actor); suppressed pairs stay broadphase pairs and generate no contacts
  void setFilterShader(const PxbFilterShaderConfig* config) { check(pxb_scene_set_filter_shader(h_, config)); }
  void setFilterData(uint32_t first, uint32_t n, const uint32_t* data4) { check(pxb_scene_set_filter_data(h_, first, n, data4)); }
  // PxShape::setContactOffset / setRestOffset per shape: 2 floats per actor (contactOffset, restOffset)
  void setShapeOffsets(uint32_t first, uint32_t n, const float* contactRest2) { check(pxb_scene_set_shape_offsets(h_, first, n, contactRest2)); }
  // PxScene::removeActor: the actors leave the simulation at the next step, indices stay valid
  void removeActors(const std::vector<uint32_t>& actors) { if (!actors.empty()) check(pxb_scene_remove_actors(h_, actors.data(), (uint32_t)actors.size())); }
  // contact reports of the last step: pairs that started / stopped touching (eNOTIFY_TOUCH_FOUND / eNOTIFY_TOUCH_LOST), (a, b) actor indices with a < b
  std::vector<uint32_t> getTouchFound() { std::vector<uint32_t> p(2 * (size_t)pxb_scene_num_touch_found(h_)); if (!p.empty()) check(pxb_scene_get_touch_found(h_, p.data())); return p; }
  std::vector<uint32_t> getTouchLost() { std::vector<uint32_t> p(2 * (size_t)pxb_scene_num_touch_lost(h_)); if (!p.empty()) check(pxb_scene_get_touch_lost(h_, p.data())); return p; }
  // PxDirectGPUAPI::copyContactData: PxGpuContactPair records into DEVICE memory (enable before the step whose contacts are wanted)
  void enableContactData(bool on = true) { check(pxb_scene_enable_contact_data(h_, on ? 1 : 0)); }
  void copyContactData(void* deviceData, uint32_t* deviceNbContactPairs, uint32_t maxPairs) { check(pxb_scene_copy_contact_data(h_, deviceData, deviceNbContactPairs, maxPairs)); }
  // fused state export: the step stores the packed State block of every dynamic body into the given device / peer-mapped / mapped pinned host buffers
  void setStateExport(const std::vector<void*>& targets, uint32_t rowOffset = 0) { check(pxb_scene_set_state_export(h_, targets.empty() ? nullptr : targets.data(), (uint32_t)targets.size(), rowOffset)); }
  void sync() { check(pxb_scene_sync(h_)); }
  uint32_t getNbPairs() { return pxb_scene_num_pairs(h_); }
  uint32_t getNbConstraints() { return pxb_scene_last_num_constraints(h_); }
  uint32_t getNbPartitions() { return pxb_scene_last_num_partitions(h_); }
  bool usesEnvironmentPath() { return pxb_scene_uses_env_path(h_) != 0; }
  PxbScene* handle() { return h_; }

 private:
  PxbScene* h_ = nullptr;
};

}  // namespace pxb
