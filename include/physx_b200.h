/* physx_b200.h -- C ABI of the B200-native rigid-body step (libphysx_b200.so).
 *
 * This is the boundary the PhysXGpu plugin shim binds (INTEGRATION.md shows the shim): plain pointers
 * and sizes, no C++ or torch types.  Each entry point names the reference interface it stands in for.
 * All functions return 0 (PXB_OK) on success or a negative PxbError; pxb_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with PXB_ERR_NO_DEVICE.
 *
 * Pointers are HOST pointers unless the name says "_device"/"dev".  Actor records are PxbActorRec below (128 bytes; the test
 * oracle reads the same layout so scenes are bit-identical on both sides of a test).
 */
#ifndef PHYSX_B200_H
#define PHYSX_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXB_API __attribute__((visibility("default")))

typedef enum {
  PXB_OK = 0,
  PXB_ERR_NO_DEVICE = -1,     /* no CUDA device / driver: the product never falls back to the CPU */
  PXB_ERR_INVALID = -2,
  PXB_ERR_CUDA = -3,          /* CUDA runtime failure; scene enters abort mode like PxCudaContext (cudamanager/src/CudaContextManager.cpp:845) */
  PXB_ERR_CAPACITY = -4,      /* pair / contact capacity exceeded (PxGpuDynamicsMemoryConfig analogue) */
  PXB_ERR_UNSUPPORTED = -5
} PxbError;

/* PxGeometryType values used by the hot path (physx/include/geometry/PxGeometry.h:48-62) */
enum { PXB_GEOM_SPHERE = 0, PXB_GEOM_PLANE = 1, PXB_GEOM_CAPSULE = 2, PXB_GEOM_BOX = 3, PXB_GEOM_CONVEXMESH = 5 };
enum { PXB_ACTOR_DYNAMIC = 1u,
       PXB_ACTOR_KINEMATIC = 2u /* together with PXB_ACTOR_DYNAMIC: PxRigidBodyFlag::eKINEMATIC.  The body has infinite mass, is moved by pxb_scene_set_kinematic_targets and stands still in steps
                                   without a target; it pushes / carries dynamic bodies with the velocity of its move (ScKinematics.cpp:44-97, DyTGSContactPrep.cpp:406-409, :724-727) and
                                   generates no pairs against statics or other kinematics (PxPairFilteringMode::eDEFAULT).  It keeps its place in the dynamic-body order.  TGS and PGS, scenes
                                   without sleeping; both paths (on the environment path the fused state export gives way to the export kernel). */,
       PXB_ACTOR_DISABLE_GRAVITY = 4u, /* PxActorFlag::eDISABLE_GRAVITY (bodyCoreComputeUnconstrainedVelocity, DyBodyCoreIntegrator.h:55-59) */
       PXB_ACTOR_GYROSCOPIC = 8u       /* PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES (copyToSolverBodyDataStep DyTGSDynamics.cpp:177-193, copyToSolverBodyData DyRigidBodyToSolverBody.cpp:53-70) */ };
enum { PXB_SOLVER_PGS = 0, PXB_SOLVER_TGS = 1 };

/* Scene description: the PxSceneDesc / PxGpuDynamicsMemoryConfig fields the hot path consumes
 * (physx/include/PxSceneDesc.h:473-487, :1023-1075). */
typedef struct {
  float    gravity[3];
  uint32_t solverType;                 /* PXB_SOLVER_TGS or PXB_SOLVER_PGS (PxSceneDesc::solverType) */
  float    bounceThresholdVelocity;    /* PxSceneDesc::bounceThresholdVelocity */
  float    frictionOffsetThreshold;    /* PxSceneDesc::frictionOffsetThreshold */
  float    frictionCorrelationDistance;/* PxSceneDesc::frictionCorrelationDistance */
  float    toleranceLength;            /* PxTolerancesScale::length */
  float    staticFriction, dynamicFriction, restitution; /* the (single) combined material */
  float    contactOffset, restOffset;  /* PxShape::setContactOffset / setRestOffset (uniform) */
  uint32_t posIters, velIters;         /* PxRigidDynamic::setSolverIterationCounts (scene max) */
  uint32_t maxActors;                  /* capacity */
  uint32_t maxPairs;                   /* PxGpuDynamicsMemoryConfig::foundLostPairsCapacity / maxRigidPatchCount analogue; 0 = 8*maxActors */
  int32_t  device;                     /* CUDA device ordinal (PxCudaContextManagerDesc) */
  uint32_t reserved[8];                /* [0] internal; [1] = PXB_FLAG_* bits; [2]/[3] = test hooks of the environment path: constraint-list slots in shared memory / constraints per CTA (0 = adaptive);
                                          [4] = sleep threshold as IEEE-754 float bits (PxRigidDynamic::setSleepThreshold, uniform; 0 = sleeping off) */
} PxbSceneDesc;
/* Scenes whose dynamic actors all carry an environment id (PxActor::setEnvironmentID, the RL many-env layout of
 * BASELINE configs 2/5) run on the environment path: one warp / one CTA per environment with solver rows in shared
 * memory (physx_b200/csrc/pxb_env.cuh).  Results are bit-identical to the device-wide path; this flag forces the latter. */
enum { PXB_FLAG_NO_ENV_PATH = 1u,
/* Documented relaxation for giant islands (BASELINE configs 3 / 4): the reference's partitioning is a sequential first-fit in
 * solver input order (DyConstraintPartition.cpp:475-568) whose dependency chains grow with the island; with this flag the
 * device-wide path colours constraints with Jones-Plassmann rounds instead (a valid, deterministic partitioning that is NOT the
 * reference's first-fit, so trajectories are comparable to the reference only statistically).  Off by default. */
       PXB_FLAG_RELAXED_PARTITIONING = 2u,
/* PxSceneFlag::eENABLE_BODY_ACCELERATIONS (PxSceneDesc.h): the step keeps the velocities it started from (integrationTGS.cu:124-131,
 * mBodySimPrevVelocitiesBufferDeviceData) so that PXB_RD_LINEAR_ACCELERATION / PXB_RD_ANGULAR_ACCELERATION can be read.  Off by default (two copies of 16 B per actor and step). */
       PXB_FLAG_BODY_ACCELERATIONS = 4u };

/* One actor = one rigid body (or static) with one shape, local shape pose = identity; 128 bytes, little endian.
 * Planes: the actor pose carries the plane frame, normal = local +X (PxPlaneGeometry).  Mass properties are explicit
 * (PxRigidBody::setMass / setMassSpaceInertiaTensor); the remaining fields are the PxRigidDynamic setters of the same name. */
typedef struct {
  uint32_t flags;       /* PXB_ACTOR_DYNAMIC or 0 (static) */
  uint32_t geomType;    /* PXB_GEOM_* */
  uint32_t envId;       /* PxActor::setEnvironmentID, 0xffffffff = none */
  uint32_t hullIdx;     /* reserved for convex meshes */
  float    pos[3];
  float    quat[4];     /* x y z w, normalised */
  float    dims[4];     /* sphere: r; capsule: r, halfHeight; box: hx hy hz */
  float    linVel[3];
  float    angVel[3];
  float    mass;
  float    inertia[3];  /* mass-space diagonal */
  float    linDamping, angDamping;
  float    maxLinVel, maxAngVel;
  float    maxDepenetrationVel;
  uint32_t materialIndex; /* index into the table of pxb_scene_set_materials (ignored while no table is set) */
  uint32_t aggregate;   /* PxAggregate membership: 0 = none, k > 0 = aggregate k; bit 31 set = the aggregate has self collisions enabled.  Members of an aggregate without self collisions
                            generate no pairs among themselves (PxGetAggregateFilterHint, BpAABBManager.cpp aggregate self-collision pairs) */
} PxbActorRec;

typedef struct PxbScene PxbScene;

/* ---- lifecycle: PxPhysXGpu factory set (physx/source/physxgpu/include/PxPhysXGpu.h:101-200;
 *      createGpuBroadPhase / createGpuNphaseImplementationContext / createGpuDynamicsContext /
 *      createGpuSimulationController in PxgPhysXGpu.cpp:153-262) ---- */
PXB_API int  pxb_scene_create(const PxbSceneDesc* desc, PxbScene** out);
PXB_API void pxb_scene_release(PxbScene* scene);
PXB_API const char* pxb_last_error(void);
PXB_API int  pxb_device_count(void);

/* PxsSimulationController::addDynamics / addPxgShape + Bp::AABBManagerBase::addBounds
 * (lowlevel/software/include/PxsSimulationController.h:122-354, lowlevelaabb/include/BpAABBManagerBase.h:175-330).
 * `recs` = nb records of 128 bytes (PxbActorRec layout).  Actor index = order of addition. */
/* ---- cooked convex hulls.  Replaces: PxsSimulationController::addPxgShape + the hull upload of PxgShapeManager (the reference GPU pipeline
 *      receives Gu::ConvexHullData cooked on the host, S/gpunarrowphase/include/PxgConvexConvexShape.h:50-65, S/geomutils/src/convex/GuConvexMeshData.h:47-175).
 *      `cooked` = nHulls records, each: PxbCookedHullHeader, float verts[nVerts][3], PxbCookedPoly polys[nPolys],
 *      uint8_t vertexRefs[nIdx] (getVertexData8, padded to 4 bytes), uint8_t facesByEdges[2 * nEdges] (getFacesByEdges8, padded to 4 bytes).
 *      A hull of more than 32 vertices also carries its hill-climbing data (Gu::BigConvexRawData, S/geomutils/src/convex/GuBigConvexData.h:54-75):
 *      PxbCookedHullHeader::reserved[0] = mSubdiv | mNbAdjVerts << 16 (0 = none), and after facesByEdges: uint8_t samples[6 * subdiv^2] (padded
 *      to 4), uint16_t valencies[nVerts][2] ({mCount, mOffset}), uint8_t adjacentVerts[nAdj] (padded to 4).
 *      Call once, before adding the actors whose PxbActorRec::hullIdx refer to it.  Limits: the reference's GPU-compatible hulls -- at most 64
 *      vertices and 64 polygons (include/cooking/PxConvexMeshDesc.h:139), at most 32 vertices per polygon; larger -> PXB_ERR_UNSUPPORTED.
 *      Identity mesh scale.  Every hull pair type is supported: plane / sphere / capsule / box / hull vs hull. ---- */
typedef struct {
  uint32_t nVerts, nPolys, nEdges, nIdx;
  float centerOfMass[3];
  float boundsCenter[3], boundsExtents[3];
  float internalRadius, internalExtents[3];
  float unitMass, unitInertiaDiag[3], unitCom[3];
  uint32_t reserved[1];
} PxbCookedHullHeader;
typedef struct { float plane[4]; uint32_t vref, nbVerts, minIndex, pad; } PxbCookedPoly;
PXB_API int  pxb_scene_set_convex_meshes(PxbScene* scene, const void* cooked, size_t bytes, uint32_t nHulls);
/* Material table (a11: materialCombiner.cuh:35 / PxsCombineMaterials, lowlevel/software/include/PxsMaterialCombiner.h:69-175).  Each actor's shape
 * refers to one entry (PxbActorRec::materialIndex); a pair's friction coefficients and restitution are combined from its two materials with the
 * larger of their PxCombineMode values (average, min, multiply, max), static friction is raised to the dynamic one, and
 * PxMaterialFlag::eDISABLE_FRICTION on either side removes the friction rows.  bits = frictionCombineMode | restitutionCombineMode << 4 |
 * flags << 8 (bit 0: eDISABLE_FRICTION).  Without a table every pair uses PxbSceneDesc's staticFriction / dynamicFriction / restitution.
 * Compliant contacts (negative restitution), eDISABLE_STRONG_FRICTION, eIMPROVED_PATCH_FRICTION -> PXB_ERR_UNSUPPORTED.  Call before the
 * first simulate. */
typedef struct { float staticFriction, dynamicFriction, restitution; uint32_t bits; } PxbMaterial;
PXB_API int  pxb_scene_set_materials(PxbScene* scene, const PxbMaterial* materials, uint32_t nb);
PXB_API int  pxb_scene_add_actors(PxbScene* scene, const void* recs, uint32_t nb);
/* Bp::AABBManagerBase::removeBounds (BpAABBManagerBase.h:191) + PxsSimulationController::removeDynamic: the listed actors leave the simulation at
 * the next step (their pairs are reported deleted / touch-lost, they are no longer integrated); actor and dynamic-body indices stay valid. */
PXB_API int  pxb_scene_remove_actors(PxbScene* scene, const uint32_t* actorIndices, uint32_t nb);
PXB_API uint32_t pxb_scene_num_actors(const PxbScene* scene);
PXB_API uint32_t pxb_scene_num_dynamic(const PxbScene* scene);

/* ---- the step: PxScene::simulate()+fetchResults() hot path = Sc::Scene collide + advance chains
 *      (simulationcontroller/src/ScPipeline.cpp:108-124), i.e. bounds -> Bp::BroadPhase::update ->
 *      PxvNphaseImplementationContext::updateContactManager -> Dy::Context::update/updatePostPartitioning ->
 *      integration.  pxb_scene_simulate only enqueues GPU work; pxb_scene_fetch_results waits for it. ---- */
PXB_API int  pxb_scene_simulate(PxbScene* scene, float dt);
PXB_API int  pxb_scene_fetch_results(PxbScene* scene, int block);

/* Constraint input order for the NEXT simulate call: n (actorA, actorB) pairs in the order the island
 * manager hands contact managers to the solver (IG::IslandSim edge lists as walked by
 * DynamicsTGSContext::prepareBodiesAndConstraints, DyTGSDynamics.cpp:822-905; the plugin shim reads them
 * from the IG::SimpleIslandManager& it is given).  n = 0 restores the canonical order (sorted pair key). */
PXB_API int  pxb_scene_set_constraint_order(PxbScene* scene, const uint32_t* pairs, uint32_t n);

/* ---- PxDirectGPUAPI mirror (physx/include/PxDirectGPUAPI.h:311-463; kernels updateBodiesAndShapes.cu:999-1253).
 *      dataType: 0 = global pose (7 floats q.xyzw,p.xyz = PxTransform), 1 = linear velocity, 2 = angular velocity,
 *      3 = force, 4 = torque (writes only).
 *      `indices` are dynamic-body indices (PxRigidDynamicGPUIndex analogue), NULL = 0..nb-1.
 *      *_device variants take device pointers and run on the scene stream without synchronising. ---- */
enum { PXB_RD_GLOBAL_POSE = 0, PXB_RD_LINEAR_VELOCITY = 1, PXB_RD_ANGULAR_VELOCITY = 2,
       PXB_RD_FORCE = 3, PXB_RD_TORQUE = 4, /* set only (PxRigidDynamicGPUAPIWriteType::eFORCE / eTORQUE, PxDirectGPUAPI.h:60-72): 3 floats per body, world frame, applied at the centre of mass by the next simulate only */
       PXB_RD_LINEAR_ACCELERATION = 5, PXB_RD_ANGULAR_ACCELERATION = 6 /* get only (PxRigidDynamicGPUAPIReadType::eLINEAR_ACCELERATION / eANGULAR_ACCELERATION; kernels getRigidDynamicLinearAcceleration /
                                                                          AngularAcceleration, updateBodiesAndShapes.cu:1063-1106): (velocity - velocity the last step started from) * (1 / dt), 3 floats; needs PXB_FLAG_BODY_ACCELERATIONS */ };
PXB_API int  pxb_get_rigid_dynamic_data(PxbScene* scene, void* data, const uint32_t* indices, int dataType, uint32_t nb);
PXB_API int  pxb_set_rigid_dynamic_data(PxbScene* scene, const void* data, const uint32_t* indices, int dataType, uint32_t nb);
/* stream-ordered host variants: PINNED host buffers, no index list, no synchronisation; complete at the next
 * pxb_scene_fetch_results / pxb_scene_sync (PxDirectGPUAPI is asynchronous as well: start / finish CUevents) */
PXB_API int  pxb_get_rigid_dynamic_data_async(PxbScene* scene, void* pinnedData, int dataType, uint32_t nb);
PXB_API int  pxb_set_rigid_dynamic_data_async(PxbScene* scene, const void* pinnedData, int dataType, uint32_t nb);
PXB_API int  pxb_scene_sync(PxbScene* scene);
PXB_API int  pxb_get_rigid_dynamic_data_device(PxbScene* scene, void* devData, const uint32_t* devIndices, int dataType, uint32_t nb);
PXB_API int  pxb_set_rigid_dynamic_data_device(PxbScene* scene, const void* devData, const uint32_t* devIndices, int dataType, uint32_t nb);
/* The same two calls with the start / finish events of PxDirectGPUAPI::getRigidDynamicData / setRigidDynamicData (PxDirectGPUAPI.h:311-370; CUevent = cudaEvent_t):
 * the work waits for `startEvent` (NULL = none) and records `finishEvent` once done; without a finish event the call synchronises, as the reference does (PxgSimulationCore.cpp:2736-2850). */
PXB_API int  pxb_get_rigid_dynamic_data_device_ev(PxbScene* scene, void* devData, const uint32_t* devIndices, int dataType, uint32_t nb, void* startEvent, void* finishEvent);
PXB_API int  pxb_set_rigid_dynamic_data_device_ev(PxbScene* scene, const void* devData, const uint32_t* devIndices, int dataType, uint32_t nb, void* startEvent, void* finishEvent);
/* PxRigidDynamic::setKinematicTarget (physx/include/PxRigidDynamic.h; NpRigidDynamic.cpp:129-158) for nb kinematic bodies: `indices` are dynamic-body indices, `poses` nb x 7 floats
 * (PxTransform: q.xyzw, p.xyz; actor poses, normalised by the call).  The next pxb_scene_simulate moves the bodies there.  The reference's GPU pipeline receives the same data through
 * PxsSimulationController::updateDynamic after Sc::Scene::kinematicsSetup (ScKinematics.cpp:117-160).  The _device variant takes device pointers, is stream-ordered on the scene stream and
 * reports a bad index through the next pxb_scene_fetch_results. */
PXB_API int  pxb_scene_set_kinematic_targets(PxbScene* scene, const uint32_t* indices, const float* poses, uint32_t nb);
PXB_API int  pxb_scene_set_kinematic_targets_device(PxbScene* scene, const uint32_t* devIndices, const float* devPoses, uint32_t nb);
/* PxRigidBody::setMass / setMassSpaceInertiaTensor at run time for nb dynamic bodies (dynamic-body indices; 4 floats per body: mass, inertia xyz; 0 = infinite): domain randomisation
 * of the mass properties between steps.  Not for kinematic bodies. */
PXB_API int  pxb_scene_set_mass_properties(PxbScene* scene, const uint32_t* indices, const float* massInertia4, uint32_t nb);
/* PxScene::setGravity (physx/include/PxScene.h; NpScene.cpp:331-342): takes effect with the next pxb_scene_simulate. */
PXB_API int  pxb_scene_set_gravity(PxbScene* scene, const float* gravity3);
/* Packed state convenience: 13 floats per dynamic body (pos3 quat4 linVel3 angVel3), dynamic-body order. */
PXB_API int  pxb_scene_get_states(PxbScene* scene, float* out);
PXB_API int  pxb_scene_set_states(PxbScene* scene, const float* in);
PXB_API int  pxb_scene_get_states_device(PxbScene* scene, float* devOut); /* same record, device pointer; stream-ordered on the scene stream (may follow pxb_scene_simulate
                                                                              without a fetch: it packs the state that step produces, e.g. into an NCCL send buffer) */
/* ---- tensor front end: device tensors in the wire formats of the reference's ovphysx bindings
 *      (ovphysx/include/ovphysx/ovphysx.h, python/ovphysx/types.py TensorType; same enum values).  Rows follow dynamic-body order or `devIdx`.
 *      Stream-ordered on the scene stream (no synchronisation); DLPack producers hand over the device pointer (physx_b200/tensor_api.py). ---- */
enum { PXB_TENSOR_RIGID_BODY_POSE = 1,      /* [N,7] (p.xyz, q.xyzw), read / write */
       PXB_TENSOR_RIGID_BODY_VELOCITY = 2,  /* [N,6] (linear, angular), read / write */
       PXB_TENSOR_RIGID_BODY_MASS = 3,      /* [N] read */
       PXB_TENSOR_RIGID_BODY_INV_MASS = 7,  /* [N] read */
       PXB_TENSOR_RIGID_BODY_FORCE = 50,    /* [N,3] world-frame force at the centre of mass, write: applied by the next simulate only */
       PXB_TENSOR_RIGID_BODY_WRENCH = 51 }; /* [N,9] (force, torque, application point in the world frame), write */
PXB_API int  pxb_tensor_read_device(PxbScene* scene, int tensorType, void* devOut, const uint32_t* devIdx, uint32_t nb);
PXB_API int  pxb_tensor_write_device(PxbScene* scene, int tensorType, const void* devIn, const uint32_t* devIdx, uint32_t nb);
/* Multi-GPU exchange helper: ONE kernel stores `bytes` from devSrc into nDst <= 8 peer-mapped device buffers over NVLink (P2P). */
PXB_API int  pxb_scatter_to_peers(PxbScene* scene, void* cudaStream, const void* devSrc, size_t bytes, const uint64_t* devDstPtrs, uint32_t nDst, uint32_t ctas);
/* Fused state export.  From the next pxb_scene_simulate on, the step's integration epilogue stores every dynamic body's packed state record
 * (13 floats, dynamic-body order, as pxb_scene_get_states) into each of nDst <= 9 device-accessible buffers at row `rowOffset`: memory of this
 * GPU, PEER-MAPPED memory of other GPUs (P2P stores over NVLink, e.g. a symmetric-memory tensor: the per-step all-gather of the many-environment
 * layout then needs no pack kernel and no copy) or MAPPED PINNED HOST memory (the device-to-host transfer overlaps the step instead of following
 * it).  Stands in for PxDirectGPUAPI::getRigidDynamicData (PxDirectGPUAPI.h:311-360; getRigidDynamicGlobalPose / LinearVelocity / AngularVelocity
 * kernels, updateBodiesAndShapes.cu:999-1253) when the consumer reads every body every step.  The targets may change every call (double
 * buffering); nDst = 0 switches the export off.  Data is complete when the step is (pxb_scene_fetch_results, or stream order on pxb_scene_stream). */
PXB_API int  pxb_scene_set_state_export(PxbScene* scene, void* const* dst, uint32_t nDst, uint32_t rowOffset);
/* Cross-GPU flags for exchanges built on the export: pxb_peer_signal stores `value` into n <= 9 (peer-mapped) 32-bit flags with ONE kernel on
 * `cudaStream` (NULL = the scene stream), ordered after everything enqueued before it at system scope; pxb_peer_wait makes the stream wait until
 * n <= 32 consecutive LOCAL flags have all reached `value` (wrap-around compare). */
PXB_API int  pxb_peer_signal(PxbScene* scene, void* cudaStream, const uint64_t* flagPtrs, uint32_t n, uint32_t value);
PXB_API int  pxb_peer_wait(PxbScene* scene, void* cudaStream, const void* devFlags, uint32_t n, uint32_t value);
PXB_API void* pxb_scene_state_device_ptr(PxbScene* scene, int which); /* 0 pos4, 1 quat4, 2 linVel4, 3 angVel4 (per ACTOR float4 arrays) */
PXB_API void* pxb_scene_stream(PxbScene* scene);                        /* cudaStream_t */

/* ---- stage-level entry points (parity tests drive these with teacher-forced inputs) ----
 * Bp::BroadPhase::update + getCreatedPairs/getDeletedPairs (lowlevelaabb/include/BpBroadPhase.h:98-222).
 * `tightBounds` = 6 floats per actor (min xyz, max xyz) or NULL to compute them from the current poses
 * (Gu::computeBounds, geomutils/src/GuBounds.cpp:354-400).  Pairs are (a,b) actor indices with a<b. */
PXB_API int  pxb_scene_compute_bounds(PxbScene* scene);
PXB_API int  pxb_scene_get_bounds(PxbScene* scene, float* out6);
PXB_API int  pxb_scene_broadphase(PxbScene* scene, const float* tightBounds);
PXB_API uint32_t pxb_scene_num_pairs(PxbScene* scene);
PXB_API int  pxb_scene_get_pairs(PxbScene* scene, uint32_t* outPairs);       /* sorted */
PXB_API uint32_t pxb_scene_num_created(PxbScene* scene);
PXB_API uint32_t pxb_scene_num_deleted(PxbScene* scene);
PXB_API int  pxb_scene_get_created(PxbScene* scene, uint32_t* outPairs);     /* sorted */
PXB_API int  pxb_scene_get_deleted(PxbScene* scene, uint32_t* outPairs);     /* sorted */
/* Touch events of the last step (a7: prepareLostFoundPairs_Stage1 / 2, gpunarrowphase/src/CUDA/cudaGJKEPA.cu:1468,1532 -> the found / lost patch
 * lists PxgNphaseImplementationContext hands to the island manager and to contact reports): pairs that started / stopped producing contacts,
 * incl. touching pairs that left the broadphase.  (a,b) actor indices with a<b, sorted.  One patch per pair, so "patch count changed" = these. */
PXB_API uint32_t pxb_scene_num_touch_found(PxbScene* scene);
PXB_API uint32_t pxb_scene_num_touch_lost(PxbScene* scene);
PXB_API int  pxb_scene_get_touch_found(PxbScene* scene, uint32_t* outPairs);
PXB_API int  pxb_scene_get_touch_lost(PxbScene* scene, uint32_t* outPairs);
/* Contacts of the last step, one 24-float record per pair in pair order:
 * [count, nx, ny, nz, 4 x (px, py, pz, separation, appliedForce)]; normal points body1 -> body0
 * (PxsContactManagerOutput / PxContactPatch + PxContact stream analogue, physx/include/PxContact.h:57-148). */
PXB_API int  pxb_scene_get_contacts(PxbScene* scene, float* out24);
/* PxDirectGPUAPI::copyContactData (physx/include/PxDirectGPUAPI.h:388-401; reference kernels compressContactStage1 / 2,
 * gpunarrowphase/src/CUDA/compressOutputContacts.cu:48-275): one PxGpuContactPair record (physx/include/PxContact.h:818-833, 80 bytes, same field
 * layout) per pair that produced contacts in the last step, in pair order, written to DEVICE memory `data` (maxPairs records); the device word
 * `nbContactPairs` receives the number of touching pairs (it can exceed maxPairs: only maxPairs records are written).  The pointers inside a record
 * point into streams owned by the scene and stay valid until the next simulate: contactPatches -> one PxContactPatch (64 B: normal, combined
 * restitution / dynamic / static friction, nbContacts, material indices), contactPoints -> nbContacts PxContact (point, separation), contactForces ->
 * nbContacts applied normal impulses, frictionPatches -> one PxFrictionPatch (52 B: world anchor positions, anchor impulses, anchor count; the
 * reference's writeBackContactBlockFriction, gpusolver/src/CUDA/solverBlockCommon.cuh:33-66).  transformCacheRef0/1 = actor indices (one shape per
 * actor), nodeIndex0/1 = PxNodeIndex with mID = PxRigidDynamicGPUIndex (0xffffffff for a static actor), actor0/1 = the actor index as a handle.
 * Stream-ordered on the scene stream (read `data` after pxb_scene_sync or on that stream).  The friction write-back costs a few loads per
 * constraint, so it is off by default: pxb_scene_enable_contact_data(scene, 1) BEFORE the step whose contacts are wanted. */
/* f1: the default simulation filter shader on the device (PxDefaultSimulationFilterShader, physxextensions/src/ExtDefaultSimulationFilterShader.cpp:238-280, without
 * the trigger branch).  Config = the extension's global state: collisionTable[g] bit h = PxGetGroupCollisionFlag(g, h); ops[3] = PxSetFilterOps (PxFilterOp: 0 AND,
 * 1 OR, 2 XOR, 3 NAND, 4 NOR, 5 NXOR, 6 SWAP_AND); filterBool = PxSetFilterBool; constants = PxSetFilterConstants (K0 word2, K0 word3, K1 word2, K1 word3).  Filter
 * data = the shape's PxFilterData (word0 = collision group 0..31, word2 / word3 = PxGroupsMask), 4 words per actor.  A pair the shader answers eSUPPRESS for stays a
 * broadphase pair (created / deleted lists unchanged) and generates no contacts, as in the reference (no contact manager).  NULL config = filtering off. */
typedef struct PxbFilterShaderConfig { uint32_t collisionTable[32]; uint32_t ops[3]; uint32_t filterBool; uint32_t constants[4]; } PxbFilterShaderConfig;
PXB_API int  pxb_scene_set_filter_shader(PxbScene* scene, const PxbFilterShaderConfig* config);
PXB_API int  pxb_scene_set_filter_data(PxbScene* scene, uint32_t firstActor, uint32_t n, const uint32_t* data4);
/* PxShape::setContactOffset / setRestOffset per shape, for actors [firstActor, firstActor + n): 2 floats each (contactOffset, restOffset).  Without this call every
 * shape has PxbSceneDesc's contactOffset / restOffset.  The broadphase inflates every bound by its own shape's contact offset (Bp contact distances,
 * BpBroadPhaseABP.cpp:1187-1197), a pair's contact distance is the sum of the two contact offsets and its rest distance the sum of the two rest offsets (PxcNpWorkUnit). */
PXB_API int  pxb_scene_set_shape_offsets(PxbScene* scene, uint32_t firstActor, uint32_t n, const float* contactRest2);
/* Local poses (a1: PxgShapeSim.shape2Actor, PxsBodyCore.body2Actor): PxShape::setLocalPose and PxRigidBody::setCMassLocalPose for actors [firstActor, firstActor + n),
 * 7 floats each (p.xyz, q.xyzw; stored normalised like the reference).  The actor keeps its pose (Sc::BodyCore::setCMassLocalPose, ScBodyCore.cpp:98-108); the body
 * frame the solver integrates becomes actorPose * body2Actor, the shape's world pose in the transform cache body2World * (body2Actor^-1 * shape2Actor)
 * (Sc::ShapeSimBase::getAbsPoseAligned, ScShapeSimBase.cpp:214-249 / updateCacheAndBound.cuh:67-160).  Every pose the API reports or accepts stays the ACTOR pose
 * (PxRigidActor::getGlobalPose); PxbActorRec.inertia is in the body frame.  Not combinable with pxb_scene_set_state_export when a body2Actor is not the identity. */
PXB_API int  pxb_scene_set_local_poses(PxbScene* scene, uint32_t firstActor, uint32_t n, const float* shape2Actor, const float* body2Actor);
typedef struct PxbGpuContactPair {
  uint8_t* contactPatches; uint8_t* contactPoints; float* contactForces; uint8_t* frictionPatches;
  uint32_t transformCacheRef0, transformCacheRef1;
  uint64_t nodeIndex0, nodeIndex1;
  uint64_t actor0, actor1;
  uint16_t nbContacts, nbPatches; uint32_t pad;
} PxbGpuContactPair;
PXB_API int  pxb_scene_enable_contact_data(PxbScene* scene, int enable);
PXB_API int  pxb_scene_copy_contact_data(PxbScene* scene, void* data, uint32_t* nbContactPairs, uint32_t maxPairs);
/* Solver statistics of the last step (PxSimulationStatistics::mNbPartitions analogue). */
PXB_API uint32_t pxb_scene_last_num_partitions(PxbScene* scene);
PXB_API uint32_t pxb_scene_last_num_constraints(PxbScene* scene);
/* Optional per-stage device timing of the last step with CUDA events on the scene stream (no reference
 * analogue: the reference has no device-side timers).  ms7 = [broadphase, narrowphase,
 * colouring, prep, solve, writeback+integration, whole step]. */
PXB_API int  pxb_scene_set_profiling(PxbScene* scene, int enable);
PXB_API int  pxb_scene_get_stage_times(PxbScene* scene, float* ms7);
/* number of kernels launched by the last pxb_scene_simulate call */
PXB_API uint32_t pxb_scene_last_num_launches(PxbScene* scene);
/* Sleeping (a18: integrateCoreParallelLaunchTGS sleepCheck / updateWakeCounter, DySleep.cpp:35-236): per DYNAMIC body the wake counter
 * (PxRigidDynamic::getWakeCounter) and 1 if asleep (isSleeping).  Islands whose bodies are all ready are put to sleep on the device. */
PXB_API int  pxb_scene_get_sleep_data(PxbScene* scene, float* wakeCounters, uint32_t* asleep);
/* 1 if the last step ran on the environment path, 0 for the device-wide path */
PXB_API int  pxb_scene_uses_env_path(PxbScene* scene);

/* ---- standalone broadphase object: a2-a6 behind the reference's own Bp::BroadPhase interface (lowlevelaabb/include/BpBroadPhase.h:98-222).
 *      The plugin shim's Bp::BroadPhase subclass (plugin/PxgB200BroadPhase.cpp) forwards Bp::BroadPhaseUpdateData (BpBroadPhaseUpdate.h:48-140)
 *      to pxb_bp_update and serves getCreatedPairs / getDeletedPairs from pxb_bp_fetch.  Objects are slots of the host arrays: bounds6 = PxBounds3
 *      per slot (tight), contactDist, groups = Bp::FilterGroup values (3-bit Bp::FilterType in the low bits), envIds or NULL, lut49 = the 7 x 7
 *      BpFilter table (BpFiltering.h:116-128) as bytes.  Pair semantics of the reference: closed-interval overlap of the bounds inflated by the
 *      contact distance, groups differ, type table, equal-or-invalid environment ids; created = new overlaps, deleted = lost overlaps between
 *      objects that are still in the broadphase (BpBroadPhase.h:160-192).  Pairs are (a, b) with a < b, sorted. ---- */
typedef struct PxbBroadPhase PxbBroadPhase;
PXB_API int  pxb_bp_create(uint32_t maxObjects, uint32_t maxPairs, int device, PxbBroadPhase** out);
PXB_API void pxb_bp_release(PxbBroadPhase* bp);
PXB_API int  pxb_bp_update(PxbBroadPhase* bp, const float* bounds6, const float* contactDist, const uint32_t* groups, const uint32_t* envIds, uint32_t capacity, const uint8_t* lut49,
                           const uint32_t* created, uint32_t nCreated, const uint32_t* updated, uint32_t nUpdated, const uint32_t* removed, uint32_t nRemoved);
PXB_API int  pxb_bp_fetch(PxbBroadPhase* bp, const uint32_t** createdPairs, uint32_t* nCreated, const uint32_t** deletedPairs, uint32_t* nDeleted);

#ifdef __cplusplus
}
#endif
#endif
