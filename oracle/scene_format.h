/* Scene interchange format shared by the oracle tools (test infrastructure).
 *
 * A scene file is:  SceneHeader | ActorRec[nActors] | PxbMaterialRec[header.reserved[2]] (material table, may be empty)
 *                   | (when header.reserved[3] == PXB_LOCAL_POSE_MAGIC) PxbLocalPoseRec[nActors]
 *                   | (when header.reserved[0] & PXB_FLAG_FILTER_SECTION) PxbFilterShaderConfig, then uint32 filterData[nActors][4] (PxFilterData word0..3 per shape)
 *                   | (when header.reserved[0] & PXB_FLAG_SHAPE_OFFSETS) float shapeOffsets[nActors][2] = PxShape::setContactOffset / setRestOffset per shape
 *                     (without the section every shape has the header's contactOffset / restOffset)
 *                   | for each hull: u32 nVerts, float xyz[nVerts]
 *                   | (when header.reserved[1] == PXB_COOKED_MAGIC) for each hull: PxbCookedHullHeader + arrays (see below)
 * The cooked section is what PxCreateConvexMesh makes of the point cloud (Gu::ConvexHullData): convex cooking is host-side work in PhysX
 * too (the GPU pipeline receives cooked hulls through PxsSimulationController::addPxgShape), so cooked hulls are an INPUT of the hot path.
 * `ref_harness cook scene.bin out.bin` writes the section from the unmodified reference cooking code.
 * All little-endian, 4-byte fields, no padding.  Python mirror: physx_b200/scenes.py (numpy dtypes).
 * One shape per actor; without the local-pose section the shape local pose and the centre-of-mass pose are identity (planes: the
 * actor pose carries the plane frame, normal = local +X as in PxPlaneGeometry, physx/include/geometry/PxPlaneGeometry.h).  With it
 * ActorRec.pos / quat are the ACTOR pose (PxRigidActor::getGlobalPose), the shape sits at shape2Actor (PxShape::setLocalPose) and the
 * body frame the solver integrates is actor pose * body2Actor (PxRigidBody::setCMassLocalPose); states are reported as actor poses.
 *
 * geomType values follow PxGeometryType (physx/include/geometry/PxGeometry.h:48-62):
 *   0 sphere, 1 plane, 2 capsule, 3 box, 5 convex mesh.
 */
#ifndef PXB_SCENE_FORMAT_H
#define PXB_SCENE_FORMAT_H
#include <stdint.h>

#define PXB_SCENE_MAGIC 0x314e4353u /* "SCN1" */

enum { PXB_GEOM_SPHERE = 0, PXB_GEOM_PLANE = 1, PXB_GEOM_CAPSULE = 2, PXB_GEOM_BOX = 3, PXB_GEOM_CONVEX = 5 };
enum { PXB_ACTOR_DYNAMIC = 1u,
       PXB_ACTOR_KINEMATIC = 2u /* with PXB_ACTOR_DYNAMIC: PxRigidBodyFlag::eKINEMATIC -- moved by PxRigidDynamic::setKinematicTarget (ScKinematics.cpp:44-97), infinite mass in the solver,
                                   no pairs against statics or other kinematics (PxPairFilteringMode::eDEFAULT, BpFiltering.cpp:36-48); it keeps its place in the dynamic-body order */,
       PXB_ACTOR_DISABLE_GRAVITY = 4u /* PxActorFlag::eDISABLE_GRAVITY: no gravity term in the unconstrained velocity (DyBodyCoreIntegrator.h:55-59) */,
       PXB_ACTOR_GYROSCOPIC = 8u /* PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES: the body's angular velocity is advanced by the torque-free gyroscopic term when the solver
                                    body is built (DyTGSDynamics.cpp:177-193, DyRigidBodyToSolverBody.cpp:53-70) */ };
enum { PXB_SOLVER_PGS = 0, PXB_SOLVER_TGS = 1 };

typedef struct {
  uint32_t magic;
  uint32_t nActors;
  uint32_t nHulls;
  uint32_t solverType;       /* PXB_SOLVER_* */
  float    gravity[3];
  float    dt;
  uint32_t posIters, velIters;
  float    staticFriction, dynamicFriction, restitution;
  float    contactOffset, restOffset;
  float    sleepThreshold;   /* 0 disables sleeping */
  float    bounceThreshold;  /* PxSceneDesc default 0.2*toleranceSpeed(10) = 2.0 */
  float    frictionOffsetThreshold; /* default 0.04 */
  float    frictionCorrelationDistance; /* default 0.025 */
  float    toleranceLength;  /* 1.0 */
  uint32_t reserved[4];
} PxbSceneHeader;

typedef struct {
  uint32_t flags;       /* PXB_ACTOR_DYNAMIC | PxRigidDynamicLockFlag bits << 8 */
  uint32_t geomType;
  uint32_t envId;       /* 0xffffffff = none */
  uint32_t hullIdx;
  float    pos[3];
  float    quat[4];     /* x y z w */
  float    dims[4];     /* sphere: r; capsule: r, halfHeight; box: hx hy hz */
  float    linVel[3];
  float    angVel[3];
  float    mass;
  float    inertia[3];  /* mass-space diagonal */
  float    linDamping, angDamping;
  float    maxLinVel, maxAngVel;
  float    maxDepenetrationVel;
  uint32_t materialIndex; /* index into the scene's material table (ignored when the table is empty: the header's material applies) */
  uint32_t aggregate;   /* PxAggregate membership: 0 = none, k > 0 = aggregate k; bit 31 set = the aggregate has self collisions enabled.  Members of an aggregate without self collisions
                            generate no pairs among themselves (PxGetAggregateFilterHint, BpAABBManager.cpp aggregate self-collision pairs) */
} PxbActorRec;

/* One PxMaterial (physx/include/PxMaterial.h): bits = frictionCombineMode | restitutionCombineMode << 4 (PxCombineMode: 0 average, 1 min,
 * 2 multiply, 3 max) | flags << 8 (bit 0 = PxMaterialFlag::eDISABLE_FRICTION).  Pairs combine as PxsCombineMaterials does
 * (lowlevel/software/include/PxsMaterialCombiner.h:69-175, rigid non-compliant branch). */
typedef struct { float staticFriction, dynamicFriction, restitution; uint32_t bits; } PxbMaterialRec;

#define PXB_COOKED_MAGIC 0x43485850u /* "PXHC" */
/* followed by: float verts[nVerts][3]; PxbCookedPoly polys[nPolys]; uint8_t vertexRefs[nIdx] (padded to 4);
 * uint8_t facesByEdges[2 * nEdges] (padded to 4)   -- Gu::ConvexHullData, physx/source/geomutils/src/convex/GuConvexMeshData.h:47-175;
 * and, when reserved[0] != 0 (hulls of more than 32 vertices; subdiv = reserved[0] & 0xffff, nAdj = reserved[0] >> 16), the hill-climbing
 * data Gu::BigConvexRawData (GuBigConvexData.h:54-75): uint8_t samples[6 * subdiv^2] (padded to 4); uint16_t valencies[nVerts][2]
 * ({mCount, mOffset}); uint8_t adjacentVerts[nAdj] (padded to 4) */
typedef struct {
  uint32_t nVerts, nPolys, nEdges, nIdx;
  float centerOfMass[3];          /* mCenterOfMass */
  float boundsCenter[3], boundsExtents[3];   /* mAABB (local bounds) */
  float internalRadius, internalExtents[3];  /* mInternal */
  float unitMass, unitInertiaDiag[3], unitCom[3];   /* PxConvexMesh::getMassInformation at density 1 (diagonal of the inertia tensor) */
  uint32_t reserved[1];
} PxbCookedHullHeader;            /* 25 words = 100 bytes */
typedef struct { float plane[4]; uint32_t vref, nbVerts, minIndex, pad; } PxbCookedPoly;   /* HullPolygonData */

/* f1: the default simulation filter shader (PxDefaultSimulationFilterShader, physxextensions/src/ExtDefaultSimulationFilterShader.cpp:238-280) and its global state:
 * PxSetGroupCollisionFlag (collisionTable[g] bit h = groups g and h collide; all ones by default), PxSetFilterOps (ops[3]: 0 AND, 1 OR, 2 XOR, 3 NAND, 4 NOR, 5 NXOR,
 * 6 SWAP_AND), PxSetFilterBool, PxSetFilterConstants (constants[0..1] = K0 as PxFilterData word2 / word3, [2..3] = K1).  Per actor: the shape's PxFilterData
 * (word0 = collision group 0..31, word2 / word3 = PxGroupsMask).  A pair the shader answers eSUPPRESS for stays a broadphase pair and generates no contacts. */
#define PXB_FLAG_FILTER_SECTION 2u   /* header.reserved[0] bit 1 */
#define PXB_FLAG_SHAPE_OFFSETS 4u    /* header.reserved[0] bit 2 */
typedef struct { uint32_t collisionTable[32]; uint32_t ops[3]; uint32_t filterBool; uint32_t constants[4]; } PxbFilterShaderConfig;   /* 160 bytes */

/* local poses: shape2Actor (p, q.xyzw), body2Actor (p, q.xyzw) */
#define PXB_LOCAL_POSE_MAGIC 0x504c5850u /* "PXLP" */
typedef struct { float shapeP[3], shapeQ[4], bodyP[3], bodyQ[4], pad[2]; } PxbLocalPoseRec;   /* 64 bytes */

/* Per-step state record written by ref_harness / oracle tools: for every DYNAMIC actor, in actor
 * order: pos[3] quat[4] linVel[3] angVel[3] = 13 floats. */
#define PXB_STATE_FLOATS 13

#endif
