/* Scene interchange format shared by the oracle tools (test infrastructure).
 *
 * A scene file is:  SceneHeader | ActorRec[nActors] | for each hull: u32 nVerts, float xyz[nVerts]
 * All little-endian, 4-byte fields, no padding.  Python mirror: physx_b200/scenes.py (numpy dtypes).
 * One shape per actor, shape local pose = identity (planes: actor pose carries the plane frame,
 * normal = local +X as in PxPlaneGeometry, physx/include/geometry/PxPlaneGeometry.h).
 *
 * geomType values follow PxGeometryType (physx/include/geometry/PxGeometry.h:48-62):
 *   0 sphere, 1 plane, 2 capsule, 3 box, 5 convex mesh.
 */
#ifndef PXB_SCENE_FORMAT_H
#define PXB_SCENE_FORMAT_H
#include <stdint.h>

#define PXB_SCENE_MAGIC 0x314e4353u /* "SCN1" */

enum { PXB_GEOM_SPHERE = 0, PXB_GEOM_PLANE = 1, PXB_GEOM_CAPSULE = 2, PXB_GEOM_BOX = 3, PXB_GEOM_CONVEX = 5 };
enum { PXB_ACTOR_DYNAMIC = 1u };
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
  float    reserved[2];
} PxbActorRec;

/* Per-step state record written by ref_harness / oracle tools: for every DYNAMIC actor, in actor
 * order: pos[3] quat[4] linVel[3] angVel[3] = 13 floats. */
#define PXB_STATE_FLOATS 13

#endif
