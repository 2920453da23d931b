// PxgB200Plugin.cpp -- libPhysXGpu_64.so: the PhysXGpu plugin boundary of the reference, implemented over libphysx_b200.so.
//
// This is host-side glue compiled against the UNMODIFIED reference headers (same compiler flags as the SDK: -fno-rtti -fno-exceptions) so that
// an unmodified PhysX application loads it through PxSetPhysXGpuLoadHook / the default dlopen("libPhysXGpu_64.so") of
// physx/source/physx/src/gpu/PxPhysXGpuModuleLoader.cpp:184-233.  What is behind the boundary today:
//
//   exports            the 12 C symbols of physxgpu/include/PxPhysXGpu.h:207-237
//   PxCudaContextManager / PxCudaContext   (physx/include/cudamanager/PxCudaContextManager.h:262-485, PxCudaContext.h:70-184) over the CUDA runtime,
//                      primary context of the chosen device -- the context libphysx_b200.so itself runs in
//   PxPhysXGpu         factory (PxPhysXGpu.h:101-200; reference implementation physxgpu/src/PxgPhysXGpu.cpp:98-262)
//   PxsMemoryManager / PxsHeapMemoryAllocatorManager / PxsKernelWranglerManager   (lowlevel/software/include/PxsMemoryManager.h, PxsHeapMemoryAllocator.h,
//                      PxsKernelWrangler.h): pinned-host and device allocators the host builds its pinned arrays on
//   Bp::BroadPhase     (lowlevelaabb/include/BpBroadPhase.h:98-222)  -> pxb_bp_update / pxb_bp_fetch: the B200 broadphase kernels, pair sets
//                      bit-identical to the reference's ABP
//   Bp::AABBManagerBase (lowlevelaabb/include/BpAABBManagerBase.h:175-390; reference: gpubroadphase/src/PxgAABBManager.cpp:501-1050): added /
//                      updated / removed handle lists from the bitmaps, BroadPhaseUpdateData, created / destroyed AABBOverlap{userData} lists
//                      per element type.  PxAggregates are flattened: their shapes enter the broadphase one by one, pairs inside an aggregate without self collisions are dropped.
//   Bp::BoundsArray    createGpuBounds: the host class on pinned memory
//
// That is the scene configuration PxBroadPhaseType::eGPU with CPU dynamics (simulationcontroller/src/ScScene.cpp:786-916 creates exactly these
// objects for it).  createGpuNphaseImplementationContext / createGpuSimulationController / createGpuDynamicsContext report an error and return
// NULL (PxSceneFlag::eENABLE_GPU_DYNAMICS through the shim is the next step, INTEGRATION.md "Status"); the particle-buffer creators return NULL
// as the interface allows.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstring>
#include <cstdio>
#include <new>

#include "foundation/PxSimpleTypes.h"
#include "foundation/PxArray.h"
#include "foundation/PxBounds3.h"
#include "PxBroadPhase.h"
#include "foundation/PxFoundation.h"
#include "foundation/PxAllocator.h"
#include "foundation/PxUserAllocated.h"
#include "foundation/PxErrors.h"
#include "foundation/PxMutex.h"
#include "cudamanager/PxCudaContextManager.h"
#include "cudamanager/PxCudaContext.h"
#include "PxPhysXGpu.h"
#include "PxsMemoryManager.h"
#include "PxsHeapMemoryAllocator.h"
#include "PxsKernelWrangler.h"
#include "BpBroadPhase.h"
#include "BpBroadPhaseUpdate.h"
#include "BpAABBManagerBase.h"
#include "BpFiltering.h"
#include "PxSceneDesc.h"

#include "../include/physx_b200.h"

using namespace physx;

#define B200_ERROR(code, ...) PxGetFoundation().error(code, PX_FL, __VA_ARGS__)

namespace
{
// ------------------------------------------------------------------------------------------------------------------------------------------
// allocators: pinned host memory (the host's "pinned" arrays: bounds, groups, contact distances, handle lists) and device memory
class B200HostAllocator : public PxsHeapMemoryAllocator
{
public:
	explicit B200HostAllocator(int device) : mDevice(device) {}
	virtual void* allocate(size_t size, int, const char*, int) PX_OVERRIDE
	{
		if(!size) return NULL;
		int prev = -1; cudaGetDevice(&prev); if(prev != mDevice) cudaSetDevice(mDevice);
		void* p = NULL;
		if(cudaHostAlloc(&p, size, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); p = NULL; B200_ERROR(PxErrorCode::eOUT_OF_MEMORY, "libPhysXGpu_64 (physx_b200): pinned host allocation of %zu bytes failed", size); }
		if(prev >= 0 && prev != mDevice) cudaSetDevice(prev);
		return p;
	}
	virtual void deallocate(void* ptr) PX_OVERRIDE { if(ptr) cudaFreeHost(ptr); }
private:
	int mDevice;
};

class B200DeviceAllocator : public PxVirtualAllocatorCallback
{
public:
	explicit B200DeviceAllocator(int device) : mDevice(device) {}
	virtual void* allocate(size_t size, int, const char*, int) PX_OVERRIDE
	{
		if(!size) return NULL;
		int prev = -1; cudaGetDevice(&prev); if(prev != mDevice) cudaSetDevice(mDevice);
		void* p = NULL;
		if(cudaMalloc(&p, size) != cudaSuccess) { cudaGetLastError(); p = NULL; B200_ERROR(PxErrorCode::eOUT_OF_MEMORY, "libPhysXGpu_64 (physx_b200): device allocation of %zu bytes failed", size); }
		if(prev >= 0 && prev != mDevice) cudaSetDevice(prev);
		return p;
	}
	virtual void deallocate(void* ptr) PX_OVERRIDE { if(ptr) cudaFree(ptr); }
private:
	int mDevice;
};

static int deviceOf(PxCudaContextManager* m);

class B200MemoryManager : public PxsMemoryManager
{
public:
	explicit B200MemoryManager(int device) : mHost(device), mDeviceAlloc(device) {}
	virtual PxVirtualAllocatorCallback* getHostMemoryAllocator() PX_OVERRIDE { return &mHost; }
	virtual PxVirtualAllocatorCallback* getDeviceMemoryAllocator() PX_OVERRIDE { return &mDeviceAlloc; }
	B200HostAllocator mHost; B200DeviceAllocator mDeviceAlloc;
};

class B200HeapMemoryAllocatorManager : public PxsHeapMemoryAllocatorManager
{
public:
	B200HeapMemoryAllocatorManager(int device, PxU64 capacity) : mHeap(device), mCapacity(capacity) { mMappedMemoryAllocators = &mHeap; }
	virtual PxU64 getDeviceMemorySize() const PX_OVERRIDE { return mCapacity; }
	virtual PxsHeapStats getDeviceHeapStats() const PX_OVERRIDE { return PxsHeapStats(); }
	virtual void flushDeferredDeallocs() PX_OVERRIDE {}
	B200HostAllocator mHeap; PxU64 mCapacity;
};

// ------------------------------------------------------------------------------------------------------------------------------------------
// Bp::BroadPhase over the standalone broadphase object of libphysx_b200.so
class B200BroadPhase : public Bp::BroadPhase
{
public:
	B200BroadPhase(int device, PxU32 maxPairs) : mBp(NULL), mDevice(device), mMaxPairs(maxPairs), mCapacity(0), mNbCreated(0), mNbDeleted(0), mFailed(false) {}
	virtual ~B200BroadPhase() { if(mBp) pxb_bp_release(mBp); }

	virtual PxBroadPhaseType::Enum getType() const PX_OVERRIDE { return PxBroadPhaseType::eGPU; }
	virtual void release() PX_OVERRIDE { PX_DELETE_THIS; }

	virtual void update(PxcScratchAllocator*, const Bp::BroadPhaseUpdateData& d, PxBaseTask*) PX_OVERRIDE
	{
		mNbCreated = mNbDeleted = 0;
		if(mFailed) return;
		if(!ensure(d.getCapacity())) return;
		PX_COMPILE_TIME_ASSERT(sizeof(PxBounds3) == 24 && sizeof(Bp::FilterGroup::Enum) == 4 && sizeof(Bp::ShapeHandle) == 4);
		PxU8 lut[49];
		const bool* l = d.getFilter().getLUT();
		for(PxU32 i = 0; i < 49; i++) lut[i] = l[i] ? 1 : 0;
		const int rc = pxb_bp_update(mBp, reinterpret_cast<const float*>(d.getAABBs()), d.getContactDistance(), reinterpret_cast<const uint32_t*>(d.getGroups()), d.getEnvIDs(), d.getCapacity(), lut,
		                             d.getCreatedHandles(), d.getNumCreatedHandles(), d.getUpdatedHandles(), d.getNumUpdatedHandles(), d.getRemovedHandles(), d.getNumRemovedHandles());
		if(rc != PXB_OK) fail("pxb_bp_update");
	}
	virtual void preBroadPhase(const Bp::BroadPhaseUpdateData&) PX_OVERRIDE {}
	virtual void fetchBroadPhaseResults() PX_OVERRIDE
	{
		mNbCreated = mNbDeleted = 0;
		if(mFailed || !mBp) return;
		const uint32_t *c = NULL, *dl = NULL; uint32_t nc = 0, nd = 0;
		if(pxb_bp_fetch(mBp, &c, &nc, &dl, &nd) != PXB_OK) { fail("pxb_bp_fetch"); return; }
		mCreated.clear(); mDeleted.clear();
		mCreated.reserve(nc); mDeleted.reserve(nd);
		for(uint32_t i = 0; i < nc; i++) mCreated.pushBack(Bp::BroadPhasePair(c[2 * i], c[2 * i + 1]));
		for(uint32_t i = 0; i < nd; i++) mDeleted.pushBack(Bp::BroadPhasePair(dl[2 * i], dl[2 * i + 1]));
		mNbCreated = nc; mNbDeleted = nd;
	}
	virtual const Bp::BroadPhasePair* getCreatedPairs(PxU32& nb) const PX_OVERRIDE { nb = mNbCreated; return mCreated.begin(); }
	virtual const Bp::BroadPhasePair* getDeletedPairs(PxU32& nb) const PX_OVERRIDE { nb = mNbDeleted; return mDeleted.begin(); }
	virtual void freeBuffers() PX_OVERRIDE { mNbCreated = mNbDeleted = 0; }
	virtual void shiftOrigin(const PxVec3&, const PxBounds3*, const PxReal*) PX_OVERRIDE {}	// bounds are re-read in full by every update
	virtual void getCaps(PxBroadPhaseCaps& caps) const PX_OVERRIDE { caps.mMaxNbRegions = 0; }

private:
	bool ensure(PxU32 capacity)
	{
		if(mBp && capacity <= mCapacity) return true;
		if(mBp)
		{
			// the object is sized once; a scene that outgrows it needs PxSceneLimits / PxGpuDynamicsMemoryConfig raised (the reference's GPU buffers behave alike)
			B200_ERROR(PxErrorCode::eINVALID_OPERATION, "libPhysXGpu_64 (physx_b200): broadphase capacity %u exceeded (%u objects): raise PxSceneLimits::maxNbBodies / maxNbStaticShapes", mCapacity, capacity);
			mFailed = true; return false;
		}
		mCapacity = PxMax(capacity * 2u, 1024u);
		if(pxb_bp_create(mCapacity, mMaxPairs, mDevice, &mBp) != PXB_OK) { mBp = NULL; fail("pxb_bp_create"); return false; }
		return true;
	}
	void fail(const char* what) { mFailed = true; B200_ERROR(PxErrorCode::eINTERNAL_ERROR, "libPhysXGpu_64 (physx_b200): %s failed: %s", what, pxb_last_error()); }

	PxbBroadPhase* mBp; int mDevice; PxU32 mMaxPairs, mCapacity, mNbCreated, mNbDeleted; bool mFailed;
	PxArray<Bp::BroadPhasePair> mCreated, mDeleted;
};

// ------------------------------------------------------------------------------------------------------------------------------------------
// Bp::AABBManagerBase (aggregates flattened): bitmaps -> handle lists -> BroadPhaseUpdateData -> overlaps by element type
class B200AABBManager : public Bp::AABBManagerBase
{
public:
	B200AABBManager(Bp::BroadPhase& bp, Bp::BoundsArray& boundsArray, PxFloatArrayPinnedSafe& contactDistance, PxU32 maxNbAggregates, PxU32 maxNbShapes, PxVirtualAllocator& allocator, PxU64 contextID,
	                PxPairFilteringMode::Enum kineKine, PxPairFilteringMode::Enum staticKine) :
		Bp::AABBManagerBase(bp, boundsArray, contactDistance, maxNbAggregates, maxNbShapes, allocator, contextID, kineKine, staticKine), mPersistentStateChanged(true) {}

	virtual void destroy() PX_OVERRIDE { PX_DELETE_THIS; }

	// PxAggregate: the grid broadphase needs no merged bound, so the shapes of an aggregate enter the broadphase one by one and the aggregate itself only keeps its
	// entry (bounds index, group, user data) and its self-collision switch; pairs between two shapes of an aggregate WITHOUT self collisions are dropped when the
	// created / deleted pairs are handed to the host (the reference computes an aggregate's self-collision pairs only when enabled, BpAABBManager.cpp Aggregate::mSelfCollisionPairs)
	struct AggregateRec { Bp::BoundsIndex index; PxU32 nbShapes; bool selfCollisions; bool used; };
	virtual Bp::AggregateHandle createAggregate(Bp::BoundsIndex index, Bp::FilterGroup::Enum group, void* userData, PxU32, PxAggregateFilterHint filterHint, PxU32) PX_OVERRIDE
	{
		Bp::AggregateHandle handle = PX_INVALID_U32;
		for(PxU32 i = 0; i < mAggregateRecs.size(); i++) if(!mAggregateRecs[i].used) { handle = i; break; }
		if(handle == PX_INVALID_U32) { handle = mAggregateRecs.size(); mAggregateRecs.pushBack(AggregateRec()); }
		AggregateRec& a = mAggregateRecs[handle];
		a.index = index; a.nbShapes = 0; a.selfCollisions = PxGetAggregateSelfCollisionBit(filterHint) != 0; a.used = true;
		initEntry(index, 0.0f, group, userData);
		mVolumeData[index].setAggregate(handle);
		mBoundsArray.setBounds(PxBounds3::empty(), index);	// (the aggregate's own bound is never in the broadphase)
		return handle;
	}
	virtual bool destroyAggregate(Bp::BoundsIndex& index_, Bp::FilterGroup::Enum& group_, Bp::AggregateHandle handle) PX_OVERRIDE
	{
		if(handle >= mAggregateRecs.size() || !mAggregateRecs[handle].used)
			return B200_ERROR(PxErrorCode::eINVALID_PARAMETER, "AABBManager::destroyAggregate - aggregateId out of bounds or already removed");
		if(mAggregateRecs[handle].nbShapes)
			return B200_ERROR(PxErrorCode::eINVALID_PARAMETER, "AABBManager::destroyAggregate - aggregate still has bounds that needs removed");
		const Bp::BoundsIndex index = mAggregateRecs[handle].index;
		index_ = index; group_ = mGroups[index];
		resetEntry(index);
		mAggregateRecs[handle].used = false;
		return true;
	}

	virtual bool addBounds(Bp::BoundsIndex index, PxReal contactDistance, Bp::FilterGroup::Enum group, void* userData, Bp::AggregateHandle aggregateHandle, Bp::ElementType::Enum volumeType, PxU32 envID) PX_OVERRIDE
	{
		if(aggregateHandle != PX_INVALID_U32 && (aggregateHandle >= mAggregateRecs.size() || !mAggregateRecs[aggregateHandle].used))
			return B200_ERROR(PxErrorCode::eINVALID_PARAMETER, "AABBManager::addBounds - aggregateId out of bounds");
		initEntry(index, contactDistance, group, userData, volumeType);
		if(mEnvIDs.size() < mVolumeData.size()) mEnvIDs.resize(mVolumeData.size(), PX_INVALID_U32);	// environment ids: PxActor::setEnvironmentID, honoured on the GPU (broadphase.cu:62-80)
		mEnvIDs[index] = PxI32(envID);
		if(aggregateHandle == PX_INVALID_U32) mVolumeData[index].setSingleActor();
		else { mVolumeData[index].setAggregated(aggregateHandle); mAggregateRecs[aggregateHandle].nbShapes++; }
		addBPEntry(index);
		mPersistentStateChanged = true;
		return true;
	}
	virtual bool removeBounds(Bp::BoundsIndex index) PX_OVERRIDE
	{
		PX_ASSERT(index < mVolumeData.size());
		if(mVolumeData[index].isAggregated()) { AggregateRec& a = mAggregateRecs[mVolumeData[index].getAggregateOwner()]; if(a.nbShapes) a.nbShapes--; }
		const bool res = removeBPEntry(index);
		resetEntry(index);
		mPersistentStateChanged = true;
		return res;
	}

	virtual void updateBPFirstPass(PxU32, Cm::FlushPool&, bool, PxBaseTask*) PX_OVERRIDE
	{
		// added: every bit of the added map; updated: changed bits of objects that are in the broadphase and were not added this frame;
		// removed: every bit of the removed map (BpAABBManager.cpp:1378-1545 without the aggregate branches)
		mAddedHandles.forceSize_Unsafe(0); mUpdatedHandles.forceSize_Unsafe(0); mRemovedHandles.forceSize_Unsafe(0);
		collect(mAddedHandleMap, mAddedHandles, false);
		if(!mOriginShifted) collect(mChangedHandleMap, mUpdatedHandles, true);
		else
		{	// after an origin shift every object in the broadphase moved
			mOriginShifted = false;
			for(PxU32 i = 0; i < mUsedSize; i++)
				if(mGroups[i] != Bp::FilterGroup::eINVALID && !mAddedHandleMap.test(i)) mUpdatedHandles.pushBack(i);
		}
		collect(mRemovedHandleMap, mRemovedHandles, false);
	}
	virtual void updateBPSecondPass(PxcScratchAllocator* scratch, PxBaseTask* continuation) PX_OVERRIDE
	{
		const Bp::BroadPhaseUpdateData updateData(mAddedHandles.begin(), mAddedHandles.size(), mUpdatedHandles.begin(), mUpdatedHandles.size(), mRemovedHandles.begin(), mRemovedHandles.size(),
		                                          mBoundsArray.begin(), mGroups.begin(), mContactDistance.begin(), reinterpret_cast<const PxU32*>(mEnvIDs.begin()), mBoundsArray.size(), mFilters,
		                                          mBoundsArray.hasChanged() || mPersistentStateChanged, false);
		mRan = mAddedHandles.size() || mUpdatedHandles.size() || mRemovedHandles.size();
		if(mRan) mBroadPhase.update(scratch, updateData, continuation);
		mPersistentStateChanged = false;
		mBoundsArray.resetChangedState();
	}
	virtual void postBroadPhase(PxBaseTask*, Cm::FlushPool&) PX_OVERRIDE
	{
		for(PxU32 i = 0; i < Bp::ElementType::eCOUNT; i++) { mCreatedOverlaps[i].forceSize_Unsafe(0); mDestroyedOverlaps[i].forceSize_Unsafe(0); }
		if(mRan)
		{
			mBroadPhase.fetchBroadPhaseResults();
			PxU32 nb = 0;
			const Bp::BroadPhasePair* pairs = mBroadPhase.getDeletedPairs(nb);
			for(PxU32 i = 0; i < nb; i++)
			{	// a pair whose volume lost its user data was removed by the host already (BpAABBManager.cpp:1611-1617)
				void* u0 = mVolumeData[pairs[i].mVolA].getUserData(); void* u1 = mVolumeData[pairs[i].mVolB].getUserData();
				if(u0 && u1 && !insideQuietAggregate(pairs[i].mVolA, pairs[i].mVolB)) output(mDestroyedOverlaps, pairs[i].mVolA, pairs[i].mVolB, u0, u1);
			}
			pairs = mBroadPhase.getCreatedPairs(nb);
			for(PxU32 i = 0; i < nb; i++)
				if(!insideQuietAggregate(pairs[i].mVolA, pairs[i].mVolB)) output(mCreatedOverlaps, pairs[i].mVolA, pairs[i].mVolB, mVolumeData[pairs[i].mVolA].getUserData(), mVolumeData[pairs[i].mVolB].getUserData());
#if PX_ENABLE_SIM_STATS
			mGpuDynamicsLostFoundPairsStats = PxMax(mGpuDynamicsLostFoundPairsStats, nb);
#endif
		}
		mAddedHandleMap.clear(); mRemovedHandleMap.clear();
		mRan = false;
	}
	virtual void reallocateChangedAABBMgActorHandleMap(const PxU32 size) PX_OVERRIDE { mChangedHandleMap.resizeAndClear(size); }
	virtual void visualize(PxRenderOutput&) PX_OVERRIDE {}
	virtual void releaseDeferredAggregateIds() PX_OVERRIDE {}
	virtual void setPersistentStateChanged() PX_OVERRIDE { mPersistentStateChanged = true; }

private:
	// both volumes are shapes of the same aggregate and that aggregate has no self collisions
	bool insideQuietAggregate(PxU32 a, PxU32 b) const
	{
		const Bp::VolumeData& va = mVolumeData[a]; const Bp::VolumeData& vb = mVolumeData[b];
		if(!va.isAggregated() || !vb.isAggregated() || va.getAggregateOwner() != vb.getAggregateOwner()) return false;
		return !mAggregateRecs[va.getAggregateOwner()].selfCollisions;
	}
	PxArray<AggregateRec> mAggregateRecs;
	template <class Map, class List> void collect(const Map& map, List& out, bool updatedPass)
	{
		const PxU32* bits = map.getWords();
		if(!bits) return;
		const PxU32 last = map.findLast();
		for(PxU32 w = 0; w <= last >> 5; ++w)
			for(PxU32 b = bits[w]; b; b &= b - 1)
			{
				const Bp::BoundsIndex handle = PxU32(w << 5 | PxLowestSetBit(b));
				if(updatedPass && (handle >= mUsedSize || mAddedHandleMap.boundedTest(handle) || mRemovedHandleMap.boundedTest(handle) || mGroups[handle] == Bp::FilterGroup::eINVALID)) continue;
				out.pushBack(handle);
			}
	}
	void output(PxArray<Bp::AABBOverlap>* overlaps, PxU32 id0, PxU32 id1, void* u0, void* u1)
	{
		const Bp::ElementType::Enum type = PxMax(mVolumeData[id0].getVolumeType(), mVolumeData[id1].getVolumeType());
		Bp::AABBOverlap o; o.mUserData0 = u0; o.mUserData1 = u1; o.mPairUserData = NULL;
		overlaps[type].pushBack(o);
	}
	bool mPersistentStateChanged, mRan = false;
};

// ------------------------------------------------------------------------------------------------------------------------------------------
// PxCudaContext / PxCudaContextManager over the CUDA runtime (primary context of one device)
static PxCUresult cuRes(cudaError_t e) { return PxCUresult(PxU32(e)); }	// cudaSuccess == CUDA_SUCCESS == 0; other codes are reported as they are

class B200CudaContext : public PxCudaContext, public PxUserAllocated
{
public:
	B200CudaContext() : mAbort(false), mLast(cudaSuccess) { mAllocatorCallback = NULL; }
	virtual void release() PX_OVERRIDE { PX_DELETE_THIS; }
	virtual PxCUresult memAlloc(CUdeviceptr* dptr, size_t n) PX_OVERRIDE { void* p = NULL; const cudaError_t e = note(cudaMalloc(&p, n)); *dptr = CUdeviceptr(size_t(p)); return cuRes(e); }
	virtual PxCUresult memFree(CUdeviceptr dptr) PX_OVERRIDE { return cuRes(note(cudaFree(reinterpret_cast<void*>(size_t(dptr))))); }
	virtual PxCUresult memHostAlloc(void** pp, size_t n, unsigned int flags) PX_OVERRIDE { return cuRes(note(cudaHostAlloc(pp, n, flags))); }
	virtual PxCUresult memFreeHost(void* p) PX_OVERRIDE { return cuRes(note(cudaFreeHost(p))); }
	virtual PxCUresult memHostGetDevicePointer(CUdeviceptr* pd, void* p, unsigned int flags) PX_OVERRIDE { void* d = NULL; const cudaError_t e = note(cudaHostGetDevicePointer(&d, p, flags)); *pd = CUdeviceptr(size_t(d)); return cuRes(e); }
	// this plugin's kernels are linked into libphysx_b200.so (no runtime-loaded modules): the module / launch entry points refuse
	virtual PxCUresult moduleLoadDataEx(CUmodule*, const void*, unsigned int, PxCUjit_option*, void**) PX_OVERRIDE { return cuRes(note(cudaErrorNotSupported)); }
	virtual PxCUresult moduleGetFunction(CUfunction*, CUmodule, const char*) PX_OVERRIDE { return cuRes(note(cudaErrorNotSupported)); }
	virtual PxCUresult moduleUnload(CUmodule) PX_OVERRIDE { return cuRes(note(cudaErrorNotSupported)); }
	virtual PxCUresult streamCreate(CUstream* s, unsigned int flags) PX_OVERRIDE { return cuRes(note(cudaStreamCreateWithFlags(reinterpret_cast<cudaStream_t*>(s), flags))); }
	virtual PxCUresult streamCreateWithPriority(CUstream* s, unsigned int flags, int prio) PX_OVERRIDE { return cuRes(note(cudaStreamCreateWithPriority(reinterpret_cast<cudaStream_t*>(s), flags, prio))); }
	virtual PxCUresult streamFlush(CUstream s) PX_OVERRIDE { const cudaError_t e = cudaStreamQuery(reinterpret_cast<cudaStream_t>(s)); return cuRes(e == cudaErrorNotReady ? cudaSuccess : note(e)); }
	virtual PxCUresult streamWaitEvent(CUstream s, CUevent e, unsigned int flags) PX_OVERRIDE { return cuRes(note(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(s), reinterpret_cast<cudaEvent_t>(e), flags))); }
	virtual PxCUresult streamWaitEvent(CUstream s, CUevent e) PX_OVERRIDE { return streamWaitEvent(s, e, 0); }
	virtual PxCUresult streamDestroy(CUstream s) PX_OVERRIDE { return cuRes(note(cudaStreamDestroy(reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult streamSynchronize(CUstream s) PX_OVERRIDE { return cuRes(note(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult eventCreate(CUevent* e, unsigned int flags) PX_OVERRIDE { return cuRes(note(cudaEventCreateWithFlags(reinterpret_cast<cudaEvent_t*>(e), flags))); }
	virtual PxCUresult eventRecord(CUevent e, CUstream s) PX_OVERRIDE { return cuRes(note(cudaEventRecord(reinterpret_cast<cudaEvent_t>(e), reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult eventQuery(CUevent e) PX_OVERRIDE { return cuRes(cudaEventQuery(reinterpret_cast<cudaEvent_t>(e))); }
	virtual PxCUresult eventSynchronize(CUevent e) PX_OVERRIDE { return cuRes(note(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(e)))); }
	virtual PxCUresult eventDestroy(CUevent e) PX_OVERRIDE { return cuRes(note(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(e)))); }
	virtual PxCUresult launchKernel(CUfunction, unsigned int, unsigned int, unsigned int, unsigned int, unsigned int, unsigned int, unsigned int, CUstream, PxCudaKernelParam*, size_t, void**, const char*, int) PX_OVERRIDE { return cuRes(note(cudaErrorNotSupported)); }
	virtual PxCUresult launchKernel(CUfunction, PxU32, PxU32, PxU32, PxU32, PxU32, PxU32, PxU32, CUstream, void**, void**, const char*, int) PX_OVERRIDE { return cuRes(note(cudaErrorNotSupported)); }
	virtual PxCUresult memcpyDtoH(void* dst, CUdeviceptr src, size_t n) PX_OVERRIDE { return cuRes(note(cudaMemcpy(dst, reinterpret_cast<const void*>(size_t(src)), n, cudaMemcpyDeviceToHost))); }
	virtual PxCUresult memcpyDtoHAsync(void* dst, CUdeviceptr src, size_t n, CUstream s) PX_OVERRIDE { return cuRes(note(cudaMemcpyAsync(dst, reinterpret_cast<const void*>(size_t(src)), n, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult memcpyHtoD(CUdeviceptr dst, const void* src, size_t n) PX_OVERRIDE { return cuRes(note(cudaMemcpy(reinterpret_cast<void*>(size_t(dst)), src, n, cudaMemcpyHostToDevice))); }
	virtual PxCUresult memcpyHtoDAsync(CUdeviceptr dst, const void* src, size_t n, CUstream s) PX_OVERRIDE { return cuRes(note(cudaMemcpyAsync(reinterpret_cast<void*>(size_t(dst)), src, n, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult memcpyDtoDAsync(CUdeviceptr dst, CUdeviceptr src, size_t n, CUstream s) PX_OVERRIDE { return cuRes(note(cudaMemcpyAsync(reinterpret_cast<void*>(size_t(dst)), reinterpret_cast<const void*>(size_t(src)), n, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult memcpyDtoD(CUdeviceptr dst, CUdeviceptr src, size_t n) PX_OVERRIDE { return cuRes(note(cudaMemcpy(reinterpret_cast<void*>(size_t(dst)), reinterpret_cast<const void*>(size_t(src)), n, cudaMemcpyDeviceToDevice))); }
	virtual PxCUresult memcpyPeerAsync(CUdeviceptr dst, CUcontext, CUdeviceptr src, CUcontext, size_t n, CUstream s) PX_OVERRIDE { return cuRes(note(cudaMemcpyAsync(reinterpret_cast<void*>(size_t(dst)), reinterpret_cast<const void*>(size_t(src)), n, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult memsetD32Async(CUdeviceptr dst, unsigned int v, size_t n, CUstream s) PX_OVERRIDE { return cuRes(note(fill32(dst, v, n, s, true))); }
	virtual PxCUresult memsetD8Async(CUdeviceptr dst, unsigned char v, size_t n, CUstream s) PX_OVERRIDE { return cuRes(note(cudaMemsetAsync(reinterpret_cast<void*>(size_t(dst)), v, n, reinterpret_cast<cudaStream_t>(s)))); }
	virtual PxCUresult memsetD32(CUdeviceptr dst, unsigned int v, size_t n) PX_OVERRIDE { return cuRes(note(fill32(dst, v, n, NULL, false))); }
	virtual PxCUresult memsetD16(CUdeviceptr dst, unsigned short v, size_t n) PX_OVERRIDE { return cuRes(note(cudaMemset2D(reinterpret_cast<void*>(size_t(dst)), 2, v & 0xff, ((v >> 8) == (v & 0xff)) ? 2 : 1, n))); }
	virtual PxCUresult memsetD8(CUdeviceptr dst, unsigned char v, size_t n) PX_OVERRIDE { return cuRes(note(cudaMemset(reinterpret_cast<void*>(size_t(dst)), v, n))); }
	virtual PxCUresult getLastError() PX_OVERRIDE { const cudaError_t e = mLast; mLast = cudaSuccess; return cuRes(e); }
	virtual void setAbortMode(bool abort) PX_OVERRIDE { mAbort = abort; }
	virtual bool isInAbortMode() PX_OVERRIDE { return mAbort; }
private:
	cudaError_t note(cudaError_t e) { if(e != cudaSuccess) { mLast = e; cudaGetLastError(); } return e; }
	// 32-bit fills: the runtime has byte fills only; a value whose four bytes are equal maps onto one, anything else goes through a host-staged copy
	cudaError_t fill32(CUdeviceptr dst, unsigned int v, size_t n, CUstream s, bool async)
	{
		void* p = reinterpret_cast<void*>(size_t(dst));
		const unsigned int b = v & 0xff;
		if(v == (b | b << 8 | b << 16 | b << 24)) return async ? cudaMemsetAsync(p, int(b), n * 4, reinterpret_cast<cudaStream_t>(s)) : cudaMemset(p, int(b), n * 4);
		unsigned int* tmp = static_cast<unsigned int*>(PX_ALLOC(n * 4, "fill32"));
		for(size_t i = 0; i < n; i++) tmp[i] = v;
		const cudaError_t e = cudaMemcpy(p, tmp, n * 4, cudaMemcpyHostToDevice);
		PX_FREE(tmp);
		return e;
	}
	bool mAbort; cudaError_t mLast;
};

class B200CudaContextManager : public PxCudaContextManager, public PxUserAllocated
{
public:
	B200CudaContextManager(int device, const cudaDeviceProp& prop, int driverVersion) : mDevice(device), mProp(prop), mDriver(driverVersion), mCtx(NULL), mConcurrent(true), mValid(true)
	{
		mCudaContext = PX_NEW(B200CudaContext)();
		cudaSetDevice(device);
		if(cudaFree(0) != cudaSuccess) { cudaGetLastError(); mValid = false; }	// creates the primary context (the one libphysx_b200.so runs in)
		typedef int (*GetCurrent)(CUcontext*);
		GetCurrent getCurrent = reinterpret_cast<GetCurrent>(dlsym(RTLD_DEFAULT, "cuCtxGetCurrent"));	// libcuda.so.1 was loaded RTLD_GLOBAL by the host's module loader
		if(getCurrent) getCurrent(&mCtx);
	}
	int device() const { return mDevice; }

	virtual CUdeviceptr getMappedDevicePtr(void* pinned) PX_OVERRIDE { void* d = NULL; if(cudaHostGetDevicePointer(&d, pinned, 0) != cudaSuccess) { cudaGetLastError(); d = NULL; } return CUdeviceptr(size_t(d)); }
	virtual void acquireContext() PX_OVERRIDE { tryAcquireContext(); }
	virtual bool tryAcquireContext() PX_OVERRIDE
	{
		// runtime API: "acquiring" = making the device (its primary context) current on this thread; recursion is counted per thread
		Tls& t = tls();
		if(t.depth++ == 0) { t.prev = -1; cudaGetDevice(&t.prev); if(t.prev != mDevice && cudaSetDevice(mDevice) != cudaSuccess) { cudaGetLastError(); return false; } }
		return true;
	}
	virtual void releaseContext() PX_OVERRIDE { Tls& t = tls(); if(t.depth && --t.depth == 0 && t.prev >= 0 && t.prev != mDevice) cudaSetDevice(t.prev); }
	virtual CUcontext getContext() PX_OVERRIDE { return mCtx; }
	virtual PxCudaContext* getCudaContext() PX_OVERRIDE { return mCudaContext; }
	virtual bool contextIsValid() const PX_OVERRIDE { return mValid; }
	virtual bool supportsArchSM10() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM11() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM12() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM13() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM20() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM30() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM35() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM50() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM52() const PX_OVERRIDE { return true; }
	virtual bool supportsArchSM60() const PX_OVERRIDE { return true; }
	virtual bool isIntegrated() const PX_OVERRIDE { return mProp.integrated != 0; }
	virtual bool canMapHostMemory() const PX_OVERRIDE { return mProp.canMapHostMemory != 0; }
	virtual int getDriverVersion() const PX_OVERRIDE { return mDriver; }
	virtual size_t getDeviceTotalMemBytes() const PX_OVERRIDE { return mProp.totalGlobalMem; }
	virtual int getMultiprocessorCount() const PX_OVERRIDE { return mProp.multiProcessorCount; }
	virtual unsigned int getClockRate() const PX_OVERRIDE { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, mDevice); return (unsigned int)khz; }
	virtual int getSharedMemPerBlock() const PX_OVERRIDE { return int(mProp.sharedMemPerBlock); }
	virtual int getSharedMemPerMultiprocessor() const PX_OVERRIDE { return int(mProp.sharedMemPerMultiprocessor); }
	virtual unsigned int getMaxThreadsPerBlock() const PX_OVERRIDE { return (unsigned int)mProp.maxThreadsPerBlock; }
	virtual const char* getDeviceName() const PX_OVERRIDE { return mProp.name; }
	virtual CUdevice getDevice() const PX_OVERRIDE { return CUdevice(mDevice); }
	virtual void setUsingConcurrentStreams(bool b) PX_OVERRIDE { mConcurrent = b; }
	virtual bool getUsingConcurrentStreams() const PX_OVERRIDE { return mConcurrent; }
	virtual void getDeviceMemoryInfo(size_t& free, size_t& total) const PX_OVERRIDE { free = total = 0; int prev = -1; cudaGetDevice(&prev); cudaSetDevice(mDevice); if(cudaMemGetInfo(&free, &total) != cudaSuccess) cudaGetLastError(); if(prev >= 0) cudaSetDevice(prev); }
	virtual CUmodule* getCuModules() PX_OVERRIDE { return NULL; }
	virtual void release() PX_OVERRIDE { mCudaContext->release(); PX_DELETE_THIS; }

protected:
	virtual void* allocDeviceBufferInternal(PxU64 n, const char*, PxI32) PX_OVERRIDE { Scope s(this); void* p = NULL; if(cudaMalloc(&p, size_t(n)) != cudaSuccess) { cudaGetLastError(); p = NULL; } return p; }
	virtual void* allocPinnedHostBufferInternal(PxU64 n, const char*, PxI32) PX_OVERRIDE { Scope s(this); void* p = NULL; if(cudaHostAlloc(&p, size_t(n), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); p = NULL; } return p; }
	virtual void freeDeviceBufferInternal(void* p) PX_OVERRIDE { Scope s(this); if(p) cudaFree(p); }
	virtual void freePinnedHostBufferInternal(void* p) PX_OVERRIDE { Scope s(this); if(p) cudaFreeHost(p); }
	virtual void clearDeviceBufferAsyncInternal(void* p, PxU32 n, CUstream st, PxI32 v) PX_OVERRIDE { Scope s(this); cudaMemsetAsync(p, v, n, reinterpret_cast<cudaStream_t>(st)); }
	virtual void copyDToHAsyncInternal(void* h, const void* d, PxU32 n, CUstream st) PX_OVERRIDE { Scope s(this); cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(st)); }
	virtual void copyHToDAsyncInternal(void* d, const void* h, PxU32 n, CUstream st) PX_OVERRIDE { Scope s(this); cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(st)); }
	virtual void copyDToDAsyncInternal(void* d, const void* src, PxU32 n, CUstream st) PX_OVERRIDE { Scope s(this); cudaMemcpyAsync(d, src, n, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(st)); }
	virtual void copyDToHInternal(void* h, const void* d, PxU32 n) PX_OVERRIDE { Scope s(this); cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost); }
	virtual void copyHToDInternal(void* d, const void* h, PxU32 n) PX_OVERRIDE { Scope s(this); cudaMemcpy(d, h, n, cudaMemcpyHostToDevice); }
	virtual void memsetD8AsyncInternal(void* d, const PxU8& v, PxU32 n, CUstream st) PX_OVERRIDE { Scope s(this); cudaMemsetAsync(d, v, n, reinterpret_cast<cudaStream_t>(st)); }
	virtual void memsetD32AsyncInternal(void* d, const PxU32& v, PxU32 n, CUstream st) PX_OVERRIDE { Scope s(this); mCudaContext->memsetD32Async(CUdeviceptr(size_t(d)), v, n, st); }

private:
	struct Tls { int depth, prev; };
	static Tls& tls() { static thread_local Tls t = {0, -1}; return t; }
	struct Scope { B200CudaContextManager* m; explicit Scope(B200CudaContextManager* m_) : m(m_) { m->acquireContext(); } ~Scope() { m->releaseContext(); } };
	int mDevice; cudaDeviceProp mProp; int mDriver; CUcontext mCtx; B200CudaContext* mCudaContext; bool mConcurrent, mValid;
};

static int deviceOf(PxCudaContextManager* m) { return m ? static_cast<B200CudaContextManager*>(m)->device() : 0; }

// ------------------------------------------------------------------------------------------------------------------------------------------
class B200KernelWranglerManager : public PxsKernelWranglerManager
{
public:
	explicit B200KernelWranglerManager(PxCudaContextManager* m) { mKernelWrangler = NULL; mCudaContextManager = m; }	// no runtime-loaded kernels: they are linked into libphysx_b200.so
};

class B200PhysXGpu : public PxPhysXGpu, public PxUserAllocated
{
public:
	B200PhysXGpu() : mWrangler(NULL) {}
	virtual void release() PX_OVERRIDE;
	virtual PxsParticleBuffer* createParticleBuffer(PxU32, PxU32, PxCudaContextManager&) PX_OVERRIDE { return NULL; }
	virtual PxsParticleAndDiffuseBuffer* createParticleAndDiffuseBuffer(PxU32, PxU32, PxU32, PxCudaContextManager&) PX_OVERRIDE { return NULL; }
	virtual PxsParticleClothBuffer* createParticleClothBuffer(PxU32, PxU32, PxU32, PxU32, PxU32, PxCudaContextManager&) PX_OVERRIDE { return NULL; }
	virtual PxsParticleRigidBuffer* createParticleRigidBuffer(PxU32, PxU32, PxU32, PxCudaContextManager&) PX_OVERRIDE { return NULL; }

	virtual PxsMemoryManager* createGpuMemoryManager(PxCudaContextManager* m) PX_OVERRIDE { return PX_NEW(B200MemoryManager)(deviceOf(m)); }
	virtual PxsHeapMemoryAllocatorManager* createGpuHeapMemoryAllocatorManager(const PxU32 heapCapacity, PxsMemoryManager*, const PxU32) PX_OVERRIDE
	{
		int dev = 0; cudaGetDevice(&dev);
		return PX_NEW(B200HeapMemoryAllocatorManager)(mWrangler ? deviceOf(mWrangler->mCudaContextManager) : dev, heapCapacity);
	}
	virtual PxsKernelWranglerManager* getGpuKernelWranglerManager(PxCudaContextManager* m) PX_OVERRIDE
	{
		if(!mWrangler) mWrangler = PX_NEW(B200KernelWranglerManager)(m);
		return mWrangler;
	}
	virtual Bp::BroadPhase* createGpuBroadPhase(const PxGpuBroadPhaseDesc&, PxsKernelWranglerManager*, PxCudaContextManager* m, PxU32, const PxGpuDynamicsMemoryConfig& config, PxsHeapMemoryAllocatorManager*, PxU64) PX_OVERRIDE
	{
		// PxGpuDynamicsMemoryConfig::foundLostPairsCapacity bounds the pairs reported per step; the persistent pair list is sized from it as well
		return PX_NEW(B200BroadPhase)(deviceOf(m), PxMax(config.foundLostPairsCapacity * 4u, 1u << 20));
	}
	virtual Bp::AABBManagerBase* createGpuAABBManager(PxsKernelWranglerManager*, PxCudaContextManager*, const PxU32, const PxGpuDynamicsMemoryConfig&, PxsHeapMemoryAllocatorManager*, Bp::BroadPhase& bp,
	                                                  Bp::BoundsArray& boundsArray, PxFloatArrayPinnedSafe& contactDistance, PxU32 maxNbAggregates, PxU32 maxNbShapes, PxVirtualAllocator& allocator, PxU64 contextID,
	                                                  PxPairFilteringMode::Enum kineKine, PxPairFilteringMode::Enum staticKine) PX_OVERRIDE
	{
		return PX_NEW(B200AABBManager)(bp, boundsArray, contactDistance, maxNbAggregates, maxNbShapes, allocator, contextID, kineKine, staticKine);
	}
	virtual Bp::BoundsArray* createGpuBounds(PxVirtualAllocator& allocator) PX_OVERRIDE { return PX_NEW(Bp::BoundsArray)(allocator); }

	virtual PxvNphaseImplementationContext* createGpuNphaseImplementationContext(PxsContext&, PxsKernelWranglerManager*, PxvNphaseImplementationFallback*, const PxGpuDynamicsMemoryConfig&, void*, void*, void*,
	                                                                             PxBoundsArrayPinned&, IG::IslandSim*, Dy::Context*, const PxU32, PxsHeapMemoryAllocatorManager*, bool) PX_OVERRIDE { return notYet("createGpuNphaseImplementationContext"), static_cast<PxvNphaseImplementationContext*>(NULL); }
	virtual PxsSimulationController* createGpuSimulationController(PxsKernelWranglerManager*, PxCudaContextManager*, Dy::Context*, PxvNphaseImplementationContext*, Bp::BroadPhase*, bool, PxsSimulationControllerCallback*, PxU32,
	                                                               PxsHeapMemoryAllocatorManager*, PxU32, PxU32, PxU32, PxU32, bool) PX_OVERRIDE { return notYet("createGpuSimulationController"), static_cast<PxsSimulationController*>(NULL); }
	virtual Dy::Context* createGpuDynamicsContext(Cm::FlushPool&, PxsKernelWranglerManager*, PxCudaContextManager*, const PxGpuDynamicsMemoryConfig&, IG::SimpleIslandManager&, PxU32, PxU32, bool, bool, bool, PxReal, PxU32,
	                                              PxvSimStats&, PxsHeapMemoryAllocatorManager*, bool, bool, PxSolverType::Enum, PxReal, bool, PxU64, bool) PX_OVERRIDE { return notYet("createGpuDynamicsContext"), static_cast<Dy::Context*>(NULL); }
private:
	static void notYet(const char* what) { B200_ERROR(PxErrorCode::eINVALID_OPERATION, "libPhysXGpu_64 (physx_b200): %s -- PxSceneFlag::eENABLE_GPU_DYNAMICS is not available through this plugin yet; use PxBroadPhaseType::eGPU with CPU dynamics, or libphysx_b200.so's own C ABI for the full GPU step", what); }
	B200KernelWranglerManager* mWrangler;
};

static B200PhysXGpu* gInstance = NULL;
void B200PhysXGpu::release() { if(mWrangler) { PX_DELETE(mWrangler); } gInstance = NULL; PX_DELETE_THIS; }
}	// namespace

// ------------------------------------------------------------------------------------------------------------------------------------------
// the 12 exports of physxgpu/include/PxPhysXGpu.h:207-237
#define B200_EXPORT extern "C" __attribute__((visibility("default")))

B200_EXPORT physx::PxPhysXGpu* PxCreatePhysXGpu()
{
	if(!gInstance) gInstance = PX_NEW(B200PhysXGpu)();
	return gInstance;
}

B200_EXPORT void PxSetPhysXGpuFoundationInstance(physx::PxFoundation& foundation) { PxSetFoundationInstance(foundation); }

B200_EXPORT physx::PxCudaContextManager* PxCreateCudaContextManager(physx::PxFoundation& foundation, const physx::PxCudaContextManagerDesc& desc, physx::PxProfilerCallback*, bool)
{
	PxSetFoundationInstance(foundation);
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); B200_ERROR(PxErrorCode::eDEBUG_WARNING, "libPhysXGpu_64 (physx_b200): no CUDA device"); return NULL; }
	if(desc.ctx && *desc.ctx) B200_ERROR(PxErrorCode::eDEBUG_INFO, "libPhysXGpu_64 (physx_b200): an application-provided CUcontext is ignored; the device's primary context is used");
	const int device = desc.deviceOrdinal >= 0 ? desc.deviceOrdinal : 0;
	if(device >= n) { B200_ERROR(PxErrorCode::eINVALID_PARAMETER, "libPhysXGpu_64 (physx_b200): device ordinal %d out of range (%d devices)", device, n); return NULL; }
	cudaDeviceProp prop;
	if(cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return NULL; }
	if(prop.major < 10) B200_ERROR(PxErrorCode::eDEBUG_WARNING, "libPhysXGpu_64 (physx_b200): device %d is sm_%d%d; the kernels of libphysx_b200.so are built for sm_100a only", device, prop.major, prop.minor);
	int driver = 0; cudaDriverGetVersion(&driver);
	B200CudaContextManager* m = PX_NEW(B200CudaContextManager)(device, prop, driver);
	if(!m->contextIsValid()) { m->release(); return NULL; }
	return m;
}

B200_EXPORT int PxGetSuggestedCudaDeviceOrdinal(physx::PxErrorCallback&)
{
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return -1; }
	return 0;
}

B200_EXPORT void PxSetPhysXGpuProfilerCallback(physx::PxProfilerCallback*) {}
// Kernel registration tables of the reference's runtime-loaded fatbins (extensions register their own CUDA modules through these): this plugin
// loads no modules, the tables are empty.
B200_EXPORT void PxGpuCudaRegisterFunction(int, const char*) {}
B200_EXPORT void** PxGpuCudaRegisterFatBinary(void*) { return NULL; }
B200_EXPORT physx::PxKernelIndex* PxGpuGetCudaFunctionTable() { return NULL; }
B200_EXPORT physx::PxU32 PxGpuGetCudaFunctionTableSize() { return 0; }
B200_EXPORT void** PxGpuGetCudaModuleTable() { return NULL; }
B200_EXPORT physx::PxU32 PxGpuGetCudaModuleTableSize() { return 0; }
B200_EXPORT physx::PxPhysicsGpu* PxGpuCreatePhysicsGpu() { return NULL; }
