# Builds the UNMODIFIED reference CPU SDK (PhysX 5.6.1) from the sources where they lie under
# $(REF) into oracle/_ref/ -- test infrastructure only (the parity oracle + CPU baseline).
# This is our own recipe (plain g++ over the reference's .cpp files); the reference's cmake build
# system is not run.  Flags mirror the reference's linux release configuration
# (physx/source/compiler/cmake/linux/CMakeLists.txt:121-176): -O3, SSE2, no fast-math, no FMA
# contraction on x86-64, NDEBUG, PVD off, static lib.
#
#   make -f oracle/ref_build.mk -j8            # -> oracle/_ref/libphysx_ref.a + oracle/_ref/ref_harness
#
# Nothing under oracle/ is linked or executed by the product path (physx_b200/); only tests/,
# __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may run ref_harness.

REF      ?= /root/reference
PX       := $(REF)/physx
S        := $(PX)/source
OUT      ?= oracle/_ref
OBJ      := $(OUT)/obj

MODULES  := foundation task common geomutils lowlevel lowlevelaabb lowleveldynamics \
            simulationcontroller physx physxextensions scenequery pvd physxcooking \
            immediatemode physxmetadata physxcharacterkinematic

# GPU=1: the same host SDK with PX_SUPPORT_GPU_PHYSX on (no -DDISABLE_CUDA_PHYSX, + physx/src/gpu: the PhysXGpu module loader), into oracle/_ref_gpu/.
# Still CPU code only -- it is the unmodified HOST the plugin shim (plugin/) is loaded into by ref_harness_gpu.
ifeq ($(GPU),1)
OUT      := oracle/_ref_gpu
OBJ      := $(OUT)/obj
EXCL     := /windows/|/omnipvd/|/device/windows|/mac/|/switch/|/android/
CUDADEF  :=
else
EXCL     := /windows/|/gpu/|/omnipvd/|/device/windows|/mac/|/switch/|/android/
CUDADEF  := -DDISABLE_CUDA_PHYSX
endif
SRCS     := $(shell find $(addprefix $(S)/,$(MODULES)) -name '*.cpp' | grep -Ev '$(EXCL)')
INCDIRS  := $(shell find $(S) -type d | grep -Ev '$(EXCL)|/gpu/|/CUDA|/gpu[a-z]*|cudamanager/src|physxgpu/src|compiler|/physxvehicle')
INCS     := -I$(PX)/include $(addprefix -I,$(INCDIRS)) -I$(PX)/pvdruntime/include

DEFS     := -DNDEBUG -DPX_SUPPORT_PVD=0 -DPX_SUPPORT_OMNI_PVD=0 -DPX_PHYSX_STATIC_LIB \
            -DPX_PUBLIC_RELEASE=1 $(CUDADEF) -DPX_NVTX=0
CXXFLAGS := -O3 -std=c++14 -fno-rtti -fno-exceptions -fno-strict-aliasing -ffunction-sections \
            -fdata-sections -fPIC -w $(DEFS)

# object name = path with slashes flattened (several modules reuse file names)
objname   = $(OBJ)/$(subst /,_,$(patsubst $(S)/%.cpp,%,$(1))).o
OBJS     := $(foreach s,$(SRCS),$(call objname,$(s)))

HARNESS_SRC := oracle/ref_harness.cpp oracle/scene_format.h

all: $(OUT)/ref_harness

$(OUT)/libphysx_ref.a: $(OBJS)
	@rm -f $@
	@ar rcs $@ $(OBJS)
	@echo "AR $@ ($(words $(OBJS)) objects)"

define RULE
$(call objname,$(1)): $(1)
	@mkdir -p $(OBJ)
	@g++ $(CXXFLAGS) $(INCS) -c $(1) -o $$@
endef
$(foreach s,$(SRCS),$(eval $(call RULE,$(s))))

$(OUT)/ref_harness: $(HARNESS_SRC) $(OUT)/libphysx_ref.a
	g++ $(CXXFLAGS) -Ioracle $(INCS) oracle/ref_harness.cpp -o $@ \
	    -Wl,--gc-sections -Wl,--start-group $(OUT)/libphysx_ref.a -Wl,--end-group -lpthread -ldl -static-libstdc++ -static-libgcc

lib: $(OUT)/libphysx_ref.a

.PHONY: all lib
