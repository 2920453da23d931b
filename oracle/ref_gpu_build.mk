# Second baseline: the reference's OWN GPU plugin, libPhysXGpu_64.so, built from the sources where they lie under $(REF) for sm_100
# only -- our recipe (plain nvcc / g++ over the reference's gpu* / cudamanager / physxgpu sources with the flags of its cmakegpu lists:
# -use_fast_math -ftz=true -prec-div=false -prec-sqrt=false); the reference's cmake is not run.  Output: oracle/_ref_gpu/reference_plugin/.
# It is loaded into the unmodified host SDK (oracle/_ref_gpu/ref_harness --gpu-plugin ... --gpu-bp --gpu-dynamics) on the GPU box and timed next to
# bench.py.  Test / measurement infrastructure only.
#   make -f oracle/ref_gpu_build.mk -j8
REF      ?= /root/reference
PX       := $(REF)/physx
S        := $(PX)/source
OUT      := oracle/_ref_gpu/reference_plugin
OBJ      := $(OUT)/obj
NVCC     ?= /usr/local/cuda/bin/nvcc
GPUMODS  := gpubroadphase gpucommon gpunarrowphase gpusimulationcontroller gpusolver gpuarticulation cudamanager physxgpu
EXCL     := /windows/
CU       := $(shell find $(addprefix $(S)/,$(GPUMODS)) -name '*.cu' | grep -Ev '$(EXCL)')
CPP      := $(shell find $(addprefix $(S)/,$(GPUMODS)) -name '*.cpp' | grep -Ev '$(EXCL)')
INCDIRS  := $(shell find $(S) -type d | grep -Ev '/windows|/omnipvd|/mac/|/switch/|/android/|compiler|/physxvehicle')
INCS     := -I$(PX)/include $(addprefix -I,$(INCDIRS)) -I$(PX)/pvdruntime/include -I/usr/local/cuda/include -Ioracle/stubs
DEFS     := -DNDEBUG -DPX_SUPPORT_PVD=0 -DPX_SUPPORT_OMNI_PVD=0 -DPX_PHYSX_STATIC_LIB -DPX_PHYSX_GPU_EXPORTS -DPX_PUBLIC_RELEASE=1 -DPX_NVTX=0 -D_CONSOLE
CXXFLAGS := -O3 -std=c++14 -fno-rtti -fno-exceptions -fno-strict-aliasing -ffunction-sections -fdata-sections -fvisibility=hidden -fPIC -w $(DEFS)
NVFLAGS  := -gencode arch=compute_100,code=sm_100 -O3 -std=c++14 -use_fast_math -ftz=true -prec-div=false -prec-sqrt=false -w $(DEFS) \
            --compiler-options=-O3,-fPIC,-msse2,-mfpmath=sse,-m64,-fvisibility=hidden,-fno-strict-aliasing
objname   = $(OBJ)/$(subst /,_,$(patsubst $(S)/%,%,$(basename $(1)))).o
CUOBJ    := $(foreach s,$(CU),$(call objname,$(s)))
CPPOBJ   := $(foreach s,$(CPP),$(call objname,$(s)))

all: $(OUT)/libPhysXGpu_64.so
define CURULE
$(call objname,$(1)): $(1)
	@mkdir -p $(OBJ)
	@$(NVCC) $(NVFLAGS) $(INCS) -c $(1) -o $$@
endef
define CPPRULE
$(call objname,$(1)): $(1)
	@mkdir -p $(OBJ)
	@g++ $(CXXFLAGS) $(INCS) -c $(1) -o $$@
endef
$(foreach s,$(CU),$(eval $(call CURULE,$(s))))
$(foreach s,$(CPP),$(eval $(call CPPRULE,$(s))))
$(OUT)/libPhysXGpu_64.so: $(CUOBJ) $(CPPOBJ) oracle/_ref_gpu/libphysx_ref.a
	g++ -shared -o $@ $(CUOBJ) $(CPPOBJ) -Wl,-Bsymbolic -Wl,--gc-sections -Wl,--exclude-libs,ALL -Wl,--start-group oracle/_ref_gpu/libphysx_ref.a -Wl,--end-group \
	    -L/usr/local/cuda/lib64 -lcudart_static -L/usr/local/cuda/lib64/stubs -lcuda -ldl -lrt -lpthread -static-libstdc++ -static-libgcc
	@echo "built $@ ($(words $(CUOBJ)) CUDA + $(words $(CPPOBJ)) C++ objects)"
.PHONY: all
