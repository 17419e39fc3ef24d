# Builds the product library parallelnbody_b200/libnbody_b200.so (CUDA, sm_100a only) and the test-only oracle.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
CSRC := parallelnbody_b200/csrc
LIB := parallelnbody_b200/libnbody_b200.so
OBJS := $(CSRC)/nbody_sim.o $(CSRC)/bh.o $(CSRC)/comm.o
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/nbody.h

ADAPTER := integration/ue4/libue4_adapter.so
all: $(LIB) oracle examples/octree_search $(ADAPTER)
$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(CSRC)/comm.o: $(CSRC)/comm.cpp $(HDRS)
	$(NVCC) $(NVFLAGS) -x cu -c -o $@ $<
$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl
oracle:
	$(MAKE) -C oracle all
# host program written against the reference's actor interface (include/nbody.hpp), linked to the product library only
examples/octree_search: examples/octree_search.cpp include/nbody.hpp include/nbody.h $(LIB)
	g++ -std=c++14 -O2 -Wall -I include $< -o $@ -L parallelnbody_b200 -lnbody_b200 -Wl,-rpath,'$$ORIGIN/../parallelnbody_b200'
# The Unreal-side adapter actor (integration/ue4/OctreeSearch.{h,cpp}: AOctreeSearch forwarding to the C ABI), compiled against
# the stand-in engine header the oracle uses and wrapped by the same C driver as the CPU actor (test infrastructure).
$(ADAPTER): integration/ue4/OctreeSearch.cpp integration/ue4/OctreeSearch.h integration/ue4/NBody.h oracle/ref_wrap.cpp oracle/shim/Engine.h include/nbody.h $(LIB)
	g++ -std=c++14 -O2 -fPIC -Wall -shared -DNBODY_B200_ADAPTER -I integration/ue4 -I oracle/shim -I include -o $@ \
	  oracle/ref_wrap.cpp integration/ue4/OctreeSearch.cpp -L parallelnbody_b200 -lnbody_b200 -Wl,-rpath,'$$ORIGIN/../../parallelnbody_b200'
clean:
	rm -f $(OBJS) $(LIB) examples/octree_search $(ADAPTER)
.PHONY: all oracle clean
