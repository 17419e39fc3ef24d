# Builds the product library parallelnbody_b200/libnbody_b200.so (CUDA, sm_100a only) and the test-only oracle.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
CSRC := parallelnbody_b200/csrc
LIB := parallelnbody_b200/libnbody_b200.so
OBJS := $(CSRC)/nbody_sim.o $(CSRC)/bh.o $(CSRC)/comm.o
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/nbody.h

all: $(LIB) oracle examples/octree_search
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
clean:
	rm -f $(OBJS) $(LIB) examples/octree_search
.PHONY: all oracle clean
