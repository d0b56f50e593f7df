# Builds, in-tree:
#   vanrijn_b200/lib/libvanrijn_cuda.so  -- the sm_100a kernels + C ABI (include/vanrijn_cuda.h)
#   vanrijn_b200/lib/libvanrijn_host.so  -- C++ host mirror of the reference API (include/vanrijn.hpp) + C shim
#   oracle/libvanrijn_oracle.so          -- CPU oracle (test infrastructure only)
# nvcc cross-compiles for sm_100a without a GPU.  -fmad=false: the reference (rustc) never fuses a*b+c.
NVCC ?= nvcc
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper without libgomp.spec); use the PATH compiler
HOSTCXX = g++
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 --extended-lambda \
          -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
LIBDIR ?= vanrijn_b200/lib
EXTRA ?=
CSRC = vanrijn_b200/csrc
CUDA_DEPS = $(CSRC)/vanrijn_cuda.cu $(CSRC)/vrj_kernels.cuh $(CSRC)/vrj_traverse.cuh $(CSRC)/vrj_device.cuh \
            $(CSRC)/rgb_basis_tables.inc include/vanrijn_cuda.h
HOST_DEPS = $(CSRC)/host/vanrijn_host.cpp $(CSRC)/host/host_capi.cpp include/vanrijn.hpp include/vanrijn_cuda.h

all: $(LIBDIR)/libvanrijn_cuda.so $(LIBDIR)/libvanrijn_host.so oracle examples

# two translation units (the render loop; the device BVH builder), compiled separately so a change to one
# does not recompile the other
OBJDIR ?= build/obj
$(OBJDIR)/vanrijn_cuda.o: $(CUDA_DEPS)
	mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(EXTRA) -c -o $@ $(CSRC)/vanrijn_cuda.cu
$(OBJDIR)/vrj_bvh_build.o: $(CSRC)/vrj_bvh_build.cu include/vanrijn_cuda.h
	mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(EXTRA) -c -o $@ $(CSRC)/vrj_bvh_build.cu

$(LIBDIR)/libvanrijn_cuda.so: $(OBJDIR)/vanrijn_cuda.o $(OBJDIR)/vrj_bvh_build.o
	mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $^ -ldl

$(LIBDIR)/libvanrijn_host.so: $(HOST_DEPS) $(LIBDIR)/libvanrijn_cuda.so
	$(HOSTCXX) -O2 -std=c++17 -fPIC -ffp-contract=off -Wall -Wextra -shared -o $@ \
	    $(CSRC)/host/vanrijn_host.cpp $(CSRC)/host/host_capi.cpp -L$(LIBDIR) -lvanrijn_cuda -pthread -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -s -C oracle

# build/vanrijn: the reference's harness (src/main.rs) over the host mirror; build/drop_in_example: doc-test + bench scene
examples: build/vanrijn build/drop_in_example
build/%: examples/%.cpp $(LIBDIR)/libvanrijn_host.so include/vanrijn.hpp
	mkdir -p build
	$(HOSTCXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -o $@ $< -L$(LIBDIR) -lvanrijn_host -lvanrijn_cuda -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'
build/vanrijn: examples/vanrijn_main.cpp $(LIBDIR)/libvanrijn_host.so include/vanrijn.hpp
	mkdir -p build
	$(HOSTCXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -o $@ $< -L$(LIBDIR) -lvanrijn_host -lvanrijn_cuda -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'

clean:
	rm -f $(LIBDIR)/*.so oracle/*.so

.PHONY: all oracle clean examples
