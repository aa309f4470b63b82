# Plain-make build of libdrv_gi for a C++ host (the reference is a C++ project; no Python needed to build or link).
# Same flags as dynamicradiancevolume_b200/build.py, which __graft_entry__.build() and the test-suite use.
#   make            libdrv_gi.so (nvcc, sm_100a) + libdrv_host.so (the uniform-block packers alone, g++)
#   make aux        the test helpers: oracle/ (CPU restatement, test infrastructure), scenes/ (procedural inputs)
#   make cpptest    tests/cpp/renderer_parity: drv::Renderer (include/drv_renderer.hpp) against the oracle
#   make check      the C++ test's host half (no GPU) — on a B200 run tests/cpp/renderer_parity without arguments
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
PKG := dynamicradiancevolume_b200
CSRC := $(PKG)/csrc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
SOURCES := ctx.cu alloc.cu rsm.cu voxel.cu gather.cu apply.cu adjacent.cu specular.cu microbench.cu host_pack.cpp
OBJS := $(addprefix $(CSRC)/,$(addsuffix .o,$(basename $(SOURCES))))
HEADERS := $(CSRC)/ctx.h $(CSRC)/device_math.cuh $(CSRC)/voxel_sample.cuh include/drv_gi.h include/drv_math.h include/drv_r11g11b10.h

all: $(PKG)/libdrv_gi.so $(PKG)/libdrv_host.so

$(CSRC)/%.o: $(CSRC)/%.cu $(HEADERS)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@
$(CSRC)/%.o: $(CSRC)/%.cpp $(HEADERS)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(PKG)/libdrv_gi.so: $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -lcudart

$(PKG)/libdrv_host.so: $(CSRC)/host_pack.cpp include/drv_gi.h include/drv_math.h
	$(CXX) -O2 -std=c++17 -fPIC -Wall -shared -o $@ $<

aux:
	$(MAKE) -C oracle
	$(MAKE) -C scenes

cpptest: all aux
	$(MAKE) -C tests/cpp

check: cpptest
	tests/cpp/renderer_parity --host

clean:
	rm -f $(OBJS) $(PKG)/libdrv_gi.so $(PKG)/libdrv_host.so
	$(MAKE) -C tests/cpp clean

.PHONY: all aux cpptest check clean
