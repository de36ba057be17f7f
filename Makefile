# Builds the product: build/libcrt_host.so (host side), build/libcrt_b200.so (sm_100a device side, the C ABI of
# include/kernels.h) and build/crt_render (the host driver, main.cpp's role).  `make oracle` builds the test oracle.
NVCC  ?= nvcc
CXX   ?= g++
PKG   := cuda-raytracing-optimized_b200
ARCH  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Iinclude -I$(PKG)/csrc
CSRC  := $(wildcard $(PKG)/csrc/*.inc $(PKG)/csrc/*.cu $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h $(PKG)/csrc/*.cpp) include/kernels.h include/rt_types.h
HOSTSRC := $(PKG)/host/host_api.cpp $(PKG)/host/bvh_builder.cpp $(PKG)/host/png_reader.cpp

.PHONY: all oracle clean
all: build/libcrt_host.so build/libcrt_b200.so build/crt_render

build/libcrt_host.so: $(HOSTSRC) $(PKG)/host/host_api.h $(PKG)/host/bvh_builder.h $(PKG)/host/png_reader.h include/rt_types.h
	@mkdir -p build
	$(CXX) -O2 -std=c++17 -fPIC -shared -Iinclude -I$(PKG)/host $(HOSTSRC) -o $@

build/libcrt_b200.so: $(CSRC)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xcompiler -pthread -shared $(PKG)/csrc/renderer.cu $(PKG)/csrc/wide_bvh.cpp -o $@ -ldl

build/crt_render: $(PKG)/host/main.cpp build/libcrt_host.so build/libcrt_b200.so
	$(CXX) -O2 -std=c++17 -Iinclude -I$(PKG)/host $< -o $@ -Lbuild -lcrt_b200 -lcrt_host -Wl,-rpath,'$$ORIGIN'

oracle: all
	$(MAKE) -C oracle all

clean:
	rm -rf build
	$(MAKE) -C oracle clean
