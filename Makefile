# Build without Python: `make` -> denovo_kmer_b200/libdkb.so (sm_100a) and the CPU oracle.
NVCC ?= nvcc
NVCCFLAGS = -shared -Xcompiler -fPIC -std=c++17 -O3 -lineinfo \
            -gencode arch=compute_100a,code=sm_100a -ccbin g++ -Xcompiler -pthread
CSRC = denovo_kmer_b200/csrc
LIB = denovo_kmer_b200/libdkb.so

all: $(LIB) oracle

$(LIB): $(CSRC)/dkb_api.cu $(CSRC)/dkb_host.cpp $(CSRC)/dkb_device.cuh $(CSRC)/dkb_scan.cuh \
        $(CSRC)/dkb_build.cuh $(CSRC)/dkb_pack.cuh include/dkb.h
	$(NVCC) $(NVCCFLAGS) -o $@ $(CSRC)/dkb_api.cu $(CSRC)/dkb_host.cpp

oracle:
	$(MAKE) -C oracle

# C++ host-layer example (include/dkb.hpp); prints gpu-ok on a B200, no-gpu elsewhere
cpp_trio: $(LIB) tests/cpp_trio.cpp include/dkb.hpp
	g++ -std=c++17 -Wall -Iinclude tests/cpp_trio.cpp -o $@ -Ldenovo_kmer_b200 -ldkb \
	    -Wl,-rpath,$(CURDIR)/denovo_kmer_b200

clean:
	rm -f $(LIB) cpp_trio
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
