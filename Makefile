# Builds the C-ABI library and the p64b command line without Python (the same commands as p64_b200/build.py).
#   make            -> p64_b200/libp64b200.so, p64_b200/p64b
#   make oracle     -> the test oracle (and, where /root/reference is mounted, oracle/_ref/*)
NVCC ?= nvcc
CXX  ?= g++
ARCH  = -gencode arch=compute_100a,code=sm_100a
CSRC  = p64_b200/csrc
SRCS  = $(CSRC)/device.cu $(CSRC)/bits.cpp $(CSRC)/encoder.cpp $(CSRC)/y4m.cpp $(CSRC)/decoder.cpp
HDRS  = $(CSRC)/kernels.cuh $(CSRC)/vlc_kernels.cuh $(CSRC)/ingest.cuh $(CSRC)/vlc_dev.h $(CSRC)/vlc_tables.h include/p64_b200.h

all: p64_b200/libp64b200.so p64_b200/p64b

p64_b200/libp64b200.so: $(SRCS) $(HDRS)
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall -shared -Xptxas -warn-spills -o $@ $(SRCS) -lpthread

p64_b200/p64b: $(CSRC)/cli.cpp p64_b200/libp64b200.so
	$(CXX) -O2 -std=c++17 -Wall -o $@ $< -Lp64_b200 -lp64b200 -Wl,-rpath,'$$ORIGIN' -lpthread

oracle: p64_b200/libp64b200.so
	$(MAKE) -C oracle

clean:
	rm -f p64_b200/libp64b200.so p64_b200/p64b

.PHONY: all oracle clean
