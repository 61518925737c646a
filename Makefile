# ehb200 — builds the CUDA library (sm_100a only) and the CPU oracle.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
CSRC      := embeddinghub_b200/csrc
OBJDIR    := build/obj
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := embeddinghub_b200/libehb200.so

all: $(LIB) oracle tests/cpp/ann_index_cases

# the reference's ANNIndex unit-test cases against the C++ drop-in twin (run by tests/test_gpu_host.py)
tests/cpp/ann_index_cases: tests/cpp/ann_index_cases.cc include/ehb200_ann_index.hpp $(LIB)
	g++ -std=c++17 -O2 -Iinclude $< -Lembeddinghub_b200 -lehb200 -Wl,-rpath,'$$ORIGIN/../../embeddinghub_b200' -o $@

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(CSRC)/kernels.h include/ehb200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -s -C oracle clean

.PHONY: all oracle clean
