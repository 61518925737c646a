# ehb200 — builds the CUDA library (sm_100a only) and the CPU oracle.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
CSRC      := embeddinghub_b200/csrc
OBJDIR    := build/obj
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := embeddinghub_b200/libehb200.so

all: $(LIB) oracle tests/cpp/ann_index_cases tests/cpp/concurrent_search tests/cpp/sharded_two_dev tests/cpp/rwlock_stress

# the reference's ANNIndex unit-test cases against the C++ drop-in twin (run by tests/test_gpu_host.py)
tests/cpp/ann_index_cases: tests/cpp/ann_index_cases.cc include/ehb200_ann_index.hpp $(LIB)
	g++ -std=c++17 -O2 -Iinclude $< -Lembeddinghub_b200 -lehb200 -Wl,-rpath,'$$ORIGIN/../../embeddinghub_b200' -o $@

# 64 pthreads issuing Q=1 searches through the C ABI (the cgo goroutine pattern; run by tests/test_gpu_round2.py)
tests/cpp/concurrent_search: tests/cpp/concurrent_search.c include/ehb200.h $(LIB)
	gcc -std=c11 -O2 -D_POSIX_C_SOURCE=200809L -Iinclude $< -Lembeddinghub_b200 -lehb200 -lpthread -lm -Wl,-rpath,'$$ORIGIN/../../embeddinghub_b200' -o $@

# host-only stress test of the reader/writer lock (run by tests/test_abi_cpu.py; no GPU needed)
tests/cpp/rwlock_stress: tests/cpp/rwlock_stress.cu $(CSRC)/index_impl.h
	$(NVCC) $(ARCH) -O2 -std=c++17 --expt-relaxed-constexpr $< -o $@ -lpthread

# n_dev = 2 through the C ABI (run by tests/test_gpu_round2.py)
tests/cpp/sharded_two_dev: tests/cpp/sharded_two_dev.c include/ehb200.h $(LIB)
	gcc -std=c11 -O2 -Iinclude $< -Lembeddinghub_b200 -lehb200 -lm -Wl,-rpath,'$$ORIGIN/../../embeddinghub_b200' -o $@

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/ehb200.h
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
