# Builds libftk_b200.so (the product: CUDA kernels + C ABI, sm_100a only) in-tree so it travels to the GPU box.
NVCC ?= /usr/local/cuda/bin/nvcc
CSRC := feature_tracker_b200/csrc
OUT := feature_tracker_b200/libftk_b200.so
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the KLT / matching numerics reproduce the reference's un-fused fp32 arithmetic bit for bit.
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude -I$(CSRC) --expt-relaxed-constexpr
SRCS := $(CSRC)/api.cu $(CSRC)/pyramid.cu $(CSRC)/klt.cu $(CSRC)/klt_basic_fastpath.cu $(CSRC)/direct_method.cu $(CSRC)/dense_flow.cu $(CSRC)/detect.cu $(CSRC)/match.cu $(CSRC)/match_mutual.cu $(CSRC)/match_cosine_tc.cu
OBJS := $(SRCS:.cu=.o)
HDRS := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/ftk_c.h

.PHONY: all clean oracle ref
all: $(OUT) tools/microbench tools/c1_latency

# measurement helpers (not part of the product library): pipe-rate microbenchmark for the roofline denominators, C1 latency probe
tools/microbench: tools/microbench.cu
	$(NVCC) $(ARCH) -O3 -o $@ $< -lcudart_static -lpthread -ldl -lrt
tools/c1_latency: tools/c1_latency.cpp include/ftk_c.h include/feature_tracker_b200/feature_tracker.h $(OUT)
	g++ -std=c++17 -O2 -Wall -Iinclude -o $@ $< -Lfeature_tracker_b200 -lftk_b200 -Wl,-rpath,'$$ORIGIN/../feature_tracker_b200'

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(OUT): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

oracle:
	$(MAKE) -C oracle oracle
ref:
	$(MAKE) -C oracle ref

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(OUT)
