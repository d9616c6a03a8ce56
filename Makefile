# Builds libaewn.so (the C-ABI CUDA library, sm_100a only), the tcgen05 probe and the C oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v
PKG       := ae-wavenet_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/aewn/libaewn.so
SRCS      := $(CSRC)/host_util.cu $(CSRC)/tgemm.cu $(CSRC)/wgrad.cu $(CSRC)/wgradw.cu $(CSRC)/misc.cu $(CSRC)/vq.cu $(CSRC)/gen.cu $(CSRC)/grcc_fwd.cu $(CSRC)/loader.cu $(CSRC)/wgradh.cu
HDRS      := $(CSRC)/ptx.cuh $(CSRC)/host_util.h include/aewn.h

all: $(LIB) probe oracle

$(LIB): $(SRCS) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS) -lcudart 2> build_ptxas.log || (cat build_ptxas.log; exit 1)
	@grep -E "error|spill" build_ptxas.log | grep -v " 0 bytes spill" || true

probe: build/aewn_probe
build/aewn_probe: $(CSRC)/probe_main.cu $(LIB)
	@mkdir -p build
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ $(CSRC)/probe_main.cu -L$(PKG)/aewn -laewn -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)/aewn'

oracle: oracle/liboracle.so
oracle/liboracle.so: oracle/vq_oracle.c
	gcc -O2 -fPIC -shared -ffp-contract=off -o $@ $< -lm

clean:
	rm -f $(LIB) build/aewn_probe oracle/liboracle.so build_ptxas.log

.PHONY: all probe oracle clean
