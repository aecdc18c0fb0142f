# Top-level build.  `make` builds the product library; `make oracle` / `make ref`
# build the CPU checkers (test infrastructure, see oracle/Makefile).
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   ?= /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -ccbin $(HOSTCXX) \
             -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-enum-compare --expt-relaxed-constexpr
LIB       := lbm_b200/liblbm_b200.so
SRC       := lbm_b200/csrc/engine.cu
HDR       := lbm_b200/csrc/kernels.cuh lbm_b200/csrc/lattice.cuh include/lbm_b200.h

all: $(LIB)

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o $@ $(SRC) -lcudart 2> lbm_b200/csrc/ptxas.log || (cat lbm_b200/csrc/ptxas.log; false)

oracle:
	$(MAKE) -C oracle

ref:
	$(MAKE) -C oracle ref

clean:
	rm -f $(LIB) lbm_b200/csrc/ptxas.log
	$(MAKE) -C oracle clean

.PHONY: all oracle ref clean
