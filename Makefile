# Builds the C-ABI shared library of the B200 PIR answer path (sm_100a only) and the CPU oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
SRC := pir_b200/csrc
OBJ := build/obj
LIB := pir_b200/lib/libpirb200.so
OBJS := $(OBJ)/kernels_ntt.o $(OBJ)/kernels_stream.o $(OBJ)/kernels_cluster.o $(OBJ)/context.o
HDRS := $(wildcard $(SRC)/*.h $(SRC)/*.cuh) include/pir_b200.h

all: $(LIB) oracle

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p pir_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build pir_b200/lib oracle/_build
.PHONY: all oracle clean
