# Builds libscgbm.so (sm_100a only) in-tree: scgbm_b200/libscgbm.so
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
SRC := $(wildcard scgbm_b200/csrc/*.cu)
OBJ := $(patsubst scgbm_b200/csrc/%.cu,build/%.o,$(SRC))
HDR := $(wildcard scgbm_b200/csrc/*.cuh) include/scgbm.h
LIB := scgbm_b200/libscgbm.so

all: $(LIB)

build/%.o: scgbm_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -ldl

clean:
	rm -rf build $(LIB)
.PHONY: all clean
