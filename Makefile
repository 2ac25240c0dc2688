# Builds libmagpo_b200.so (hand-written sm_100a kernels + C ABI) and the oracle's C helpers.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
SRC_DIR   := magpo_b200/csrc
OBJ_DIR   := build/obj
LIB       := magpo_b200/lib/libmagpo_b200.so
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,$(OBJ_DIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(SRC_DIR)/*.cuh) include/magpo_b200.h

all: $(LIB)

$(OBJ_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(OBJ_DIR)
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl

clean:
	rm -rf build $(LIB)

.PHONY: all clean
