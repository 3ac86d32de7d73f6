# Builds the C-ABI library (tensor_ops_b200/libtops_b200.so) and the bring-up probe, sm_100a only.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Iinclude
CSRC := tensor_ops_b200/csrc
OBJ := build/gemm_sm100.o build/gemm_sm100_inst_kk.o build/gemm_sm100_inst_kmn.o build/gemm_sm100_inst_mnk.o build/gemm_sm100_inst_mnmn.o build/kernels.o build/split_f16.o build/api.o
LIB := tensor_ops_b200/libtops_b200.so

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

GEMM_OBJ := build/gemm_sm100.o build/gemm_sm100_inst_kk.o build/gemm_sm100_inst_kmn.o build/gemm_sm100_inst_mnk.o build/gemm_sm100_inst_mnmn.o
probe: $(GEMM_OBJ) build/split_f16.o tools/gemm_probe.cu
	$(NVCC) $(NVFLAGS) -o tools/gemm_probe tools/gemm_probe.cu $(GEMM_OBJ) build/split_f16.o -lcudart

clean:
	rm -rf build $(LIB) tools/gemm_probe
