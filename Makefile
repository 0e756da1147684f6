# Builds the product library (CUDA, sm_100a only) in-tree: zkp_subnet_b200/libzkp_b200.so
NVCC ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --diag-suppress 550
SRC := zkp_subnet_b200/csrc/zkp_b200.cu
HDR := $(wildcard zkp_subnet_b200/csrc/*.cuh zkp_subnet_b200/csrc/*.h zkp_subnet_b200/csrc/*.hpp zkp_subnet_b200/csrc/host/*.hpp include/*.h)
LIB := zkp_subnet_b200/libzkp_b200.so
# CPython-side wire codec helper (List[str] <-> bytes); plain g++, no CUDA, loaded with ctypes.PyDLL
WIRE := zkp_subnet_b200/_zkp_wire.so
PYINC := $(shell python -c "import sysconfig; print(sysconfig.get_paths()['include'])")

all: $(LIB) $(WIRE)

$(WIRE): zkp_subnet_b200/csrc/wire_py.cpp zkp_subnet_b200/csrc/codec.hpp
	g++ -O3 -std=c++17 -shared -fPIC -I$(PYINC) -o $@ $< -lpthread

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRC) 2> build/ptxas_zkp_b200.log || (cat build/ptxas_zkp_b200.log; exit 1)
	@grep -E "error|warning" build/ptxas_zkp_b200.log | grep -v "pragma" | head -20 || true

microbench: tools/microbench.cu $(HDR)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/microbench tools/microbench.cu

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB) $(WIRE) build/*
.PHONY: all oracle clean microbench
