# Builds libgpet_b200.so (sm_100a only) and the gPET-compatible CLI.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-ffp-contract=off -Xptxas -v $(EXTRA)
CSRC := gpet_b200/csrc
OBJ := build/abi.o build/digitizer.o build/transport.o build/host_io.o build/planner.o
LIB := gpet_b200/libgpet_b200.so

all: $(LIB) bin/gpet_b200

# transport.cu: no implicit FMA contraction, so that the staged kernels and the fused front end (same device functions,
# different surrounding code) produce bit-identical photons, and the arithmetic is the oracle's (-ffp-contract=off);
# every intended FMA is spelled fmaf() in the source
build/transport.o: NVFLAGS += -fmad=false

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.hpp $(CSRC)/*.cuh include/*.h)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

build/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.hpp $(CSRC)/*.cuh include/*.h)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

bin/gpet_b200: $(CSRC)/cli_main.cpp $(LIB)
	@mkdir -p bin
	g++ -O2 -std=c++17 -Iinclude $< -o $@ -Lgpet_b200 -lgpet_b200 -Wl,-rpath,'$$ORIGIN/../gpet_b200'

clean:
	rm -rf build bin $(LIB)
