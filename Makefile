# Build of the B200-native PNOL hot path. Everything is compiled for sm_100a only.
#   make            -> product libraries + oracle restatement
#   make ref        -> oracle/_ref (verbatim reference, needs /root/reference)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
PKG       := parallelnonlinearoptimizationlibrary_b200
CSRC      := $(PKG)/csrc
HOSTSRC   := $(PKG)/host
LIBDIR    := $(PKG)/lib
BUILD     := build

NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
             -Xcompiler -fPIC,-ffp-contract=off -Iinclude -I$(CSRC) -Xptxas -v
CXXFLAGS  := -O2 -std=c++17 -fPIC -ffp-contract=off -Iinclude -I$(HOSTSRC) -Wall

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(CU_SRCS))
HOST_SRCS := $(wildcard $(HOSTSRC)/*.cpp)
HOST_OBJS := $(patsubst $(HOSTSRC)/%.cpp,$(BUILD)/host_%.o,$(HOST_SRCS))

all: $(LIBDIR)/libpnol_b200.so $(LIBDIR)/libpnol_b200_host.so oracle ubench

# tuning aid (FP64 pipe microbenchmark), not part of the library
ubench: tools/ubench/fp64_ubench
tools/ubench/fp64_ubench: tools/ubench/fp64_ubench.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -o $@ $<

$(BUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard include/*.h) $(wildcard include/pnol/*) $(wildcard include/pnol/device/*)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIBDIR)/libpnol_b200.so: $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(CU_OBJS) -ldl -lpthread

$(BUILD)/host_%.o: $(HOSTSRC)/%.cpp $(wildcard include/*.h) $(wildcard include/pnol/*)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIBDIR)/libpnol_b200_host.so: $(HOST_OBJS) $(LIBDIR)/libpnol_b200.so
	$(CXX) -shared -o $@ $(HOST_OBJS) -L$(LIBDIR) -lpnol_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle port

ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf $(BUILD) $(LIBDIR)/*.so oracle/*.so oracle/_ref

.PHONY: all oracle ref clean ubench
