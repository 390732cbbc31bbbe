# Builds the B200-native shared library (C ABI in include/scanner_b200.h), the host-side
# plugin-surface library and the oracle (test infrastructure).  sm_100a only.
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
CSRC      := scanner_b200/csrc
BUILD     := build
LIB       := scanner_b200/libscanner_b200.so

KOBJS := $(BUILD)/scn_k_byte.o $(BUILD)/scn_k_short.o $(BUILD)/scn_k_shortc.o $(BUILD)/scn_k_float.o
OBJS  := $(BUILD)/scn_api.o $(BUILD)/scn_records.o $(BUILD)/scn_large.o $(BUILD)/scn_hackrf.o $(BUILD)/scn_convert.o $(BUILD)/scn_exchange.o $(BUILD)/scn_nccl.o $(BUILD)/scn_cluster.o $(KOBJS)
HOSTSRC := $(wildcard $(CSRC)/host/*.cpp)
HOSTOBJS := $(patsubst $(CSRC)/host/%.cpp,$(BUILD)/host_%.o,$(HOSTSRC))
KHDRS := $(CSRC)/scn_fft.cuh $(CSRC)/scn_kernel.cuh $(CSRC)/scn_wpt.cuh $(CSRC)/scn_p64.cuh $(CSRC)/scn_wconst.cuh $(CSRC)/scn_dispatch.h $(CSRC)/scn_timedomain.cuh include/scanner_b200.h

TOOL      := scanner_b200/scan_b200

SELFTEST  := scanner_b200/host_selftest

all: $(LIB) $(TOOL) $(SELFTEST) oracle

$(SELFTEST): $(CSRC)/tools/host_selftest.cpp $(LIB) $(wildcard $(CSRC)/host/*.h)
	$(CXX) -O1 -std=c++17 -Iinclude -I$(CSRC)/host -o $@ $< -Lscanner_b200 -lscanner_b200 -lpthread -Wl,-rpath,'$$ORIGIN'

$(TOOL): $(CSRC)/tools/scan_b200.cpp $(LIB) $(wildcard $(CSRC)/host/*.h)
	$(CXX) -O2 -std=c++17 -Iinclude -I$(CSRC)/host -o $@ $< -Lscanner_b200 -lscanner_b200 -lpthread -Wl,-rpath,'$$ORIGIN'

$(BUILD):
	mkdir -p $(BUILD)

$(BUILD)/%.o: $(CSRC)/%.cu $(KHDRS) | $(BUILD)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(BUILD)/host_%.o: $(CSRC)/host/%.cpp $(wildcard $(CSRC)/host/*.h) include/scanner_b200.h | $(BUILD)
	$(CXX) -O2 -std=c++17 -fPIC -Wall -Iinclude -I$(CSRC)/host -c $< -o $@

$(LIB): $(OBJS) $(HOSTOBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) $(HOSTOBJS) -cudart shared -lpthread -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(BUILD) $(LIB) $(TOOL) $(SELFTEST)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
