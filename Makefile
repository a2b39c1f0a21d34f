# Builds the B200-native query library, the synthetic-corpus generator and the CPU oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
PROFILE ?= 0
ifeq ($(PROFILE),1)
EXTRA := -DFG_PROFILE_PHASES
endif
NVFLAGS := $(EXTRA) -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude -Ifugu_b200/csrc
CSRC := fugu_b200/csrc

all: fugu_b200/libfugu_gpu.so fugu_b200/libfugu_host.so fugu_b200/synth/libfugu_synth.so oracle/liboracle.so

$(CSRC)/fg_kernels.o: $(CSRC)/fg_kernels.cu $(CSRC)/fg_internal.h $(CSRC)/fg_device.h $(CSRC)/fg_ptx.h
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@
$(CSRC)/fg_lead.o: $(CSRC)/fg_lead.cu $(CSRC)/fg_internal.h $(CSRC)/fg_device.h $(CSRC)/fg_ptx.h
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@
$(CSRC)/fg_api.o: $(CSRC)/fg_api.cu $(CSRC)/fg_internal.h $(CSRC)/fg_error.h $(CSRC)/fg_pool.h include/fugu_gpu.h
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(CSRC)/fg_host.o: $(CSRC)/fg_host.cpp $(CSRC)/fg_pool.h $(CSRC)/fg_unicode_tables.h include/fugu_gpu.h include/fugu_host.h
	$(CXX) -O2 -std=c++17 -fPIC -Wall -Iinclude -c $< -o $@

# the device library: CUDA kernels + the C ABI of include/fugu_gpu.h
fugu_b200/libfugu_gpu.so: $(CSRC)/fg_kernels.o $(CSRC)/fg_lead.o $(CSRC)/fg_api.o
	$(NVCC) -shared $(ARCH) -o $@ $^ -lpthread -ldl

# the host library (include/fugu_host.h): planner, tokenizer, dataset builder, micro-batcher. No CUDA code and no
# link-time dependency on the device library (bound with dlopen on first use)
fugu_b200/libfugu_host.so: $(CSRC)/fg_host.o
	$(CXX) -shared -Wl,--exclude-libs,ALL -o $@ $^ -lpthread -ldl

fugu_b200/synth/libfugu_synth.so: fugu_b200/synth/synth.cpp
	$(CXX) -O3 -march=x86-64-v2 -std=c++17 -shared -fPIC -pthread -o $@ $<

oracle/liboracle.so: oracle/oracle.cpp oracle/orc.py
	python -c 'import sys; sys.path.insert(0, "."); from oracle import orc; orc.build()'
	@touch $@

clean:
	rm -f $(CSRC)/*.o fugu_b200/libfugu_gpu.so fugu_b200/libfugu_host.so fugu_b200/synth/libfugu_synth.so oracle/liboracle.so

.PHONY: all clean
