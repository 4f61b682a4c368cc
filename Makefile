# Builds the product library (CUDA, sm_100a only) and, for the test-suite, the CPU oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH = -gencode arch=compute_100a,code=sm_100a
CSRC = resampler_b200/csrc
OUT = resampler_b200/lib
NVFLAGS = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
          -Xcompiler -Wall -Xptxas -v
HDRS = $(CSRC)/planner.h $(CSRC)/fir_common.h $(CSRC)/fir_kernels.h $(CSRC)/filter_design.h \
       $(CSRC)/sm100_ptx.cuh $(CSRC)/pcm_ingest.h $(CSRC)/filter_design_device.h $(CSRC)/fir_submit.h $(CSRC)/sinf_glibc.h \
       include/resampler_b200.h

all: $(OUT)/libresampler_b200.so oracle

$(OUT)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OUT)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OUT)/filter_design.o: $(CSRC)/filter_design.cpp $(HDRS)
	@mkdir -p $(OUT)
	g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -Wall -c $< -o $@

$(OUT)/libresampler_b200.so: $(OUT)/fir_api.o $(OUT)/fir_kernels.o $(OUT)/fir_fast.o $(OUT)/fir_tensor.o $(OUT)/fir_tc2.o $(OUT)/fir_submit.o $(OUT)/filter_design_device.o $(OUT)/fft_resampler.o $(OUT)/pcm_ingest.o $(OUT)/microbench.o $(OUT)/filter_design.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(OUT) oracle/_build

.PHONY: all oracle clean
