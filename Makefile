# Build libdspfront.so (CUDA, sm_100a only) and the C oracle.
NVCC ?= nvcc
CSRC := dsp_audioreclabs_b200/csrc
SRCS := $(CSRC)/capi.cu $(CSRC)/frontend_pcm.cu $(CSRC)/frontend_pipe.cu $(CSRC)/frontend_exact.cu $(CSRC)/knn.cu $(CSRC)/knn_dense.cu $(CSRC)/knn_tc16.cu $(CSRC)/mfcc_dtw.cu $(CSRC)/misc.cu
HDRS := $(wildcard $(CSRC)/*.cuh) include/dspfront.h
OBJS := $(SRCS:.cu=.o) $(CSRC)/wavio.o
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr
LIB := dsp_audioreclabs_b200/libdspfront.so

all: $(LIB)

%.o: %.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

%.o: %.cpp $(HDRS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-Wall -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS) -lcudart

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB)

.PHONY: all clean
