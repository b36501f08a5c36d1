# Builds libmpn_b200.so (hand-written CUDA for sm_100a; C ABI in include/mpn_abi.h),
# plus the test oracle (oracle/). The reference's own Makefile:1-6 builds libnms.so the
# same way (a .so next to the sources, loaded by relative path).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v
CSRC := multipathnet_b200/csrc
SRCS := $(CSRC)/abi.cu $(CSRC)/nms.cu $(CSRC)/roi.cu $(CSRC)/elementwise.cu $(CSRC)/preproc.cu $(CSRC)/post.cu $(CSRC)/dist.cu $(CSRC)/conv_simt.cu $(CSRC)/gemm_tc.cu $(CSRC)/model.cu
OBJS := $(SRCS:.cu=.o)
LIB := multipathnet_b200/libmpn_b200.so

all: $(LIB) oracle
$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/conv_gemm.cuh $(CSRC)/roi.cuh $(CSRC)/image_scale.cuh include/mpn_abi.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)
# nms.cu must keep the reference's unfused fp32 op order: explicit *_rn intrinsics + -fmad=false
$(CSRC)/nms.o: NVFLAGS += -fmad=false
# preproc.cu reproduces image.scale's unfused fp32 arithmetic (image_scale.cuh uses *_rn intrinsics; belt and braces)
$(CSRC)/preproc.o: NVFLAGS += -fmad=false
$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -ldl
oracle:
	$(MAKE) -C oracle
clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB); $(MAKE) -C oracle clean
.PHONY: all oracle clean
