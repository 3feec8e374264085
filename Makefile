# Build of the B200-native path-tracing hot path.
#   make lib      path_tracer_b200/lib/libptb200.so   (nvcc, sm_100a only)
#   make oracle   CPU oracles (test infrastructure, see oracle/Makefile)
#   make          both
NVCC ?= nvcc
CSRC := path_tracer_b200/csrc
LIB := path_tracer_b200/lib/libptb200.so
# Strict IEEE-754 binary32: no FMA contraction, exact division and sqrt, denormals kept.
NVFLAGS := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
           -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
           -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude -I$(CSRC) $(PT_DEFS)
SRCS := $(CSRC)/pt_kernel.cu $(CSRC)/pt_api.cu $(CSRC)/pt_pack.cpp
HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/pt_abi.h

all: lib oracle

lib: $(LIB)

$(LIB): $(SRCS) $(HDRS)
	mkdir -p $(dir $(LIB))
	$(NVCC) $(NVFLAGS) -shared -cudart static -o $@ $(SRCS)

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB) oracle/libpt_oracle.so

.PHONY: all lib oracle clean
