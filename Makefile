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
SRCS := $(CSRC)/pt_wave.cu $(CSRC)/pt_lane.cu $(CSRC)/pt_api.cu $(CSRC)/pt_pack.cpp
OBJS := $(patsubst $(CSRC)/%,build/obj/%.o,$(SRCS))
HDRS := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/pt_abi.h

all: lib oracle host

lib:
	$(MAKE) -j4 $(LIB)

build/obj/%.o: $(CSRC)/% $(HDRS)
	mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -x cu -c -o $@ $<

$(LIB): $(OBJS)
	mkdir -p $(dir $(LIB))
	$(NVCC) $(NVFLAGS) -shared -cudart static -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle

# Host side: the reference's UNMODIFIED application (src/main.cpp, compiled in place) on top of the
# C++ host mirror in path_tracer_b200/include + compat and libptb200.so.  Only where the reference
# tree exists; the binary travels to the GPU box.
REFERENCE ?= /root/reference
HOST_HDRS := $(wildcard path_tracer_b200/include/pt/*.hpp path_tracer_b200/compat/*.hpp path_tracer_b200/compat/*/*.h*) include/ptscene_io.hpp
ifneq ($(wildcard $(REFERENCE)/src/main.cpp),)
host: build/sycl-rt-b200 build/sycl-rt-b200-st
build/sycl-rt-b200: $(REFERENCE)/src/main.cpp $(HOST_HDRS) $(LIB)
	mkdir -p build
	g++ -std=c++20 -O2 -ffp-contract=off -w -DOUTPUT_WIDTH=800 -DOUTPUT_HEIGHT=480 \
	    -Ipath_tracer_b200/compat -Ipath_tracer_b200/include -Iinclude $< -o $@ \
	    -Lpath_tracer_b200/lib -lptb200 -Wl,-rpath,'$$ORIGIN/../path_tracer_b200/lib'
# the same application built the reference's FPGA way (CMake option USE_SINGLE_TASK), small: the mode is one serial chain
build/sycl-rt-b200-st: $(REFERENCE)/src/main.cpp $(HOST_HDRS) $(LIB)
	mkdir -p build
	g++ -std=c++20 -O2 -ffp-contract=off -w -DUSE_SINGLE_TASK -DOUTPUT_WIDTH=32 -DOUTPUT_HEIGHT=24 \
	    -Ipath_tracer_b200/compat -Ipath_tracer_b200/include -Iinclude $< -o $@ \
	    -Lpath_tracer_b200/lib -lptb200 -Wl,-rpath,'$$ORIGIN/../path_tracer_b200/lib'
else
host:
	@echo "host: $(REFERENCE) not present; keeping prebuilt build/sycl-rt-b200 as is"
endif

clean:
	rm -f $(LIB) oracle/libpt_oracle.so

.PHONY: all lib oracle host clean
