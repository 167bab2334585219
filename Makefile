# Builds the C-ABI library (sm_100a only) and the host-side record reader.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
SRC := $(wildcard avsr_tf1_b200/csrc/*.cu)
OBJ := $(SRC:.cu=.o)
LIB := avsr_tf1_b200/lib/libavsr_b200.so

IOLIB := avsr_tf1_b200/lib/libavsr_io.so
CXX ?= g++

all: $(LIB) $(IOLIB)

# host-only input pipeline (TFRecord / SequenceExample reader + padded-batch assembler), include/avsr_io.h
$(IOLIB): avsr_tf1_b200/csrc_host/tfrecord.cc include/avsr_io.h
	mkdir -p avsr_tf1_b200/lib
	$(CXX) -O3 -std=c++17 -fPIC -Wall -shared -pthread -o $@ $<

%.o: %.cu avsr_tf1_b200/csrc/common.cuh avsr_tf1_b200/csrc/ap4_common.cuh include/avsr_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p avsr_tf1_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

clean:
	rm -f $(OBJ) $(LIB) $(IOLIB)

# trace variant of the library (tools/ap4d_trace.py): the two-product attention kernel stamps clock64 at its
# synchronisation points.  Never loaded by the product (AVSR_B200_LIB points the tool at it).
TRACELIB := avsr_tf1_b200/lib/libavsr_b200_trace.so
trace: $(TRACELIB)
$(TRACELIB): $(OBJ)
	$(NVCC) $(NVFLAGS) -DAP4D_TRACE -c avsr_tf1_b200/csrc/attn_persist4d.cu -o avsr_tf1_b200/csrc/attn_persist4d.trace.o
	$(NVCC) $(ARCH) -shared -o $@ $(filter-out avsr_tf1_b200/csrc/attn_persist4d.o,$(OBJ)) avsr_tf1_b200/csrc/attn_persist4d.trace.o
