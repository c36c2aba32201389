# tests/cpp/reference_suite.mk -- TEST INFRASTRUCTURE: the drop-in proof.
# Compiles the REFERENCE's own sources, unchanged and from where they lie (REF=/root/reference),
# against THIS repo's include/ tree and libloopsb200.so:
#   * every unittests/test_*.cu  (Catch2 replaced by tests/cpp/catch2_mini, nothing else)
#   * every examples/spmv/*.cu as .f32 and .f64 (cxxopts replaced by tests/cpp/stubs/cxxopts.hpp),
#     plus examples/{range,saxpy,spmm}
# Outputs only into tests/_refsuite/ (git-ignored, NOT gpurun-ignored: the binaries travel to
# the GPU box, where tests/test_gpu_reference_suite.py runs them). The reference's own build
# system (cmake + FetchContent) is not used.
REF   ?= /root/reference
ROOT  := $(abspath $(dir $(lastword $(MAKEFILE_LIST)))/../..)
OUT   := $(ROOT)/tests/_refsuite
NVCC  ?= nvcc
FLAGS := -std=c++17 -O1 --expt-extended-lambda --expt-relaxed-constexpr -w \
         -gencode arch=compute_100a,code=sm_100a -I$(ROOT)/include \
         -L$(ROOT)/loops_b200 -lloopsb200 -Xlinker -rpath -Xlinker '$$ORIGIN/../../loops_b200'

UNIT  := $(basename $(notdir $(wildcard $(REF)/unittests/test_*.cu)))
EXS   := $(basename $(notdir $(wildcard $(REF)/examples/spmv/*.cu)))
HDRS  := $(shell find $(ROOT)/include -type f) \
         $(ROOT)/tests/cpp/catch2_mini/catch2/catch_test_macros.hpp $(ROOT)/tests/cpp/stubs/cxxopts.hpp

all: $(addprefix $(OUT)/unit.,$(UNIT)) $(addprefix $(OUT)/loops.spmv.,$(addsuffix .f32,$(EXS))) \
     $(addprefix $(OUT)/loops.spmv.,$(addsuffix .f64,$(EXS))) $(OUT)/loops.range $(OUT)/loops.saxpy $(OUT)/loops.spmm.thread_mapped

$(OUT)/unit.%: $(REF)/unittests/%.cu $(HDRS) | $(ROOT)/loops_b200/libloopsb200.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -I$(ROOT)/tests/cpp/catch2_mini -I$(REF)/unittests $< -o $@

$(OUT)/loops.spmv.%.f32: $(REF)/examples/spmv/%.cu $(HDRS) | $(ROOT)/loops_b200/libloopsb200.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -DLOOPS_VALUE_T=float -I$(ROOT)/tests/cpp/stubs -I$(REF)/examples/spmv $< -o $@

$(OUT)/loops.spmv.%.f64: $(REF)/examples/spmv/%.cu $(HDRS) | $(ROOT)/loops_b200/libloopsb200.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -DLOOPS_VALUE_T=double -I$(ROOT)/tests/cpp/stubs -I$(REF)/examples/spmv $< -o $@

$(OUT)/loops.range: $(REF)/examples/range/range.cu $(HDRS) | $(ROOT)/loops_b200/libloopsb200.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -I$(ROOT)/tests/cpp/stubs $< -o $@
$(OUT)/loops.saxpy: $(REF)/examples/saxpy/saxpy.cu $(HDRS) | $(ROOT)/loops_b200/libloopsb200.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -I$(ROOT)/tests/cpp/stubs $< -o $@
$(OUT)/loops.spmm.thread_mapped: $(REF)/examples/spmm/thread_mapped.cu $(HDRS) | $(ROOT)/loops_b200/libloopsb200.so
	@mkdir -p $(OUT)
	$(NVCC) $(FLAGS) -I$(ROOT)/tests/cpp/stubs -I$(REF)/examples/spmm $< -o $@

clean:
	rm -rf $(OUT)
.PHONY: all clean
