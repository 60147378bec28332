# Builds the product library (CUDA kernels + C ABI) for sm_100a, in-tree:
#   odr_audioenc_b200/libtoolame_b200.so
NVCC ?= nvcc
SRC := odr_audioenc_b200/csrc
OUT := odr_audioenc_b200/libtoolame_b200.so
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
           -Xcompiler -fPIC,-Wall,-Wextra,-fvisibility=hidden -Xptxas -v
HDRS := $(wildcard $(SRC)/*.h) include/toolame.h include/toolame_b200.h include/dab_framing_b200.h

all: $(OUT) examples/dabenc
CSRC := $(SRC)/mp2_kernels.cu $(SRC)/mp2_batch.cpp $(SRC)/toolame_shim.cpp $(SRC)/dab_framing.cpp
$(OUT): $(CSRC) $(HDRS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)
examples/dabenc: examples/dabenc.cpp $(OUT) $(HDRS)
	g++ -O2 -std=c++17 -Wall -Wextra -o $@ examples/dabenc.cpp -Lodr_audioenc_b200 -ltoolame_b200 -Wl,-rpath,'$$ORIGIN/../odr_audioenc_b200'
clean:
	rm -f $(OUT) examples/dabenc
