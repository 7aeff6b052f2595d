# Builds the product library (sm_100a only) and, through oracle/Makefile, the test oracle.
NVCC     ?= /usr/local/cuda/bin/nvcc
PKG      := rs-sync_b200
SRC      := $(PKG)/csrc
OUT      ?= $(PKG)/lib
EXTRA    ?=
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -fmad=false -std=c++17 -Iinclude -I$(SRC) $(EXTRA) \
            -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-unused-function
OBJS     := $(OUT)/engine.o $(OUT)/capi.o $(OUT)/host_ingest.o $(OUT)/cxx_dropin.o

.PHONY: all lib oracle ref clean
all: lib oracle
lib: $(OUT)/librssync_b200.so

$(OUT)/engine.o: $(SRC)/engine.cu $(SRC)/engine.h $(SRC)/device_math.cuh $(SRC)/rng.h $(SRC)/spec_trig.h $(SRC)/gyro_scan.h
	@mkdir -p $(OUT)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OUT)/engine.ptxas.log || (cat $(OUT)/engine.ptxas.log; false)

$(OUT)/%.o: $(SRC)/%.cpp $(SRC)/engine.h $(SRC)/host_ingest.h $(SRC)/rng.h $(SRC)/spec_trig.h $(SRC)/gyro_scan.h $(SRC)/nccl_dyn.h include/rssync_b200.h include/rssync.h
	@mkdir -p $(OUT)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OUT)/librssync_b200.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static

oracle:
	$(MAKE) -C oracle oracle
ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf $(OUT)
	$(MAKE) -C oracle clean
