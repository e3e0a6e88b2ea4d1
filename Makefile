# Top-level build.  `make` builds the sm_100a shared library (CUDA kernels + C ABI), the host-side C++ front end
# and the oracle checker library.  nvcc cross-compiles without a GPU.  Artefacts stay in-tree (git-ignored).
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
PKG      := colibri-core_b200
CSRC     := $(PKG)/csrc
LIBDIR   := $(PKG)/lib
BINDIR   := $(PKG)/bin
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v
LIB      := $(LIBDIR)/libcolibri_b200.so

.PHONY: all lib oracle ref host clean
all: lib oracle host

lib: $(LIB)

$(LIBDIR)/kernels.o: $(CSRC)/kernels.cu $(CSRC)/kernels.h $(CSRC)/device_utils.cuh $(CSRC)/spooky.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/kernels.ptxas.log || (cat $(LIBDIR)/kernels.ptxas.log; false)

$(LIBDIR)/engine.o: $(CSRC)/engine.cu $(CSRC)/relations.h $(CSRC)/kernels.h $(CSRC)/shard.h $(CSRC)/engine_common.h include/colibri_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/engine.ptxas.log || (cat $(LIBDIR)/engine.ptxas.log; false)

$(LIBDIR)/shard.o: $(CSRC)/shard.cu $(CSRC)/shard.h $(CSRC)/kernels.h $(CSRC)/engine_common.h include/colibri_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/shard.ptxas.log || (cat $(LIBDIR)/shard.ptxas.log; false)

$(LIBDIR)/shard_p2p.o: $(CSRC)/shard_p2p.cu $(CSRC)/shard.h $(CSRC)/kernels.h $(CSRC)/engine_common.h include/colibri_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/shard_p2p.ptxas.log || (cat $(LIBDIR)/shard_p2p.ptxas.log; false)

$(LIBDIR)/partition.o: $(CSRC)/partition.cu $(CSRC)/kernels.h $(CSRC)/device_utils.cuh $(CSRC)/spooky.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/partition.ptxas.log || (cat $(LIBDIR)/partition.ptxas.log; false)

$(LIBDIR)/relations.o: $(CSRC)/relations.cu $(CSRC)/relations.h $(CSRC)/kernels.h $(CSRC)/engine_common.h $(CSRC)/device_utils.cuh include/colibri_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/relations.ptxas.log || (cat $(LIBDIR)/relations.ptxas.log; false)

$(LIBDIR)/shard_multi.o: $(CSRC)/shard_multi.cu $(CSRC)/shard.h $(CSRC)/kernels.h $(CSRC)/engine_common.h include/colibri_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/shard_multi.ptxas.log || (cat $(LIBDIR)/shard_multi.ptxas.log; false)

$(LIBDIR)/index.o: $(CSRC)/index.cu $(CSRC)/kernels.h $(CSRC)/device_utils.cuh
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/index.ptxas.log || (cat $(LIBDIR)/index.ptxas.log; false)

$(LIBDIR)/shard_kernels.o: $(CSRC)/shard_kernels.cu $(CSRC)/kernels.h $(CSRC)/device_utils.cuh
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/shard_kernels.ptxas.log || (cat $(LIBDIR)/shard_kernels.ptxas.log; false)

$(LIBDIR)/pattern_index.o: $(CSRC)/pattern_index.cu $(CSRC)/kernels.h $(CSRC)/device_utils.cuh $(CSRC)/spooky.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/pattern_index.ptxas.log || (cat $(LIBDIR)/pattern_index.ptxas.log; false)

$(LIBDIR)/flexgrams.o: $(CSRC)/flexgrams.cu $(CSRC)/kernels.h $(CSRC)/engine_common.h $(CSRC)/device_utils.cuh include/colibri_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/flexgrams.ptxas.log || (cat $(LIBDIR)/flexgrams.ptxas.log; false)

$(LIBDIR)/model_io.o: $(CSRC)/model_io.cu $(CSRC)/kernels.h $(CSRC)/engine_common.h include/colibri_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/model_io.ptxas.log || (cat $(LIBDIR)/model_io.ptxas.log; false)

$(LIB): $(LIBDIR)/kernels.o $(LIBDIR)/engine.o $(LIBDIR)/shard.o $(LIBDIR)/shard_p2p.o $(LIBDIR)/shard_multi.o $(LIBDIR)/partition.o $(LIBDIR)/relations.o $(LIBDIR)/index.o $(LIBDIR)/shard_kernels.o $(LIBDIR)/pattern_index.o $(LIBDIR)/model_io.o $(LIBDIR)/flexgrams.o
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $^

oracle:
	$(MAKE) -C oracle oracle

ref:
	$(MAKE) -C oracle ref

host: lib
	@if [ -f $(PKG)/host/Makefile ]; then $(MAKE) -C $(PKG)/host; fi

clean:
	rm -rf $(LIBDIR) $(BINDIR)
	$(MAKE) -C oracle clean
