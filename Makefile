# minimod-b200 build.  Product artefacts go to minimod_b200/lib and minimod_b200/bin (git-ignored,
# but they travel to the GPU box).  `make emul` builds the CPU SIMT-emulation of the kernel sources
# used by the CPU-only tests (tests/kernel_emul; never shipped).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v -Xlinker -Bsymbolic
CXXFLAGS  := -O2 -g -std=c++17 -Wall -fPIC
CSRC      := minimod_b200/csrc
HOST      := minimod_b200/host
LIBDIR    := minimod_b200/lib
BINDIR    := minimod_b200/bin
EMUL      := tests/kernel_emul

all: lib host oracle emul emul-cli

lib: $(LIBDIR)/libminimod_cuda.so
$(LIBDIR)/libminimod_cuda.so: $(wildcard $(CSRC)/*) include/minimod_cuda.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/mmc_api.cu 2> $(LIBDIR)/ptxas.log || (cat $(LIBDIR)/ptxas.log; false)
	@grep -E "registers|spill|error" $(LIBDIR)/ptxas.log | head -20 || true

emul: $(EMUL)/_build/libminimod_emul.so
$(EMUL)/_build/libminimod_emul.so: $(wildcard $(CSRC)/*) $(EMUL)/cuda_emul.cpp $(EMUL)/cuda_emul.h include/minimod_cuda.h
	@mkdir -p $(EMUL)/_build
	$(CXX) $(CXXFLAGS) -DMMC_EMUL -I $(EMUL) -x c++ $(CSRC)/mmc_api.cu $(EMUL)/cuda_emul.cpp -shared -Wl,-Bsymbolic -o $@

HOST_SRCS := $(wildcard $(HOST)/*.cpp)
host: $(LIBDIR)/libminimod_host.so $(BINDIR)/minimod
$(LIBDIR)/libminimod_host.so: $(filter-out $(HOST)/main.cpp,$(HOST_SRCS)) $(wildcard $(HOST)/*.h) include/minimod_cuda.h
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -I include -shared -o $@ $(filter-out $(HOST)/main.cpp,$(HOST_SRCS)) -lz -lpthread -ldl
$(BINDIR)/minimod: $(HOST)/main.cpp $(LIBDIR)/libminimod_host.so $(LIBDIR)/libminimod_cuda.so
	@mkdir -p $(BINDIR)
	$(CXX) $(CXXFLAGS) -I include -o $@ $(HOST)/main.cpp -L$(LIBDIR) -lminimod_host -lminimod_cuda -Wl,-rpath,'$$ORIGIN/../lib' -lz -lpthread

# the same CLI linked against the SIMT emulation of the kernels (CPU-only CI; tests only)
emul-cli: $(EMUL)/_build/minimod_emul
$(EMUL)/_build/minimod_emul: $(HOST)/main.cpp $(LIBDIR)/libminimod_host.so $(EMUL)/_build/libminimod_emul.so
	$(CXX) $(CXXFLAGS) -I include -DMINIMOD_VERSION=\"v0.5.0-b200-emul\" -o $@ $(HOST)/main.cpp -L$(LIBDIR) -lminimod_host -L$(EMUL)/_build -lminimod_emul -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../../../minimod_b200/lib' -lz -lpthread

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(LIBDIR) $(BINDIR) $(EMUL)/_build
	$(MAKE) -C oracle clean

.PHONY: all lib emul emul-cli host oracle clean
