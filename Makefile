# zkcnn_b200 build.  Product: zkcnn_b200/lib/libzkcnn_b200.so (nvcc, sm_100a only).
#   make lib      CUDA library behind include/zkcnn_b200.h
#   make emu      test-only host build of the same sources on the CUDA emulator (tests/emu) -- never shipped
#   make host     stand-alone host side (prover/verifier/circuit builder) -> zkcnn_b200/lib/libzkcnn_host.so + zkcnn_prove
#   make oracle   reference build + oracle restatement (test infrastructure, needs /root/reference for the _ref part)
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
CSRC      := zkcnn_b200/csrc
LIBDIR    := zkcnn_b200/lib
EMUDIR    := tests/emu/_build
CSRC_DEPS := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.hpp $(CSRC)/*.cu) include/zkcnn_b200.h

.PHONY: all lib emu host host_emu oracle clean
all: lib host

lib: $(LIBDIR)/libzkcnn_b200.so
$(LIBDIR)/libzkcnn_b200.so: $(CSRC_DEPS)
	mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -Xcompiler -Wno-unknown-pragmas -shared -cudart static -Xlinker -soname=libzkcnn_b200.so -Xlinker -Bsymbolic $(CSRC)/capi.cu -o $@

emu: $(EMUDIR)/libzkcnn_b200_emu.so
$(EMUDIR)/libzkcnn_b200_emu.so: $(CSRC_DEPS) tests/emu/cuda_emu.cpp tests/emu/cuda_emu.hpp
	mkdir -p $(EMUDIR)
	$(CXX) -O2 -g -std=c++17 -fPIC -DZK_EMU -Wall -Wno-unknown-pragmas -Wno-unused-function -shared -Wl,-Bsymbolic -Wl,-soname=libzkcnn_b200_emu.so -x c++ $(CSRC)/capi.cu -x none tests/emu/cuda_emu.cpp -o $@ -lpthread

# ---- stand-alone host side: circuit compiler, witness generator, protocol driver, C entry points (include/zkcnn_host.h)
HOST      := zkcnn_b200/host
HOST_SRCS := $(HOST)/zk_types.cpp $(HOST)/neuralNetwork.cpp $(HOST)/verifier.cpp $(HOST)/prover.cpp $(HOST)/polyProver.cpp \
             $(HOST)/transcript.cpp $(HOST)/host_api.cpp
HOST_DEPS := $(wildcard $(HOST)/*.hpp $(HOST)/*.cpp) include/zkcnn_b200.h include/zkcnn_host.h $(CSRC)/mont.cuh $(CSRC)/g1.cuh
HOSTFLAGS := -O3 -g -std=c++17 -fPIC -Wall -Wno-unknown-pragmas -pthread

host: $(LIBDIR)/libzkcnn_host.so $(LIBDIR)/zkcnn_prove
$(LIBDIR)/libzkcnn_host.so: $(HOST_DEPS) $(LIBDIR)/libzkcnn_b200.so
	$(CXX) $(HOSTFLAGS) -shared -Wl,-Bsymbolic -Wl,-soname=libzkcnn_host.so $(HOST_SRCS) -L$(LIBDIR) -lzkcnn_b200 -Wl,-rpath,'$$ORIGIN' -o $@
$(LIBDIR)/zkcnn_prove: $(HOST)/main.cpp $(LIBDIR)/libzkcnn_host.so
	$(CXX) $(HOSTFLAGS) $(HOST)/main.cpp -L$(LIBDIR) -lzkcnn_host -lzkcnn_b200 -Wl,-rpath,'$$ORIGIN' -o $@

# test-only: the same host side on top of the emulator build of the kernels
host_emu: $(EMUDIR)/libzkcnn_host_emu.so $(EMUDIR)/zkcnn_prove_emu
$(EMUDIR)/libzkcnn_host_emu.so: $(HOST_DEPS) $(EMUDIR)/libzkcnn_b200_emu.so
	$(CXX) $(HOSTFLAGS) -shared -Wl,-Bsymbolic -Wl,-soname=libzkcnn_host_emu.so $(HOST_SRCS) -L$(EMUDIR) -lzkcnn_b200_emu -Wl,-rpath,'$$ORIGIN' -o $@
$(EMUDIR)/zkcnn_prove_emu: $(HOST)/main.cpp $(EMUDIR)/libzkcnn_host_emu.so
	$(CXX) $(HOSTFLAGS) $(HOST)/main.cpp -L$(EMUDIR) -lzkcnn_host_emu -lzkcnn_b200_emu -Wl,-rpath,'$$ORIGIN' -o $@

oracle:
	$(MAKE) -C oracle ref dropin dropin_emu

clean:
	rm -rf $(LIBDIR) $(EMUDIR)
