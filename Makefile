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

.PHONY: all lib emu host oracle clean
all: lib host

lib: $(LIBDIR)/libzkcnn_b200.so
$(LIBDIR)/libzkcnn_b200.so: $(CSRC_DEPS)
	mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -Xcompiler -Wno-unknown-pragmas -shared -cudart static -Xlinker -soname=libzkcnn_b200.so $(CSRC)/capi.cu -o $@

emu: $(EMUDIR)/libzkcnn_b200_emu.so
$(EMUDIR)/libzkcnn_b200_emu.so: $(CSRC_DEPS) tests/emu/cuda_emu.cpp tests/emu/cuda_emu.hpp
	mkdir -p $(EMUDIR)
	$(CXX) -O2 -g -std=c++17 -fPIC -DZK_EMU -Wall -Wno-unknown-pragmas -Wno-unused-function -shared -Wl,-soname=libzkcnn_b200_emu.so -x c++ $(CSRC)/capi.cu -x none tests/emu/cuda_emu.cpp -o $@ -lpthread

clean:
	rm -rf $(LIBDIR) $(EMUDIR)
