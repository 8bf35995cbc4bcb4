# Native build of av1dec_b200 (run by __graft_entry__.build()).
#
#   make            product: engine (nvcc, sm_100a) + host decoder library + drop-in CLI
#   make emu        TEST-ONLY host emulation of the kernels (tests/emu/), for CPU debugging
#   make oracle     the unmodified reference as the checker (oracle/_ref/), needs /root/reference
#
# The product links the reference's FRONT END (parser / entropy decoder / block syntax) from the
# sources where they lie under $(REF): the north-star keeps that front end on the host.  Its
# objects are cached in av1dec_b200/_frontend/ (git-ignored, shipped to the GPU box with the
# snapshot).  When $(REF) is absent (GPU box) the prebuilt objects / libraries are used as is.

REF      ?= /root/reference
NVCC     ?= nvcc
CXX      ?= g++
CC       ?= gcc
PKG      := av1dec_b200
LIB      := $(PKG)/lib
FE       := $(PKG)/_frontend
EMU      := tests/emu

NVFLAGS  := -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC
# -fno-aggressive-loop-optimizations: see SURVEY.md section 0 fact 3 (the reference indexes 7-entry arrays 1..7)
FEFLAGS  := -std=c++14 -O3 -fno-aggressive-loop-optimizations -fPIC -w -ffunction-sections -fdata-sections
# the reference's decode() tree walk (the CPU pixel path's entry points, virtual so the vtables keep
# them alive): weakened in the front-end objects so that host/pixel_path_guard.cpp's abort() traps
# replace them; everything below them is then unreferenced and dropped by --gc-sections
PIXEL_ROOTS := _ZN7YamiAv14Tile6decodeERSt10shared_ptrIN4Yami8YuvFrameEERKSt6vectorIS4_SaIS4_EE \
    _ZN7YamiAv110SuperBlock6decodeERSt10shared_ptrIN4Yami8YuvFrameEERKSt6vectorIS4_SaIS4_EE \
    _ZN7YamiAv19Partition6decodeERSt10shared_ptrIN4Yami8YuvFrameEERKSt6vectorIS4_SaIS4_EE \
    _ZN7YamiAv15Block6decodeERSt10shared_ptrIN4Yami8YuvFrameEERKSt6vectorIS4_SaIS4_EE
HOSTFLAGS:= -std=c++17 -O2 -fno-aggressive-loop-optimizations -fPIC -ffunction-sections -Wall -Wno-unused-function -Wno-unknown-pragmas
REFINC   := -I$(REF)/aom -I$(REF)/common -I$(REF)/interface -I$(REF)/decoder

CSRC     := $(PKG)/csrc
KERNELS  := recon postfilter deblock cdef lr engine
KHDRS    := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/av1b200.h include/av1b200_format.h

# reference front-end translation units (parse side). NOT included: Av1Decoder, IntraPredict,
# LoopFilter, Cdef, LoopRestoration, VideoFrame -- the CPU pixel path.
FE_UNITS := BitReader Block Cdfs EntropyDecoder InterPredict Parser Partition SuperBlock SymbolDecoder Tile TransformBlock
FE_OBJS  := $(patsubst %,$(FE)/%.o,$(FE_UNITS)) $(FE)/entropymode.o

HOST_SRCS:= emitter decoder capi yami_adapter pixel_path_guard
HOST_HDRS:= $(wildcard $(PKG)/host/*.h) include/av1b200_decoder.h $(KHDRS)

HAVE_REF := $(wildcard $(REF)/decoder/Parser.cpp)

all: product
product: $(LIB)/libav1b200.so $(LIB)/libav1b200dec.so $(PKG)/bin/av1dec tests/native/ivd_drive

# ---------------------------------------------------------------- engine (CUDA)
$(LIB)/obj/%.o: $(CSRC)/%.cu $(KHDRS)
	@mkdir -p $(LIB)/obj
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB)/libav1b200.so: $(patsubst %,$(LIB)/obj/%.o,$(KERNELS))
	$(NVCC) -shared -cudart static -o $@ $^

ifneq ($(HAVE_REF),)
# ---------------------------------------------------------------- reference front end
$(FE)/%.o: $(REF)/decoder/%.cpp
	@mkdir -p $(FE)
	$(CXX) $(FEFLAGS) $(REFINC) -c $< -o $@
	objcopy $(patsubst %,--weaken-symbol=%,$(PIXEL_ROOTS)) $@

$(FE)/entropymode.o: $(REF)/aom/entropymode.c
	@mkdir -p $(FE)
	$(CC) -O3 -fPIC -w $(REFINC) -c $< -o $@

# ---------------------------------------------------------------- host decoder library
$(LIB)/obj/host_%.o: $(PKG)/host/%.cpp $(HOST_HDRS)
	@mkdir -p $(LIB)/obj
	$(CXX) $(HOSTFLAGS) -I$(PKG)/host $(REFINC) -c $< -o $@

$(LIB)/libav1b200dec.so: $(patsubst %,$(LIB)/obj/host_%.o,$(HOST_SRCS)) $(FE_OBJS) $(LIB)/libav1b200.so $(PKG)/host/exports.map
	$(CXX) -shared -o $@ $(patsubst %,$(LIB)/obj/host_%.o,$(HOST_SRCS)) $(FE_OBJS) -L$(LIB) -lav1b200 -Wl,-rpath,'$$ORIGIN' -Wl,--no-undefined -Wl,--gc-sections -Wl,--version-script=$(PKG)/host/exports.map -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64

# ---------------------------------------------------------------- drop-in CLI: the reference's own tests/*.cpp, unchanged
$(PKG)/bin/av1dec: $(LIB)/libav1b200dec.so $(REF)/tests/Av1Dec.cpp
	@mkdir -p $(PKG)/bin
	$(CXX) $(FEFLAGS) -I$(PKG)/host $(REFINC) -I$(REF)/tests -o $@ $(REF)/tests/Av1Dec.cpp $(REF)/tests/DecodeInput.cpp \
	    $(REF)/tests/DecodeOutput.cpp -x c $(REF)/tests/md5.c -x none -L$(LIB) -lav1b200dec -lav1b200 -Wl,-rpath,'$$ORIGIN/../lib'

# ---------------------------------------------------------------- test program: a libyami-style client of IVideoDecoder
tests/native/ivd_drive: tests/native/ivd_drive.cpp
	$(CXX) -std=c++14 -O1 -w $(REFINC) -o $@ $< -ldl

# ---------------------------------------------------------------- test-only emulation
emu: $(EMU)/libav1b200_emu.so $(EMU)/libav1b200dec_emu.so $(EMU)/av1dec_emu

$(EMU)/obj/%.o: $(CSRC)/%.cu $(KHDRS)
	@mkdir -p $(EMU)/obj
	$(CXX) -x c++ -DAV1B_EMU $(HOSTFLAGS) -c $< -o $@

$(EMU)/libav1b200_emu.so: $(patsubst %,$(EMU)/obj/%.o,$(KERNELS))
	$(CXX) -shared -o $@ $^

$(EMU)/libav1b200dec_emu.so: $(patsubst %,$(LIB)/obj/host_%.o,$(HOST_SRCS)) $(FE_OBJS) $(EMU)/libav1b200_emu.so
	$(CXX) -shared -o $@ $(patsubst %,$(LIB)/obj/host_%.o,$(HOST_SRCS)) $(FE_OBJS) -L$(EMU) -lav1b200_emu -Wl,-rpath,'$$ORIGIN' -Wl,--no-undefined -Wl,--gc-sections -Wl,--version-script=$(PKG)/host/exports.map

$(EMU)/av1dec_emu: $(EMU)/libav1b200dec_emu.so $(REF)/tests/Av1Dec.cpp
	$(CXX) $(FEFLAGS) -I$(PKG)/host $(REFINC) -I$(REF)/tests -o $@ $(REF)/tests/Av1Dec.cpp $(REF)/tests/DecodeInput.cpp \
	    $(REF)/tests/DecodeOutput.cpp -x c $(REF)/tests/md5.c -x none -L$(EMU) -lav1b200dec_emu -lav1b200_emu -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle REF=$(REF)
else
$(LIB)/libav1b200dec.so $(PKG)/bin/av1dec tests/native/ivd_drive:
	@test -f $@ || (echo "error: $@ missing and $(REF) not available to build it" && false)
emu oracle:
	@echo "$(REF) not available: using prebuilt artefacts"
endif

clean:
	rm -rf $(LIB) $(FE) $(PKG)/bin $(EMU)/obj $(EMU)/*.so $(EMU)/av1dec_emu tests/native/ivd_drive

.PHONY: all product emu oracle clean
