# Build recipe (also driven by __graft_entry__.build()).
#   make lib      -> birdnest/audio_b200/libbnflac.so   (the product: sm_100a CUDA + C ABI)
#   make oracle   -> oracle/_build/liboracle.so, flac_oracle   (test infrastructure)
#   make corpus   -> corpus/_build/libbncorpus.so, bncorpus     (synthetic stream generator)
#   make ref      -> oracle/_ref/refflac + LibFlac.dll copy      (only where /root/reference exists)
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall -Xptxas -v
CSRC := birdnest/audio_b200/csrc
LIB := birdnest/audio_b200/libbnflac.so

all: lib oracle corpus host shim

lib: $(LIB)
# The device code is three translation units (front kernels; decode variants with 64-bit / 32-bit accumulation) that share
# kernels_common.cuh and compile in parallel under `make -j`; engine.cu (host runtime, seconds) only calls the launch_* host
# wrappers, so no relocatable device code is needed.  Objects and the ptxas -v logs live in csrc/_obj (git-ignored).
OBJ := $(CSRC)/_obj
DEVHDR := $(CSRC)/bnflac_dev.h $(CSRC)/kernels_common.cuh include/bnflac.h
$(OBJ)/kernels.o: $(CSRC)/kernels.cu $(DEVHDR)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $(CSRC)/kernels.cu 2> $(OBJ)/kernels.log || (cat $(OBJ)/kernels.log; exit 1)
	@grep -E "error|warning" $(OBJ)/kernels.log | grep -v "ptxas info" || true
$(OBJ)/kernels_decode_%.o: $(CSRC)/kernels_decode_%.cu $(CSRC)/kernels_decode.cuh $(DEVHDR)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(OBJ)/kernels_decode_$*.log || (cat $(OBJ)/kernels_decode_$*.log; exit 1)
	@grep -E "error|warning" $(OBJ)/kernels_decode_$*.log | grep -v "ptxas info" || true
$(OBJ)/engine.o: $(CSRC)/engine.cu $(CSRC)/bnflac_dev.h include/bnflac.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $(CSRC)/engine.cu 2> $(OBJ)/engine.log || (cat $(OBJ)/engine.log; exit 1)
	@grep -E "error|warning" $(OBJ)/engine.log | grep -v "ptxas info" || true
$(OBJ)/encoder.o: $(CSRC)/encoder.cu $(CSRC)/encoder_kernels.cuh $(CSRC)/md5.hpp include/bnflac.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $(CSRC)/encoder.cu 2> $(OBJ)/encoder.log || (cat $(OBJ)/encoder.log; exit 1)
	@grep -E "error|warning" $(OBJ)/encoder.log | grep -v "ptxas info" || true
LIBOBJ := $(OBJ)/kernels.o $(OBJ)/kernels_decode_wide.o $(OBJ)/kernels_decode_narrow.o $(OBJ)/engine.o $(OBJ)/encoder.o
$(LIB): $(LIBOBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(LIBOBJ)

host: birdnest/audio_b200/flacdecoder_demo
birdnest/audio_b200/flacdecoder_demo: $(CSRC)/flac_decoder.hpp $(CSRC)/flac_decoder_demo.cpp $(LIB)
	g++ -O2 -std=c++17 -Wall -Iinclude -o $@ $(CSRC)/flac_decoder_demo.cpp -Lbirdnest/audio_b200 -lbnflac -Wl,-rpath,'$$ORIGIN'

# the libFLAC 1.2.1 stream-decoder symbols the unmodified C# binds, replayed from a bnflac handle
SHIM := birdnest/audio_b200/libLibFlac.so
shim: $(SHIM)
$(SHIM): $(CSRC)/libflac_shim.cpp $(CSRC)/md5.hpp include/bnflac_legacy.h include/bnflac.h $(LIB)
	g++ -O2 -std=c++17 -Wall -fPIC -shared -Iinclude -I$(CSRC) -o $@ $(CSRC)/libflac_shim.cpp -Lbirdnest/audio_b200 -lbnflac -Wl,-rpath,'$$ORIGIN'

oracle: oracle/_build/liboracle.so oracle/_build/flac_oracle
oracle/_build/liboracle.so: oracle/flac_oracle.c oracle/flac_oracle.h
	@mkdir -p oracle/_build
	gcc -O3 -fPIC -shared -Wall -o $@ oracle/flac_oracle.c
oracle/_build/flac_oracle: oracle/flac_oracle.c oracle/flac_oracle.h
	@mkdir -p oracle/_build
	gcc -O3 -Wall -DFO_MAIN -o $@ oracle/flac_oracle.c

corpus: corpus/_build/libbncorpus.so corpus/_build/bncorpus
corpus/_build/libbncorpus.so: corpus/bncorpus.c corpus/bncorpus.h
	@mkdir -p corpus/_build
	gcc -O3 -ffp-contract=off -fPIC -shared -w -o $@ corpus/bncorpus.c -lm -lpthread
corpus/_build/bncorpus: corpus/bncorpus.c corpus/bncorpus.h
	@mkdir -p corpus/_build
	gcc -O3 -ffp-contract=off -w -DBNC_MAIN -o $@ corpus/bncorpus.c -lm -lpthread

ref:
	$(MAKE) -C oracle/refdll

clean:
	rm -rf $(LIB) $(SHIM) oracle/_build corpus/_build birdnest/audio_b200/flacdecoder_demo

.PHONY: all lib oracle corpus ref clean host shim

# the encoder kernels compiled for the CPU (one pthread per CUDA thread): logic check without a GPU, tests/test_encode_emu_cpu.py
emu: tools/emu/_build/libencemu.so
tools/emu/_build/libencemu.so: tools/emu/enc_emu.cpp tools/emu/cuda_emu.h birdnest/audio_b200/csrc/encoder_kernels.cuh
	@mkdir -p tools/emu/_build
	g++ -O1 -std=c++17 -fPIC -shared -pthread -o $@ tools/emu/enc_emu.cpp
