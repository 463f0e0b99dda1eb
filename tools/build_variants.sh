#!/bin/bash
# usage: tools/build_variants.sh name1="-DX=1 -DY=2" name2="..."   -> _build/variants/<name>.so (parallel builds)
mkdir -p _build/variants
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared"
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  ( $NV $flags -Xptxas -v -o _build/variants/$name.so birdnest/audio_b200/csrc/kernels.cu birdnest/audio_b200/csrc/kernels_decode_wide.cu birdnest/audio_b200/csrc/kernels_decode_narrow.cu birdnest/audio_b200/csrc/engine.cu birdnest/audio_b200/csrc/encoder.cu > _build/variants/$name.log 2>&1 || echo "BUILD FAILED $name" ) &
done
wait
ls -la _build/variants/*.so
