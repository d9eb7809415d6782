#!/bin/bash
# Build a variant of the library that differs in csrc/stft_pair.cu's compile flags only:
#   scripts/build_variant.sh NAME -DPAIR_WARPS=8 ...   ->  torchaudio_contrib_b200/lib/variants/libtac_NAME.so
# (other objects are reused from the last `python build_native.py`); load it with TAC_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p torchaudio_contrib_b200/lib/variants
obj=torchaudio_contrib_b200/lib/variants/stft_pair_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 1886 "$@" \
  -c torchaudio_contrib_b200/csrc/stft_pair.cu -o $obj
others=$(ls torchaudio_contrib_b200/lib/obj/*.o | grep -v stft_pair.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o torchaudio_contrib_b200/lib/variants/libtac_$name.so $obj $others -lcuda
echo torchaudio_contrib_b200/lib/variants/libtac_$name.so
