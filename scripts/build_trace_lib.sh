# timing build of the library: clock / globaltimer stamps compiled into stft2048_kernel (-DTAC_K1_TRACE_BUILD)
set -e
cd "$(dirname "$0")/.."
OUT=torchaudio_contrib_b200/lib/trace
mkdir -p $OUT
for f in torchaudio_contrib_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 1886 \
       -DTAC_K1_TRACE_BUILD -c $f -o $OUT/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libtac_b200_trace.so $OUT/*.o -lcuda
ls -la $OUT/libtac_b200_trace.so
