#!/bin/sh
# Build libmdsf.so in-tree for sm_100a (cross-compiles without a GPU).  The three translation units compile in
# parallel: the C ABI + prep/bin kernels, the splat instantiations, the FFT pass instantiations.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${MDSF_OUT:-$HERE/../libmdsf.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OBJ="${MDSF_OBJDIR:-$HERE/build}"
mkdir -p "$OBJ"
FLAGS="-O3 -std=c++17 -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -diag-suppress 128,177 ${MDSF_NVCC_FLAGS:-}"
pids=""
for tu in mdsf_api mdsf_splat_inst mdsf_pass_inst; do
    "$NVCC" $FLAGS -c "$HERE/$tu.cu" -o "$OBJ/$tu.o" &
    pids="$pids $!"
done
for p in $pids; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ/mdsf_api.o" "$OBJ/mdsf_splat_inst.o" "$OBJ/mdsf_pass_inst.o" \
    -lcufft -Xlinker -rpath=/usr/local/cuda/lib64
echo "built $OUT"
# host-side trajectory ingest helpers (no CUDA dependency)
"${CXX:-g++}" -O2 -std=c++17 -shared -fPIC -o "$HERE/../libmdsf_io.so" "$HERE/mdsf_io.cpp" -lz -lpthread
echo "built $HERE/../libmdsf_io.so"
