#!/bin/sh
# Build libmdsf.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${MDSF_OUT:-$HERE/../libmdsf.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -std=c++17 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo \
    ${MDSF_NVCC_FLAGS:-} -o "$OUT" "$HERE/mdsf_api.cu" -lcufft -Xlinker -rpath=/usr/local/cuda/lib64
echo "built $OUT"
# host-side trajectory ingest helpers (no CUDA dependency)
"${CXX:-g++}" -O2 -std=c++17 -shared -fPIC -o "$HERE/../libmdsf_io.so" "$HERE/mdsf_io.cpp" -lz -lpthread
echo "built $HERE/../libmdsf_io.so"
