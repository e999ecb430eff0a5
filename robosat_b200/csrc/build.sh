#!/bin/sh
# Build librsb200.so in-tree for sm_100a. Usage: sh robosat_b200/csrc/build.sh
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
OBJS=""
for f in rsb_host rsb_conv rsb_conv_row rsb_elementwise rsb_loss rsb_train rsb_wgrad; do
  $NVCC $FLAGS -c $f.cu -o $f.o 2> $f.ptxas.log || { cat $f.ptxas.log; exit 1; }
  OBJS="$OBJS $f.o"
done
# host-side PNG codec (plain C++ over zlib)
${CXX:-g++} -O3 -std=c++17 -fPIC -I/usr/local/cuda/include -c rsb_png.cpp -o rsb_png.o
${CXX:-g++} -O3 -std=c++17 -fPIC -c rsb_inflate.cpp -o rsb_inflate.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../librsb200.so $OBJS rsb_png.o rsb_inflate.o -cudart static -lz -lpthread
# bring-up probes: their own library, not part of the product (include/rsb200_debug.h)
$NVCC $FLAGS -c rsb_debug.cu -o rsb_debug.o 2> rsb_debug.ptxas.log || { cat rsb_debug.ptxas.log; exit 1; }
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../librsb200_debug.so rsb_debug.o rsb_host.o -cudart static
echo "built $(cd .. && pwd)/librsb200.so"
