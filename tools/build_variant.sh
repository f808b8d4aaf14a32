#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..." file.cu [file.cu ...]: a copy of the library with the named sources recompiled
# under extra flags -> gpurun_variants/libmvsf_b200_NAME.so (select it with MVSF_LIB_PATH; A-B measurements on the GPU box)
set -e
cd "$(dirname "$0")/.."
NAME=$1; FLAGS=$2; shift 2
C=mvsformerplusplus_b200/csrc
mkdir -p gpurun_variants /tmp/variant_$NAME
python -m mvsformerplusplus_b200.build >/dev/null
OBJS=""
for o in api geometry warp_corr warp_tile vis_cnn costreg_unet costreg_tr fmt linear_tc conv3d_tc; do
  if [[ " $* " == *" $o.cu "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr --extended-lambda $FLAGS -c $C/$o.cu -o /tmp/variant_$NAME/$o.o
    OBJS="$OBJS /tmp/variant_$NAME/$o.o"
  else OBJS="$OBJS $C/$o.o"; fi
done
nvcc -shared -o gpurun_variants/libmvsf_b200_$NAME.so $OBJS -gencode arch=compute_100a,code=sm_100a -lcudart_static -lrt -lpthread -ldl
echo gpurun_variants/libmvsf_b200_$NAME.so
