#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "-DFLAG ..." : libpcvae variant with one translation unit rebuilt with extra flags
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
C=vae_posterior_consistency_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -I include -I $C $3 -c $C/$2 -o /tmp/variant_$1.o
OBJS=$(ls $C/*.o | grep -v "${2%.cu}.o")
nvcc -shared -o tools/variants/lib$1.so $OBJS /tmp/variant_$1.o -gencode arch=compute_100a,code=sm_100a --cudart shared -Xlinker -rpath,/usr/local/cuda/lib64
echo built tools/variants/lib$1.so
