#!/bin/bash
# Builds the in-tree CUDA library for sm_100a. Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
PKG=video_diffusion_nnx_b200
SRC=$(ls $PKG/csrc/*.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -Iinclude "$@" \
  -o $PKG/libvdn.so $SRC
echo "built $PKG/libvdn.so"
