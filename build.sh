#!/bin/bash
# Builds the in-tree CUDA library for sm_100a (one object per source, compiled in parallel).
# Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
PKG=video_diffusion_nnx_b200
OBJ=$PKG/csrc/.obj
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude $*"
pids=()
for f in $PKG/csrc/*.cu; do
  o=$OBJ/$(basename ${f%.cu}).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ $PKG/csrc/vdn_common.cuh -nt $o ] || [ $PKG/csrc/vdn_host.h -nt $o ] || [ include/vdn.h -nt $o ]; then
    nvcc $FLAGS -c -o $o $f &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $PKG/libvdn.so $OBJ/*.o -ldl
echo "built $PKG/libvdn.so"
# XLA-FFI shim (ffi/vdn_ffi.cc): real handlers where jaxlib's headers exist, a stub (vdn_ffi_available() == 0) elsewhere
XLA_INC=$(python -c "import jax.ffi; print(jax.ffi.include_dir())" 2>/dev/null || true)
if [ -n "$XLA_INC" ]; then
  g++ -O2 -std=c++17 -shared -fPIC ffi/vdn_ffi.cc -Iinclude -I"$XLA_INC" -I/usr/local/cuda/include -L$PKG -lvdn -Wl,-rpath,'$ORIGIN' -o $PKG/libvdn_ffi.so
else
  g++ -O2 -std=c++17 -shared -fPIC ffi/vdn_ffi.cc -Iinclude -o $PKG/libvdn_ffi.so
fi
echo "built $PKG/libvdn_ffi.so (xla headers: ${XLA_INC:-none})"
