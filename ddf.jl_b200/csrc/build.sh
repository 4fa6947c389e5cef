#!/bin/bash
# Build libddf_b200.so for sm_100a (nvcc cross-compiles without a GPU).
# Usage: build.sh [-j N]   -- output: ddf.jl_b200/libddf_b200.so
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libddf_b200.so"
OBJ="$HERE/_obj"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# DDF_EXTRA_FLAGS: extra nvcc flags for experiments, e.g. DDF_EXTRA_FLAGS=-DDDF_PIPE_TRACE (delete _obj/wave.o first:
# objects are only rebuilt when a source is newer).  Unset for the shipped library.
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v ${DDF_EXTRA_FLAGS:-}"
pids=()
for f in core mesh geometry ops wave comm solve lookup; do
  src="$HERE/$f.cu"; obj="$OBJ/$f.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/common.cuh" -nt "$obj" ] || [ "$HERE/mesh.cuh" -nt "$obj" ] || [ "$HERE/sell.cuh" -nt "$obj" ] || [ "$HERE/stage.cuh" -nt "$obj" ] || [ "$HERE/wavefront2.cuh" -nt "$obj" ] || [ "$HERE/../../include/ddf_b200.h" -nt "$obj" ]; then
    ( $NVCC $FLAGS -c "$src" -o "$obj" > "$OBJ/$f.log" 2>&1 || { cat "$OBJ/$f.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/core.o "$OBJ"/mesh.o "$OBJ"/geometry.o "$OBJ"/ops.o "$OBJ"/wave.o "$OBJ"/comm.o "$OBJ"/solve.o "$OBJ"/lookup.o -ldl
echo "built $OUT"
