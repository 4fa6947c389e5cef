#!/bin/bash
# libddf_b200_trace.so = the library with wave.cu built with -DDDF_PIPE_TRACE (pipeline time stamps; tools/pipe_trace.py, pipe2_trace.py)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)/../ddf.jl_b200/csrc"
bash "$HERE/build.sh" > /dev/null
mkdir -p "$HERE/_obj/trace"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DDDF_PIPE_TRACE -c "$HERE/wave.cu" -o "$HERE/_obj/trace/wave.o"
O="$HERE/_obj"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/../libddf_b200_trace.so" "$O"/core.o "$O"/mesh.o "$O"/geometry.o "$O"/ops.o "$O"/trace/wave.o "$O"/comm.o "$O"/solve.o "$O"/lookup.o -ldl
echo "built $HERE/../libddf_b200_trace.so"
