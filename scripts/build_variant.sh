#!/bin/bash
# Build a tuning variant of the library: scripts/build_variant.sh <name> "<extra nvcc flags>"  ->  experiments/libmlsim_<name>.so
# (git-ignored; travels to the GPU box with the gpurun snapshot).  Used with scripts/ab_solver.py / scripts/ab_step.py.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/experiments/build_$name
mkdir -p $out
for f in plan solver_kernels conv_kernels datagen_kernels; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -cudart static "$@" \
       -c $root/ml_accelerated_simulation_b200/csrc/$f.cu -o $out/$f.o &
done
wait
nvcc -shared -cudart static -o $root/experiments/libmlsim_$name.so $out/*.o 2>/dev/null
rm -rf $out
echo $root/experiments/libmlsim_$name.so
