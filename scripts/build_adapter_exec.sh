#!/bin/bash
# builds tests/adapter_exec.cc + chroma_adapter/*.cc + tests/mock_chroma into $1 (default gpurun_out/adapter_exec)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-$ROOT/gpurun_out/adapter_exec}
T=$(mktemp -d)
mkdir -p $T/actions/ferm/invert/b200_solvers
cp $ROOT/chroma_adapter/*.h $ROOT/chroma_adapter/*.cc $T/actions/ferm/invert/b200_solvers/
g++ -std=c++11 -O1 -g -o $OUT -I $T -I $ROOT/tests/mock_chroma -I $ROOT/include $ROOT/tests/adapter_exec.cc $ROOT/tests/mock_chroma/mock_qdp.cc \
    $T/actions/ferm/invert/b200_solvers/*.cc -L $ROOT/chroma_b200 -lb200clover -L $ROOT/oracle -loracle \
    -Wl,-rpath,$ROOT/chroma_b200 -Wl,-rpath,$ROOT/oracle -Wl,-rpath-link,/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -fopenmp
rm -rf $T
echo $OUT
