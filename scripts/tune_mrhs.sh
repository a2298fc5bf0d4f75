#!/bin/bash
# Build + time the multi-RHS kernels under different (rhs per CTA : min CTAs/SM) ON THE GPU BOX.
# usage: scripts/tune_mrhs.sh "4:1 6:1 ..." "4:4 12:1 ..."   (fp64 variants, fp32 variants) -> gpurun_out/tune_mrhs.txt
mkdir -p gpurun_out
: > gpurun_out/tune_mrhs.txt
for v in $1; do
  n=${v%%:*}; mb=${v##*:}
  B200_MRHS_NRB=$n B200_MRHS_MINB=$mb python -m chroma_b200.build --force > /dev/null 2>gpurun_out/tune_build.err || { echo "d $v build failed" >> gpurun_out/tune_mrhs.txt; continue; }
  echo "fp64 NRB=$n minb=$mb: $(PROF_PREC=double PROF_LATT=${PROF_LATT:-32,32,32,64} python scripts/prof_mrhs.py 2>&1 | tr '\n' ' ')" >> gpurun_out/tune_mrhs.txt
done
for v in $2; do
  n=${v%%:*}; mb=${v##*:}
  B200_MRHS_NRB_F=$n B200_MRHS_MINB_F=$mb python -m chroma_b200.build --force > /dev/null 2>gpurun_out/tune_build.err || { echo "f $v build failed" >> gpurun_out/tune_mrhs.txt; continue; }
  echo "fp32 NRB=$n minb=$mb: $(PROF_PREC=single PROF_LATT=${PROF_LATT:-32,32,32,64} python scripts/prof_mrhs.py 2>&1 | tr '\n' ' ')" >> gpurun_out/tune_mrhs.txt
done
cat gpurun_out/tune_mrhs.txt
