#!/bin/bash
# Time the batched (multi-RHS) operator for tuning variants of the library that were BUILT ON THE CPU BOX
# (B200_BUILD_TAG=<tag> python -m chroma_b200.build; see chroma_b200/build.py) -- the GPU box only measures.
# usage: scripts/tune_mrhs_variants.sh "<tags for fp64>" "<tags for fp32>"   ('default' = the product library)
mkdir -p gpurun_out
out=gpurun_out/tune_mrhs_variants.txt
: > $out
for prec in double single; do
  if [ $prec = double ]; then tags="$1"; else tags="$2"; fi
  for t in $tags; do
    tag=$t; [ $t = default ] && tag=""
    echo "$prec $t: $(B200_LIB_TAG=$tag PROF_PREC=$prec PROF_LATT=${PROF_LATT:-48,48,48,48} PROF_REPS=5 python scripts/prof_mrhs.py 2>&1 | tr '\n' ' ')" >> $out
  done
done
cat $out
