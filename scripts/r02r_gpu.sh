#!/bin/bash
# round-2 GPU job R (1 GPU): ncu --set full of the batched kernels (lean prologue), product build and the depth-2 / 12-RHS-per-CTA variant
mkdir -p gpurun_out
for tag in "" d2n12; do
  export B200_LIB_TAG=$tag
  n=${tag:-product}
  PROF_LATT=48,48,48,48 ncu --set full --clock-control none --import-source on -k regex:dslash_mrhs_kernel -c 2 -f -o gpurun_out/r02r_$n \
      python scripts/prof_mrhs.py > gpurun_out/r02r_ncu_$n.log 2>&1; echo "ncu $n rc=$?"
  ncu -i gpurun_out/r02r_$n.ncu-rep --page raw --csv > gpurun_out/r02r_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02r_$n.ncu-rep --page details > gpurun_out/r02r_${n}_details.txt 2>/dev/null
  ncu -i gpurun_out/r02r_$n.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02r_${n}_source.csv.gz
  rm -f gpurun_out/r02r_$n.ncu-rep
done
ls -la gpurun_out | grep r02r
