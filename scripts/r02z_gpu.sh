#!/bin/bash
# round-2 GPU job Z (1 GPU): ncu --set full of the batched kernels INSIDE the running CG loop (EPI_AINV, EPI_M_NORM, EPI_M_CG)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:dslash_mrhs_kernel --launch-skip 8 -c 8 -f -o gpurun_out/r02z \
    python bench.py --no-cpu --no-fp32 --no-solve --steps 4 --warmup 2 --lattice 48 48 48 48 > gpurun_out/r02z_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02z.ncu-rep --page raw --csv > gpurun_out/r02z_raw.csv 2>/dev/null
ncu -i gpurun_out/r02z.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02z_source.csv.gz
rm -f gpurun_out/r02z.ncu-rep
ls -la gpurun_out | grep r02z
