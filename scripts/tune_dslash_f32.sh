#!/bin/bash
# Build + bench the fp32 dslash kernel under different (block, min-blocks) launch bounds ON THE GPU BOX.
# usage: scripts/tune_dslash_f32.sh "128:1 128:2 ..."   -> gpurun_out/tune_f32.txt
mkdir -p gpurun_out
: > gpurun_out/tune_f32.txt
for v in $1; do
  blk=${v%%:*}; mb=${v##*:}
  B200_DSLASH_BLOCK_F=$blk B200_DSLASH_MINBLOCKS_F=$mb python -m chroma_b200.build --force > /dev/null 2>gpurun_out/tune_build.err || { echo "$v build failed" >> gpurun_out/tune_f32.txt; continue; }
  python bench.py --prec single --steps 10 --warmup 3 --no-cpu > gpurun_out/tune_f32_${blk}_${mb}.json 2>/dev/null
  python - "$v" gpurun_out/tune_f32_${blk}_${mb}.json >> gpurun_out/tune_f32.txt <<'PY'
import json,sys
d=json.load(open(sys.argv[2]))
r=d["roofline"]
print(sys.argv[1], "cg_ms=%.3f cg_gflops=%.0f  M_ms=%.3f M_frac=%.3f  ainv_ms=%.3f  epiM_ms=%.3f" % (d["ms_per_step"], d["value"], d["clover_dslash"]["ms_per_apply"], d["clover_dslash"]["frac_of_peak"], list(r["other_kernels"].values())[0]["ms_per_launch"], r["ms_per_launch"]))
PY
done
cat gpurun_out/tune_f32.txt
