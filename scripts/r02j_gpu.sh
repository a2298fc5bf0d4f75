#!/bin/bash
# round-2 GPU job J (1 GPU): final validation -- the driver's own sequence: pytest -m gpu, smoke, bench, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -4 gpurun_out/r02j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02j_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02j_bench_reference.json 2> gpurun_out/r02j_bench_reference.err; echo "ref rc=$?"
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02j_bench_1gpu.json 2> gpurun_out/r02j_bench_1gpu.err ) 2>&1 | grep real; echo "bench rc=$?"
for th in 8 12 16; do B200_COPY_THREADS=$th python scripts/time_setup.py > gpurun_out/r02j_setup_threads$th.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r02j_setup_threads$th.json'));print($th, d['field_upload_pageable_ms'], d['field_download_pageable_ms'], d['load_gauge_ms'])"; done
nproc
