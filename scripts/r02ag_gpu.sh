#!/bin/bash
# round-2 GPU job AG (1 GPU): the one-device multi-rank cases (split lattices, batched fields included) on the last build
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "one_device" > gpurun_out/r02ag_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02ag_pytest.log
