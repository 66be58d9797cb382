#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -5 gpurun_out/pytest_gpu2.log
timeout 1500 python tools/tune2.py all > gpurun_out/tune2.txt 2>&1
grep "== best" gpurun_out/tune2.txt
