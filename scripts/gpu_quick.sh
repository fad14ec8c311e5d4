#!/bin/bash
# quick GPU check of selected test files: bash scripts/gpu_quick.sh tests/a.py tests/b.py ...
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest "$@" -m gpu -q --timeout 240 2>&1 | tail -150 > gpurun_out/pytest_quick.log
tail -40 gpurun_out/pytest_quick.log
