#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_family.py -x -q > gpurun_out/pytest_family.log 2>&1; echo "pytest family rc=$?"; tail -12 gpurun_out/pytest_family.log
