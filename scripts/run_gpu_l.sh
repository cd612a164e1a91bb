#!/usr/bin/env bash
# development helper (under gpurun): both branches in one call, sequential vs overlapped
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_family.py -x -q -k "whole_generate or synthetic" > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_l.log
timeout 900 python scripts/fam_bench.py --pairs 512 --haps 16000 --snps 60000 --K 100 --iters 1 --mixed 12000 > gpurun_out/fam_mixed_c5.log 2>&1; echo "rc=$?"
timeout 600 python scripts/fam_bench.py --pairs 1184 --haps 24000 --iters 2 --mixed 18944 > gpurun_out/fam_mixed_c2.log 2>&1; echo "rc=$?"
cat gpurun_out/fam_mixed_c5.log gpurun_out/fam_mixed_c2.log
