#!/usr/bin/env bash
# development helper (under gpurun): related-branch parity + timing with virtual right messages
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_family.py -x -q > gpurun_out/pytest_family.log 2>&1; echo "pytest family rc=$?"; tail -5 gpurun_out/pytest_family.log
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 --check 24 > gpurun_out/fam_t1.log 2>&1; echo "rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 7104 --haps 30000 --K 100 > gpurun_out/fam_t3.log 2>&1; echo "rc=$?"
timeout 600 python scripts/fam_bench.py --pairs 1024 --haps 4400 --snps 60000 --K 100 --iters 1 --check 2 > gpurun_out/fam_c5.log 2>&1; echo "rc=$?"
cat gpurun_out/fam_t1.log gpurun_out/fam_t3.log gpurun_out/fam_c5.log
