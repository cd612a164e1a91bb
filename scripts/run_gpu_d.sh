#!/usr/bin/env bash
# development helper (under gpurun): related-branch parity + phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_family.py -x -q > gpurun_out/pytest_family.log 2>&1; echo "pytest family rc=$?"
tail -5 gpurun_out/pytest_family.log
timeout 300 python scripts/fam_bench.py --check 16 --snps 4000 --pairs 64 --haps 2000 2>&1 | grep -v "kgw phases" > gpurun_out/fam_check.log; echo "fam_check rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 --iters 1 2>&1 | grep -v "kgw phases" > gpurun_out/fam_ph1.log; echo "rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 --K 100 --iters 1 2>&1 | grep -v "kgw phases" > gpurun_out/fam_ph3.log; echo "rc=$?"
cat gpurun_out/fam_check.log gpurun_out/fam_ph1.log gpurun_out/fam_ph3.log
