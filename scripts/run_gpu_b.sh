#!/usr/bin/env bash
# development helper (under gpurun): related-branch parity + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_family.py -x -q > gpurun_out/pytest_family.log 2>&1; echo "pytest family rc=$?"
tail -5 gpurun_out/pytest_family.log
timeout 300 python scripts/fam_bench.py --check 16 --snps 4000 --pairs 64 --haps 2000 > gpurun_out/fam_check.log 2>&1; echo "fam_check rc=$?"
timeout 300 python scripts/fam_bench.py --unrelated > gpurun_out/fam_bench.log 2>&1; echo "fam_bench rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 > gpurun_out/fam_bench2.log 2>&1; echo "fam_bench2 rc=$?"
cat gpurun_out/fam_check.log gpurun_out/fam_bench.log gpurun_out/fam_bench2.log
