#!/usr/bin/env bash
# development helper (under gpurun): related-branch parity, timing, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_family.py -x -q > gpurun_out/pytest_family.log 2>&1; echo "pytest family rc=$?"
tail -3 gpurun_out/pytest_family.log
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 > gpurun_out/fam_t1.log 2>&1; echo "rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 --K 100 > gpurun_out/fam_t2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:family_kernel -c 1 -f -o gpurun_out/prof_family \
  python scripts/fam_bench.py --pairs 2368 --haps 12000 --iters 1 > gpurun_out/ncu_fam.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/fam_t1.log gpurun_out/fam_t2.log
