#!/usr/bin/env bash
# development helper (under gpurun): related-branch phases, scaling with the number of families, ncu capture
mkdir -p gpurun_out
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 --iters 1 > gpurun_out/fam_ph1.log 2>&1; echo "rc=$?"
timeout 400 python scripts/fam_bench.py --pairs 7104 --haps 30000 --iters 1 > gpurun_out/fam_ph2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:family_kernel -c 1 -f -o gpurun_out/prof_family \
  python scripts/fam_bench.py --pairs 2368 --haps 12000 --iters 1 > gpurun_out/ncu_fam.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/fam_ph1.log gpurun_out/fam_ph2.log
