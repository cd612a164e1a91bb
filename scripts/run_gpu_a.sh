#!/usr/bin/env bash
# development helper (under gpurun): bench line, related-branch timing, ncu launch list + one full capture
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python scripts/fam_bench.py --check 4 --snps 4000 --pairs 64 --haps 2000 > gpurun_out/fam_check.log 2>&1; echo "fam_check rc=$?"
python scripts/fam_bench.py --unrelated > gpurun_out/fam_bench.log 2>&1; echo "fam_bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sampler_kernel -c 1 -f -o gpurun_out/prof \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
cat gpurun_out/bench.json; cat gpurun_out/fam_check.log gpurun_out/fam_bench.log
