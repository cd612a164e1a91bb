#!/usr/bin/env bash
# usage: scripts/gpu_sweep.sh "<bench opts 1>" "<bench opts 2>" ...   (development helper, run under gpurun)
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for opt in "$@"; do
  echo "== bench.py $opt" >> gpurun_out/sweep.txt
  python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e $opt > gpurun_out/sweep_line.json 2> gpurun_out/sweep_err.txt
  if [ -s gpurun_out/sweep_line.json ]; then
    python -c "import json; d=json.load(open('gpurun_out/sweep_line.json')); print('%.3fe9 hap-SNPs/s  %.2f ms  frac %.3f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac']), d['kernel'])" >> gpurun_out/sweep.txt
  else
    tail -5 gpurun_out/sweep_err.txt >> gpurun_out/sweep.txt
  fi
done
cat gpurun_out/sweep.txt
