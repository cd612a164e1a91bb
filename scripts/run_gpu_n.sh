#!/usr/bin/env bash
# development helper (under gpurun): whole GPU suite + mixed runs after the pool / budget changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python scripts/fam_bench.py --pairs 512 --haps 16000 --snps 60000 --K 100 --iters 1 --mixed 12000 > gpurun_out/fam_mixed_c5.log 2>&1; echo "rc=$?"
timeout 600 python scripts/fam_bench.py --pairs 1184 --haps 24000 --iters 2 --mixed 18944 > gpurun_out/fam_mixed_c2.log 2>&1; echo "rc=$?"
cat gpurun_out/fam_mixed_c5.log gpurun_out/fam_mixed_c2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n.json'))
for k in ('value','e2e','related_branch'): print(k, d.get(k))
"
