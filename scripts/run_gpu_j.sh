#!/usr/bin/env bash
# development helper (under gpurun): BGEN + related-branch parity, timing, bench legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bgen.py tests/test_gpu_family.py -x -q > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_j.log
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 --check 24 > gpurun_out/fam_t1.log 2>&1; echo "rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 7104 --haps 30000 > gpurun_out/fam_t2.log 2>&1; echo "rc=$?"
cat gpurun_out/fam_t1.log gpurun_out/fam_t2.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_j.json'))
for k in ('value','bgen_ingest','related_branch'): print(k, d.get(k))
"
