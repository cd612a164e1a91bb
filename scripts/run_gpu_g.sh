#!/usr/bin/env bash
# development helper (under gpurun): occupancy A/B of the family kernel, C5-shaped related pairs, bench line
mkdir -p gpurun_out
timeout 300 python scripts/fam_bench.py --pairs 7104 --haps 30000 > gpurun_out/fam_mb4.log 2>&1; echo "rc=$?"
timeout 600 python scripts/fam_bench.py --pairs 512 --haps 2400 --snps 60000 --K 100 --iters 1 > gpurun_out/fam_c5.log 2>&1; echo "rc=$?"
cp knockoffgwas_b200/lib/libkgw.so /tmp/libkgw_mb4.so
cp gpurun_mb5_libkgw.so knockoffgwas_b200/lib/libkgw.so
timeout 300 python scripts/fam_bench.py --pairs 7104 --haps 30000 > gpurun_out/fam_mb5.log 2>&1; echo "rc=$?"
cp gpurun_mb6_libkgw.so knockoffgwas_b200/lib/libkgw.so
timeout 300 python scripts/fam_bench.py --pairs 7104 --haps 30000 > gpurun_out/fam_mb6.log 2>&1; echo "rc=$?"
cp /tmp/libkgw_mb4.so knockoffgwas_b200/lib/libkgw.so
for f in mb4 mb5 mb6 c5; do echo "== $f"; cat gpurun_out/fam_$f.log; done
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
for k in ('value','ms_per_step','e2e','bgen_ingest','related_branch','bed_writer'): print(k, d.get(k))
"
