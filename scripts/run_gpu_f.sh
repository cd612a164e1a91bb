#!/usr/bin/env bash
# development helper (under gpurun): BGEN ingest tests, related-branch parity, occupancy A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bgen.py tests/test_gpu_family.py -x -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_f.log
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 > gpurun_out/fam_mb4.log 2>&1; echo "rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 7104 --haps 30000 --K 100 > gpurun_out/fam_mb4b.log 2>&1; echo "rc=$?"
cp knockoffgwas_b200/lib/libkgw.so /tmp/libkgw_mb4.so; cp gpurun_mb6_libkgw.so knockoffgwas_b200/lib/libkgw.so
timeout 300 python scripts/fam_bench.py --pairs 2368 --haps 12000 --check 8 > gpurun_out/fam_mb6.log 2>&1; echo "rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 3552 --haps 16000 > gpurun_out/fam_mb6c.log 2>&1; echo "rc=$?"
timeout 300 python scripts/fam_bench.py --pairs 7104 --haps 30000 --K 100 > gpurun_out/fam_mb6b.log 2>&1; echo "rc=$?"
cp /tmp/libkgw_mb4.so knockoffgwas_b200/lib/libkgw.so
for f in mb4 mb4b mb6 mb6c mb6b; do echo "== $f"; cat gpurun_out/fam_$f.log; done
