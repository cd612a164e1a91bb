#!/usr/bin/env bash
# development helper (under gpurun): both branches in one call, sequential vs overlapped, after the memory split
mkdir -p gpurun_out
timeout 900 python scripts/fam_bench.py --pairs 512 --haps 16000 --snps 60000 --K 100 --iters 1 --mixed 12000 > gpurun_out/fam_mixed_c5.log 2>&1; echo "rc=$?"
timeout 900 python scripts/fam_bench.py --pairs 1024 --haps 16000 --snps 60000 --K 100 --iters 1 --mixed 12000 > gpurun_out/fam_mixed_c5b.log 2>&1; echo "rc=$?"
cat gpurun_out/fam_mixed_c5.log gpurun_out/fam_mixed_c5b.log
