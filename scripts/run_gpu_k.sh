#!/usr/bin/env bash
# development helper (under gpurun --gpus 2): two-rank bench (both arms) + single-GPU ncu of the family kernel
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"
tail -2 gpurun_out/bench_2gpu.err; cut -c1-700 gpurun_out/bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "2gpu ref rc=$?"
cut -c1-500 gpurun_out/bench_2gpu_ref.json
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:family_kernel -c 1 -f -o gpurun_out/prof_family2 \
  python scripts/fam_bench.py --pairs 2368 --haps 12000 --iters 1 > gpurun_out/ncu_fam2.log 2>&1; echo "ncu rc=$?"
