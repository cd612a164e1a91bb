mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
scripts/gpu_sweep.sh "" "--block-loci 8" "--mean-group 10" "--skip-sweeps 6" "--skip-sweeps 5" "--skip-sweeps 3" "--haps 87500 --snps 60000 --steps 2 --warmup 1"
