mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
scripts/gpu_sweep.sh "" "--mean-group 10" "--skip-sweeps 5" "--skip-sweeps 3" "--haps 56832 --steps 3"
