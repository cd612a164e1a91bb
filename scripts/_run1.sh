mkdir -p gpurun_out
./scripts/ubench/lat > gpurun_out/ub_lat.txt 2>&1
./scripts/ubench/thr > gpurun_out/ub_thr.txt 2>&1
scripts/gpu_sweep.sh "--haps 18944" "--haps 20000" "--haps 14208 --lanes 4" "--haps 37888" "--haps 18944 --mean-group 10" "--haps 18944 --skip-sweeps 6" "--haps 18944 --skip-sweeps 5" "--haps 18944 --skip-sweeps 3" "--haps 87500 --snps 60000 --steps 2 --warmup 1"
