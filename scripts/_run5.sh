mkdir -p gpurun_out
python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_d.json 2> gpurun_out/bench_ref_d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
tail -c 3000 gpurun_out/bench_d.json; tail -c 600 gpurun_out/bench_ref_d.json; tail -5 gpurun_out/bench_d.err
