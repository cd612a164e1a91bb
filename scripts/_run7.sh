mkdir -p gpurun_out
python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_e.json 2> gpurun_out/bench_ref_e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sampler_kernel -s 1 -c 1 -o gpurun_out/prof_v34 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_v34.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench_e.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
tail -3 gpurun_out/bench_e.err
