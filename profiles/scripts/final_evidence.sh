#!/bin/bash
# Final evidence pass of a round (run on the GPU box from the repo root): full bench line, ncu launch list of the bench
# command, DRAM traffic of every convolution launch of one step, CUPTI timeline, microbenchmarks.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
L=$(python -c "import json;d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1]);print(int(d['roofline']['launches_per_step']))")
echo "conv launches per step: $L"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"igemm|thin_" -s $((2 * L)) -c $L --csv --log-file gpurun_out/r2_step_traffic.csv python profiles/scripts/prof_step.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 600 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-gpu-reference --no-graph > gpurun_out/r2_launches_bench.log 2>&1
python profiles/scripts/gaps.py > gpurun_out/r2_step_timeline.txt 2>&1
(python profiles/scripts/bench_igemm.py; python profiles/scripts/bench_thin.py; python profiles/scripts/bench_ew.py) > gpurun_out/r2_microbench.txt 2>&1
wc -l gpurun_out/r2_step_traffic.csv gpurun_out/r2_launches_final.csv
