#!/bin/bash
# final evidence run of the round: full GPU suite, default bench (all legs), launch list, smoke, per-layer table
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "default bench wall: $(( $(date +%s) - T0 )) s"; cut -c1-260 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
timeout 90 python tools/profile_layers.py > gpurun_out/final_layers.txt 2>&1; tail -1 gpurun_out/final_layers.txt
for i in 1 2; do timeout 100 python bench.py --quick --no-cpu-baseline --no-torch-baseline --steps 20 > gpurun_out/final_quick$i.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/final_quick$i.json'));print('quick$i', round(d['value'],2), round(d['ms_per_step'],2), d['host_enqueue_ms_per_step'], d['clocks'])"; done
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 700 --csv --log-file gpurun_out/final_launches.csv python bench.py --quick --no-cpu-baseline --no-torch-baseline --steps 2 --warmup 3 > gpurun_out/final_ncu_bench.log 2>&1; wc -l gpurun_out/final_launches.csv
