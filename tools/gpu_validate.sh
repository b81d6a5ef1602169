#!/bin/bash
# One GPU-box pass over the evidence of a build (run under gpurun): full GPU suite, smoke, default bench, per-layer table,
# launch list.  Outputs land in gpurun_out/; copy what is to be kept into profiles/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; tail -3 gpurun_out/tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/bench.json
timeout 90 python tools/profile_layers.py > gpurun_out/layers.txt 2>&1; tail -1 gpurun_out/layers.txt
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 700 --csv --log-file gpurun_out/launches.csv \
  python bench.py --quick --no-cpu-baseline --no-torch-baseline --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/summarise_launches.py gpurun_out/launches.csv "ncu launch list of bench.py --quick" > gpurun_out/launches_summary.md
