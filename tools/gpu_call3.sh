#!/bin/bash
# validation of the 2x2-block up2_add_split and the vectorised CBAM kernels; gru diagnostics (gate epilogue vs 192-column tiles)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c3_tests.log 2>&1; tail -5 gpurun_out/c3_tests.log
timeout 100 python tools/bench_conv.py --cases gru,gru_n192,gru_gates > gpurun_out/c3_conv.txt 2>&1; cat gpurun_out/c3_conv.txt
timeout 100 python bench.py --quick --no-cpu-baseline --no-torch-baseline > gpurun_out/c3_bench_quick.json 2> gpurun_out/c3_bench_quick.err; cut -c1-330 gpurun_out/c3_bench_quick.json; tail -2 gpurun_out/c3_bench_quick.err
timeout 90 python tools/profile_layers.py > gpurun_out/c3_layers.txt 2>&1; grep -E "^ew:|^# all" gpurun_out/c3_layers.txt
