#!/bin/bash
# validation: pixel-per-thread image im2col, register-resident up4_softmax, fast gate non-linearities
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c4_tests.log 2>&1; tail -5 gpurun_out/c4_tests.log
timeout 100 python tools/bench_conv.py --cases gru_n192,gru_gates > gpurun_out/c4_conv.txt 2>&1; cat gpurun_out/c4_conv.txt
timeout 100 python bench.py --quick --no-cpu-baseline --no-torch-baseline > gpurun_out/c4_bench_quick.json 2> gpurun_out/c4_bench_quick.err; cut -c1-330 gpurun_out/c4_bench_quick.json; tail -2 gpurun_out/c4_bench_quick.err
timeout 90 python tools/profile_layers.py > gpurun_out/c4_layers.txt 2>&1; head -8 gpurun_out/c4_layers.txt; grep -E "^ew:|^# all" gpurun_out/c4_layers.txt
timeout 120 python tools/bench_helpers.py > gpurun_out/c4_helpers.txt 2>&1; grep -E "im2col|softmax|cbam" gpurun_out/c4_helpers.txt
