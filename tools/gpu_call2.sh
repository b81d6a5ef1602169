#!/bin/bash
# validation of the 8-warp conv epilogue: tests + conv micro-benchmarks + quick bench + layer profile
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c2_tests.log 2>&1; tail -5 gpurun_out/c2_tests.log
timeout 150 python tools/bench_conv.py --cases g16_1x1,g8_1x1_res,stem_1x1,ds_1x1,up84_c2_head,gru,up84_c1,up84_c1_actlo,fuser_c2,res2_c3_precise > gpurun_out/c2_conv.txt 2>&1; cat gpurun_out/c2_conv.txt
timeout 100 python bench.py --quick --no-cpu-baseline --no-torch-baseline > gpurun_out/c2_bench_quick.json 2> gpurun_out/c2_bench_quick.err; cut -c1-330 gpurun_out/c2_bench_quick.json; tail -2 gpurun_out/c2_bench_quick.err
timeout 90 python tools/profile_layers.py > gpurun_out/c2_layers.txt 2>&1; tail -2 gpurun_out/c2_layers.txt
