#!/bin/bash
# profiling call: tests + micro-benchmarks + ncu source-level captures of the shallow-K conv and the in-step read kernels
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/c1_tests.log 2>&1; tail -3 gpurun_out/c1_tests.log
timeout 120 python tools/bench_conv.py --cases g16_1x1,g8_1x1_res,stem_1x1,ds_1x1,up84_c2_head,gru,up84_c1 > gpurun_out/c1_conv.txt 2>&1; cat gpurun_out/c1_conv.txt
timeout 100 python bench.py --quick --no-cpu-baseline --no-torch-baseline > gpurun_out/c1_bench_quick.json 2> gpurun_out/c1_bench_quick.err; cut -c1-330 gpurun_out/c1_bench_quick.json
timeout 120 python tools/bench_helpers.py 2>&1 | grep -i "im2col" > gpurun_out/c1_helpers.txt; cat gpurun_out/c1_helpers.txt
timeout 150 ncu --set full --import-source on --clock-control none -k regex:conv_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/c1_conv_g16 -f python tools/bench_conv.py --cases g16_1x1 > gpurun_out/c1_ncu_g16.log 2>&1; tail -2 gpurun_out/c1_ncu_g16.log
timeout 150 ncu --set full --import-source on --clock-control none -k regex:conv_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/c1_conv_stem -f python tools/bench_conv.py --cases stem_1x1 > gpurun_out/c1_ncu_stem.log 2>&1; tail -2 gpurun_out/c1_ncu_stem.log
timeout 200 ncu --set full --import-source on --clock-control none -k 'regex:readout_sparse|simtopk_kernel|merge_kernel|thr_floor' --launch-skip 20 --launch-count 4 -o gpurun_out/c1_read -f python bench.py --quick --no-cpu-baseline --no-torch-baseline --steps 3 --warmup 3 > gpurun_out/c1_ncu_read.log 2>&1; tail -2 gpurun_out/c1_ncu_read.log
ls -la gpurun_out/*.ncu-rep
