#!/bin/bash
# DEV: one gpurun call that produces the round's evidence: GPU tests, bench lines (C3 default, C5), ncu captures and launch lists.
# Every ncu command runs only after the same plain command has exited 0.  Output: gpurun_out/f_*
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/f_tests.log
timeout 600 python bench.py > gpurun_out/f_c3.json 2> gpurun_out/f_c3.err
timeout 300 python bench.py --config c5 --steps 4 --no-cpu-baseline --no-latency > gpurun_out/f_c5.json 2> gpurun_out/f_c5.err
timeout 200 python tools/profile_solve.py 148 300 > gpurun_out/f_prof.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:acb_solve -s 2 -c 1 -o gpurun_out/f_solve python tools/profile_solve.py 148 300 > /dev/null 2>&1
timeout 200 python tools/gpu_c5.py 128 > gpurun_out/f_c5tool.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_cols_it|k_level" -s 45 -c 3 -o gpurun_out/f_gen python tools/gpu_c5.py 128 > /dev/null 2>&1
timeout 200 python bench.py --steps 2 --warmup 3 --batch 1184 --no-cpu-baseline --no-latency > gpurun_out/f_c3_small.json 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_c3_launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 1184 --no-cpu-baseline --no-latency > /dev/null 2>&1
timeout 200 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/f_c5_small.json 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/f_c5_launches.csv \
    python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline --no-latency > /dev/null 2>&1
tail -3 gpurun_out/f_tests.log
cut -c1-300 gpurun_out/f_c3.json
cut -c1-300 gpurun_out/f_c5.json
cat gpurun_out/f_prof.log gpurun_out/f_c5tool.log
ls -la gpurun_out
