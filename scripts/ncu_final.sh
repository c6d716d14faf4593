#!/bin/bash
# run on the GPU box (via gpurun): launch list of one bench step + a full capture of the roofline kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_rows -s 6 -c 1 -f \
    -o gpurun_out/final_roi_align_fwd_rows python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
