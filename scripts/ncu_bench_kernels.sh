#!/bin/bash
# run on the GPU box (via gpurun): launch list of one bench step + full captures of the hot kernels
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for k in roi_align_fwd_sep nms_fused embed_match_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -f -o gpurun_out/final_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
# the bit-exact forward (bench --math exact) and the backward (not part of the inference step)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_march -s 6 -c 2 -f \
    -o gpurun_out/final_roi_align_fwd_march python bench.py --math exact --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for res in 7 14; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_march -s 1 -c 1 -f \
      -o gpurun_out/final_roi_align_bwd_march_$res python scripts/ncu_bwd.py $res > /dev/null 2>&1
done
# the post-processing kernels around the NMS (module-level scripts)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rpn_front_kernel -s 2 -c 1 -f \
    -o gpurun_out/final_rpn_front_kernel python scripts/perf_rpn_front.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"cand_kernel|select_detections_kernel|seg_scan_kernel" -s 8 -c 4 -f \
    -o gpurun_out/final_box_post python scripts/perf_modules.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paste_masks_kernel -s 2 -c 1 -f \
    -o gpurun_out/final_paste_masks_kernel python scripts/perf_masks.py > /dev/null 2>&1
ls -la gpurun_out
