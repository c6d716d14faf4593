#!/bin/bash
# run on the GPU box (via gpurun): round-2 evidence -- launch list of the bench step + full captures of the hot kernels
mkdir -p gpurun_out
rm -f gpurun_out/final_*
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline-all --sustain 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_bench_launches.csv $B > /dev/null 2>&1
for k in roi_align_fwd_rows nms_fused roi_order_kernel select_topk; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/final_$k $B > /dev/null 2>&1
done
# the tcgen05 kernels at the BASELINE config sizes
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o gpurun_out/final_tc_gemm_softmax_cfg5_top1 python scripts/ncu_match.py cfg5 top > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o gpurun_out/final_tc_gemm_softmax_cfg5_probs python scripts/ncu_match.py cfg5 probs > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o gpurun_out/final_tc_gemm_softmax_cfg3_probs python scripts/ncu_match.py cfg3 probs > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o gpurun_out/final_tc_gemm_softmax_wide_lvis python scripts/ncu_match.py lvis probs > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -f -o gpurun_out/final_tc_gemm_linear_emb_pred python scripts/ncu_match.py linear > /dev/null 2>&1
# backward (not part of the inference step) and the bf16 forward
for res in 7 14; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_taprow -s 1 -c 1 -f -o gpurun_out/final_roi_align_bwd_taprow_$res python scripts/ncu_bwd.py $res > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_rows -s 1 -c 1 -f -o gpurun_out/final_roi_align_fwd_rows_bf16 python scripts/ncu_rows2.py 0 step bf16 > /dev/null 2>&1
# the tiled gather at BASELINE config #1 (the reference's shipped C4 pooler)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_tile -s 3 -c 1 -f -o gpurun_out/final_roi_align_fwd_tile python scripts/perf_config1.py 0 > /dev/null 2>&1
ls -la gpurun_out | head -40
