#!/bin/bash
# Turns the artifacts of scripts/gpu_check.sh (gpurun_out/) into the tracked summaries under
# profiles/ (run here, no GPU needed: ncu only reads the reports).
set -e
R=${1:-r02}
O=gpurun_out
P=profiles
cp $O/bench.json $P/${R}_bench_n1.json
cp $O/bench_reference.json $P/${R}_bench_reference_n1.json
cp $O/launches.csv $P/${R}_launches.csv
[ -f $O/launches_images32.csv ] && cp $O/launches_images32.csv $P/${R}_launches_images32.csv
[ -f $O/bench_images32.json ] && cp $O/bench_images32.json $P/${R}_bench_n1_images32.json
cp $O/microbench.jsonl $P/${R}_microbench.jsonl
{ cat $O/host.txt; cat $O/gpu.csv; } > $P/${R}_host.txt
EXTRA="dram__bytes_read.sum dram__bytes_write.sum bank_conflicts_pipe_lsu_mem_shared_op_ld.sum l1tex__data_pipe_lsu_wavefronts.sum.pct l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum lts__t_sectors_srcunit_tex_op_read.sum lts__t_sectors_srcunit_tex_op_write.sum sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active smsp__thread_inst_executed_per_inst_executed.ratio"
for k in roi_fwd_sorted roi_bin_sort roi_fwd_band roi_fwd roi_fwd_tiled proposals proposals_images32 nms iou_f64 iou_f32 detect roi_bwd; do
  f=$O/prof_$k.ncu-rep
  [ -f $f ] || continue
  out=$P/${R}_ncu_$k.txt
  echo "# ncu --set full --clock-control none, scripts/gpu_check.sh (B200, $R final code); see scripts/ncu_summary.py" > $out
  python scripts/ncu_summary.py $f >> $out
  python scripts/ncu_metrics.py $f $EXTRA | grep -v "^###" >> $out
done
ls -la $P
