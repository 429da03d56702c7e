#!/bin/bash
# ncu --set full capture of the RoI-pool forward kernel on the C4 workload (bench.py shapes)
mkdir -p gpurun_out
for v in 1 2; do
WSSDL_ROI_FWD_VPT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd -s 2 -c 1 \
   -o gpurun_out/prof_roi_fwd_vpt$v -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vpt$v.log 2>&1
done
ls -la gpurun_out
