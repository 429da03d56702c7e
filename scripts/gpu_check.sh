#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), per-kernel timings, ncu launch list
# + full captures of the kernels named in DESIGN.md.  Everything lands in gpurun_out/.
# Every step runs under its own `timeout`.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
if [ "${SKIP_TESTS}" != "1" ]; then
  timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 --timeout 600 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if [ "${SKIP_BENCH}" != "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  echo "bench reference rc=$?"; cat gpurun_out/bench_reference.json
  timeout 300 python bench.py --images 32 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_images32.json 2> gpurun_out/bench_images32.err
  echo "bench 32 images rc=$?"; cat gpurun_out/bench_images32.json
fi
if [ "${SKIP_MICRO}" != "1" ]; then
  rm -f gpurun_out/microbench.jsonl
  timeout 1500 python scripts/microbench.py --cpu --out gpurun_out/microbench.jsonl > gpurun_out/microbench.log 2>&1
  echo "microbench rc=$?"
  timeout 600 python scripts/microbench.py --only roicmp --out gpurun_out/microbench.jsonl >> gpurun_out/microbench.log 2>&1
fi
if [ "${SKIP_NCU}" != "1" ]; then
  NCU="ncu --clock-control none"
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 60 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
  timeout 300 $NCU --metrics gpu__time_duration.sum -c 60 --csv \
      --log-file gpurun_out/launches_images32.csv python bench.py --images 32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline >> gpurun_out/ncu_list.log 2>&1
  FULL="$NCU --set full --import-source on -f"
  B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
  timeout 600 $FULL -k regex:roi_pool_fwd_bins -s 2 -c 1 -o gpurun_out/prof_roi_fwd_sorted $B > gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:roi_bin_sort -s 2 -c 1 -o gpurun_out/prof_roi_bin_sort $B >> gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:proposals_kernel -s 2 -c 1 -o gpurun_out/prof_proposals $B >> gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:proposals_kernel -s 2 -c 1 -o gpurun_out/prof_proposals_images32 $B --images 32 >> gpurun_out/ncu_full.log 2>&1
  WSSDL_ROI_FWD_KERNEL=band timeout 600 $FULL -k regex:roi_pool_fwd_band -s 2 -c 1 -o gpurun_out/prof_roi_fwd_band $B >> gpurun_out/ncu_full.log 2>&1
  WSSDL_ROI_FWD_KERNEL=direct timeout 600 $FULL -k regex:roi_pool_fwd_kernel -s 2 -c 1 -o gpurun_out/prof_roi_fwd $B >> gpurun_out/ncu_full.log 2>&1
  WSSDL_ROI_FWD_KERNEL=tiled timeout 600 $FULL -k regex:roi_pool_fwd_tiled -s 2 -c 1 -o gpurun_out/prof_roi_fwd_tiled $B >> gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:"nms_mask|nms_sweep" -s 2 -c 2 -o gpurun_out/prof_nms \
      python scripts/ncu_targets.py nms >> gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:bbox_overlaps -s 2 -c 1 -o gpurun_out/prof_iou_f64 \
      python scripts/ncu_targets.py iou >> gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:bbox_overlaps -s 5 -c 1 -o gpurun_out/prof_iou_f32 \
      python scripts/ncu_targets.py iou >> gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:detect_postprocess -s 1 -c 1 -o gpurun_out/prof_detect \
      python scripts/ncu_targets.py detect >> gpurun_out/ncu_full.log 2>&1
  timeout 600 $FULL -k regex:roi_pool_bwd -s 1 -c 1 -o gpurun_out/prof_roi_bwd \
      python scripts/ncu_targets.py bwd >> gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
