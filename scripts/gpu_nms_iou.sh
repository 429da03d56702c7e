#!/bin/bash
# One gpurun call for the NMS / bbox_overlaps work: parity tests + per-kernel timings.
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nms_gpu.py tests/test_bbox_gpu.py tests/test_proposal_gpu.py -q -m gpu --maxfail=10 --timeout 600 > gpurun_out/pytest_nms_iou.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_nms_iou.log
rm -f gpurun_out/microbench_nms_iou.jsonl
timeout 900 python scripts/microbench.py --only nms,iou ${MB_ARGS} --out gpurun_out/microbench_nms_iou.jsonl > gpurun_out/microbench_nms_iou.log 2>&1
echo "microbench rc=$?"; tail -3 gpurun_out/microbench_nms_iou.log
