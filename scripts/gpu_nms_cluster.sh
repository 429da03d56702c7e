#!/bin/bash
# Stand-alone NMS: single-CTA sweep vs the cluster sweep: parity tests (both variants, under a
# timeout) and the C5 timing sweep.
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nms_gpu.py tests/test_detect_gpu.py -q -m gpu --maxfail=5 --timeout 300 > gpurun_out/pytest_nms.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_nms.log
for v in 0 1; do
  echo "== WSSDL_NMS_SWEEP_CLUSTER=$v"
  WSSDL_NMS_SWEEP_CLUSTER=$v timeout 600 python scripts/microbench.py --only nms 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if d.get('op') == 'nms' and d['thresh'] == 0.7:
        print('N=%-7d clustered=%-5s kept=%-6d %8.4f ms' % (d['N'], d['clustered'], d['kept'], d['ms']))
"
done
