#!/bin/bash
# Proposals kernel, one CTA per image vs clusters of 2 / 4 / 8 CTAs per image: parity tests (all
# variants, under a timeout: a cluster-barrier bug would hang) and timings.
set +e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_proposal_gpu.py tests/test_pipeline_gpu.py -q -m gpu --maxfail=5 --timeout 120 > gpurun_out/pytest_prop.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_prop.log
for v in 0 2 4 8; do
  echo "== WSSDL_PROPOSALS_CLUSTER=$v"
  WSSDL_PROPOSALS_CLUSTER=$v timeout 300 python scripts/microbench.py --only proposals,train 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if d.get('op') in ('proposals', 'train_step_device_kernels'):
        print('%-28s %-44s B=%-3d %8.4f ms' % (d['op'], d['tag'], d['B'], d['ms']))
"
done
