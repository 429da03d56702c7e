#!/bin/bash
# One gpurun call for the RoI-pool forward kernels: parity tests, direct / tiled / band / sorted
# timings, and (unless SKIP_NCU=1) a full ncu capture of one of them (NCU_KERNEL=band|sorted|...,
# NCU_REGEX = its kernel-name regex) on the bench (C4) workload.
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_roi_pool_gpu.py tests/test_pipeline_gpu.py -q -m gpu --maxfail=10 --timeout 600 > gpurun_out/pytest_roi.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_roi.log
rm -f gpurun_out/microbench_roicmp.jsonl
timeout 900 python scripts/microbench.py --only roicmp --out gpurun_out/microbench_roicmp.jsonl > gpurun_out/microbench_roicmp.log 2>&1
echo "microbench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/microbench_roicmp.jsonl'):
    d = json.loads(l)
    if d.get('op') == 'roi_pool_fwd':
        print('%-50s %8.4f ms %8.1f GB/s  %.3f of measured' % (d['tag'], d['ms'], d['gbs'], d['frac_measured']))
PY
if [ "${SKIP_NCU}" != "1" ]; then
  WSSDL_ROI_FWD_KERNEL=${NCU_KERNEL:-band} timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_REGEX:-roi_pool_fwd_band} -s 2 -c 1 \
      -o gpurun_out/prof_roi_fwd_band -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_band.log 2>&1
  tail -3 gpurun_out/ncu_band.log
fi
ls -la gpurun_out
