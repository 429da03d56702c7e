#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list + full capture of the
# dominant kernel.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 --timeout 600 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
if [ "${SKIP_BENCH}" != "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [ "${SKIP_NCU}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fwd -s 2 -c 2 \
      -o gpurun_out/prof_roi_fwd -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:proposals_kernel -s 2 -c 1 \
      -o gpurun_out/prof_proposals -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
  ls -la gpurun_out
fi
