#!/bin/bash
# Scaling run on one 8-GPU box: bench.py at the given N values (strong scaling of the 256-image
# batch; the weak-scaling value is a key of the same line).
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.csv 2>&1
nvidia-smi topo -m > gpurun_out/scale_topo.txt 2>&1
free -g > gpurun_out/scale_host.txt; nproc >> gpurun_out/scale_host.txt; lscpu | grep -E "NUMA|Model name|Socket" >> gpurun_out/scale_host.txt
avail=$(awk '/MemAvailable/{print int($2/1048576)}' /proc/meminfo)
for N in ${SCALE_NS:-8 4 2}; do
  extra=""
  # the e2e leg pins 16.5 GB of host memory per rank
  if [ "$avail" -lt $((N * 24)) ]; then extra="--no-e2e"; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + N)) bench.py --gpus $N --steps 10 --warmup 3 $extra \
      > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  echo "N=$N rc=$? (MemAvailable ${avail} GB $extra)"; tail -1 gpurun_out/bench_n$N.json | python -c "
import json,sys
b=json.loads(sys.stdin.read())
print('value',round(b['value']),'e2e',round(b['e2e']['value']),'dev',round(b['e2e_features_on_device']['value']),'numa',b.get('numa_node_rank0'))"
done
