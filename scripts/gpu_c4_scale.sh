#!/bin/bash
# BASELINE config C4 (8 GiB of gzip members, the file on rank 0's GPU) at 8 / 4 / 2 / 1 GPUs of one box.
# Usage (from the build box): gpurun --gpus 8 --timeout 1200 -- scripts/gpu_c4_scale.sh [MiB]
MIB=${1:-8192}
mkdir -p gpurun_out
for n in ${2:-8 4 2 1}; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --workload c4 --c4-mib $MIB --steps 3 --warmup 3 2> gpurun_out/c4_n$n.err | grep "^{" > gpurun_out/c4_n$n.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n \
      bench.py --gpus $n --workload c4 --c4-mib $MIB --steps 3 --warmup 3 2> gpurun_out/c4_n$n.err | grep "^{" > gpurun_out/c4_n$n.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c4_n$n.json"))
    print("N=%d  %.1f GB/s  %.2f ms/step  phases %s  collective share %.2f  (generated in %.0f s)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["phases_ms"], d["collective_share"], d["generated_in_s"]))
except Exception as e:
    print("N=$n failed:", e); print(open("gpurun_out/c4_n$n.err").read()[-1500:])
PY
done
