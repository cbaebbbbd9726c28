#!/bin/bash
# On an 8-GPU box: the 2-GPU parity test of the C4 path, then C4 at the given GPU counts; optionally the C2 headline at 8 GPUs
# (e2e vs the box's pinned-copy ceiling).  Usage: gpurun --gpus 8 --timeout 1500 -- scripts/gpu_multi8.sh "8 4 2" [c2]
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
scripts/gpu_c4_scale.sh 8192 "${1:-8 2}"
if [ "$2" = "c2" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
  bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/c2_n8.err | grep "^{" > gpurun_out/c2_n8.json
python - <<PY
import json
d = json.load(open("gpurun_out/c2_n8.json"))
print("C2 N=8: value %.1f GB/s, e2e %s" % (d["value"], json.dumps(d["e2e"])))
PY
fi
