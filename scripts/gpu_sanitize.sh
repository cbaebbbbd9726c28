#!/bin/bash
# the compute-sanitizer half of gpu_multi.sh on one GPU
sed -n '/== compute-sanitizer memcheck/,$p' scripts/gpu_multi.sh > /tmp/san_part.sh
bash /tmp/san_part.sh
