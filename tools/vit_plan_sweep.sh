#!/bin/bash
# Sweep (token tile width : K splits) of one GEMM of the ViT-L/14 forward; prints the forward time per setting.
# usage: tools/vit_plan_sweep.sh BATCH   (run under gpurun)
B=${1:-6}
run() { VFMREG_VIT_PLAN="$1" python tools/bench_kernels.py vitl $B 2>/dev/null | grep "^vit vitl14" | sed "s/^/$1  /"; }
run "none:0:0"
for nt in 256 224 192 160 128 96; do run "qkv:$nt:1"; done
for nt in 256 224 192 160 128; do run "fc1:$nt:1"; done
for sp in 1 2 3 4; do for nt in 256 192 128 96; do run "fc2:$nt:$sp"; done; done
for sp in 1 2 3; do for nt in 256 192 128 96; do run "proj:$nt:$sp"; done; done
