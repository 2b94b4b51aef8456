#!/bin/bash
# Focused re-sweep of the (tile width : K splits) plans at HEAD (the residual + LayerNorm kernel reads every partial sum, so the
# forward time is what decides).  usage: tools/vit_plan_sweep2.sh BATCH
B=${1:-6}
run() { VFMREG_VIT_PLAN="$1" python tools/bench_kernels.py vitl $B 2>/dev/null | grep "^vit vitl14" | sed "s/^/$1  /" | cut -c1-70; }
run "none:0:0"
for p in proj:128:1 proj:96:1 proj:192:1 proj:224:2 proj:256:2 proj:160:2; do run $p; done
for p in fc2:192:2 fc2:256:2 fc2:128:1 fc2:96:1 fc2:160:2 fc2:224:2; do run $p; done
for p in qkv:256:1 qkv:160:1 qkv:192:1; do run $p; done
for p in fc1:192:1 fc1:256:1 fc1:224:1 fc1:160:1; do run $p; done
run "proj:128:1,fc2:192:2"
run "proj:96:1,fc2:192:2"
