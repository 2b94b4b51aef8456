#!/bin/bash
# Round-2 measurement pass (run under gpurun, one GPU): bench line, launch list, ncu --set full captures.
# usage: tools/gpu_profile.sh TAG
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --scenes-per-step 1 --lanes 1 --distinct-pairs 0 --no-cpu-baseline --no-extras > $O/${TAG}_launches.log 2>&1
FULL="$NCU --set full --import-source on"
ONE_LANE=1 $FULL -k regex:match_tc3 -s 2 -c 2 -f -o $O/${TAG}_ncu_match python tools/one_register.py configs1 1 > $O/${TAG}_ncu.log 2>&1
ONE_LANE=1 $FULL -k regex:score_kernel -s 1 -c 1 -f -o $O/${TAG}_ncu_score python tools/one_register.py configs1 1 >> $O/${TAG}_ncu.log 2>&1
$FULL -k regex:match_tc3 -s 1 -c 1 -f -o $O/${TAG}_ncu_refshape python tools/one_register.py refshape 2 >> $O/${TAG}_ncu.log 2>&1
$FULL -k regex:match_tc3 -s 0 -c 1 -f -o $O/${TAG}_ncu_c3 python tools/one_register.py configs3 1 >> $O/${TAG}_ncu.log 2>&1
$FULL -k regex:vit_gemm -s 100 -c 5 -f -o $O/${TAG}_ncu_vitgemm python tools/one_vit.py vitl14 6 >> $O/${TAG}_ncu.log 2>&1
$FULL -k regex:attention -s 24 -c 1 -f -o $O/${TAG}_ncu_attn python tools/one_vit.py vitl14 6 >> $O/${TAG}_ncu.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${TAG}_launches_vitl_b6.csv python tools/one_vit.py vitl14 6 >> $O/${TAG}_ncu.log 2>&1
python tools/bench_kernels.py > $O/${TAG}_kernel_bench.txt 2>&1
ls -la $O | tail -20
