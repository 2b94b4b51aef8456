python -m pytest tests/test_gpu_vit.py tests/test_gpu_project.py -x -q 2>&1 | tail -2
python tools/bench_kernels.py vit 2>&1 | grep "^vit vit" > gpurun_out/r2n_vit_bench.txt; cat gpurun_out/r2n_vit_bench.txt
NCU="ncu --clock-control none --set full --import-source on"
$NCU -k regex:vit_gemm -s 110 -c 6 -f -o gpurun_out/r2n_ncu_vitgemm_b48 python tools/one_vit.py vitl14 48 > gpurun_out/r2n_ncu.log 2>&1
$NCU -k regex:attention_tc -s 24 -c 1 -f -o gpurun_out/r2n_ncu_attn_b48 python tools/one_vit.py vitl14 48 >> gpurun_out/r2n_ncu.log 2>&1
$NCU -k regex:vit_gemm -s 110 -c 6 -f -o gpurun_out/r2n_ncu_vitgemm_b6 python tools/one_vit.py vitl14 6 >> gpurun_out/r2n_ncu.log 2>&1
$NCU -k regex:attention_tc -s 24 -c 1 -f -o gpurun_out/r2n_ncu_attn_b6 python tools/one_vit.py vitl14 6 >> gpurun_out/r2n_ncu.log 2>&1
