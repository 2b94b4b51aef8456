# ViT-L/14: parity, timing, and ncu --set full captures of the GEMM epilogue variants and the attention kernel (48 / 6 images)
python -m pytest tests/test_gpu_vit.py -x -q 2>&1 | tail -2
for b in 6 12 48; do python tools/bench_kernels.py vitl $b 2>&1 | grep "^vit "; done | tee gpurun_out/r2h_vit_bench.txt
NCU="ncu --clock-control none --set full --import-source on"
# launches of one layer in the middle of the forward: qkv, proj, fc1, fc2 (vit_gemm) and the attention between them
$NCU -k regex:vit_gemm -s 41 -c 4 -f -o gpurun_out/r2h_ncu_vitgemm_b48 python tools/one_vit.py vitl14 48 > gpurun_out/r2h_ncu.log 2>&1
$NCU -k regex:attention_tc -s 10 -c 1 -f -o gpurun_out/r2h_ncu_attn_b48 python tools/one_vit.py vitl14 48 >> gpurun_out/r2h_ncu.log 2>&1
$NCU -k regex:layernorm -s 20 -c 1 -f -o gpurun_out/r2h_ncu_ln_b48 python tools/one_vit.py vitl14 48 >> gpurun_out/r2h_ncu.log 2>&1
$NCU -k regex:attention_tc -s 10 -c 1 -f -o gpurun_out/r2h_ncu_attn_b6 python tools/one_vit.py vitl14 6 >> gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_ncu.log
