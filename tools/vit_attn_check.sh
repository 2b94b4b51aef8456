# after a change to the attention kernel: parity, timing at 6 / 48 images, in-situ trace
out=gpurun_out/r2g_attn_check.txt; : > $out
timeout 600 python -m pytest tests/test_gpu_vit.py -x -q 2>&1 | tail -2 | tee -a $out
for b in 6 48; do timeout 300 python tools/bench_kernels.py vitl $b 2>&1 | grep "^vit " | tee -a $out; done
VFMREG_LIB=vfm_registration_b200/libvfmreg_b200_trace.so python tools/vit_trace.py vitl14 48 > gpurun_out/r2g_trace_b48.txt 2>&1; tail -11 gpurun_out/r2g_trace_b48.txt | tee -a $out
VFMREG_LIB=vfm_registration_b200/libvfmreg_b200_trace.so python tools/vit_trace.py vitl14 6 > gpurun_out/r2g_trace_b6.txt 2>&1; tail -11 gpurun_out/r2g_trace_b6.txt | tee -a $out
