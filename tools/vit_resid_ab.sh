# A/B of the fused residual epilogue (EPI_F32_RESID) against partial sums + fold, ViT-L/14 at several batch sizes
out=gpurun_out/r2g_resid_ab.txt; : > $out
timeout 600 python -m pytest tests/test_gpu_vit.py -x -q 2>&1 | tail -2 | tee -a $out
for b in 6 12 24 48; do
  for r in 1 0; do
    echo "# B=$b VFMREG_VIT_RESID=$r" >> $out
    VFMREG_VIT_RESID=$r VFMREG_VIT_VERBOSE=1 timeout 300 python tools/bench_kernels.py vitl $b 2>&1 | grep -i "plan\|split\|^vit " >> $out
  done
done
cat $out
