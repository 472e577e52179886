mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -3) > gpurun_out/r2_t_kernels3.log 2>&1; tail -2 gpurun_out/r2_t_kernels3.log
: > gpurun_out/r2_cmp_libs2.log
for i in 1 2 3; do
  for lib in r2start 1f0fec5 HEAD; do
    if [ $lib = HEAD ]; then L=$PWD/pfd_b200/libpfd_b200.so; else L=$PWD/tools/oldlib/libpfd_b200_$lib.so; fi
    echo "== $lib round $i" >> gpurun_out/r2_cmp_libs2.log
    PFD_B200_LIB=$L timeout 200 python tools/ab_unet.py --rounds 4 --reps 20 2>&1 | grep "^default" >> gpurun_out/r2_cmp_libs2.log
  done
done
cat gpurun_out/r2_cmp_libs2.log
