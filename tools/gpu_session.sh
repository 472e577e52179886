mkdir -p gpurun_out
(timeout 240 python -m pytest tests/test_kernels_gpu.py -x -q -s -k "cta_pair" 2>&1 | tail -25) > gpurun_out/r2_t_pair.log 2>&1; tail -25 gpurun_out/r2_t_pair.log
if grep -q "passed" gpurun_out/r2_t_pair.log && ! grep -q "failed" gpurun_out/r2_t_pair.log; then
  timeout 400 python tools/gemm_pair_ab.py > gpurun_out/r2_gemm_pair_ab.log 2>&1; tail -16 gpurun_out/r2_gemm_pair_ab.log
  timeout 300 python tools/ab_unet.py --env-variant pair=gemm_pair:1 > gpurun_out/r2_ab_pair.log 2>&1; tail -4 gpurun_out/r2_ab_pair.log
fi
(timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -5) > gpurun_out/r2_t_kernels2.log 2>&1; tail -3 gpurun_out/r2_t_kernels2.log
