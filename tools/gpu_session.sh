mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -30) > gpurun_out/r2_t_kernels.log 2>&1
tail -12 gpurun_out/r2_t_kernels.log
timeout 300 python tools/gemm_epi_ab.py > gpurun_out/r2_gemm_epi_ab.log 2>&1; tail -20 gpurun_out/r2_gemm_epi_ab.log
timeout 300 python tools/xattn_perf.py --rounds 3 > gpurun_out/r2_xattn_perf2.log 2>&1; grep -A8 "cross L0 512" gpurun_out/r2_xattn_perf2.log | head -12
timeout 400 python tools/ab_unet.py --env-variant notma=gemm_tma_epi:0 --env-variant noshort=xattn_short:0 > gpurun_out/r2_ab_tma.log 2>&1; tail -5 gpurun_out/r2_ab_tma.log
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2_gputest3.log 2>&1; tail -8 gpurun_out/r2_gputest3.log
