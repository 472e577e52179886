mkdir -p gpurun_out
timeout 120 python tools/sk_debug.py > gpurun_out/r2_sk_debug2.log 2>&1; sed -n 1,12p gpurun_out/r2_sk_debug2.log
(timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "stream_k or tma_store or cta_pair" 2>&1 | tail -4) > gpurun_out/r2_t_sk2.log 2>&1; tail -3 gpurun_out/r2_t_sk2.log
timeout 400 python tools/gemm_pair_ab.py --option gemm_streamk --on 1 --off 0 > gpurun_out/r2_gemm_sk_ab2.log 2>&1; tail -16 gpurun_out/r2_gemm_sk_ab2.log | cut -c1-150
timeout 300 python tools/ab_unet.py --env-variant nosk=gemm_streamk:0 > gpurun_out/r2_ab_sk2.log 2>&1; tail -4 gpurun_out/r2_ab_sk2.log | cut -c1-200
