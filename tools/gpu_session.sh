mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -s -k "polynomial" 2>&1 | grep -E "flash poly|passed|failed|Error" | tail -20) > gpurun_out/r2_t_f16x2.log 2>&1; tail -20 gpurun_out/r2_t_f16x2.log
timeout 300 python tools/xattn_perf.py --rounds 3 > gpurun_out/r2_xattn_perf4.log 2>&1; grep -v XATTN_RESULT gpurun_out/r2_xattn_perf4.log
timeout 300 python tools/ab_unet.py --env-variant f16x2=flash_poly_mod:1 > gpurun_out/r2_ab_f16x2.log 2>&1; tail -4 gpurun_out/r2_ab_f16x2.log | cut -c1-200
