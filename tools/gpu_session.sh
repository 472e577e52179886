mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "short_key or flash" 2>&1 | tail -5) > gpurun_out/r2_t_short3.log 2>&1; tail -3 gpurun_out/r2_t_short3.log
timeout 300 python tools/xattn_perf.py --rounds 3 > gpurun_out/r2_xattn_perf3.log 2>&1; head -16 gpurun_out/r2_xattn_perf3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"xattn_short" -c 2 -o gpurun_out/r2_xattn2 python tools/ncu_xattn.py > gpurun_out/r2_ncu_xattn2.log 2>&1; tail -2 gpurun_out/r2_ncu_xattn2.log
