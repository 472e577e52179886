mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_split_gpu.py -q -s 2>&1 | tail -6) > gpurun_out/r2_t_split.log 2>&1; tail -4 gpurun_out/r2_t_split.log
timeout 400 python bench.py --gpus 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_bench_2gpu_weak.json 2> gpurun_out/r2_bench_2gpu_weak.err; cut -c1-250 gpurun_out/r2_bench_2gpu_weak.json
timeout 400 python bench.py --gpus 2 --split --batch 8 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_bench_2gpu_split.json 2> gpurun_out/r2_bench_2gpu_split.err; cut -c1-250 gpurun_out/r2_bench_2gpu_split.json
(timeout 400 python -m pytest tests/test_configs_gpu.py tests/test_pipeline_gpu.py -q -s 2>&1 | grep -E "\[parity\]|passed|failed") > gpurun_out/r2_parity_full.log 2>&1; tail -2 gpurun_out/r2_parity_full.log
