#!/bin/bash
# Measurement pass of a round (run on the GPU box via gpurun): GPU test suite with the parity prints, the bench line of
# every BASELINE config, ncu launch list, ncu full-set captures of the hot kernels, per-shape GEMM table of one UNet
# evaluation.  Everything lands in gpurun_out/; the summaries worth keeping are copied to profiles/ by hand.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "\[parity\]|\[short-key|\[cta pair\]|\[stream-K\]|\[flash poly\]|\[smoke\]|passed|failed|error" ) > gpurun_out/r2_gputest_final.log 2>&1; tail -3 gpurun_out/r2_gputest_final.log
timeout 600 python bench.py > gpurun_out/r2_bench_c2_final.json 2> gpurun_out/r2_bench_c2_final.err; cut -c1-300 gpurun_out/r2_bench_c2_final.json
for c in 3 4 5 1; do timeout 500 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2_bench_c${c}_final.json 2> gpurun_out/r2_bench_c${c}_final.err; cut -c1-160 gpurun_out/r2_bench_c${c}_final.json; done
timeout 500 bash tools/profile_launches.sh > gpurun_out/r2_prof.log 2>&1; head -14 gpurun_out/launches_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"xattn_short|flash_attn" -o gpurun_out/r2_xattn_final python tools/ncu_xattn.py > gpurun_out/r2_ncu_xattn_final.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|flash_attn" -o gpurun_out/r2_targets2_final python tools/ncu_targets2.py > gpurun_out/r2_ncu_targets2_final.log 2>&1
PFD_GEMM_TRACE=1 timeout 400 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/unet_eval_launches.csv python tools/unet_eval_profile.py > gpurun_out/unet_eval.out 2> gpurun_out/unet_eval_trace.log
python tools/gemm_breakdown.py gpurun_out/unet_eval_launches.csv gpurun_out/unet_eval_trace.log > gpurun_out/r2_gemm_breakdown.txt 2>&1; head -12 gpurun_out/r2_gemm_breakdown.txt
