#!/bin/bash
# One GPU-box pass: parity tests, bench (N=1), ncu launch list, ncu full captures of the three kernels.
# Usage: gpurun --timeout 1700 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
export LAS_PARITY_REPORT=$PWD/gpurun_out/parity_shapes_$TAG.jsonl; rm -f $LAS_PARITY_REPORT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:speller_decode_persistent -s 9 -c 1 -f -o gpurun_out/prof_decoder_$TAG \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-pipeline > gpurun_out/ncu_dec_$TAG.log 2>&1; echo "ncu dec rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_recurrence_cluster -s 9 -c 3 -f -o gpurun_out/prof_rec_$TAG \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-pipeline > gpurun_out/ncu_rec_$TAG.log 2>&1; echo "ncu rec rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 6 -c 3 -f -o gpurun_out/prof_gemm_$TAG \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-pipeline > gpurun_out/ncu_gemm_$TAG.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out | tail -12
# the fused generic decoder step (shipped 1024x2 speller): launch list of a few steps + one full capture per kernel
LAS_PROBE_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 400 -c 60 --csv \
  --log-file gpurun_out/launches_yaml_$TAG.csv python tools/gen_step_probe.py 16 64 > gpurun_out/ncu_yaml_list_$TAG.log 2>&1; echo "ncu yaml list rc=$?"
LAS_PROBE_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gen_cell_step -s 200 -c 2 -f -o gpurun_out/prof_gencell_$TAG \
  python tools/gen_step_probe.py 16 64 > gpurun_out/ncu_gencell_$TAG.log 2>&1; echo "ncu gen cell rc=$?"
LAS_PROBE_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attend_cluster -s 100 -c 1 -f -o gpurun_out/prof_attend_$TAG \
  python tools/gen_step_probe.py 16 64 > gpurun_out/ncu_attend_$TAG.log 2>&1; echo "ncu attend rc=$?"
# (the c4 decoder -- cooperative launch + cluster dimension -- cannot be profiled: ncu reports LaunchFailed; its numbers come from CUDA events)
