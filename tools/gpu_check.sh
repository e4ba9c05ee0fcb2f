#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list + one full capture of the dominant kernel.
# usage: tools/gpu_check.sh <tag> [kernel-regex]
tag="${1:-run}"; kre="${2:-frame_ll}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_$tag.log
tail -2 gpurun_out/bench_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"frame_ll|w8_gem|attn_decode|sample_kernel|advance|next_input|rmsnorm|tapgemm|dwconv|window_attn|snake|clamp|rvq|rope|act_prep" --csv --log-file gpurun_out/launches_$tag.csv python tools/ncu_step.py > gpurun_out/ncu_list_$tag.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s 1 -c 1 -o gpurun_out/full_$tag -f python tools/ncu_step.py > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
