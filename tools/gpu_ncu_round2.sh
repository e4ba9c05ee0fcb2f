#!/bin/bash
# Round-2 ncu evidence, run on the GPU box (gpurun -- bash tools/gpu_ncu_round2.sh).  Reports stay in /tmp; only summaries and
# raw-page CSVs go to gpurun_out/ (the merge-back limit is 64 MiB).  Every ncu run follows a plain run of the same command.
set -x
S="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --metrics sm__inst_executed_pipe_tc.sum,sm__inst_executed_pipe_tensor.sum,l1tex__m_xbar2l1tex_read_bytes_pipe_tma.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__mem_tensor_cycles_active.avg,l1tex__data_pipe_tc_wavefronts.sum"
python tools/ncu_step.py > gpurun_out/r2_ncu_plain1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:frame_ll -s 1 -c 2 -o /tmp/r2_full_frame_ll -f python tools/ncu_step.py > gpurun_out/r2_ncu1.log 2>&1
ncu -i /tmp/r2_full_frame_ll.ncu-rep --page raw --csv > gpurun_out/r2_frame_ll_raw.csv
python tools/ncu_summary.py /tmp/r2_full_frame_ll.ncu-rep gpurun_out/r02_ncu_full_frame_ll_talker_step.json "ncu --set full, frame_ll_kernel stack mode (talker decode step, ctx 300), round 2 tree"
cp /tmp/r2_full_frame_ll.ncu-rep gpurun_out/
python tools/gemm_ncu.py > gpurun_out/r2_ncu_plain2.log 2>&1 && \
  ncu $S --clock-control none -k regex:w8_gemm_tc -c 18 -o /tmp/r2_w8 -f python tools/gemm_ncu.py > gpurun_out/r2_ncu2.log 2>&1
ncu -i /tmp/r2_w8.ncu-rep --page raw --csv > gpurun_out/r2_w8_gemm_tc_raw.csv
python tools/codec_probe.py 32 96 > gpurun_out/r2_ncu_plain3.log 2>&1 && \
  ncu $S --clock-control none -k regex:tapgemm_tc -s 111 -c 111 -o /tmp/r2_tap -f python tools/codec_probe.py 32 96 > gpurun_out/r2_ncu3.log 2>&1
ncu -i /tmp/r2_tap.ncu-rep --page raw --csv > gpurun_out/r2_tapgemm_tc_raw.csv
# prompt pass of 64 x 300 tokens: which kernel takes what share (launch list, durations only)
python tools/bs64_probe.py 64 300 1 4 > gpurun_out/r2_ncu_plain4.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn_prefill|w8_gemm_tc|act_prep|kv_write|attn_decode" -c 800 --csv --log-file gpurun_out/r2_launches_prefill_bs64.csv python tools/bs64_probe.py 64 300 1 4 > gpurun_out/r2_ncu4.log 2>&1
ls -la gpurun_out/ /tmp/*.ncu-rep; du -sh gpurun_out
