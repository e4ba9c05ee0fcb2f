#!/bin/bash
# Round-2 ncu evidence for the kernels changed late in the round (cluster split-K GEMM, fp16 tap-GEMM), run on the GPU box.
# Reports stay in /tmp; only raw-page CSVs go to gpurun_out/.  Every ncu run follows a plain run of the same command.
set -x
S="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --metrics sm__inst_executed_pipe_tc.sum,l1tex__m_xbar2l1tex_read_bytes_pipe_tma.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum"
python tools/gemm_ncu.py > gpurun_out/r2b_plain2.log 2>&1 && \
  ncu $S --clock-control none -k regex:w8_gemm_tc -c 27 -o /tmp/r2b_w8 -f python tools/gemm_ncu.py > gpurun_out/r2b_ncu2.log 2>&1
ncu -i /tmp/r2b_w8.ncu-rep --page raw --csv > gpurun_out/r2b_w8_gemm_tc_raw.csv
python tools/codec_probe.py 32 96 > gpurun_out/r2b_plain3.log 2>&1 && \
  ncu $S --clock-control none -k regex:"tapgemm_tc|conv_out_clamp" -s 73 -c 73 -o /tmp/r2b_tap -f python tools/codec_probe.py 32 96 > gpurun_out/r2b_ncu3.log 2>&1
ncu -i /tmp/r2b_tap.ncu-rep --page raw --csv > gpurun_out/r2b_tapgemm_tc_raw.csv
du -sh gpurun_out
