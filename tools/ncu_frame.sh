#!/bin/bash
# ncu --set full of the launches of ONE 4K frame (the second of two), raw metrics to CSV for tools/ncu_summarise.py
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv3x3|autoexposure|input_process' -s 17 -c 17 \
  -f -o gpurun_out/frame_ncu python tools/one_frame.py > gpurun_out/frame_ncu.log 2>&1
tail -3 gpurun_out/frame_ncu.log
ncu -i gpurun_out/frame_ncu.ncu-rep --page raw --csv > gpurun_out/frame_ncu_raw.csv 2>/dev/null
wc -l gpurun_out/frame_ncu_raw.csv; ls -la gpurun_out/frame_ncu.ncu-rep
