#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun from the repo root): launch list of two 10 M-event steps, one
# `--set full` capture of the first launch of every parse / emit kernel at 4 M events, exported as CSV (the .ncu-rep
# stays in /tmp: it is larger than what gpurun_out/ may carry).  Numbers printed by runs under ncu are never bench values.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python tools/profile_step.py 10000000 2 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.csv
gzip -f gpurun_out/r02_launches.csv
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:k_parse<\(int\)[012]|k_emit_fused<\(int\)[012]' -c 12 -o /tmp/r02_full -f \
    python tools/profile_step.py 4000000 1 > gpurun_out/r02_ncu_capture.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > /tmp/r02_ncu_full_raw.csv
python tools/ncu_summary.py traffic /tmp/r02_ncu_full_raw.csv 4000000 > gpurun_out/r02_ncu_traffic.json
gzip -c /tmp/r02_ncu_full_raw.csv > gpurun_out/r02_ncu_full_raw.csv.gz
head -c 600 gpurun_out/r02_ncu_traffic.json
