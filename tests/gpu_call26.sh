#!/bin/bash
# ncu with source counters on the LSTM launches of the warm step (token-level x5, frame-level x1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tests/profile_step.py f16 > gpurun_out/plain26.log 2>&1 &&
ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section SchedulerStats --section LaunchStats --section Occupancy \
    --clock-control none --import-source on -k regex:lstm_cluster -s 6 -c 6 -o /tmp/prof26 python tests/profile_step.py f16 > gpurun_out/ncu26.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu26.log
ncu -i /tmp/prof26.ncu-rep --page raw --csv > gpurun_out/prof26_raw.csv 2>/dev/null
for i in 1 5; do
  ncu -i /tmp/prof26.ncu-rep --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | gzip > gpurun_out/prof26_src_$i.csv.gz
done
ls -la gpurun_out/prof26*
