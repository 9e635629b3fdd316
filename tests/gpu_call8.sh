#!/bin/bash
# ncu with source counters on stage-0 k7/k11 launches (old kernel) and the first stage-1 launches (weights-as-A kernel), warm step;
# the report is turned into CSV pages on the box (the .ncu-rep itself is too large to bring back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tests/profile_step.py f16 > gpurun_out/plain8.log 2>&1 &&
ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section SchedulerStats --section LaunchStats --section Occupancy \
    --clock-control none --import-source on -k regex:conv_umma -s 249 -c 22 -o /tmp/prof8 python tests/profile_step.py f16 > gpurun_out/ncu8.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu8.log; ls -la /tmp/prof8.ncu-rep
ncu -i /tmp/prof8.ncu-rep --page raw --csv > gpurun_out/prof8_raw.csv 2>/dev/null
for i in 0 1 6 7 13 14 20 21; do
  ncu -i /tmp/prof8.ncu-rep --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | gzip > gpurun_out/prof8_src_$i.csv.gz
done
ls -la gpurun_out/prof8*
