#!/bin/bash
# ncu evidence of the current binary: (1) launch list of one warm step, (2) DRAM bytes / tensor-pipe share / time of every conv launch of that
# step, (3) --set full of 42 resblock launches of the weights-as-A kernel in the warm step (stage 0: k3 x6, k7 x6, k11 x6; stage 1: noise_res k11 x6,
# k3 x6, k7 x6, k11 x6).  Reports are turned into CSV
# on the box (the .ncu-rep files are too large to bring back).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tests/profile_step.py f16"
$CMD > gpurun_out/plain_final.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_final.log; exit 1; }
cat gpurun_out/plain_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 293 -c 293 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list exit $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
    --clock-control none -k regex:conv_umma -s 144 -c 144 --csv --log-file gpurun_out/r02_conv_launches.csv $CMD > gpurun_out/ncu_c.log 2>&1
echo "conv metrics exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv_umma_wa -s 54 -c 42 -o /tmp/prof_wa $CMD > gpurun_out/ncu_f.log 2>&1
echo "set full exit $?"
ncu -i /tmp/prof_wa.ncu-rep --page raw --csv > gpurun_out/r02_wa_full_raw.csv 2>/dev/null
ncu -i /tmp/prof_wa.ncu-rep --page details --csv 2>/dev/null | gzip > gpurun_out/r02_wa_full_details.csv.gz
ls -la gpurun_out/r02_* /tmp/prof_wa.ncu-rep
