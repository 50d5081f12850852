#!/bin/bash
# ncu evidence for the round (run under gpurun, 1 GPU): (1) launch list of a short bench
# run, (2) full captures of the SpMV kernel, the triangular-sweep kernel and the AMG smoother.
# Each ncu pass only runs after the same command exited 0 without ncu.
mkdir -p gpurun_out
GRID=${GRID:-256}
CMD="python bench.py --grid ${GRID} --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-2000} -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit: $?"
$CMD > gpurun_out/prof_plain2.json 2>> gpurun_out/prof_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 20 -c 1 \
    -o gpurun_out/prof_spmv -f $CMD > gpurun_out/ncu_spmv.log 2>&1
echo "spmv capture exit: $?"
$CMD > gpurun_out/prof_plain3.json 2>> gpurun_out/prof_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:tri_ -s 20 -c 1 \
    -o gpurun_out/prof_tri -f $CMD > gpurun_out/ncu_tri.log 2>&1
echo "tri capture exit: $?"
AMG="python scripts/amg_prof.py ${AMG_GRID:-128} 2"
$AMG > gpurun_out/amg_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gs_ -s 28 -c 8 \
    -o gpurun_out/prof_gs -f $AMG > gpurun_out/ncu_gs.log 2>&1
echo "gs capture exit: $?"
ls -la gpurun_out/ | tail -30
