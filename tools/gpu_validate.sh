#!/bin/bash
# One-GPU validation pass run under gpurun: GPU tests, the bench line, the ncu launch list of the bench command and
# --set full captures of the step's largest kernels.  Outputs go to gpurun_out/ (tag = $1).
TAG=${1:-r02}
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/${TAG}_gputest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
STEP="python tools/profile_step.py --spp 64 --reps 1"
$STEP > gpurun_out/${TAG}_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_raygen|k_trace_rec' -c 6 -o gpurun_out/${TAG}_prof_step -f $STEP > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -5 gpurun_out/${TAG}_gputest.log; cat gpurun_out/${TAG}_bench.json | cut -c1-3000; tail -3 gpurun_out/${TAG}_bench.err
