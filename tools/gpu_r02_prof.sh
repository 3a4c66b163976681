#!/bin/bash
# round 2 profiling session (one GPU): the bench line, then - each only after the same command exited 0 without ncu -
# the ncu launch list of the bench command and one full capture of the dominant kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
python bench.py > gpurun_out/${TAG}_bench_c4_n1.json 2> gpurun_out/${TAG}_bench_c4_n1.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c4.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_c4.log 2>&1
echo "launch list rc=$?"
python tools/tc_probe.py 0 4000 > gpurun_out/${TAG}_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mxf4_match -s 3 -c 1 -f -o gpurun_out/${TAG}_mxf4 \
  python tools/tc_probe.py 0 4000 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
cat gpurun_out/${TAG}_probe.log
