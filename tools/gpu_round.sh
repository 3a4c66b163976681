#!/bin/bash
# One GPU-box session: parity tests, the bench line, then (only after those exit 0) the ncu launch list and one
# full capture of the dominant kernel.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench_n1.json
python bench.py --keyframes 6000 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --keyframes 6000 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
python tools/tc_probe.py 0 4000 > gpurun_out/${TAG}_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mxf4_match -s 3 -c 1 -f -o gpurun_out/${TAG}_mxf4 \
  python tools/tc_probe.py 0 4000 > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
cat gpurun_out/${TAG}_probe.log
python tools/l2_probe.py 400 > gpurun_out/${TAG}_l2probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:l2_match -s 1 -c 1 -f -o gpurun_out/${TAG}_l2 \
  python tools/l2_probe.py 400 > gpurun_out/${TAG}_ncu3.log 2>&1
echo "l2 capture rc=$?"
cat gpurun_out/${TAG}_l2probe.log
