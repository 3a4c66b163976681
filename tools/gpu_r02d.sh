#!/bin/bash
# round 2, final session: the every-kernel-family case, then the GPU suite, smoke, the driver's bench command,
# its ncu launch list and one full capture of the dominant kernel (each ncu pass only after the same command exited 0 without it)
set -u
mkdir -p gpurun_out
TAG=r02h
# (compute-sanitizer is closed on this pool - "runs under it have left GPUs needing a reset" - so the small every-kernel-family case
# only runs plain; bad accesses are hunted with the kernels' own bounded waits / traps and the oracle comparisons)
python tools/sanitize_case.py > gpurun_out/${TAG}_sanitize_plain.log 2>&1; echo "plain rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c4_n1.json 2> gpurun_out/${TAG}_bench_c4_n1.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_c4.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_c4.log 2>&1
echo "launch list rc=$?"
