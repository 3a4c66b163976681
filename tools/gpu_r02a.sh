#!/bin/bash
# round 2, first GPU session: parity tests (incl. the new robustness / frame tests), frame latency, default bench
set -u
mkdir -p gpurun_out build
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest.log
tail -25 gpurun_out/r02a_pytest.log
gcc -O2 -I include tools/frame_latency.c slam_loop_closing_b200/csrc/liblc_b200.so -Wl,-rpath,$PWD/slam_loop_closing_b200/csrc -o build/frame_latency \
  && ./build/frame_latency 97 2000 > gpurun_out/r02a_frame_latency.txt 2>&1; echo "frame_latency rc=$?"; cat gpurun_out/r02a_frame_latency.txt
gcc -O2 -I include tools/loop_latency.c slam_loop_closing_b200/csrc/liblc_b200.so -Wl,-rpath,$PWD/slam_loop_closing_b200/csrc -o build/loop_latency \
  && ./build/loop_latency 200 > gpurun_out/r02a_loop_latency.txt 2>&1; cat gpurun_out/r02a_loop_latency.txt
gcc -O2 -I include tools/pair_latency.c slam_loop_closing_b200/csrc/liblc_b200.so -Wl,-rpath,$PWD/slam_loop_closing_b200/csrc -o build/pair_latency \
  && ./build/pair_latency > gpurun_out/r02a_pair_latency.txt 2>&1; cat gpurun_out/r02a_pair_latency.txt
python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r02a_bench_n1.json; tail -5 gpurun_out/r02a_bench_n1.err
