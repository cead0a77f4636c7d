#!/bin/bash
# Runs the C++ mirror (tests + benchmarks) on NP GPUs of one box: NP=2 scripts/run_cpp_checks.sh
cd "$(dirname "$0")/.."
cd distributed-join_b200
TR="python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1"
echo "== single rank analytical"; timeout 120 bin/compare_against_analytical 2>&1 | tail -8
echo "== 2 rank analytical"; timeout 120 $TR --master-port 29601 bin/compare_against_analytical 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W09" | tail -10
echo "== 1 rank single-gpu comparison"; timeout 300 bin/compare_against_single_gpu 2>&1 | tail -22
echo "== $NP rank single-gpu comparison"; timeout 300 $TR --master-port 29606 bin/compare_against_single_gpu 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W09" | tail -22
echo "== 2 rank shuffle test"; timeout 120 $TR --master-port 29602 bin/test_shuffle_on 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W09" | tail -4
echo "== 2 rank join bench"; timeout 200 $TR --master-port 29603 bin/distributed_join --build-table-nrows 100000000 --probe-table-nrows 100000000 --nvlink-domain-size ${NP:-2} --iterations 3 --report-timing 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W09" | tail -24
echo "== 2 rank a2a"; timeout 200 $TR --master-port 29604 bin/all_to_all --max-size 4096000000 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W09" | tail -16
echo "== 2 rank shuffle_on bench"; timeout 200 $TR --master-port 29605 bin/shuffle_on --nrows 100000000 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|W09" | tail -5
