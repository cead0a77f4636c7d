#!/bin/bash
# One-box multi-GPU validation + numbers: NP=8 scripts/run_n8.sh   (writes gpurun_out/n${NP}_*.log)
cd "$(dirname "$0")/.."
NP=${NP:-8}
mkdir -p gpurun_out
TRP="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
F='^\*\|OMP_NUM\|^$\|W09'
echo "== parity"; timeout 300 $TRP --master-port 29701 tests/test_multi_gpu.py 2>&1 | grep -v "$F" | tail -8
echo "== timeline"; timeout 300 $TRP --master-port 29709 scripts/trace_timeline.py 2>&1 | grep "trace rank 0"
echo "== bench"; timeout 400 $TRP --master-port 29702 bench.py --gpus $NP --steps 5 --warmup 3 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/n${NP}_bench.json | cut -c1-3000
cd distributed-join_b200
TR="$TRP --no-python"
echo "== all_to_all (config 3)"; timeout 300 $TR --master-port 29703 bin/all_to_all --max-size 8192000000 2>&1 | grep -v "$F" | tee ../gpurun_out/n${NP}_all_to_all.log | tail -16
echo "== shuffle_on (config 4)"; timeout 300 $TR --master-port 29704 bin/shuffle_on --nrows ${SHUFFLE_ROWS:-400000000} --iterations 3 2>&1 | grep -v "$F" | tee ../gpurun_out/n${NP}_shuffle_on.log | tail -4
echo "== distributed_join (config 2 per-rank sizes)"; timeout 300 $TR --master-port 29705 bin/distributed_join --build-table-nrows 100000000 --probe-table-nrows 100000000 --nvlink-domain-size $NP --iterations 4 --report-timing 2>&1 | grep -v "$F" | grep "Rank 0\|Elasped\|benchmark" | tee ../gpurun_out/n${NP}_join_cfg2.log | tail -12
echo "== distributed_join (config 5: duplicates, sel 0.9, odf 4)"; timeout 400 $TR --master-port 29706 bin/distributed_join --build-table-nrows 100000000 --probe-table-nrows 400000000 --selectivity 0.9 --duplicate-build-keys --over-decomposition-factor 4 --nvlink-domain-size $NP --iterations 3 2>&1 | grep -v "$F" | grep "Elasped\|benchmark\|ERROR" | tee ../gpurun_out/n${NP}_join_cfg5.log | tail -6
