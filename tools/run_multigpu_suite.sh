#!/bin/bash
# tools/run_multigpu_suite.sh -- everything that needs more than one GPU, in ONE gpurun call (box time is charged
# per GPU): the device-set tests on distinct devices, the library's own multi-GPU path, the 8-link PCIe ceiling,
# and bench.py across processes.  Usage (from the repo root, on the box): [QUICK=1] bash tools/run_multigpu_suite.sh <N>
# (QUICK skips the PCIe probe and the PDL-off bench run)
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2_mg_gpus.txt
nvidia-smi topo -m > $OUT/r2_mg_topo.txt 2>&1
echo "== device-set tests on $N GPUs"
timeout 600 python -m pytest tests/test_gpu_runtime.py tests/test_dropin_cpp.py -m gpu -q > $OUT/r2_mg_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/r2_mg_pytest.log
echo "== library device set (one process, $N GPUs)"
timeout 900 python tools/bench_multidev.py --devices all --reps 10 > $OUT/r2_multidev_n$N.jsonl 2> $OUT/r2_multidev_n$N.err; echo "multidev rc=$?"; tail -3 $OUT/r2_multidev_n$N.err
echo "== PCIe ceiling, $N ranks at once"
[ -n "$QUICK" ] || timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/pcie_probe.py --bind > $OUT/r2_pcie_probe_n$N.json 2> $OUT/r2_pcie_probe_n$N.err; echo "probe rc=$?"; tail -2 $OUT/r2_pcie_probe_n$N.err
echo "== bench.py, $N ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err; echo "bench rc=$?"; tail -3 $OUT/r2_bench_n$N.err
[ -n "$QUICK" ] || timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 3 --pdl 0 --no-e2e > $OUT/r2_bench_n${N}_pdl0.json 2> $OUT/r2_bench_n${N}_pdl0.err; echo "bench pdl0 rc=$?"
if [ "$N" = "8" ]; then
  for M in 2 4; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 2953$M bench.py --gpus $M --steps 20 --warmup 3 --no-e2e > $OUT/r2_bench_n$M.json 2> $OUT/r2_bench_n$M.err; echo "bench n=$M rc=$?"
  done
fi
echo done
