#!/usr/bin/env bash
# Soak test of the opt-in exchange path: N ranks, K jobs back to back, each under its own limit.
# Usage (on a multi-GPU box): tools/exchange_soak.sh [GPUS=8] [JOBS=10] [GAP_SECONDS=1] [COMM_SMS=32]
# One line per job in gpurun_out/exchange_soak.log: job index, exit code (124 = hit the limit, i.e. hung), seconds, ms/step.
set -u
GPUS=${1:-8}; JOBS=${2:-10}; GAP=${3:-1}; SMS=${4:-32}
mkdir -p gpurun_out
for i in $(seq 1 "$JOBS"); do
  t0=$(date +%s)
  DAVF_NCCL_REGISTER=1 DAVF_NCCL_HIGH_PRIORITY=1 DAVF_COMM_SMS=$SMS DAVF_BENCH_WATCHDOG_S=150 timeout 200 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port $((29600 + i)) \
    bench.py --gpus "$GPUS" --steps 30 --warmup 5 --skip-eager > gpurun_out/soak_$i.json 2> gpurun_out/soak_$i.err
  rc=$?
  ms=$(python -c "import json,sys; print(json.loads(open('gpurun_out/soak_$i.json').read().strip().splitlines()[-1])['ms_per_step'])" 2>/dev/null || echo -)
  echo "job $i rc=$rc seconds=$(( $(date +%s) - t0 )) ms_per_step=$ms sms=$SMS gap=$GAP" | tee -a gpurun_out/exchange_soak.log
  sleep "$GAP"
done
