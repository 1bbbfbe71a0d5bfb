#!/bin/bash
# First GPU call of the next session: confirm the kernels / paths written after the round-1 GPU
# budget was spent, and time them.  Run from the repo root on a B200:  bash tests/next_session.sh
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 400 python tests/gpu_probe.py dropout_bf16 gru_cluster_exact 2>&1 | tail -6
cp gpurun_out/probe.json gpurun_out/probe_next_session.json 2>/dev/null
timeout 300 python tests/bench_trainer.py 256 10 > gpurun_out/bench_trainer.jsonl 2> gpurun_out/bench_trainer.err
tail -3 gpurun_out/bench_trainer.jsonl
timeout 400 python tests/bench_configs.py > gpurun_out/configs_l2_gru.jsonl 2> gpurun_out/configs_l2_gru.err
M3T_GRU_CLUSTER=1 timeout 400 python tests/bench_configs.py > gpurun_out/configs_cluster_gru.jsonl 2> gpurun_out/configs_cluster_gru.err
grep '"config": 5\|"config": 1' gpurun_out/configs_l2_gru.jsonl gpurun_out/configs_cluster_gru.jsonl | cut -c1-260
