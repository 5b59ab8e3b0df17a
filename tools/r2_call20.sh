#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/check_time_shard.py 2>&1 | grep -v "Warn\|OMP\|\*\*\*" | tail -6 > gpurun_out/r2c20_timeshard.log
cat gpurun_out/r2c20_timeshard.log
timeout 300 python -m pytest tests/test_gpu_train_loop.py -m gpu -q --timeout 200 2>&1 | tail -3
