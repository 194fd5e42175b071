#!/bin/bash
# BASELINE.json configs at 8 GPUs (run through `gpurun --gpus 8`): Pushing 8 x 4096 (headline, weak scaling), Stacking 32 768 envs
# (config 4), the seven-config mix 65 536 envs (config 5).  One JSON line each into gpurun_out/r2_configs/.
mkdir -p gpurun_out/r2_configs
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps $2 --warmup 10 $3 > gpurun_out/r2_configs/$4.json 2> gpurun_out/r2_configs/$4.err; tail -c 600 gpurun_out/r2_configs/$4.json | head -c 300; echo; }
run 29511 100 "" pushing_n8
run 29512 100 "--workload stacking" stacking_n8
run 29513 60 "--workload mixed7" mixed7_n8
