#!/bin/bash
# The headline bench line at N GPUs of one box, as the driver launches it (gpurun --gpus N -- bash tools/scale_lines.sh N).
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2y_batch64k_${N}gpu.json 2> gpurun_out/r2y_batch64k_${N}gpu.err; echo rc=$?
