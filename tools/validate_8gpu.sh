#!/bin/bash
# Round-2 validation on an 8-GPU box (gpurun --gpus 8 -- bash tools/validate_8gpu.sh): the C++ caller of mpc_solve_batch_multi on
# 8 GPUs (64K problems per GPU) with pageable and with page-locked caller arrays, and the headline bench's timed region with
# NCCL's gather kernels on fewer channels (do they get in the way of the persistent grid of the next solve?).
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
g++ -O2 -std=c++11 -I include -I /usr/local/cuda/include -o /tmp/test_multi_gpu tests/cpp/test_multi_gpu.cpp -L carnd-mpc-project_b200 -lmpc_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/carnd-mpc-project_b200 -Wl,-rpath,/usr/local/cuda/lib64 && python -c "
import json,sys
import importlib.util
spec = importlib.util.spec_from_file_location('wl', 'carnd-mpc-project_b200/workloads.py'); wl = importlib.util.module_from_spec(spec); spec.loader.exec_module(wl)
json.dump(wl.reference_data()['configs']['stable'], open('/tmp/config-stable.json','w'))
" && { /tmp/test_multi_gpu /tmp/config-stable.json 524288 8; /tmp/test_multi_gpu /tmp/config-stable.json 524288 8 pinned; } > gpurun_out/r2t_cpp_multi8.log 2>&1; echo cpp rc=$?
for ch in default 1 2; do
  if [ $ch = default ]; then unset NCCL_MAX_NCHANNELS NCCL_MIN_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch NCCL_MIN_NCHANNELS=$ch; fi
  $TR bench.py --gpus 8 --steps 30 --warmup 5 --only-timed > gpurun_out/r2t_b8_ch$ch.json 2> gpurun_out/r2t_b8_ch$ch.err; echo ch=$ch rc=$?
done
MPC_BENCH_NO_GATHER=1 $TR bench.py --gpus 8 --steps 30 --warmup 5 --only-timed > gpurun_out/r2t_b8_nogather.json 2> gpurun_out/r2t_b8_nogather.err; echo nogather rc=$?
