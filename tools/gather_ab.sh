for g in overlap serial overlap serial; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $1 --steps 30 --warmup 5 --only-timed --gather $g > gpurun_out/r2w_$1gpu_$g.json 2>> gpurun_out/r2w_$1gpu.err; python - <<PY
import json
d=json.load(open('gpurun_out/r2w_$1gpu_$g.json')); print('$1 gpus', '$g', round(d['ms_per_step'],3), [round(x,3) for x in d['per_rank_solve_ms']])
PY
done
