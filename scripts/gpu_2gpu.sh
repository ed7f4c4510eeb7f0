#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
tail -n 5 gpurun_out/bench_2gpu.err
cat gpurun_out/bench_2gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('n_gpus', d['n_gpus'], 'render %.1f Mrays/s e2e %.1f  train %.1f Mrays/s  nerf %.2f Mrays/s' % (d['value']/1e6, d['e2e']['value']/1e6, d['train']['value']/1e6, d['nerf']['value']/1e6))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 | tail -1 | cut -c1-200
