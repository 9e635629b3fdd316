#!/bin/bash
# the driver's N>1 launch of both arms (weak scaling: a replica of configs[1] per rank), N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 3 \
  > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
echo "own arm exit $?"; tail -1 gpurun_out/scale_n$N.json | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 1 \
  > gpurun_out/scale_ref_n$N.json 2> gpurun_out/scale_ref_n$N.err
echo "reference arm exit $?"; grep '^{' gpurun_out/scale_ref_n$N.json | tail -1 | cut -c1-330
