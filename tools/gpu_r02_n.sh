#!/bin/bash
# round 2, GPU call N (8 GPUs, strict timeouts): numpy-API e2e, mapped host outputs vs one D2H copy, per-rank times
O=gpurun_out/r02n; mkdir -p $O
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 tools/e2e_modes.py > $O/e2e_modes_8gpu.json 2> $O/e2e_modes_8gpu.err; echo "rc=$?"
cat $O/e2e_modes_8gpu.json
nproc; nvidia-smi topo -m > $O/topo8.txt 2>&1
