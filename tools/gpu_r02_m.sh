#!/bin/bash
# round 2, GPU call M (2 GPUs): numpy-API e2e with mapped host outputs vs one D2H copy, at N=1 and N=2; GPU tests
O=gpurun_out/r02m; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ho in mapped copy; do
  CUDA_VISIBLE_DEVICES=0 timeout 120 python bench.py --steps 200 --warmup 20 --no-extra --host-outputs $ho > $O/bench1_$ho.json 2> $O/bench1_$ho.err
  python -c "import json;d=json.load(open('$O/bench1_$ho.json'));print('N=1 $ho value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$((RANDOM%10)) bench.py --gpus 2 --steps 200 --warmup 20 --no-extra --host-outputs $ho > $O/bench2_$ho.json 2> $O/bench2_$ho.err
  python -c "import json;d=json.load(open('$O/bench2_$ho.json'));print('N=2 $ho value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
done
nproc; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14 > $O/topo.txt
