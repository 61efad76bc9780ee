#!/bin/bash
# round 2, GPU call L (8 GPUs, strict timeouts): serialised throughput with the statistics exchange every 4th step
O=gpurun_out/r02l; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29561 bench.py --gpus 8 --steps 200 --warmup 20 --no-extra --no-e2e --stats-sync-every 4 > $O/bench8_k4.json 2> $O/bench8_k4.err; echo "rc=$?"
python -c "import json;d=json.load(open('$O/bench8_k4.json'));print('N=8 K=4 value %.4g serialized %.4g' % (d['value'], d['value_serialized']), d['clocks'])"
timeout 150 $TR --master-port 29562 bench.py --gpus 8 --steps 200 --warmup 20 --no-extra > $O/bench8.json 2> $O/bench8.err; echo "rc=$?"
python -c "import json;d=json.load(open('$O/bench8.json'));print('N=8 K=1 value %.4g serialized %.4g e2e %.4g' % (d['value'], d['value_serialized'], d['e2e']['value']), d['clocks'])"
