#!/bin/bash
# round 2, GPU call P (8 GPUs, strict timeout): the default bench line on the final tree
O=gpurun_out/r02p; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 > $O/bench8.json 2> $O/bench8.err; echo "rc=$?"
python -c "import json;d=json.load(open('$O/bench8.json'));print('N=8 value %.4g serialized %.4g flushed %.4g e2e %.4g' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value']), d['clocks']); print({k:(v['value'],v['value_serialized']) for k,v in d['extra'].items()})"
