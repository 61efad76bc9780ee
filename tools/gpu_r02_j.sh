#!/bin/bash
# round 2, GPU call J (2 GPUs, strict timeouts): data-parallel training tool end to end (incl. the final evaluation on
# rank 0), bench clocks sampling under torchrun
O=gpurun_out/r02j; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29551 tools/train_ppo.py --envs 512 --steps 2000000 --seed 0 --out $O/ppo_2gpu_2M.json > $O/ppo_2gpu.log 2>&1; echo "train rc=$?"; grep -E "^done|evaluation \(" $O/ppo_2gpu.log | cut -c1-300
timeout 240 $TR --master-port 29552 bench.py --gpus 2 --steps 200 --warmup 20 > $O/bench2.json 2> $O/bench2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j/bench2.json'))
print('N=2 value %.4g serialized %.4g e2e %.4g kernel_ms %.4f' % (d['value'], d['value_serialized'], d['e2e']['value'], d['roofline']['kernel_ms']), d['clocks'])
PY
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
