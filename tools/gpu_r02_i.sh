#!/bin/bash
# round 2, GPU call I (8 GPUs): statistics exchange + data-parallel learner checks, the 8-GPU bench line, configs[4] run
O=gpurun_out/r02i; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/multi_gpu_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -4 | tee $O/multi_gpu_check.log
timeout 900 $TR --master-port 29542 bench.py --gpus 8 --steps 200 --warmup 20 > $O/bench8.json 2> $O/bench8.err; tail -2 $O/bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i/bench8.json'))
print('N=8 value %.4g serialized %.4g flushed %.4g e2e %.4g kernel_ms %.4f exchange %s' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value'], d['roofline']['kernel_ms'], d['config']['statistics_exchange']))
print({k:(v['value'],v['value_serialized']) for k,v in d['extra'].items()})
PY
timeout 900 $TR --master-port 29543 tools/train_ppo.py --envs 512 --steps 16000000 --seed 0 --out $O/ppo_8gpu_16M.json > $O/ppo_8gpu.log 2>&1; tail -4 $O/ppo_8gpu.log
timeout 600 $TR --master-port 29544 bench.py --gpus 8 --steps 200 --warmup 20 --envs-per-gpu 8192 --no-extra --no-e2e > $O/bench8_8192.json 2> $O/bench8_8192.err
python -c "import json;d=json.load(open('gpurun_out/r02i/bench8_8192.json'));print('N=8 x 8192 value %.4g serialized %.4g' % (d['value'], d['value_serialized']))"
