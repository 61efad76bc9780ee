#!/bin/bash
# round 2, GPU call D (2 GPUs): tests, single-GPU bench with graph-replayed e2e, peer-exchange check, 2-GPU bench
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench1.json 2> $O/bench1.err; tail -3 $O/bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d/bench1.json'))
print('N=1 value %.4g serialized %.4g flushed %.4g e2e %.4g kernel_ms %.4f launches %d' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))
print({k:(v['value'],v['value_serialized']) for k,v in d['extra'].items()})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py 2>&1 | tail -8 | tee $O/multi_gpu_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 200 --warmup 20 > $O/bench2.json 2> $O/bench2.err; tail -3 $O/bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d/bench2.json'))
print('N=2 value %.4g serialized %.4g e2e %.4g kernel_ms %.4f exchange %s' % (d['value'], d['value_serialized'], d['e2e']['value'], d['roofline']['kernel_ms'], d['config']['statistics_exchange']))
print({k:(v['value'],v['value_serialized']) for k,v in d['extra'].items()})
PY
