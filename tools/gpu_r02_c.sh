#!/bin/bash
# round 2, GPU call C: statistics epilogue + fused VecNormalize exchange kernel + host path; register-relief kernel
O=gpurun_out/r02c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 | tee $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c/bench.json'))
print('value %.4g serialized %.4g flushed %.4g e2e %.4g kernel_ms %.4f launches %d' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))
print(d['e2e']); print(d['extra']); print(d['episode_stats'])
PY
