#!/bin/bash
# round 2, GPU call H: final single-GPU validation (median opt-in path, v1 removed), default bench line, ncu capture
O=gpurun_out/r02h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 1200 python bench.py > $O/bench1.json 2> $O/bench1.err; tail -3 $O/bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h/bench1.json'))
print('N=1 value %.4g serialized %.4g flushed %.4g e2e %.4g kernel_ms %.4f launches %d' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))
print({k:(v['value'],v['value_serialized']) for k,v in d['extra'].items()}, d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; python -c "import json;d=json.load(open('$O/bench_ref.json'));print('reference arm', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['kind'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mimic_step -s 6 -c 1 -o $O/prof python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2> $O/ncu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
