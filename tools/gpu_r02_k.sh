#!/bin/bash
# round 2, GPU call K: zero-copy host outputs of the numpy API
O=gpurun_out/r02k; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $O/tests.log
timeout 300 python tools/e2e_breakdown.py > $O/e2e_breakdown.json 2> $O/e2e_breakdown.err; cat $O/e2e_breakdown.json; tail -3 $O/e2e_breakdown.err
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-extra > $O/bench1.json 2> $O/bench1.err; tail -3 $O/bench1.err
python -c "
import json
d=json.load(open('$O/bench1.json'))
print('N=1 value %.4g serialized %.4g e2e %.4g kernel_ms %.4f' % (d['value'], d['value_serialized'], d['e2e']['value'], d['roofline']['kernel_ms']), d['e2e'], d['clocks'])"
