#!/bin/bash
# round 2, GPU call O (1 GPU): final tree - GPU tests, smoke, default bench, e2e breakdown, reference arm
O=gpurun_out/r02o; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/bench1.json'));print('N=1 value %.4g serialized %.4g flushed %.4g e2e %.4g kernel_ms %.4f' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value'], d['roofline']['kernel_ms']), d['clocks'], d['cpu_baseline']['value'])"
timeout 100 python tools/e2e_modes.py > $O/e2e_modes_1gpu.json 2> $O/e2e_modes_1gpu.err; cat $O/e2e_modes_1gpu.json
timeout 100 python tools/e2e_breakdown.py > $O/e2e_breakdown.json 2> $O/e2e_breakdown.err; tail -c 600 $O/e2e_breakdown.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 400 $O/bench_ref.json
