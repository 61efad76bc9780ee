#!/bin/bash
# round 2, GPU call T (1 GPU): final tree - GPU tests, smoke, default bench line, ncu full capture + launch list
O=gpurun_out/r02t; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
timeout 400 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/bench1.json'));print('N=1 value %.4g serialized %.4g flushed %.4g e2e %.4g kernel_ms %.4f' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value'], d['roofline']['kernel_ms']), d['clocks']); print({k:(v['value'],v['value_serialized']) for k,v in d['extra'].items()})"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mimic_step -s 6 -c 1 -o $O/prof python bench.py --steps 5 --warmup 3 --pre-warmup 0 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2> $O/ncu.err; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches.csv python bench.py --steps 12 --warmup 3 --pre-warmup 0 --no-cpu-baseline --no-extra > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 100 python tools/e2e_breakdown.py > $O/e2e_breakdown.json 2> $O/e2e_breakdown.err
