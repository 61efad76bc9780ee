#!/bin/bash
# round 2, GPU call E: tests; barrier placement A/B on the fast-path kernel; e2e breakdown; ncu capture
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-extra > $O/bench_$name.json 2> $O/bench_$name.err
  python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name value %.4g ser %.4g kernel_ms %.4f' % (d['value'],d['value_serialized'],d['roofline']['kernel_ms']))"
}
run bar1 DRLOCO_B200_STAGE_BARRIER=1
run bar2 DRLOCO_B200_STAGE_BARRIER=2
run bar0 DRLOCO_B200_STAGE_BARRIER=0
run bar2_b64 DRLOCO_B200_STAGE_BARRIER=2 DRLOCO_B200_BLOCK=64
run bar1_again DRLOCO_B200_STAGE_BARRIER=1
timeout 600 python tools/e2e_breakdown.py > $O/e2e_breakdown.json 2> $O/e2e_breakdown.err; cat $O/e2e_breakdown.json | head -40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mimic_step -s 6 -c 1 -o $O/prof python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2> $O/ncu.err
DRLOCO_B200_STAGE_BARRIER=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mimic_step -s 6 -c 1 -o $O/prof_bar2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>> $O/ncu.err
