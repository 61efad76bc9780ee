#!/bin/bash
# round 2, GPU call G: v1 path removed + model block trimmed: tests; A/B of the median-torque history; serialized chain
O=gpurun_out/r02g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-extra > $O/bench_$name.json 2> $O/bench_$name.err
  python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name value %.4g ser %.4g flushed %.4g kernel_ms %.4f' % (d['value'],d['value_serialized'],d['value_l2_flushed'],d['roofline']['kernel_ms']))"
}
run default A=1
run nomedian DRLOCO_B200_MEDIAN_TORQUE=0
run default2 A=1
run nomedian2 DRLOCO_B200_MEDIAN_TORQUE=0
timeout 600 python tools/e2e_breakdown.py > $O/e2e_breakdown.json 2> $O/e2e_breakdown.err; head -14 $O/e2e_breakdown.json
