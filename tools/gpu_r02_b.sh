#!/bin/bash
# round 2, GPU call B: parity tests on the second optimisation batch + CTA-size / barrier sweep
O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
run() { # name, env assignments...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_$name.json 2> $O/bench_$name.err
  python -c "import json;d=json.load(open('$O/bench_$name.json'));print('$name value %.4g kernel_ms %.4f flushed %.4g' % (d['value'],d['roofline']['kernel_ms'],d['value_l2_flushed']))"
}
run b128 DRLOCO_B200_BLOCK=128
run b128_nobar DRLOCO_B200_BLOCK=128 DRLOCO_B200_STAGE_BARRIER=0
run b64 DRLOCO_B200_BLOCK=64
run b64_nobar DRLOCO_B200_BLOCK=64 DRLOCO_B200_STAGE_BARRIER=0
run b32 DRLOCO_B200_BLOCK=32 DRLOCO_B200_STAGE_BARRIER=0
run b96 DRLOCO_B200_BLOCK=96
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --envs-per-gpu 65536 > $O/bench_65536.json 2> $O/bench_65536.err
python -c "import json;d=json.load(open('$O/bench_65536.json'));print('65536 value %.4g kernel_ms %.4f' % (d['value'],d['roofline']['kernel_ms']))"
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --envs-per-gpu 8192 > $O/bench_8192.json 2> $O/bench_8192.err
python -c "import json;d=json.load(open('$O/bench_8192.json'));print('8192 value %.4g kernel_ms %.4f' % (d['value'],d['roofline']['kernel_ms']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mimic_step -s 6 -c 1 -o $O/prof python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu.err
