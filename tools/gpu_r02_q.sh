#!/bin/bash
# round 2, GPU call Q (1 GPU): two-level statistics sum vs the single last-block pass (previous build in build_ab/)
O=gpurun_out/r02q; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
  for v in new prev; do
    if [ $v = prev ]; then export DRLOCO_B200_LIB=$PWD/build_ab/libdrloco_b200_prev.so; else unset DRLOCO_B200_LIB; fi
    timeout 200 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err
    python -c "import json;d=json.load(open('$O/bench_${v}_$rep.json'));x=d['extra'];print('$v $rep value %.4g ser %.4g flushed %.4g kernel_ms %.4f | cfg2 %.4g kernel %.4f | cfg3 %.4g' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['roofline']['kernel_ms'], x['configs[2]']['value_serialized'], x['configs[2]']['kernel_ms'], x['configs[3]']['value_serialized']))"
  done
done
