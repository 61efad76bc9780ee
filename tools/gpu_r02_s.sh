#!/bin/bash
# round 2, GPU call S (1 GPU): 256-thread CTAs (two per SM; developer build -DDRL_STEP_MAX_BLOCK=256) vs the shipped 128
O=gpurun_out/r02s; mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 100 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline > $O/$name.json 2> $O/$name.err
  python -c "import json;d=json.load(open('$O/$name.json'));x=d['extra'];print('$name value %.4g ser %.4g flushed %.4g kernel_ms %.4f | cfg2 %.4g | cfg3 %.4g' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['roofline']['kernel_ms'], x['configs[2]']['value_serialized'], x['configs[3]']['value_serialized']))"
}
B=$PWD/build_ab/libdrloco_b200_b256.so
run shipped_128 A=1
run b256_256 DRLOCO_B200_LIB=$B DRLOCO_B200_BLOCK=256
run b256_128 DRLOCO_B200_LIB=$B DRLOCO_B200_BLOCK=128
run shipped_128_again A=1
run b256_256_again DRLOCO_B200_LIB=$B DRLOCO_B200_BLOCK=256
