#!/bin/bash
# round 2, GPU call R (1 GPU): CTA shape x barrier placement re-checked on the final kernel (developer env hooks)
O=gpurun_out/r02r; mkdir -p $O
for blk in 128 96 64; do
  for bar in 1 0 2; do
    DRLOCO_B200_BLOCK=$blk DRLOCO_B200_STAGE_BARRIER=$bar timeout 100 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-extra > $O/b_${blk}_$bar.json 2> $O/b_${blk}_$bar.err
    python -c "import json;d=json.load(open('$O/b_${blk}_$bar.json'));print('block $blk barrier $bar value %.4g ser %.4g flushed %.4g kernel_ms %.4f' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['roofline']['kernel_ms']))"
  done
done
