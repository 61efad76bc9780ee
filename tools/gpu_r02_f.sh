#!/bin/bash
# round 2, GPU call F: bank-conflict layout, median torque, single-copy e2e; PPO learning-curve comparison leg
O=gpurun_out/r02f; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench1.json 2> $O/bench1.err; tail -3 $O/bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f/bench1.json'))
print('N=1 value %.4g serialized %.4g flushed %.4g e2e %.4g kernel_ms %.4f launches %d' % (d['value'], d['value_serialized'], d['value_l2_flushed'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches']))
print({k:(v['value'],v['value_serialized']) for k,v in d['extra'].items()}, d['e2e'])
PY
timeout 600 python tools/e2e_breakdown.py > $O/e2e_breakdown.json 2> $O/e2e_breakdown.err; head -14 $O/e2e_breakdown.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mimic_step -s 6 -c 1 -o $O/prof python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2> $O/ncu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
# the learner of configs[4] with the reference's rollout shape (8 envs x 2048 steps per update), GPU env: the curve the
# CPU-oracle run of the same learner (profiles/r02_ppo_cpu_oracle_2M.json) is compared with
timeout 1500 python tools/train_ppo.py --envs 8 --steps 2000000 --reference-hypers --seed 0 --no-eval --out $O/ppo_b200_8env_2M.json > $O/ppo_b200_8env.log 2>&1; tail -2 $O/ppo_b200_8env.log
