#!/bin/bash
# round 2, GPU call A: MuJoCo availability probe, parity tests for both generations of the dynamics evaluation,
# A/B bench, launch list + one full ncu capture of the new kernel.  Everything is logged under gpurun_out/r02a/.
O=gpurun_out/r02a; mkdir -p $O
{ echo "== python -m pip install mujoco (the box has no network) =="; timeout 60 python -m pip install mujoco 2>&1 | tail -4;
  echo "== offline wheelhouse =="; timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse mujoco 2>&1 | tail -3;
  echo "== import =="; python -c "import mujoco" 2>&1 | tail -1; python -c "import mujoco_py" 2>&1 | tail -1; } > $O/mujoco_probe.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
echo "== FD=2 tests =="; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 | tee $O/tests_fd2.log
echo "== FD=2 all tests (no -x) =="; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/tests_fd2_all.log
echo "== FD=1 tests =="; DRLOCO_B200_FD=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/tests_fd1.log; tail -3 $O/tests_fd1.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/smoke.log
for v in 2 1; do
  DRLOCO_B200_FD=$v timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_fd$v.json 2> $O/bench_fd$v.err
  python -c "import json;d=json.load(open('$O/bench_fd$v.json'));print('FD=$v value',d['value'],'kernel_ms',d['roofline']['kernel_ms'],d['episode_stats'])"
done
DRLOCO_B200_STAGE_BARRIER=0 timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_fd2_nobar.json 2> $O/bench_fd2_nobar.err
python -c "import json;d=json.load(open('$O/bench_fd2_nobar.json'));print('FD=2 nobarrier value',d['value'],'kernel_ms',d['roofline']['kernel_ms'])"
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --envs-per-gpu 65536 > $O/bench_fd2_65536.json 2> $O/bench_fd2_65536.err
python -c "import json;d=json.load(open('$O/bench_fd2_65536.json'));print('FD=2 65536 value',d['value'],'kernel_ms',d['roofline']['kernel_ms'])"
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e --env-id MimicWalker165cm65kg --envs-per-gpu 16384 > $O/bench_fd2_w165.json 2> $O/bench_fd2_w165.err
python -c "import json;d=json.load(open('$O/bench_fd2_w165.json'));print('FD=2 w165 value',d['value'],'kernel_ms',d['roofline']['kernel_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 45 --csv --log-file $O/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mimic_step -s 6 -c 1 -o $O/prof_fd2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> $O/ncu.err
ls -la $O
