"""oracle: CPU restatement of the reference's DeepMimic env-step path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (drloco_b200) never does; it fails loudly when its CUDA library is missing.

  physics.py      ctypes front-end of liboracle.so (walker_physics.c, float64 MuJoCo-pipeline restatement; PARITY UNPINNED)
  env_oracle.py   numpy float64 restatement of MimicEnv / ref_trajecs / Monitor / VecNormalize logic, pinned against
                  golden vectors produced by running the reference's own Python (tools/gen_golden.py)
"""
