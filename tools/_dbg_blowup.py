import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drloco_b200.vec_env import B200MimicVecEnv
def P(*a):
    print(*a, flush=True)
which = sys.argv[1]
n = 16
env = B200MimicVecEnv("StraightMimicWalker", num_envs=n, seed=5)
env.reset()
q, v, c = env.get_state()
P("reset ok")
if which == "vel":
    v[3, 0] = 3e11
elif which == "nan":
    q[5, 7] = np.nan
elif which == "eval":
    env.env_method("activate_evaluation")
    for k in range(3):
        env.reset(); torch.cuda.synchronize(); P("eval reset", k, env.get_state()[2][:2])
    sys.exit(0)
env.set_state(q, v, c)
P("state set")
obs, rew, done, infos = env.step(np.zeros((n, 8), np.float32))
torch.cuda.synchronize()
P("step ok", done, rew[:8])
P(env.stats()["blowups"])
