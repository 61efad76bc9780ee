"""Developer sweep (GPU box): forward-pass parity, a short rollout, and step time for several CTA shapes / batch sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_check import speed, rollout, forward_pieces
forward_pieces()
rollout(steps=12)
for blk in (64,128):
    speed(4096, block=blk)
speed(65536, block=64); speed(65536, block=128); speed(16384, block=64); speed(16384, block=128); speed(8192, block=64); speed(8192, block=128)
speed(4096, integrator="euler")
