"""Drop-in for R2D2/ReplayMemory.py."""
from distributed_rl_b200.r2d2 import Replay  # noqa: F401
Replay_Server = Replay
