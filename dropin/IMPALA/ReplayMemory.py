"""Drop-in for IMPALA/ReplayMemory.py."""
from distributed_rl_b200.impala import Replay  # noqa: F401
