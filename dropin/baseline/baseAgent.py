"""Drop-in for baseline/baseAgent.py (node types of the shipped configs only)."""
from distributed_rl_b200.agent import GraphAgent as baseAgent  # noqa: F401
