"""Drop-in for baseline/PER.py."""
from distributed_rl_b200.per import PER  # noqa: F401
