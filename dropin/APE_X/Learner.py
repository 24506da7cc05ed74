"""Drop-in for APE_X/Learner.py: `Learner()` (no arguments) + `.run()`, as
run_learner.py:15-18 uses it, backed by distributed_rl_b200.apex."""
from distributed_rl_b200.apex import ApexConfig, Learner as _Learner
from APE_X.ReplayMemory import Replay, _connect  # noqa: F401  (same import the reference has, :4)


class Learner(_Learner):
    def __init__(self):
        cfg = ApexConfig.from_configuration()
        super().__init__(cfg, connect=_connect(cfg.REDIS_SERVER))
