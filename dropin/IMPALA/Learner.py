"""Drop-in for IMPALA/Learner.py."""
from distributed_rl_b200.impala import ImpalaConfig, Learner as _Learner


class Learner(_Learner):
    def __init__(self):
        from APE_X.ReplayMemory import _connect
        cfg = ImpalaConfig.from_configuration()
        super().__init__(cfg, connect=_connect(cfg.REDIS_SERVER))
