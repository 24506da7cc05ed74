"""Drop-in for APE_X/ReplayMemory.py: `Replay()` thread with sample/update/lock."""
from distributed_rl_b200.apex import ApexConfig, Replay as _Replay


def _connect(host):
    """redis.StrictRedis(host, 6379) like APE_X/ReplayMemory.py:32 — or None when the
    redis package / server is absent (pre-filled or in-process use)."""
    try:
        import redis
        c = redis.StrictRedis(host=host, port=6379)
        c.ping()
        return c
    except Exception as e:  # noqa: BLE001
        print(f"[b2rl] no Redis at {host}:6379 ({type(e).__name__}); running without the actor wire")
        return None


class Replay(_Replay):
    def __init__(self):
        cfg = ApexConfig.from_configuration()
        super().__init__(cfg, connect=_connect(cfg.REDIS_SERVER))


from distributed_rl_b200.replay_server import Replay_Server as _ReplayServerClient


class Replay_Server(_ReplayServerClient):
    """Consumer of a stand-alone ReplayServer (:170-257)."""

    def __init__(self):
        import configuration as C
        cfg = ApexConfig.from_configuration()
        super().__init__(cfg, connect=_connect(cfg.REDIS_SERVER),
                         connect_push=_connect(getattr(C, "REDIS_SERVER_PUSH", cfg.REDIS_SERVER)))
