"""In-memory stand-in for redis.StrictRedis with REAL list semantics (LRANGE / LTRIM index rules,
MULTI pipelines executed atomically under one lock) — the subset the learners and actors use.
Test infrastructure only."""
import threading


def _norm(i, n):
    return i + n if i < 0 else i


class _Pipeline:
    def __init__(self, db):
        self._db, self._cmds = db, []

    def __getattr__(self, name):
        def queue(*a, **kw):
            self._cmds.append((name, a, kw))
            return self
        return queue

    def execute(self):
        with self._db._mu:
            out = [getattr(self._db, "_" + name)(*a, **kw) for name, a, kw in self._cmds]
        self._cmds = []
        return out


def _locked(name):
    def f(self, *a, **kw):
        with self._mu:
            self.calls.append(name)
            return getattr(self, "_" + name)(*a, **kw)
    return f


class FakeRedis:
    def __init__(self, host=None, port=6379, **kw):
        self._s, self._mu = {}, threading.RLock()
        self.calls = []

    # -- raw commands (no locking; callers hold _mu) ------------------------------------------
    def _lrange(self, key, a, b):
        lst = self._s.get(key, [])
        n = len(lst)
        a, b = max(_norm(a, n), 0), min(_norm(b, n), n - 1)
        return list(lst[a:b + 1]) if a <= b else []

    def _ltrim(self, key, a, b):
        lst = self._s.get(key, [])
        n = len(lst)
        a2, b2 = max(_norm(a, n), 0), min(_norm(b, n), n - 1)
        kept = lst[a2:b2 + 1] if a2 <= b2 else []
        if kept:
            self._s[key] = kept
        else:
            self._s.pop(key, None)
        return True

    def _delete(self, *keys):
        return sum(self._s.pop(k, None) is not None for k in keys)

    def _rpush(self, key, *vals):
        self._s.setdefault(key, []).extend(vals)
        return len(self._s[key])

    def _llen(self, key):
        return len(self._s.get(key, []))

    def _set(self, key, val):
        self._s[key] = val
        return True

    def _get(self, key):
        return self._s.get(key)

    # -- public surface -------------------------------------------------------------------------
    def pipeline(self):
        return _Pipeline(self)

    lrange = _locked("lrange")
    ltrim = _locked("ltrim")
    delete = _locked("delete")
    rpush = _locked("rpush")
    llen = _locked("llen")
    set = _locked("set")
    get = _locked("get")

    def scan(self):
        with self._mu:
            return (0, list(self._s.keys()))

    def ping(self):
        return True
