"""TEST INFRASTRUCTURE ONLY -- empty stand-in so `import redis` at
/root/reference/bigsi/storage/redis.py:1 succeeds; the redis backend is never
instantiated by the oracle harness."""


class StrictRedis:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("redis is not available in the oracle harness")


Redis = StrictRedis
