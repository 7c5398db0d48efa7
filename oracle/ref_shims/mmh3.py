"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party `mmh3` wheel (2.5.1,
/root/reference/.conda/mmh3/meta.yaml:1-5) so the UNMODIFIED reference package can
be imported in the build container.  Backed by scikit-learn's independent
MurmurHash3_x86_32 so golden vectors do not depend on this repo's own murmur3.
Call site: /root/reference/bigsi/bloom/bloomfilter.py:5-6."""
from sklearn.utils import murmurhash3_32


def hash(key, seed=0, signed=True):
    if isinstance(key, str):
        key = key.encode("utf-8")
    return int(murmurhash3_32(bytes(key), seed=seed, positive=not signed))
