"""bigsi_b200: B200-native BIGSI query engine (hand-written sm_100a CUDA behind a C ABI).

Importing the package never compiles or falls back to anything: the CUDA library
(bigsi_b200/libbigsi_b200.so) must have been built (python -m bigsi_b200.build) and is loaded
on first use; every compute call fails loudly without it or without a CUDA device.
"""
from ._lib import MODE_AND, MODE_COUNTS, BigsiB200Error, device_count  # noqa: F401
from .bigsi import BIGSI, BigsiQueryResult, DEFAULT_CONFIG  # noqa: F401
from .bloom import BloomFilter, generate_hashes, load_bitarray  # noqa: F401
from .index import DeviceIndex, hash_kmers, hash_kmers_dev, kmers_to_array, threshold_dev  # noqa: F401
from .metadata import DELETION_SPECIAL_SAMPLE_NAME, SampleMetadata  # noqa: F401
from .scoring import Scorer  # noqa: F401
from .utils import canonical, convert_query_kmer, reverse_comp, seq_to_kmers  # noqa: F401

__version__ = "0.1.0"
