"""trgt-denovo_b200: B200-native read<->allele alignment hot path of trgt-denovo.

Host-side mirror (Python, over ctypes) of the reference interface for this path.  The package
directory name contains a hyphen; import it through the top-level shim `trgt_denovo_b200`.

Everything that computes goes through the C ABI of `libtrgt_denovo_b200.so`
(include/trgt_denovo_b200.h) and therefore through the sm_100a CUDA kernels.  There is no CPU
path: if the library is missing or no CUDA device is present, construction fails loudly.
"""
from .binding import (  # noqa: F401
    AlleleOut,
    Config,
    Counters,
    Heuristic,
    LIB_PATH,
    Penalties,
    TdbError,
    TrioLocus,
    device_count,
    lib,
)
from .aligner import (  # noqa: F401
    AlignmentStatus,
    Context,
    PackedInput,
    pack_sequences,
    WFAligner,
    create_aligner_with_scoring,
)
from .batch import SeqBatch  # noqa: F401
from . import rows, shard  # noqa: F401  (per-locus pipeline above the hot path; locus sharding)
