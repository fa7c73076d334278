"""stringsearch_b200 -- B200-native suffix-array engine behind the stringsearch API.

Python mirror of the reference crates for one hot path:

    divsufsort.sort / sort_in_place        (crates/divsufsort/src/lib.rs:20-29)
    sacabase.SuffixArray, StringIndex      (crates/sacabase/src/lib.rs)
    sacapart.PartitionedSuffixArray        (crates/sacapart/src/lib.rs:26-98)
      + DistributedPartitionedSuffixArray (one rank per GPU), ReplicatedSuffixArray (query parallelism)
    divsuftest  (bench | run | verify)     (crates/divsuftest/src/main.rs; C++: stringsearch_b200/host)
    divsufsort.bwt / inverse_bwt           (c-sources/divsufsort.c:372-405, utils.c:111-156)
    divsufsort.lcp / sort_with_lcp         (LCP array; no counterpart in the reference)

Everything computes on the GPU through libgsa.so (include/gsa.h).  Importing the package
without the built library raises ImportError; there is no CPU fallback.
"""
from . import _native  # noqa: F401  (raises loudly when libgsa.so is missing)
from . import divsufsort, sacabase, sacapart  # noqa: F401
from ._native import GsaError, BuildStats  # noqa: F401

__all__ = ["divsufsort", "sacabase", "sacapart", "GsaError", "BuildStats"]
