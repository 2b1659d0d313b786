"""classeq2_b200 - B200-native placement path of classeq (k-mer extraction + murmur3 hashing,
k-mer index probing, per-node hit counting, one-vs-rest tree descent) behind the reference's
``place_sequences`` / model-type API surface.  Compute lives in ``libclasseq_b200.so``
(hand-written sm_100a CUDA, C ABI in ``include/classeq_b200.h``); importing this package without
that library raises - there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (loads the shared library or raises)
from .engine import (BatchResult, Index, PlaceParams, ResidentBatch, debug_kmer_hashes,  # noqa: F401
                     filter_sequence, host_murmur3_h1, make_batch)
from .model import BuiltModel, Clade, FlatModel, KmersMap, Tree  # noqa: F401
from .build import map_kmers_to_tree, tree_from_newick  # noqa: F401
from .placement import (PlacementTime, load_annotations, load_database, place_sequences,  # noqa: F401
                        place_sequences_native, read_fasta, save_database)

__all__ = ["Index", "ResidentBatch", "PlaceParams", "BatchResult", "Clade", "KmersMap", "Tree",
           "FlatModel", "BuiltModel", "debug_kmer_hashes", "host_murmur3_h1", "filter_sequence", "make_batch", "place_sequences", "place_sequences_native", "load_database",
           "save_database", "map_kmers_to_tree", "tree_from_newick", "load_annotations", "read_fasta", "PlacementTime"]
