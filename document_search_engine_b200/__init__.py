"""B200-native BM25F query-scoring engine behind the Whoosh searcher surface.

Drop-in for one hot path of CodeOptimist/document-search-engine: BM25F scoring of
parsed query terms over inverted-index posting lists plus top-k collection
(SURVEY.md §8).  Host code is Python; the GPU is reached through the C ABI in
``include/bm25f.h`` (``csrc/libbm25f.so``) via ctypes.  There is no CPU fallback.
"""
from .index import FlatIndex, Schema
from .query import And, DateRange, Every, Not, NullQuery, Or, Phrase, Prefix, QueryParser, Term, Wildcard
from .scoring import AscDateBM25F, BM25F, DateBM25F, DescDateBM25F, WeightingModel

__all__ = ["FlatIndex", "Schema", "And", "Or", "Term", "Every", "Not", "NullQuery", "Prefix", "Wildcard", "DateRange", "Phrase", "QueryParser",
           "BM25F", "DateBM25F", "DescDateBM25F", "AscDateBM25F", "WeightingModel"]
__version__ = "0.1.0"
