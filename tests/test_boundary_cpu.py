"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/bm25f.h declares,
reports its ABI version, and the product path fails loudly when no device is present (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from document_search_engine_b200 import _ffi
from document_search_engine_b200 import And, Or, Term
from document_search_engine_b200.variants import Variants, expand_with_map

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bm25f.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bm25f_[a-z_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_ffi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(_ffi.LIB_PATH)


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_ffi.EXPORTS)
    src = open(HEADER).read()
    assert int(re.search(r"#define BM25F_ABI_VERSION (\d+)", src).group(1)) == _ffi.ABI_VERSION
    assert int(re.search(r"#define BM25F_MAX_K\s+(\d+)", src).group(1)) == _ffi.MAX_K
    assert int(re.search(r"#define BM25F_MAX_LEAVES_PER_QUERY (\d+)", src).group(1)) == _ffi.MAX_LEAVES_PER_QUERY


def test_library_exports_every_declared_symbol(lib):
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing
    lib.bm25f_abi_version.restype = ctypes.c_int
    assert lib.bm25f_abi_version() == _ffi.ABI_VERSION


def test_struct_layouts_match_header():
    # field order / count of the ctypes mirrors vs the header's typedefs
    src = open(HEADER).read()
    bodies = {name: body for body, name in re.findall(r"typedef struct \{([^}]*)\} (\w+);", src)}
    for cname, cls in (("bm25f_index_desc", _ffi.IndexDesc), ("bm25f_options", _ffi.Options),
                       ("bm25f_query_batch", _ffi.QueryBatchDesc), ("bm25f_stats", _ffi.Stats)):
        body = bodies[cname]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = re.findall(r"(\w+)\s*;", body)
        assert names == [n for n, _ in cls._fields_], cname


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from document_search_engine_b200.corpus import make_corpus
    ix = make_corpus(50, 40, 1, device="cpu")
    with pytest.raises(_ffi.EngineError):
        ix.searcher()


def test_bad_arguments_do_not_crash(lib):
    lib.bm25f_create.restype = ctypes.c_int
    lib.bm25f_last_error.restype = ctypes.c_char_p
    out = ctypes.c_void_p()
    assert lib.bm25f_create(None, 0, None, ctypes.byref(out)) == -1
    assert b"null" in lib.bm25f_last_error()
    d = _ffi.IndexDesc(_ffi.ABI_VERSION + 7, 1, 0, 0, 0, 0, None, None, None, None, None, None)
    assert lib.bm25f_create(ctypes.byref(d), 0, None, ctypes.byref(out)) == -5       # ABI mismatch
    lib.bm25f_destroy.restype = None
    lib.bm25f_destroy(None)


def test_variants_table(tmp_path):
    p = tmp_path / "uk_us_variations.txt"
    p.write_text("colour color\nhonour honor\n\nanalyse analyze\n", encoding="utf-8")
    v = Variants.load(str(p))
    assert v.uk_variations["colour"] == "color" and v.us_variations["analyze"] == "analyse"
    assert v.uk_us_variations == {"colour", "color", "honour", "honor", "analyse", "analyze"}
    assert v.other("color") == "colour" and v.other("grey") is None
    # the reference's own use: substitute only when the variant occurs in the index (my_flask.py:253-256)
    assert v.substitute("colour", lambda w: 3 if w == "color" else 0) == "color"
    assert v.substitute("colour", lambda w: 0) == "colour"
    # config-3 rewrite: AND of OR-groups
    q = v.expand(And([Term("exact", "colour"), Term("exact", "grey"), Term("exact", "honor", boost=2.0)]))
    assert str(q) == str(And([Or([Term("exact", "colour"), Term("exact", "color")]), Term("exact", "grey"),
                              Or([Term("exact", "honor", boost=2.0), Term("exact", "honour", boost=2.0)])]))
    q = v.expand(Term("exact", "colour"), only_if=lambda f, w: False)
    assert str(q) == str(Term("exact", "colour"))
    q = expand_with_map(And([Term("body", 60), Term("body", 61)]), lambda r: r ^ 1)
    assert str(q) == str(And([Or([Term("body", 60), Term("body", 61)]), Or([Term("body", 61), Term("body", 60)])]))
