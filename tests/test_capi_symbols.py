"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/qb200.h declares; without a GPU every entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from qdrant_b200 import build

    build.build()
    from qdrant_b200 import _capi

    return _capi


def header_functions():
    src = open(os.path.join(ROOT, "include", "qb200.h")).read()
    return sorted(set(re.findall(r"QB_API[^;(]*?\b(qb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    fns = header_functions()
    for must in ("qb_storage_create_dense", "qb_storage_create_sq8", "qb_storage_create_pq", "qb_storage_create_bq", "qb_scorer_create",
                 "qb_scorer_create_internal", "qb_score_points", "qb_score_point", "qb_score_internal", "qb_search_batch", "qb_rescore"):
        assert must in fns


def test_library_exports_every_declared_symbol(capi):
    L = C.CDLL(capi.LIB_PATH)
    for name in header_functions():
        assert hasattr(L, name), f"{name} declared in include/qb200.h but not exported"


def test_python_binding_covers_header(capi):
    assert sorted(capi.SIGNATURES.keys()) == header_functions()
    assert capi.lib().qb_abi_version() == 2


def test_no_silent_cpu_fallback(capi):
    """Without a CUDA device creation must fail with QB_ERR_NO_DEVICE — never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    n = C.c_int32(-1)
    assert capi.lib().qb_device_count(C.byref(n)) == capi.QB_ERR_NO_DEVICE
    assert n.value == 0
    from qdrant_b200.scorer import DenseVectorStorage, Distance

    with pytest.raises(capi.QbError) as ei:
        DenseVectorStorage(np.zeros((4, 32), np.float32), Distance.Dot)
    assert ei.value.status == capi.QB_ERR_NO_DEVICE
    assert b"no CPU fallback" in capi.lib().qb_last_error()


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (tier rule: the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "qdrant_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.replace("test_product_does_not_import_oracle", ""), f"{f} mentions oracle"


def test_custom_query_flat_layouts():
    """Host-side mirrors of RecoQuery / DiscoverQuery / ContextQuery flatten their vectors in the order qb_scorer_create_custom
    documents (include/qb200.h) — the same order the reference's flat_iter() yields (reco_query.rs:26-28, discover_query.rs:38-42,
    context_query.rs:94-96)."""
    import numpy as np

    from qdrant_b200 import scorer as qb

    v = [np.full(4, i, np.float32) for i in range(7)]
    vecs, n_a, n_b = qb.RecoBestScoreQuery(qb.RecoQuery([v[0], v[1]], [v[2]])).flat()
    assert (n_a, n_b) == (2, 1) and [int(x[0]) for x in vecs] == [0, 1, 2]
    assert qb.RecoSumScoresQuery(qb.RecoQuery([], [v[3]])).flat()[1:] == (0, 1)
    vecs, n_a, n_b = qb.DiscoverQuery(v[6], [qb.ContextPair(v[0], v[1]), qb.ContextPair(v[2], v[3])]).flat()
    assert (n_a, n_b) == (2, 0) and [int(x[0]) for x in vecs] == [6, 0, 1, 2, 3]
    vecs, n_a, n_b = qb.ContextQuery([qb.ContextPair(v[4], v[5])]).flat()
    assert (n_a, n_b) == (1, 0) and [int(x[0]) for x in vecs] == [4, 5]
    assert int(qb.RecoBestScoreQuery.kind) == 1 and int(qb.RecoSumScoresQuery.kind) == 2 and int(qb.DiscoverQuery.kind) == 3 and int(qb.ContextQuery.kind) == 4


def test_rust_ffi_declares_every_exported_function():
    """bindings/rust/src/ffi.rs is generated from the header (tools/gen_rust_ffi.py): it must cover the whole C ABI, with matching arity."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fns, status, abi = gen.parse()
    assert sorted(f[0] for f in fns) == header_functions()
    rs = open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    assert f"QB200_ABI_VERSION: i32 = {abi}" in rs
    for name, params, ret in fns:
        m = re.search(r"pub fn " + name + r"\(([^)]*)\)", rs)
        assert m, f"{name} missing from ffi.rs (run tools/gen_rust_ffi.py)"
        got = [a for a in m.group(1).split(",") if a.strip()]
        assert len(got) == len(params), f"{name}: {len(got)} parameters in ffi.rs, {len(params)} in the header"
