"""GPU: the C++ host-side mirror (qdrant_b200/host/qdrant_b200.hpp: RawScorerBuilder / FilteredScorer /
BatchFilteredSearcher over the C ABI) gives the oracle's results bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import assert_topk_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror(tmp_path, oracle):
    from qdrant_b200 import build

    build.build()
    exe = os.path.join(ROOT, "qdrant_b200", "lib", "host_selftest")
    n, dim, nq, top = 90_000, 96, 3, 10
    rng = np.random.default_rng(11)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    queries = rng.standard_normal((nq, dim)).astype(np.float32)
    base.tofile(tmp_path / "base.f32")
    queries.tofile(tmp_path / "q.f32")
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(tmp_path / "base.f32"), str(tmp_path / "q.f32"), str(n), str(dim), str(nq), str(top), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(out, dtype=np.uint8)
    off = 0
    want = oracle.scan_f32(oracle.DOT, base, queries, top)
    for q in range(nq):
        c = int(raw[off:off + 4].view(np.uint32)[0]); off += 4
        got = raw[off:off + c * 8].view(oracle.SCORED); off += c * 8
        assert_topk_equal(got, want[q], None, f"cpp q={q}")
    kept = [i for i in range(48) if i % 3 != 0][:32]
    for q in range(nq):
        s = raw[off:off + 128].view(np.float32); off += 128
        np.testing.assert_array_equal(s, oracle.score_points_f32(oracle.DOT, base, queries[q], np.array(kept, np.uint32)))
    si = raw[off:off + 4].view(np.float32)[0]
    assert si == oracle.similarity_f32(oracle.DOT, base[1], base[2])
