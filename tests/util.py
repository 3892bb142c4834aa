"""Shared helpers for the parity tests."""
import numpy as np


def assert_topk_equal(got, want, all_scores=None, what=""):
    """Tie-aware comparison of two descending top-k lists of ScoredPointOffset.

    ScoredPointOffset orders by score only (lib/common/common/src/types.rs:21-25): the reference's heap keeps an
    unspecified subset of equal-score points at the k-th boundary and an unspecified order inside equal-score runs.
    So: score arrays must be identical (bit-exact); ids must match as sets inside every equal-score run above the
    boundary score; at the boundary score the ids may differ but must really have that score (all_scores, if given).
    """
    assert got.size == want.size, f"{what}: count {got.size} != {want.size}"
    if got.size == 0:
        return
    gs, ws = got["score"], want["score"]
    assert np.array_equal(gs.view(np.uint32) if False else gs, ws), f"{what}: scores differ\n got {gs[:10]}\nwant {ws[:10]}"
    assert np.all(gs[:-1] >= gs[1:]), f"{what}: not sorted descending"
    boundary = gs[-1]
    for sc in np.unique(gs):
        gi = set(got["idx"][gs == sc].tolist())
        wi = set(want["idx"][ws == sc].tolist())
        assert len(gi) == int(np.sum(gs == sc)), f"{what}: duplicate ids"
        if sc != boundary:
            assert gi == wi, f"{what}: id sets differ at score {sc}: {sorted(gi)[:8]} vs {sorted(wi)[:8]}"
        elif all_scores is not None:
            for i in gi:
                assert all_scores[i] == sc, f"{what}: id {i} reported with score {sc} but oracle says {all_scores[i]}"


def pack_bitmap(mask: np.ndarray) -> np.ndarray:
    n = mask.size
    bits = np.zeros(((n + 63) // 64) * 64, dtype=bool)
    bits[:n] = mask
    return np.packbits(bits, bitorder="little").view(np.uint64).copy()
