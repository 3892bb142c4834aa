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
    gs, ws = np.ascontiguousarray(got["score"]), np.ascontiguousarray(want["score"])
    # BIT PATTERNS, not values: -0.0 must come back as -0.0.  NaN payloads are not observable through f32 comparison (x86 produces
    # the negative default NaN, the GPU the canonical positive one), so NaNs only have to sit at the same positions.
    nan = np.isnan(ws)
    assert np.array_equal(np.isnan(gs), nan), f"{what}: NaN positions differ\n got {gs[:10]}\nwant {ws[:10]}"
    assert np.array_equal(gs.view(np.uint32)[~nan], ws.view(np.uint32)[~nan]), f"{what}: score bits differ\n got {gs[:10]}\nwant {ws[:10]}"
    fin = gs[~nan]
    assert np.all(fin[:-1] >= fin[1:]) and (not nan.any() or nan[: int(nan.sum())].all()), f"{what}: not sorted descending (NaN first, OrderedFloat)"
    boundary = gs[-1]
    for sc in np.unique(gs[~nan]):
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
