"""The prefilter kernels (qb_prefilter.cu, pq_scan8 / pq_scan16 in qb_quant.cu) may only DROP a row when its exact score cannot
reach the threshold.  Each of them rests on an inequality between the exact f32 score the oracle computes (the reference's summation
order) and the compact score the kernel computes; this file restates the kernels' arithmetic in numpy — same quantisation steps, same
constants — and checks the inequality row by row on random and adversarial inputs, on the CPU.  (The GPU tests check the end result:
bit-identical top-k; these check the margin itself, with nothing left to chance by a lucky threshold.)"""
import numpy as np
import pytest

F = np.float32


def _q8_rows(x):
    """f32_to_q8_rows_kernel: c = rint(x * (127 / max)), s_r = max / 127; rows below 1e-30 keep all-zero codes and s_r = 2 * max."""
    mx = np.abs(x).max(axis=1).astype(F)
    tiny = ~(mx >= F(1e-30))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        sr = np.where(tiny, mx * F(2), (mx / F(127)).astype(F)).astype(F)
        inv = np.where(tiny, F(0), (F(127) / mx).astype(F)).astype(F)
    c = np.clip(np.rint((x * inv[:, None]).astype(F)), -127, 127).astype(np.int64)
    return c, sr


def _q8_query(q):
    """dense_q8_filter_kernel prologue: q ~ s_q (h + l / 254)."""
    qmax = F(np.abs(q).max())
    sq = (qmax / F(127)).astype(F) if qmax > 0 else F(0)
    inv = (F(127) / qmax).astype(F) if qmax > 0 else F(0)
    y = (q * inv).astype(F)
    h = np.clip(np.rint(y), -127, 127).astype(F)
    l = np.clip(np.rint(((y - h).astype(F) * F(254)).astype(F)), -127, 127)
    return sq, h.astype(np.int64), l.astype(np.int64)


@pytest.mark.parametrize("case", ["gauss", "unit", "spiky", "mixed_scale", "denormal", "sparse_query", "flat"])
@pytest.mark.parametrize("dim", [64, 200, 768, 1000])
def test_int8_plane_upper_bound_covers_the_exact_score(oracle, case, dim):
    rng = np.random.default_rng(dim + sum(map(ord, case)))
    n = 3000
    x = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal(dim).astype(F)
    if case == "unit":
        x = oracle.preprocess_rows_f32(oracle.COSINE, x); q = oracle.preprocess_f32(oracle.COSINE, q)
    elif case == "spiky":
        x[:, rng.integers(0, dim, 3)] *= F(300.0)           # a few huge coordinates: everything else quantises to ~0
    elif case == "mixed_scale":
        x *= (10.0 ** rng.uniform(-8, 8, (n, 1))).astype(F)
    elif case == "denormal":
        x[: n // 2] *= F(1e-38); x[n // 2 : n // 2 + 10] = 0
    elif case == "sparse_query":
        q[rng.random(dim) < 0.9] = 0
    elif case == "flat":
        x = np.sign(x).astype(F) * F(0.37); q = np.sign(q).astype(F)   # every |x_i| equal: codes +-127, rounding error at its bound
    exact = oracle.score_points_f32(oracle.DOT, x, q, np.arange(n, dtype=np.uint32)).astype(np.float64)
    c, sr = _q8_rows(x)
    sq, h, l = _q8_query(q)
    H, L = c @ h, c @ l
    q1, qn = np.abs(q).astype(np.float64).sum(), np.sqrt((q.astype(np.float64) ** 2).sum())
    mxn = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1)).max()
    e_q = q1 * (0.5 + 2.0 ** -13) + float(sq) * dim * 0.27
    slack = (dim * 2.0 ** -22 + 2.0 ** -17) * (1 + np.sqrt(dim) / 127) * qn * mxn + 1e-37
    up = sr.astype(np.float64) * (float(sq) * (H + L / 254.0) + e_q)
    worst = (exact - slack - up).max()
    assert worst <= 0, f"{case} dim={dim}: exact exceeds the kernel's upper bound by {worst}"
    # and the bound is not vacuous: on ordinary data it stays within a few percent of the score spread
    if case in ("gauss", "unit"):
        assert np.median(up - exact) < 0.5 * exact.std()


@pytest.mark.parametrize("dim", [64, 768, 1536])
def test_bf16_plane_margin_covers_the_exact_score(oracle, dim):
    rng = np.random.default_rng(dim)
    n = 3000
    x = (rng.standard_normal((n, dim)) * 10.0 ** rng.uniform(-3, 3, (n, 1))).astype(F)
    q = rng.standard_normal(dim).astype(F)
    exact = oracle.score_points_f32(oracle.DOT, x, q, np.arange(n, dtype=np.uint32)).astype(np.float64)
    u = x.view(np.uint32).astype(np.uint64)
    bf = (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(F)     # round to nearest even, as bf16_rne
    approx = bf.astype(np.float64) @ q.astype(np.float64)
    qn, mxn = np.sqrt((q.astype(np.float64) ** 2).sum()), np.sqrt((x.astype(np.float64) ** 2).sum(axis=1)).max()
    eps = (2.0 ** -9 * (1 + 2.0 ** -10) + dim * 2.0 ** -22) * qn * mxn
    assert np.abs(approx - exact).max() <= eps


@pytest.mark.parametrize("case", ["gauss", "wide", "cancel", "flat"])
def test_pq_u8_table_threshold_never_drops_a_qualifying_row(oracle, case):
    """pq16_prep_kernel: u = rint((lut - min_j) / step), T = floor((thr - sum_j min_j - margin) / step) - 1: exact >= thr  =>  sum u >= T.
    pq_scan8: |bf16-table sum - exact| <= (2^-9 + 2^-15) sum_j max|lut_j|."""
    rng = np.random.default_rng(sum(map(ord, case)))
    m, K, n = 96, 256, 4000
    lut = rng.standard_normal((m, K)).astype(F)
    if case == "wide":
        lut *= (10.0 ** rng.uniform(-6, 6, (m, 1))).astype(F)
    elif case == "cancel":
        lut[: m // 2] *= F(1e4); lut[m // 2 :] = -lut[: m // 2]
    elif case == "flat":
        lut[:] = F(0.25)
    codes = rng.integers(0, K, (n, m), dtype=np.uint8)
    if case == "cancel":
        codes[:, m // 2 :] = codes[:, : m // 2]
    pq = oracle.PQ(m * 4, 4, np.zeros((K, m * 4), F), codes, oracle.QD_DOT, False)
    exact = np.array([pq.score(lut, i) for i in range(n)], dtype=np.float64)
    mn = lut.min(axis=1)
    rng_j = (lut.max(axis=1) - mn).astype(F)
    step = max(F(rng_j.max() / F(255)), F(1e-30))
    inv = F(1) / step
    u = np.clip(np.rint(((lut - mn[:, None]).astype(F) * inv).astype(F)), 0, 255).astype(np.int64)
    usum = u[np.arange(m)[None, :], codes].sum(axis=1)
    A = float(np.abs(lut).max(axis=1).astype(np.float64).sum()) * 1.0001
    base = float(mn.astype(np.float64).sum())
    margin = m * float(step) * (0.5 + 1 / 16384) + 4 * m * 2.0 ** -24 * A
    for thr in np.quantile(exact, [0.5, 0.9, 0.999]):
        T = np.floor((thr - base - margin) / float(step)) - 1
        qualifies = exact >= thr
        assert (usum[qualifies] >= T).all(), f"{case}: a row with exact >= thr fails the integer test"
    u32 = lut.view(np.uint32).astype(np.uint64)
    bf = (((u32 + 0x7FFF + ((u32 >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(F)
    approx = bf.astype(np.float64)[np.arange(m)[None, :], codes].sum(axis=1)
    assert np.abs(approx - exact).max() <= (2.0 ** -9 + 2.0 ** -15) * A
