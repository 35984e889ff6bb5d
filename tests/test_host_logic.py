"""Host-side logic that needs no GPU: scalar algebra, slab planning, the fixed-order reduction,
the shell closure and the batch partition."""
import numpy as np
import pytest

from irrotational_warp_lab_b200 import backend, sharding, sweeps
from irrotational_warp_lab_b200._capi import FLUSH_PLANES, MAX_SHELLS, ROW_WIDTH
from oracle import oracle_np


def test_summary_metrics_and_tail_match_reference_formulas():
    m = backend.summary_metrics(0.9257104946542771, -0.8737748504578291)
    assert m["e_abs"] == 0.9257104946542771 + 0.8737748504578291
    assert m["ratio_e_pos_over_abs_e_neg"] == 0.9257104946542771 / 0.8737748504578291
    assert backend.summary_metrics(0.0, 0.0) == {"e_abs": 0.0, "neg_fraction": None, "net_over_abs": None,
                                                 "ratio_e_pos_over_abs_e_neg": None}
    assert backend.summary_metrics(1.0, 0.0)["ratio_e_pos_over_abs_e_neg"] is None
    assert backend.two_point_tail_extrapolation(0.9, 0.92, 40.0, 60.0) == oracle_np.two_point_tail(0.9, 0.92, 40.0, 60.0)
    # SURVEY App. B, n = 100: net/abs at infinity
    p = backend.two_point_tail_extrapolation(0.9058670090877987, 0.9203491050140727, 40.0, 60.0)
    q = backend.two_point_tail_extrapolation(-0.8019799711638814, -0.851025159615422, 40.0, 60.0)
    assert abs(abs(p + q) / (abs(p) + abs(q)) - 0.00010417053546587333) < 1e-15


@pytest.mark.parametrize("n0", [3, 15, 16, 17, 100, 101, 256, 1000, 1024])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_slab_bounds_cover_the_axis(n0, world):
    b = sharding.slab_bounds(n0, world)
    assert b[0][0] == 0 and b[-1][1] == n0
    for (a0, a1), (b0, b1) in zip(b, b[1:]):
        assert a1 == b0 and a0 <= a1
    for a0, a1 in b:
        assert a0 % FLUSH_PLANES == 0
        r0, r1 = sharding.rows_of_slab(a0, a1)
        assert r1 - r0 == (a1 - a0 + FLUSH_PLANES - 1) // FLUSH_PLANES
    rows = sorted(r for a0, a1 in b for r in range(*sharding.rows_of_slab(a0, a1)))
    assert rows == list(range(sharding.num_flush_blocks(n0)))


def test_reduce_table_is_fixed_order_and_exact_for_counts():
    rng = np.random.default_rng(0)
    nrows = 64
    table = np.zeros((nrows, ROW_WIDTH))
    exp_pos = 0.0
    tot = 0
    for r in range(nrows):
        e_pos, e_neg = rng.random() * 10.0 ** rng.integers(-12, 3), -rng.random()
        cnt = int(rng.integers(0, 2**40))
        table[r] = sharding.pack_row(e_pos, e_neg, cnt, 2, 3, [e_pos / 2, 1.0], [e_neg / 2, -1.0], [cnt // 2, 7])
        exp_pos += e_pos
        tot += cnt
    out = sharding.reduce_table_host(table, 2, n_points=tot + 5 * nrows)
    assert out["e_pos"] == exp_pos                      # same left-to-right order
    assert out["cnt_pos"] == tot and out["cnt_neg"] == 2 * nrows and out["cnt_zero"] == 3 * nrows
    assert out["cnt_nan"] == 0
    assert out["shell_cnt"][1] == 7 * nrows and len(out["shell_pos"]) == 2
    # splitting the rows between "ranks" and concatenating is the identity: P-independent by construction
    parts = np.concatenate([table[:10], table[10:37], table[37:]])
    assert sharding.reduce_table_host(parts, 2)["e_pos"] == out["e_pos"]
    assert ROW_WIDTH == 6 + 3 * MAX_SHELLS


def test_non_finite_points_poison_the_sums_like_the_reference_masks():
    """The reference multiplies by its masks (reproduce_rodal_exact.py:57-61, 220-224): NaN * 0 = inf * 0 = NaN."""
    import math
    e = np.array([1.0, -2.0, np.nan, 3.0])
    with np.errstate(invalid="ignore"):
        ref = (float(np.sum(e * (e > 0))), float(np.sum(e * (e < 0))))
    assert all(math.isnan(v) for v in ref)
    table = sharding.pack_row(4.0, -2.0, 2, 1, 0, [1.0], [-2.0], [3])[None, :]
    out = sharding.reduce_table_host(table, 1, n_points=4)              # one point is none of pos / neg / zero
    assert out["cnt_nan"] == 1
    assert all(math.isnan(out[k]) for k in ("e_pos", "e_neg")) and math.isnan(out["shell_pos"][0])
    # +inf inside shell 0, outside shell 1
    e = np.array([1.0, -2.0, np.inf]); m0 = np.array([1, 1, 1]); m1 = np.array([1, 1, 0])
    with np.errstate(invalid="ignore"):
        ref = [float(np.sum(e * (e > 0))), float(np.sum(e * (e < 0))),
               float(np.sum(e * (e > 0) * m0)), float(np.sum(e * (e > 0) * m1)), float(np.sum(e * (e < 0) * m0))]
    res = {"cnt_nan": 0, "e_pos": math.inf, "e_neg": -2.0, "shell_pos": [math.inf, 1.0], "shell_neg": [-2.0, -2.0]}
    sharding.apply_reference_nonfinite(res)
    got = [res["e_pos"], res["e_neg"], res["shell_pos"][0], res["shell_pos"][1], res["shell_neg"][0]]
    for g, r in zip(got, ref):
        assert (math.isnan(g) and math.isnan(r)) or g == r, (got, ref)
    fin = {"cnt_nan": 0, "e_pos": 1.0, "e_neg": -1.0, "shell_pos": [0.5], "shell_neg": [-0.5]}
    assert sharding.apply_reference_nonfinite(dict(fin)) == fin


def test_partition_points_round_robin():
    for world in (1, 2, 3, 8):
        seen = sorted(q for r in range(world) for q in sweeps.partition_points(20, r, world))
        assert seen == list(range(20))
    assert sweeps.partition_points(20, 3, 8) == [3, 11, 19]
    assert sweeps.partition_points(2, 5, 8) == []


def test_shell_closure_looks_up_known_radii_and_relaunches_for_new_ones():
    calls = []

    def evaluate(radii, out=None):
        calls.append(list(radii))
        return {"shell_pos": [r * 2 for r in radii], "shell_neg": [-r for r in radii]}

    clo = backend._ShellClosure(evaluate, {40.0: (1.0, -1.0), 60.0: (2.0, -2.0)})
    assert clo(40.0) == (1.0, -1.0) and clo(60) == (2.0, -2.0) and calls == []
    assert clo(55.5) == (111.0, -55.5) and calls == [[55.5]]
    assert clo(55.5) == (111.0, -55.5) and len(calls) == 1          # cached after the first relaunch


def test_sweep_record_algebra_matches_reference_loop():
    rec = sweeps._sweep_record(1.0, 0.9058670090877987, -0.8019799711638814, 0.9203491050140727, -0.851025159615422,
                               40.0, 60.0, 0.1)
    assert rec["ratio_r2"] == 0.9203491050140727 / 0.851025159615422
    assert abs(rec["tail_net_frac"] - 0.00010417053546587333) < 1e-15
    assert set(rec) == {"v", "e_pos_inf", "e_neg_inf", "e_net_inf", "e_abs_inf", "tail_net_frac", "ratio_r2", "timing_s"}
    assert sweeps._sweep_record(0.0, 0, 0, 0, 0, 40.0, 60.0, 0.0)["ratio_r2"] is None


def test_grid_spacing_is_taken_from_linspace_not_from_the_formula():
    x, dx = backend._grid(-66.0, 66.0, 100, "float64")
    assert dx == 1.3333333333333286 and dx != 132.0 / 99
    x32, dx32 = backend._grid(-66.0, 66.0, 8, "float32")
    assert dx32 == 18.85714340209961 and x32.dtype == np.float32


def test_tail_fit_restatement():
    """Reference tests/test_tail.py restated: constant field averages to itself, a 1/r^4 field is recovered,
    n <= 2 does not converge; plus a cross-check of the binning against a direct per-bin mean."""
    from irrotational_warp_lab_b200 import tail
    x = np.linspace(-10, 10, 101)
    X, Y = np.meshgrid(x, x)
    r_bins, avg = tail.radial_average_z0(5.0 * np.ones_like(X), x, x, n_bins=20)
    assert np.allclose(avg[avg > 0], 5.0, atol=1e-10) and len(r_bins) == 20
    R = np.sqrt(X ** 2 + Y ** 2) + 1e-6
    res = tail.compute_tail_correction(2.0 / R ** 4, x, x, n_bins=30)
    assert abs(res.exponent - 4.0) < 0.1 and res.amplitude > 0 and res.tail_integral_neg == 0.0
    assert res.tail_integral_pos > 0 and res.tail_uncertainty >= 0 and res.fit_r_max == res.r_bins[-1]
    assert tail.tail_integral_2d(1.0, 2.0, 10.0) == (0.0, 0.0) and tail.tail_integral_2d(1.0, 1.5, 10.0) == (0.0, 0.0)
    pos, neg = tail.tail_integral_2d(1.0, 3.0, 10.0)
    assert abs(pos - 2.0 * np.pi * (1.0 / 10.0)) < 1e-12 and neg == 0.0
    pos, neg = tail.tail_integral_2d(-1.0, 4.0, 10.0)
    assert pos == 0.0 and abs(neg - 2.0 * np.pi * (1.0 / (2.0 * 100.0))) < 1e-12
    assert tail.fit_power_law_decay(np.array([1.0, 2.0]), np.array([1.0, 0.5]), 0.0, 3.0) == (0.0, 0.0, np.inf)
    # the closed last edge: the farthest point belongs to no bin
    edges = np.linspace(0, R.max(), 31)
    sel = (R >= edges[-2]) & (R < edges[-1])
    assert avg.shape == (20,) and not sel[R == R.max()].any()


def test_tail_integral_3d_algebra():
    """4 pi A int_R^inf r^(2-n) dr = 4 pi A R^(3-n) / (n - 3) for n > 3; divergent otherwise (twin of tail.py:134-161)."""
    from irrotational_warp_lab_b200 import tail
    pos, neg = tail.tail_integral_3d(2.0, 6.0, 10.0)
    assert neg == 0.0 and pos == pytest.approx(4.0 * np.pi * 2.0 * 10.0 ** -3 / 3.0, rel=1e-15)
    pos, neg = tail.tail_integral_3d(-2.0, 6.0, 10.0)
    assert pos == 0.0 and neg == pytest.approx(4.0 * np.pi * 2.0 * 10.0 ** -3 / 3.0, rel=1e-15)
    assert tail.tail_integral_3d(1.0, 3.0, 10.0) == (0.0, 0.0)
    # numerical quadrature of the same integrand
    r = np.geomspace(10.0, 1e7, 200001)
    f = 4.0 * np.pi * 2.0 * r ** (2.0 - 6.0)
    assert float(np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(r))) == pytest.approx(tail.tail_integral_3d(2.0, 6.0, 10.0)[0], rel=1e-6)
