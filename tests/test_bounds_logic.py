"""CPU property test of the pruning bounds the exact searches rest on (tc_search.cuh: cell_coord,
row_gap, own_row_gap, axis_bound), restated here in numpy f32 with the device's operation order:
for ANY query - inside the grid or cells outside it, as ICP sources are - and any indexed point,
the per-axis bound of the point's row never exceeds the true per-axis distance.  A bound that did
would let the search skip the true nearest neighbour."""
import numpy as np

F = np.float32


def _cell_coord(x, o, inv, n):
    u = ((x - o).astype(F) * inv).astype(F)          # xmul(xsub(x, o), inv)
    c = np.clip(np.floor(u).astype(np.int64), 0, n - 1)
    return c, u


def _row_gap(d, f, outside_aware):
    own = np.maximum(F(0), np.maximum(f - F(1), -f)).astype(F) if outside_aware else np.zeros_like(f)
    below = (f + (-d - 1).astype(F)).astype(F)
    above = (d.astype(F) - f).astype(F)
    return np.where(d == 0, own, np.where(d < 0, below, above)).astype(F)


def _axis_bound(du, cell, mag):
    return np.maximum(F(0), ((du * cell).astype(F) * F(0.99999)).astype(F) - (mag * F(4.8e-7)).astype(F))


def _check(rng, n_trials, outside_cells, outside_aware):
    worst = 0.0
    for _ in range(40):
        n = int(rng.integers(1, 2000))
        cell = F(10.0 ** rng.uniform(-3, 1))
        inv = F(1.0) / cell
        o = F(rng.uniform(-1000, 1000))
        extent = F((n - rng.uniform(0, 1)) * float(cell))           # bbox extent: last cell partly filled
        m = n_trials // 40
        p = (o + rng.uniform(0, 1, m).astype(F) * extent).astype(F)  # indexed points: inside the bbox
        # adversarial points: exactly on cell faces
        k = m // 4
        p[:k] = (o + (rng.integers(0, n, k).astype(F) * cell).astype(F)).astype(F)
        p = np.minimum(p, (o + extent).astype(F))
        q = (o + rng.uniform(-outside_cells, n + outside_cells, m).astype(F) * cell).astype(F)
        q[k:2 * k] = p[k:2 * k]                                       # queries on indexed points
        cq, uq = _cell_coord(q, o, inv, n)
        cp, _ = _cell_coord(p, o, inv, n)
        f = (uq - cq.astype(F)).astype(F)
        mag = (extent + np.abs(q - o).astype(F)).astype(F)           # g.ex + fabsf(qx - g.ox)
        bound = _axis_bound(_row_gap(cp - cq, f, outside_aware), cell, mag).astype(np.float64)
        true = np.abs(p.astype(np.float64) - q.astype(np.float64))
        slack = bound - true
        worst = max(worst, float(slack.max()))
        assert np.all(bound <= true), (n, float(cell), float(slack.max()))
    return worst


def test_row_bounds_never_exceed_the_true_axis_distance_inside_the_grid():
    _check(np.random.default_rng(1), 400_000, outside_cells=0.0, outside_aware=False)


def test_row_bounds_hold_for_queries_outside_the_grid():
    """ICP sources: up to dozens of cells outside the target's bounding box, with the own-row gap
    (own_row_gap) that round 2 added for `Best1` searches - and also without it (kNN external
    queries), where the own row's bound is simply zero."""
    rng = np.random.default_rng(2)
    _check(rng, 400_000, outside_cells=40.0, outside_aware=True)
    _check(rng, 200_000, outside_cells=40.0, outside_aware=False)


def test_own_row_gap_is_what_makes_the_top_plane_prunable():
    # a query 15 cells above the grid: without the own-row gap the bound of its (clamped) own row
    # is 0; with it, 15 cells minus the rounding slack
    n, cell, o = 7, F(0.2), F(0.0)
    q = np.array([o + F(22) * cell], F)
    cq, uq = _cell_coord(q, o, F(1) / cell, n)
    f = (uq - cq.astype(F)).astype(F)
    mag = (F(n) * cell + np.abs(q - o)).astype(F)
    assert cq[0] == n - 1
    assert _axis_bound(_row_gap(np.array([0]), f, False), cell, mag)[0] == 0
    b = _axis_bound(_row_gap(np.array([0]), f, True), cell, mag)[0]
    assert 14.9 * 0.2 < b <= 15.0 * 0.2
