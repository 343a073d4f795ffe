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


def _d2(p, q):
    """dist2_exact: (dx*dx + dy*dy) + dz*dz in f32, no FMA."""
    d = (p - q).astype(F)
    return ((d[..., 0] * d[..., 0]).astype(F) + (d[..., 1] * d[..., 1]).astype(F)).astype(F) + \
        (d[..., 2] * d[..., 2]).astype(F)


def test_icp_keep_rule_never_keeps_a_match_that_a_search_would_change():
    """tc_icp.cu keeps last iteration's match without a search when
         (d + |s - q_full|) (1 + 2e-5) + 3e-7 |s|_1 < a (1 - 2e-5),
    a = distance from q_full (the position at the last full search) to the nearest OTHER point.
    Whenever the rule fires, the brute-force nearest neighbour at s under the reference's order
    (smallest f32 d2, strict <: ties go to the lower index) must be that same point - including
    near-ties constructed to sit on the edge of the rule."""
    rng = np.random.default_rng(7)
    fired = 0
    for trial in range(300):
        n = 64
        scale = F(10.0 ** rng.uniform(-1, 3))
        tgt = (rng.normal(size=(n, 3)) * scale).astype(F)
        q_full = (tgt[rng.integers(0, n, 256)] + (rng.normal(size=(256, 3)) * scale * 0.2)).astype(F)
        d2_full = _d2(tgt[None, :, :], q_full[:, None, :])                  # [256, n]
        order = np.lexsort((np.arange(n)[None, :].repeat(256, 0), d2_full), axis=1)
        match = order[:, 0]
        a = np.sqrt(np.take_along_axis(d2_full, order[:, 1:2], 1)[:, 0]).astype(F)  # nearest other
        # move the query: random steps, plus steps aimed at the runner-up sized to graze the rule
        step = (rng.normal(size=(256, 3)) * scale * 0.05).astype(F)
        runner = tgt[order[:, 1]]
        aim = runner - q_full
        aim /= np.maximum(np.linalg.norm(aim, axis=1, keepdims=True), 1e-20)
        d0 = np.sqrt(d2_full[np.arange(256), match])
        graze = ((a - d0) * 0.5 * rng.uniform(0.9, 1.1, 256))[:, None] * aim
        step[::2] = graze[::2].astype(F)
        s = (q_full + step).astype(F)
        d2_now = _d2(tgt[None, :, :], s[:, None, :])
        d = np.sqrt(d2_now[np.arange(256), match]).astype(F)
        m = (s - q_full).astype(F)
        moved = np.sqrt(((m[:, 0] * m[:, 0]).astype(F) + (m[:, 1] * m[:, 1]).astype(F)).astype(F) +
                        (m[:, 2] * m[:, 2]).astype(F)).astype(F)
        lhs = ((d + moved).astype(F) * F(1.00002)).astype(F) + \
            (F(3e-7) * np.abs(s).sum(1).astype(F)).astype(F)
        keep = (a > 0) & (lhs < (a * F(0.99998)).astype(F))
        brute = np.lexsort((np.arange(n)[None, :].repeat(256, 0), d2_now), axis=1)[:, 0]
        assert np.array_equal(brute[keep], match[keep])
        fired += int(keep.sum())
    assert fired > 10_000   # the rule does fire (about half of the trials)


def test_box_query_cells_cover_every_point_within_the_radius():
    """box_cells (tc_search.cuh): the cells [cell(q - r'), cell(q + r')] per axis, r' = sqrt(r2)
    inflated by 1e-5 relative + 1e-6 (|q|_1 + cell), hold every indexed point within distance
    sqrt(r2) of the query - the exactness of seeded ICP searches, radius queries and the radius
    outlier count rests on it.  Points are placed at |p - q| <= r on the axis, many exactly at r."""
    rng = np.random.default_rng(11)
    for _ in range(60):
        n = int(rng.integers(1, 3000))
        cell = F(10.0 ** rng.uniform(-3, 1))
        inv = F(1.0) / cell
        o = F(rng.uniform(-500, 500))
        m = 5000
        q = (o + rng.uniform(-5, n + 5, m).astype(F) * cell).astype(F)
        other = (rng.uniform(-100, 100, (m, 2))).astype(F)            # the query's other two coordinates
        r = (10.0 ** rng.uniform(-4, 1, m)).astype(F) * cell
        r2 = (r * r).astype(F)
        frac = rng.uniform(-1, 1, m)
        frac[: m // 3] = np.sign(frac[: m // 3])                       # exactly at the radius
        p = (q + (frac * r.astype(np.float64)).astype(F)).astype(F)
        # only points whose f32 squared distance really is <= r2 count as "within"
        within = ((p - q).astype(F) * (p - q).astype(F)).astype(F) <= r2
        rr = (np.sqrt(r2).astype(F) * F(1.00001)).astype(F) + \
            (F(1e-6) * (np.abs(q) + np.abs(other).sum(1).astype(F) + cell).astype(F)).astype(F) + F(1e-30)
        lo, _ = _cell_coord((q - rr).astype(F), o, inv, n)
        hi, _ = _cell_coord((q + rr).astype(F), o, inv, n)
        cp, _ = _cell_coord(p, o, inv, n)
        ok = (cp >= lo) & (cp <= hi)
        assert np.all(ok[within]), (n, float(cell))
