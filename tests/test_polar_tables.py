"""CPU check of K1's bin tables (scl_polar_tables, host code): the kernel never evaluates atanf, sqrt or the double-precision
index formulas of descriptor.h:1352-1374,1425-1435 per point; it searches threshold tables bisected from those formulas.
Here the same lookups are done with numpy in float32 (s = fl(fl(x*x) + fl(y*y)), t = fl(|y| / |x|)) and compared with the
oracle's per-point (ring, sector) on random clouds and on points a few ulps either side of every ring and sector boundary.
No GPU needed: the -m gpu tests compare the kernel itself with the same oracle."""
import numpy as np
import pytest

from oracle_lib import Oracle


def _bins_from_tables(tab, pts, S):
    x, y = pts[:, 0].astype(np.float32), pts[:, 1].astype(np.float32)
    with np.errstate(all="ignore"):
        s = (x * x + y * y).astype(np.float32)                      # float32 products and sum, as the reference's float expression
        ring = 1 + np.searchsorted(tab["ring_thr"], s, side="right")
        xn, yn = x < 0, y < 0
        qd = np.where(xn, np.where(yn, 2, 1), np.where(yn, 3, 0))
        num = np.where(qd == 3, -y, y).astype(np.float32)
        den = np.where(qd == 1, -x, x).astype(np.float32)
        t = (num / den).astype(np.float32)
    sector = np.zeros(len(x), np.int64)
    for q in range(4):
        m = qd == q
        cnt = np.searchsorted(tab["sec_thr"][q], t[m], side="right")
        sector[m] = tab["sec_base"][q] + tab["sec_dir"][q] * cnt
    sector[np.isnan(t)] = 1
    ok = ~(np.isnan(x) | np.isnan(y)) & ~(s > tab["s_max"])
    return np.where(ok, ring, 0).astype(np.int32), np.where(ok, sector, 0).astype(np.int32)


def _edge_points(R, S, rng):
    out = []
    f32 = np.float32
    for k in range(1, R + 1):                                       # ring boundaries (and the radius limit at k = R)
        r = f32(80.0 * k / R)
        for a in rng.uniform(0, 2 * np.pi, 16):
            x0, y0 = f32(r * np.cos(a)), f32(r * np.sin(a))
            xs = [x0]
            for _ in range(4):
                xs.append(np.nextafter(xs[-1], f32(np.inf)))
            for _ in range(4):
                xs.insert(0, np.nextafter(xs[0], f32(-np.inf)))
            out += [(x, y0, 1.0) for x in xs]
        for d in range(-3, 4):
            x = r
            for _ in range(abs(d)):
                x = np.nextafter(x, f32(np.inf if d > 0 else -np.inf))
            out += [(x, 0.0, 1.0), (0.0, -x, 1.0), (-x, 0.0, 1.0), (0.0, x, 1.0)]
    for k in range(S):                                              # sector boundaries
        a = 2 * np.pi * k / S
        for rad in (0.5, 7.0, 33.0, 79.0):
            p = np.array([rad * np.cos(a), rad * np.sin(a), 1.0], np.float32)
            out += [tuple(p), tuple(np.nextafter(p, f32(np.inf))), tuple(np.nextafter(p, f32(-np.inf)))]
    out += [(0.0, 0.0, 1.0), (-0.0, 3.0, 1.0), (3.0, -0.0, 1.0), (1e-38, 1e-38, 1.0), (80.0, 0.0, 1.0), (80.00001, 0.0, 1.0), (0.0, -80.0, 1.0)]
    return np.array(out, np.float32)


@pytest.mark.parametrize("rs", [(20, 60), (40, 120), (10, 36)])
def test_tables_reproduce_the_oracle_bins(rs):
    from scl_slam_b200 import engine
    R, S = rs
    tab = engine.polar_tables(R, S, 80.0)
    assert len(tab["ring_thr"]) == R - 1 and np.all(np.diff(tab["ring_thr"]) > 0)
    for q in range(4):
        assert np.all(np.diff(tab["sec_thr"][q]) > 0)
    rng = np.random.default_rng(17)
    rad = np.exp(rng.uniform(np.log(1e-4), np.log(95.0), 300000))
    ang = rng.uniform(0, 2 * np.pi, rad.size)
    bulk = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-3, 15, rad.size)], 1).astype(np.float32)
    pts3 = np.concatenate([_edge_points(R, S, rng), bulk])
    pts = np.zeros((len(pts3), 4), np.float32); pts[:, :3] = pts3
    o = Oracle(num_ring=R, num_sector=S)
    _, oring, osector = o.make_scancontext(pts, want_bins=True)
    ring, sector = _bins_from_tables(tab, pts, S)
    assert np.array_equal(ring, oring), (int((ring != oring).sum()), pts[ring != oring][:5], ring[ring != oring][:5], oring[ring != oring][:5])
    assert np.array_equal(sector, osector), (int((sector != osector).sum()), pts[sector != osector][:5])


def test_tables_reject_unsupported_geometries():
    from scl_slam_b200 import engine
    with pytest.raises(RuntimeError):
        engine.polar_tables(80, 60, 80.0)          # 79 ring thresholds: more than the kernel's 63-entry table
    with pytest.raises(RuntimeError):
        engine.polar_tables(20, 60, 0.0)
