"""Seeded multi-robot save sequences for the row-key candidate stage (lidar_iris_descriptor, descriptor.h:1047-1267):
trajectories whose second half revisits the first (near-duplicate keys, small compare() distance), exact duplicate keys
(libnabo's self-match rule), robots of different sizes (one below numCandidates + 1: the early returns)."""
import numpy as np


def make(seed, rows=80, sizes=(260, 180, 7), dup_every=37):
    """-> list of (key[rows] float32, robot, index, feature) in save order, robots interleaved."""
    rng = np.random.default_rng(seed)
    per_robot = []
    for r, n in enumerate(sizes):
        base = rng.uniform(0.5, 6.0, rows).astype(np.float32)
        walk = np.cumsum(rng.normal(0, 0.08, (n, rows)), axis=0).astype(np.float32)
        keys = np.abs(base[None, :] + walk).astype(np.float32)
        place = np.arange(n, dtype=np.float32)
        half = n // 2
        for t in range(half, n):                       # revisit an earlier place of this robot or of robot 0
            src = t - half
            if t % 3 == 0:
                keys[t] = keys[src] + rng.normal(0, 0.01, rows).astype(np.float32)
                place[t] = place[src]
            if dup_every and t % dup_every == 0:
                keys[t] = keys[src]                    # an exact duplicate: d2 = 0 is skipped by libnabo
                place[t] = place[src]
        if r > 0 and len(per_robot[0][0]) > 20:        # other robots pass through robot 0's places too
            k0, p0 = per_robot[0]
            for t in range(0, n, 5):
                s = (t * 7) % len(k0)
                keys[t] = k0[s] + rng.normal(0, 0.01, rows).astype(np.float32)
                place[t] = p0[s] + 1000.0 * 0          # same place id as robot 0's entry
        feat = (place * 0.5 + rng.normal(0, 0.05, n)).astype(np.float32)
        per_robot.append((keys, feat))
    order = []
    cursors = [0] * len(sizes)
    while any(c < n for c, n in zip(cursors, sizes)):
        r = int(rng.integers(0, len(sizes)))
        if cursors[r] < sizes[r]:
            i = cursors[r]
            order.append((per_robot[r][0][i].copy(), r, i, float(per_robot[r][1][i])))
            cursors[r] += 1
    return order


def run(oracle_factory, saves, robot_num, this_id, **params):
    """Replays `saves` and then every intra / inter query; -> dict of result arrays."""
    o = oracle_factory(robot_num=robot_num, this_id=this_id, **params)
    n_own = 0
    for key, r, i, f in saves:
        o.save(key, r, i, f)
        n_own += r == this_id
    K = params.get("num_candidates", 10)
    intra = [o.detect_intra(p) for p in range(n_own)]
    inter = [o.detect_inter(g) for g in range(len(saves))]

    def pack(rs):
        return dict(id=np.array([x[0] for x in rs], np.int32), bias=np.array([x[1] for x in rs], np.float32),
                    n=np.array([x[2] for x in rs], np.int32),
                    cand=np.stack([x[3] if x[2] else np.full(K, -1, np.int32) for x in rs]),
                    d2=np.stack([x[4] if x[2] else np.full(K, np.inf, np.float32) for x in rs]))
    return {"intra": pack(intra), "inter": pack(inter), "index": np.array([o.get_index(g) for g in range(len(saves))], np.int32),
            "sizes": np.array([o.size(-1)] + [o.size(r) for r in range(robot_num)], np.int32)}
