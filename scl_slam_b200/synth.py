"""Seeded synthetic inputs for the Scan Context loop-closure path.

The reference ships no data (its only "test" replays external rosbags,
/root/reference/launch/test_distributed_loop.launch:37-42), so every input shape the benches
and parity tests need is generated here:

  world / lidar_dirs / scan / trajectory : D1 (VLP-16), D2 (HDL-64), D4 (Livox Horizon) clouds
  desc_db / desc_queries                 : D3 descriptor database + perturbed, rotated queries

Everything is a pure function of its seed. The descriptor database generator is written with
device-agnostic torch ops and an integer hash instead of a device RNG, so the same code fills
a 1M-entry database directly in HBM on the GPU box and a 2k-entry one on the CPU in tests.
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# integer-hash uniform numbers (identical on CPU and CUDA)
# --------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def _hash32(x):
    """lowbias32 on an int64 tensor holding 32-bit values."""
    x = x & _M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32
    x = x ^ (x >> 16)
    return x


def _uniform(idx, salt):
    """U[0,1) float32 from an int64 index tensor and an integer salt."""
    h = _hash32(idx * 0x9E3779B1 + (salt * 0x85EBCA6B & _M32))
    h = _hash32(h + salt)
    return (h >> 8).to(torch.float32) * (1.0 / 16777216.0)


# --------------------------------------------------------------------------------------------
# D3: descriptor database and queries
# --------------------------------------------------------------------------------------------
def desc_db(n, num_ring=20, num_sector=60, seed=3, device="cpu", start=0, chunk=65536, out=None):
    """n descriptors [n, R, S] float32: a smooth random height field per entry (4 sinusoid
    products over (ring, sector) + per-bin noise) scaled into [0, 25] m, radial occlusion runs
    (bins beyond a per-sector start ring zeroed), whole sectors zeroed with p = 0.05 (exercises
    the zero-norm skip of distDirectSC, descriptor.h:1523). `start` offsets the entry ids so
    shards of one database can be generated independently."""
    R, S = num_ring, num_sector
    dev = torch.device(device)
    if out is None:
        out = torch.empty((n, R, S), dtype=torch.float32, device=dev)
    r = torch.arange(R, device=dev, dtype=torch.float32).view(1, 1, R) / R
    s = torch.arange(S, device=dev, dtype=torch.float32).view(1, 1, S) / S
    bins = torch.arange(R * S, device=dev, dtype=torch.int64).view(1, R * S)
    secs = torch.arange(S, device=dev, dtype=torch.int64).view(1, S)
    comp = torch.arange(4, device=dev, dtype=torch.int64).view(1, 4)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        e = torch.arange(start + c0, start + c1, device=dev, dtype=torch.int64).view(-1, 1) + seed * 0x1000003
        ek = e * 4 + comp                                        # [m,4]
        amp = 1.0 + 3.0 * _uniform(ek, 11)
        fr = torch.floor(1.0 + 3.0 * _uniform(ek, 12)) * 0.5     # half-integer ring frequencies
        fs = torch.floor(1.0 + 4.0 * _uniform(ek, 13))           # integer sector frequencies (periodic)
        pr = _uniform(ek, 14)
        ps = _uniform(ek, 15)
        fr_t = torch.sin(2 * math.pi * (fr.unsqueeze(-1) * r + pr.unsqueeze(-1)))   # [m,4,R]
        fs_t = torch.sin(2 * math.pi * (fs.unsqueeze(-1) * s + ps.unsqueeze(-1)))   # [m,4,S]
        field = torch.einsum("mk,mkr,mks->mrs", amp, fr_t, fs_t)                    # [m,R,S]
        noise = _uniform(e * (R * S) + bins, 21).view(-1, R, S)
        # per-entry radial profile: what makes ring keys (row means) discriminative at 1M entries
        rings = torch.arange(R, device=dev, dtype=torch.int64).view(1, R)
        profile = 10.0 * (_uniform(e * R + rings, 24) - 0.5)                        # [m,R]
        h = (12.5 + profile.unsqueeze(-1) + 1.6 * field + 4.0 * (noise - 0.5)).clamp_(0.0, 25.0)
        # radial occlusion: per (entry, sector) a start ring beyond which nothing is seen
        u = _uniform(e * S + secs, 22)                                               # [m,S]
        start_ring = torch.where(u > 0.6, torch.full_like(u, float(R)), torch.floor(R * (0.25 + 1.25 * u)))
        ring_id = torch.arange(R, device=dev, dtype=torch.float32).view(1, R, 1)
        h = torch.where(ring_id >= start_ring.unsqueeze(1), torch.zeros_like(h), h)
        # whole-sector dropout
        dead = _uniform(e * S + secs, 23) < 0.05
        h = torch.where(dead.unsqueeze(1), torch.zeros_like(h), h)
        out[c0:c1] = h
    return out


def desc_queries(db, q, seed=4, noise_sigma=0.05, dropout=0.02):
    """q queries derived from database entries: entry `src[i]` column-rotated
    right by `shift[i]` sectors, N(0, sigma^2) added on non-zero bins, 2 % of bins dropped.
    Returns (queries [q,R,S] float32, src int64 [q], shift int64 [q]); queries live on db.device."""
    n, R, S = db.shape
    dev = db.device
    i = torch.arange(q, device=dev, dtype=torch.int64) + seed * 0x2000003
    src = (_uniform(i, 31).to(torch.float64) * n).to(torch.int64).clamp_(0, n - 1)
    shift = (_uniform(i, 32).to(torch.float64) * S).to(torch.int64).clamp_(0, S - 1)
    base = db[src]                                                        # [q,R,S]
    cols = (torch.arange(S, device=dev).view(1, S) - shift.view(-1, 1)) % S
    rot = torch.gather(base, 2, cols.view(q, 1, S).expand(q, R, S))       # out[:, :, c] = in[:, :, (c - shift) % S]
    bins = torch.arange(R * S, device=dev, dtype=torch.int64).view(1, R * S)
    u1 = _uniform(i.view(-1, 1) * (R * S) + bins, 33).clamp_min_(1e-7)
    u2 = _uniform(i.view(-1, 1) * (R * S) + bins, 34)
    g = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2 * math.pi * u2)
    noisy = rot + noise_sigma * g.view(q, R, S)
    noisy = torch.where(rot != 0, noisy, rot)
    drop = _uniform(i.view(-1, 1) * (R * S) + bins, 35).view(q, R, S) < dropout
    noisy = torch.where(drop, torch.zeros_like(noisy), noisy)
    return noisy.contiguous(), src, shift


# --------------------------------------------------------------------------------------------
# D1 / D2 / D4: clouds from a box world
# --------------------------------------------------------------------------------------------
def make_world(seed=1, n_boxes=600, area=400.0):
    """Ground plane z = 0 plus n axis-aligned boxes (w, d in U[2,20] m, h in U[2,15] m)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-area / 2, area / 2, size=(n_boxes, 2))
    wd = rng.uniform(2.0, 20.0, size=(n_boxes, 2))
    h = rng.uniform(2.0, 15.0, size=(n_boxes, 1))
    lo = np.concatenate([c - wd / 2, np.zeros((n_boxes, 1))], axis=1)
    hi = np.concatenate([c + wd / 2, h], axis=1)
    return lo.astype(np.float32), hi.astype(np.float32)


def lidar_dirs(kind="vlp16", n_az=None, seed=7):
    """Unit ray directions [P,3] in the sensor frame."""
    if kind == "vlp16":
        elev = np.deg2rad(np.arange(-15.0, 15.1, 2.0))
        n_az = n_az or 1800
    elif kind == "hdl64":
        elev = np.deg2rad(np.linspace(-24.9, 2.0, 64))
        n_az = n_az or 1875
    elif kind == "livox":
        # Livox Horizon: 81.7 x 25.1 deg forward FoV, non-repetitive rosette-like pattern
        n = n_az or 24000
        rng = np.random.default_rng(seed)
        t = np.arange(n) / n
        ph = rng.uniform(0, 2 * np.pi)
        az = np.deg2rad(81.7 / 2) * np.sin(2 * np.pi * 61.0 * t + ph) * np.cos(2 * np.pi * 7.0 * t)
        el = np.deg2rad(25.1 / 2) * np.sin(2 * np.pi * 97.0 * t + 0.5 * ph)
        d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
        return d.astype(np.float32)
    else:
        raise ValueError(kind)
    az = np.linspace(0, 2 * np.pi, n_az, endpoint=False)
    e, a = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)], axis=-1).reshape(-1, 3)
    return d.astype(np.float32)


def scan(world, pose, dirs, seed=0, sensor_height=1.65, max_range=100.0, sigma=0.02, dropout=0.05,
         box_radius=120.0):
    """Ray-cast one scan. pose = (x, y, yaw). Returns [P',4] float32 x,y,z,intensity in the
    sensor frame (ground at z = -sensor_height, matching LIDAR_HEIGHT, descriptor.h:1312)."""
    lo, hi = world
    x, y, yaw = pose
    rng = np.random.default_rng(seed)
    cy, sy = math.cos(yaw), math.sin(yaw)
    d = dirs.astype(np.float32)
    dw = np.stack([cy * d[:, 0] - sy * d[:, 1], sy * d[:, 0] + cy * d[:, 1], d[:, 2]], axis=1)
    o = np.array([x, y, sensor_height], np.float32)
    t = np.full(d.shape[0], np.inf, np.float32)
    down = dw[:, 2] < -1e-6
    t[down] = -sensor_height / dw[down, 2]
    near = (np.abs((lo[:, 0] + hi[:, 0]) / 2 - x) < box_radius) & (np.abs((lo[:, 1] + hi[:, 1]) / 2 - y) < box_radius)
    inside = (lo[:, 0] < x) & (x < hi[:, 0]) & (lo[:, 1] < y) & (y < hi[:, 1]) & (hi[:, 2] > sensor_height)
    near &= ~inside                     # a box the sensor stands in is transparent
    blo, bhi = lo[near], hi[near]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dw
        for c0 in range(0, d.shape[0], 8192):
            sl = slice(c0, min(d.shape[0], c0 + 8192))
            t1 = (blo[None, :, :] - o[None, None, :]) * inv[sl, None, :]
            t2 = (bhi[None, :, :] - o[None, None, :]) * inv[sl, None, :]
            tmin = np.nanmax(np.minimum(t1, t2), axis=2)
            tmax = np.nanmin(np.maximum(t1, t2), axis=2)
            hit = (tmax >= np.maximum(tmin, 0.0))
            th = np.where(hit, np.maximum(tmin, 0.0), np.inf).min(axis=1) if blo.shape[0] else np.full(tmin.shape[0], np.inf)
            t[sl] = np.minimum(t[sl], th.astype(np.float32))
    t = t + rng.normal(0.0, sigma, size=t.shape).astype(np.float32)
    keep = np.isfinite(t) & (t < max_range) & (t > 0.5) & (rng.uniform(size=t.shape) >= dropout)
    p = d[keep] * t[keep, None]
    inten = rng.uniform(0, 255, size=(p.shape[0], 1)).astype(np.float32)
    return np.concatenate([p, inten], axis=1).astype(np.float32)


def trajectory(n_keyframes=2000, spacing=1.0, width=300.0, height=200.0, seed=1, reverse_fraction=0.15):
    """Two laps of a rounded rectangle at `spacing` m per keyframe; lap 2 is offset laterally by
    U[-0.5, 0.5] m and a segment of it is driven in reverse (yaw + 180 deg) to exercise shifts.
    Returns [n,3] float64 (x, y, yaw)."""
    rng = np.random.default_rng(seed)
    per_lap = n_keyframes // 2
    scale = per_lap * spacing / (2 * (width + height))
    w, h = width * scale, height * scale
    rad = 0.15 * min(w, h)
    # perimeter parameterisation of a rounded rectangle centred at the origin
    segs = [(w - 2 * rad), (math.pi / 2) * rad, (h - 2 * rad), (math.pi / 2) * rad] * 2
    perim = sum(segs)

    def at(sdist):
        sdist = sdist % perim
        corners = [(w / 2 - rad, -h / 2 + rad), (w / 2 - rad, h / 2 - rad), (-w / 2 + rad, h / 2 - rad), (-w / 2 + rad, -h / 2 + rad)]
        starts = [(-w / 2 + rad, -h / 2, 0.0), None, (w / 2, -h / 2 + rad, math.pi / 2), None,
                  (w / 2 - rad, h / 2, math.pi), None, (-w / 2, h / 2 - rad, -math.pi / 2), None]
        for k, L in enumerate(segs):
            if sdist <= L or k == len(segs) - 1:
                if k % 2 == 0:
                    x0, y0, yaw = starts[k]
                    return x0 + math.cos(yaw) * sdist, y0 + math.sin(yaw) * sdist, yaw
                cx, cy = corners[k // 2]
                a0 = -math.pi / 2 + (k // 2) * math.pi / 2
                a = a0 + sdist / rad
                return cx + rad * math.cos(a), cy + rad * math.sin(a), a + math.pi / 2
            sdist -= L
        raise AssertionError

    out = np.zeros((n_keyframes, 3))
    rev0 = int(per_lap * 0.3)
    rev1 = rev0 + int(per_lap * reverse_fraction)
    for i in range(n_keyframes):
        lap, k = divmod(i, per_lap)
        x, y, yaw = at(k * perim / per_lap)
        if lap >= 1:
            off = rng.uniform(-0.5, 0.5)
            x += -math.sin(yaw) * off
            y += math.cos(yaw) * off
            if rev0 <= k < rev1:
                yaw += math.pi
        out[i] = (x, y, (yaw + math.pi) % (2 * math.pi) - math.pi)
    return out


def to_pcl_xyzi(points):
    """[P,4] x,y,z,intensity -> [P,8] float32 with the 32-byte pcl::PointXYZI layout
    (x,y,z,pad,intensity,pad,pad,pad) that makeScancontext receives (descriptor.h:1404)."""
    p = np.zeros((points.shape[0], 8), np.float32)
    p[:, 0:3] = points[:, 0:3]
    p[:, 3] = 1.0
    p[:, 4] = points[:, 3] if points.shape[1] > 3 else 0.0
    return p


# --------------------------------------------------------------------------------------------
# the same ray caster in torch (runs on the GPU box): batches of scans for the C1 / C4 / C5 benches
# --------------------------------------------------------------------------------------------
def scan_batch_torch(world, poses, dirs, seed=0, device="cuda", sensor_height=1.65, max_range=100.0, sigma=0.02, dropout=0.05,
                     box_radius=120.0, pcl_layout=True, rays_per_chunk=1 << 15):
    """Ray-casts len(poses) scans of the box world on `device`. Same geometry as scan() (ground plane + axis-aligned
    boxes, range noise, dropout), noise from the integer hash (so CPU and CUDA give the same clouds for a seed).
    Returns (points, offsets): points [sum P_i, 8] float32 in the pcl::PointXYZI layout (or [.., 4] x, y, z, intensity
    with pcl_layout=False) in the SENSOR frame, offsets int32 [n + 1]."""
    dev = torch.device(device)
    lo = torch.as_tensor(world[0], device=dev)
    hi = torch.as_tensor(world[1], device=dev)
    d = torch.as_tensor(dirs, device=dev, dtype=torch.float32)
    poses = np.asarray(poses, np.float64).reshape(-1, 3)
    clouds, counts = [], []
    P = d.shape[0]
    for si, (x, y, yaw) in enumerate(poses):
        cy, sy = math.cos(yaw), math.sin(yaw)
        dw = torch.stack([cy * d[:, 0] - sy * d[:, 1], sy * d[:, 0] + cy * d[:, 1], d[:, 2]], dim=1)
        o = torch.tensor([x, y, sensor_height], device=dev, dtype=torch.float32)
        t = torch.full((P,), float("inf"), device=dev)
        down = dw[:, 2] < -1e-6
        t = torch.where(down, -sensor_height / dw[:, 2].clamp(max=-1e-6), t)
        cx, cyb = (lo[:, 0] + hi[:, 0]) / 2, (lo[:, 1] + hi[:, 1]) / 2
        near = ((cx - x).abs() < box_radius) & ((cyb - y).abs() < box_radius)
        inside = (lo[:, 0] < x) & (x < hi[:, 0]) & (lo[:, 1] < y) & (y < hi[:, 1]) & (hi[:, 2] > sensor_height)
        near &= ~inside
        blo, bhi = lo[near], hi[near]
        if blo.shape[0]:
            inv = 1.0 / dw
            for c0 in range(0, P, rays_per_chunk):
                sl = slice(c0, min(P, c0 + rays_per_chunk))
                t1 = (blo[None] - o[None, None]) * inv[sl, None]
                t2 = (bhi[None] - o[None, None]) * inv[sl, None]
                tmin = torch.nan_to_num(torch.minimum(t1, t2), nan=-float("inf")).amax(dim=2)
                tmax = torch.nan_to_num(torch.maximum(t1, t2), nan=float("inf")).amin(dim=2)
                hit = tmax >= tmin.clamp(min=0.0)
                th = torch.where(hit, tmin.clamp(min=0.0), torch.full_like(tmin, float("inf"))).amin(dim=1)
                t[sl] = torch.minimum(t[sl], th)
        idx = torch.arange(P, device=dev, dtype=torch.int64) + (seed * 1000003 + si) * 0x10001
        u1 = _uniform(idx, 41).clamp_min(1e-7)
        u2 = _uniform(idx, 42)
        t = t + sigma * torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2 * math.pi * u2)
        keep = torch.isfinite(t) & (t < max_range) & (t > 0.5) & (_uniform(idx, 43) >= dropout)
        p = d[keep] * t[keep, None]
        inten = 255.0 * _uniform(idx[keep], 44)
        if pcl_layout:
            c = torch.zeros((p.shape[0], 8), device=dev, dtype=torch.float32)
            c[:, 0:3] = p
            c[:, 3] = 1.0
            c[:, 4] = inten
        else:
            c = torch.cat([p, inten[:, None]], dim=1)
        clouds.append(c)
        counts.append(c.shape[0])
    offsets = np.zeros(len(counts) + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    pts = torch.cat(clouds) if clouds else torch.zeros((0, 8 if pcl_layout else 4), device=dev)
    return pts.contiguous(), offsets


def desc_db_trajectory(n, num_ring=20, num_sector=60, seed=5, device="cpu", start=0, run=16):
    """A database ordered like a trajectory (the robustness arm of the bench): every `run` consecutive entries are one place
    seen from slightly different poses — the entry of desc_db at index i // run with fresh N(0, 0.1^2) noise on its
    non-empty bins, 1 % bin dropout and a yaw drift of up to one sector — so neighbouring keys are near-duplicates
    (what D1 / D2 look like to the ring-key search)."""
    R, S = num_ring, num_sector
    dev = torch.device(device)
    g0, g1 = start // run, (start + n + run - 1) // run
    places = desc_db(g1 - g0, R, S, seed=seed, device=dev, start=g0)
    i = torch.arange(start, start + n, device=dev, dtype=torch.int64)
    base = places[(i // run - g0)]
    drift = ((_uniform(i, 51) * 3).to(torch.int64) - 1).view(-1, 1)                    # -1, 0, +1 sectors
    cols = (torch.arange(S, device=dev).view(1, S) - drift) % S
    rot = torch.gather(base, 2, cols.view(n, 1, S).expand(n, R, S))
    bins = torch.arange(R * S, device=dev, dtype=torch.int64).view(1, R * S)
    u1 = _uniform(i.view(-1, 1) * (R * S) + bins, 52).clamp_min_(1e-7)
    u2 = _uniform(i.view(-1, 1) * (R * S) + bins, 53)
    g = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2 * math.pi * u2)
    out = torch.where(rot != 0, rot + 0.1 * g.view(n, R, S), rot)
    drop = _uniform(i.view(-1, 1) * (R * S) + bins, 54).view(n, R, S) < 0.01
    return torch.where(drop, torch.zeros_like(out), out).contiguous()


def desc_db_smooth(n, num_ring=20, num_sector=60, seed=5, device="cpu", start=0, lap=400_000, chunk=65536):
    """A database that is ONE long smooth trajectory (the second robustness arm of the bench): keyframe i sees the scene of
    desc_db entry `seed` with every ring raised or lowered by a slowly varying amount (three sinusoids per ring with periods of
    3 000 - 30 000 keyframes, amplitudes of 0.5 - 2 m), plus N(0, 0.02^2) per non-empty bin; the robot drives the same loop again
    every `lap` keyframes, so places are revisited. Neighbouring ring keys differ by millimetres: the K nearest keys of a query
    are its temporal neighbours, in the same key tile, and thousands of keys lie within any loose bound."""
    R, S = num_ring, num_sector
    dev = torch.device(device)
    base = desc_db(1, R, S, seed=seed, device=dev)[0]                                     # [R,S]
    rk = torch.arange(R * 3, device=dev, dtype=torch.int64) + seed * 0x3000005
    amp = (0.5 + 1.5 * _uniform(rk, 61)).view(1, R, 3)
    per = (3000.0 + 27000.0 * _uniform(rk, 62)).view(1, R, 3)
    ph = _uniform(rk, 63).view(1, R, 3)
    bins = torch.arange(R * S, device=dev, dtype=torch.int64).view(1, R * S)
    out = torch.empty((n, R, S), dtype=torch.float32, device=dev)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        i = torch.arange(start + c0, start + c1, device=dev, dtype=torch.int64)
        t = (i % lap).to(torch.float64).view(-1, 1, 1)
        drift = (amp.double() * torch.sin(2 * math.pi * (t / per.double() + ph.double()))).sum(-1).float()      # [m,R]
        u1 = _uniform(i.view(-1, 1) * (R * S) + bins, 64).clamp_min_(1e-7)
        u2 = _uniform(i.view(-1, 1) * (R * S) + bins, 65)
        g = (torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2 * math.pi * u2)).view(-1, R, S)
        v = base.unsqueeze(0) + drift.unsqueeze(-1) + 0.02 * g
        out[c0:c1] = torch.where(base.unsqueeze(0) != 0, v.clamp_min(0.05), torch.zeros_like(v))
    return out
