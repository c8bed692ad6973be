"""The closed-form key-tile map of the attention kernel (struct TileMap in must3r_b200/csrc/attention.cu) restated in
Python and checked against the straightforward walk it replaced: same tiles, same order, same first key, valid count and
mask flag, for two key segments and a skip range anywhere (inside one segment, across both, empty).  The CUDA struct
itself is exercised by the GPU attention tests; this pins the arithmetic, including the edge cases those shapes miss."""
import random

BN = 128


def walk(nk0, nk1, lo, hi):
    out = []
    for seg, (n, base) in enumerate(((nk0, 0), (nk1, nk0))):
        t = 0
        while t * BN < n:
            l0, l1 = t * BN, min(t * BN + BN, n)
            g0, g1 = base + l0, base + l1
            if not (g0 >= lo and g1 <= hi):                       # fully masked tiles are never visited
                out.append((seg, t, g0, l1 - l0, (l1 - l0 < BN) or (g0 < hi and g1 > lo)))
            t += 1
    return out


def masked_run(n, base, T, lo, hi):
    l, h = lo - base, hi - base
    a = 0 if l <= 0 else (l + BN - 1) // BN
    b = T if h >= n else (0 if h <= 0 else h // BN)
    return a, max(0, min(b, T) - a)


def tile_map(nk0, nk1, lo, hi):
    T0, T1 = (nk0 + BN - 1) // BN, (nk1 + BN - 1) // BN
    a0 = c0 = a1 = c1 = 0
    if hi > lo:
        a0, c0 = masked_run(nk0, 0, T0, lo, hi)
        a1, c1 = masked_run(nk1, nk0, T1, lo, hi)
    s0, ns = (a0 if c0 > 0 else T0 + a1), c0 + c1
    out = []
    for i in range(T0 + T1 - ns):
        u = i + (ns if i >= s0 else 0)
        seg = 1 if u >= T0 else 0
        t = u - (T0 if seg else 0)
        n = nk1 if seg else nk0
        nv = min(BN, n - t * BN)
        g0 = (nk0 if seg else 0) + t * BN
        out.append((seg, t, g0, nv, (nv < BN) or (g0 < hi and g0 + nv > lo)))
    return out


def test_closed_form_tile_map_equals_walk():
    rng = random.Random(2)
    cases = [(768, 768, 768, 1536), (768, 1536, 768 + 768, 768 + 1536), (0, 600, 0, 300), (700, 600, 650, 900),
             (15360, 768, 15360, 16128), (129, 0, 0, 0), (128, 128, 0, 256 - 1), (300, 300, 100, 500)]
    for _ in range(20000):
        nk0 = rng.choice([0, 1, 127, 128, 129, 256, 300, 700, 768, 1536, rng.randint(0, 3000)])
        nk1 = rng.choice([0, 1, 128, 300, 600, 768, rng.randint(0, 1500)])
        if nk0 + nk1 == 0:
            continue
        if rng.random() < 0.2:
            lo = hi = 0
        else:
            lo = rng.randint(0, nk0 + nk1 - 1)
            hi = rng.randint(lo + 1, nk0 + nk1)
        cases.append((nk0, nk1, lo, hi))
    for nk0, nk1, lo, hi in cases:
        assert tile_map(nk0, nk1, lo, hi) == walk(nk0, nk1, lo, hi), (nk0, nk1, lo, hi)
