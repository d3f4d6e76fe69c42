"""Seeded fuzz of the neighbour search against the oracle: small random clouds that aim at the rarely taken branches
(boxes of 1..14 cell columns with a sliver last column, unwrapped and negative x, points on cell boundaries, on the
periodic seam, coincident points, pairs at exactly the cut-off distance, dense clusters, a single fluid particle).
Every case must reproduce the reference's seven pair columns bit for bit and in order (nb:312-392).

Also run against the CPU functional simulator by tests/test_cpusim.py (written after round 1's GPU budget was
spent; the arithmetic that decides membership is IEEE and unfused on both sides).
"""
import numpy as np
import pytest

from parity_util import PAIR_NAMES

pytestmark = pytest.mark.gpu


def _cloud(rng, kind, n, DL, height, h):
    cell = 2.0 * h
    if kind == "uniform":
        pos = rng.random((n, 2)) * np.array([DL, height])
    elif kind == "clusters":
        k = int(rng.integers(1, 5))
        centres = rng.random((k, 2)) * np.array([DL, height])
        pos = centres[rng.integers(0, k, n)] + rng.normal(0.0, 0.6 * h, (n, 2))
        pos[:, 1] = np.abs(pos[:, 1])
    elif kind == "lattice":
        m = max(1, int(np.sqrt(n)))
        gx, gy = np.meshgrid((np.arange(m) + 0.5) * DL / m, (np.arange(-(-n // m)) + 0.5) * height / m)
        pos = np.column_stack([gx.ravel(), gy.ravel()])[:n] + rng.normal(0.0, 0.05 * h, (n, 2))
    elif kind == "boundaries":       # coordinates on multiples of the cell size and of h
        pos = np.column_stack([rng.integers(0, max(2, int(DL / cell) + 2), n) * cell,
                               rng.integers(0, max(2, int(height / h) + 1), n) * h])
        pos += (rng.random((n, 2)) < 0.5) * rng.normal(0.0, 0.3 * h, (n, 2))
    elif kind == "seam":             # hugging x = 0 and x = DL from both sides, some exactly on them
        side = rng.integers(0, 4, n)
        x = np.where(side == 0, rng.random(n) * cell, np.where(side == 1, DL - rng.random(n) * cell,
                     np.where(side == 2, 0.0, DL)))
        pos = np.column_stack([x + (rng.random(n) < 0.3) * rng.normal(0.0, 1e-9, n), rng.random(n) * height])
    else:                            # "threshold": partners at exactly the cut-off, and coincident points
        base = rng.random((n, 2)) * np.array([DL, height])
        half = n // 2
        pos = base.copy()
        pos[half:half + half // 2, 0] = base[:half // 2, 0] + cell          # r == 2h exactly (up to rounding)
        pos[half:half + half // 2, 1] = base[:half // 2, 1]
        pos[half + half // 2:] = base[:n - half - half // 2]                # coincident (r = 0: never a pair)
    if rng.random() < 0.4:           # unwrapped input: whole periods added to or taken from x
        pos[:, 0] += rng.integers(-2, 3, pos.shape[0]) * DL
    return np.asfortranarray(pos[:n].astype(np.float64))


KINDS = ("uniform", "clusters", "lattice", "boundaries", "seam", "threshold")


def cases(block, n_cases=30):
    """(label, pos, n_fluid, n_total, h, DL) of the seeded cases of one block."""
    rng = np.random.default_rng(9000 + block)
    for case in range(n_cases):
        kind = KINDS[(block + case) % len(KINDS)]
        h = float(rng.choice([0.02, 0.05, 0.13, 0.2]))
        ncols = int(rng.integers(1, 15))
        DL = (ncols - float(rng.choice([0.0, 0.37, 0.999]))) * 2.0 * h if ncols > 1 else float(rng.uniform(0.5, 1.9)) * h
        height = float(rng.integers(1, 7)) * 2.0 * h * float(rng.uniform(0.3, 1.0))
        n = int(rng.integers(2, 400))
        nf = int(rng.integers(1, n + 1))
        pos = _cloud(rng, kind, n, DL, height, h)
        yield f"block {block} case {case} ({kind}, n={n}, nf={nf}, h={h}, DL={DL})", pos, nf, n, h, DL


@pytest.mark.parametrize("block", range(6))
def test_neighbor_fuzz_bit_exact(sphb, oracle_be, block):
    for label, pos, nf, n, h, DL in cases(block):
        want = oracle_be.neighbor_search(pos, nf, n, h, DL)
        got = sphb.sph_neighbor_search_mex(pos, nf, n, h, DL)
        for k, name in enumerate(PAIR_NAMES):
            assert np.array_equal(got[k], want[k]), f"{label}, {want[0].size} pairs: {name}"


def _ulps(x, k):
    x = np.float64(x)
    for _ in range(abs(k)):
        x = np.nextafter(x, np.inf if k > 0 else -np.inf)
    return float(x)


@pytest.mark.parametrize("seed", [3, 4])
def test_neighbor_fuzz_ulps_around_thresholds(sphb, oracle_be, seed):
    """Coordinates within three ulps of the values where the reference's arithmetic changes branch (0, the cell
    size and its multiples, DL - 2h, DL, DL + 2h, -2h), box lengths within three ulps of a whole number of cells."""
    rng = np.random.default_rng(seed)
    for case in range(150):
        h = float(rng.choice([0.05, 0.1, 0.13, 0.0325, 0.065]))
        cell = 2.0 * h
        ncols = int(rng.integers(3, 12))
        DL = _ulps(ncols * cell, int(rng.integers(-3, 4)))
        n = int(rng.integers(20, 200))
        nf = int(rng.integers(1, n + 1))
        crit = [0.0, cell, DL - cell, DL, 2 * cell, DL - 2 * cell, (ncols - 1) * cell, (ncols - 2) * cell, 3 * cell,
                DL + cell, -cell, DL / 2]
        x = np.array([_ulps(rng.choice(crit), int(rng.integers(-3, 4))) for _ in range(n)])
        k = rng.random(n) < 0.3
        x[k] = rng.random(int(k.sum())) * DL
        y = np.array([_ulps(rng.choice([0.0, cell, 2 * cell, h, 0.5 * h]), int(rng.integers(-2, 3))) for _ in range(n)])
        k = rng.random(n) < 0.3
        y[k] = rng.random(int(k.sum())) * int(rng.integers(1, 5)) * cell
        pos = np.asfortranarray(np.column_stack([x, y]))
        want = oracle_be.neighbor_search(pos, nf, n, h, DL)
        got = sphb.sph_neighbor_search_mex(pos, nf, n, h, DL)
        for kk, name in enumerate(PAIR_NAMES):
            assert np.array_equal(got[kk], want[kk]), f"seed {seed} case {case} (h={h}, {ncols} columns, DL={DL!r}): {name}"


def test_neighbor_wrapped_x_equal_to_DL(sphb, oracle_be):
    """nb:270-272 wraps a tiny negative x to DL itself (x - floor(x / DL) DL rounds to DL), which files the particle
    in column 0 with x_w = DL: its neighbours in column 1 find it through the minimum image.  (Found by the fuzz.)"""
    h, DL = 0.1, 1.4000000000000001
    pos = np.asfortranarray(np.array([[0.20000000000000004, 0.4], [-1.5e-323, 0.4], [0.7, 0.4], [-1e-17, 0.45],
                                      [0.1, 0.5], [1.3, 0.45]]))
    for nf in (6, 4, 2):
        want = oracle_be.neighbor_search(pos, nf, 6, h, DL)
        got = sphb.sph_neighbor_search_mex(pos, nf, 6, h, DL)
        assert want[0].size >= 2
        for k, name in enumerate(PAIR_NAMES):
            assert np.array_equal(got[k], want[k]), (nf, name)
