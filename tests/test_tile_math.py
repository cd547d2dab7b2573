"""
Host-side statement of the arithmetic behind the column / tile walk (csrc/isy_raster.cuh: build_phases,
isy_layout.cuh: copy_columns), checked for the properties the kernel relies on -- no GPU needed:

  * with H evenly spaced headings h0 + k*delta, every (ray, heading) query falls into exactly one azimuth column, every
    column holds every ray exactly once, at heading (c - b_j) mod H;
  * the fp32 query azimuth fl(phase_j + coloff[c]) is within 1e-6 rad of the reference's wrapped yaw_j + heading_k
    (the budget kEdgeEps leaves for it is 1.9e-6);
  * the columns a copy is filed under (its azimuth box with the margins of the kernel, columns -1 and H aliasing H - 1
    and 0) contain every query whose reference azimuth lies inside the box.
"""
import numpy as np
import pytest

from oracle import isy_oracle as o

PI, TWO_PI = np.pi, 2 * np.pi
MARGIN = np.float32(1e-4)          # kBinMargin


def wrap(y):                        # (-pi, pi], as the kernel and SciPy's as_euler
    y = np.where(y > PI, y - TWO_PI, y)
    return np.where(y <= -PI, y + TWO_PI, y)


def build_phases(yaw, h0, H):
    """fp64 -> fp32 tables of build_phases: phase[j], colb[j], coloff[c + 1] for c = -1 .. H"""
    delta = TWO_PI / H
    x = yaw + h0 + PI
    b = np.floor(x * (H / TWO_PI))
    phi = x - b * delta
    low, high = phi < 0, phi >= delta
    phi = np.where(low, phi + delta, np.where(high, phi - delta, phi))
    b = np.where(low, b - 1, np.where(high, b + 1, b))
    colb = np.mod(b.astype(np.int64), H)
    coloff = ((np.arange(H + 2) - 1) * delta - PI).astype(np.float32)
    return phi.astype(np.float32), colb, coloff


def copy_columns(plo, phi, H):
    """(c_lo, number of columns) of a copy with azimuth box [plo, phi] (float32, margin included)"""
    pi_f = np.float32(PI)
    inv_delta = np.float32(1.0) / (np.float32(TWO_PI) / np.float32(H))
    glo = max(np.float32(plo), -pi_f - MARGIN) - MARGIN
    ghi = min(np.float32(phi), pi_f + MARGIN) + MARGIN
    if not glo <= ghi:
        return 0, 0
    c_lo = int(np.floor((glo + pi_f) * inv_delta))
    c_hi = int(np.floor((ghi + pi_f) * inv_delta))
    return c_lo, min(c_hi - c_lo + 1, H)


@pytest.mark.parametrize("R,H,first_deg", [(1000, 36, -180.), (1000, 36, 0.3), (400, 6, -180.), (1280, 28, 179.999),
                                           (999, 5, 17.2), (2500, 12, 7.5), (64, 128, -3.)])
def test_every_query_has_one_column_and_the_right_azimuth(R, H, first_deg):
    _, yaw = o.fibonacci_eye(R)
    heads = wrap(np.deg2rad(first_deg + np.arange(H) * 360. / H))
    hs = np.sort(heads)                                  # the kernel's sorted slots: hs[k] = h0 + k*delta
    h0, delta = hs[0], TWO_PI / H
    assert np.abs(hs - (h0 + np.arange(H) * delta)).max() < 5e-7
    phase, colb, coloff = build_phases(yaw, h0, H)
    assert (phase >= 0).all() and (phase <= np.float32(delta)).all()
    k = np.arange(H)
    col = (colb[:, None] + k[None, :]) % H               # column of (ray j, sorted heading k)
    # every column holds every ray exactly once
    assert (np.sort(col, axis=1) == k[None, :]).all()
    # the heading the walk derives for (ray j, column c) is the one that put it there
    for c in (0, H // 2, H - 1):
        ks = (c - colb) % H
        assert (col[np.arange(R), ks] == c).all()
    # fp32 query azimuth against the reference's composition
    f = phase[:, None] + coloff[col + 1]                 # float32 + float32
    assert f.dtype == np.float32
    ref = wrap(yaw[:, None] + hs[None, :])
    d = np.abs(f.astype(np.float64) - ref)
    d = np.minimum(d, TWO_PI - d)                        # (queries next to the seam may sit on its other side)
    assert d.max() < 1e-6
    # the queries whose fp32 azimuth is more than 2 pi away from the reference's are all within 3e-6 rad of the seam,
    # where the kernel redoes them in fp64
    far = np.abs(f.astype(np.float64) - ref) > 1.0
    assert (np.abs(np.abs(f[far]) - np.float32(PI)) < 3e-6).all()


@pytest.mark.parametrize("H", [5, 12, 36, 128])
def test_the_columns_of_a_copy_hold_every_query_inside_its_box(H):
    rng = np.random.RandomState(H)
    delta = TWO_PI / H
    for _ in range(4000):
        centre = rng.uniform(-PI - 0.5, PI + 0.5)
        half = rng.choice([1e-4, 1e-2, 0.1, 0.7, 2.0]) * rng.uniform(0.2, 1.0)
        lo, hi = centre - half, centre + half             # the copy's azimuth range (a seam copy reaches beyond +-pi)
        plo, phi = np.float32(lo) - MARGIN, np.float32(hi) + MARGIN
        c_lo, nc = copy_columns(plo, phi, H)
        cols = {(c_lo + i) % H for i in range(nc)}
        # queries of the domain (-pi, pi] inside the range, and the column each one is tested in
        ys = rng.uniform(max(lo, -PI + 1e-9), min(hi, PI), 8) if max(lo, -PI) < min(hi, PI) else np.array([])
        for y in ys:
            c = int(np.floor((y + PI) / delta)) % H
            assert c in cols, (lo, hi, y, c, cols)
        if nc:       # first column -1 .. H (the aliases of H - 1 and 0), never more than H columns
            assert -1 <= c_lo <= H and 1 <= nc <= H
