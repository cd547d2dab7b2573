"""
View-batch generators: the (position, heading) sets of the reference's data-collection
simulations, produced in one vectorised step instead of one Python iteration per view.

They restate the index maths of ``VisualFamiliarityDataCollectionSimulation._step``
(reference sim/simulation.py:2660-2707) and of ``col2x / row2y / ori2yaw``
(reference sim/_helpers.py:1253-1361, float32 arithmetic included) and return arrays in the
layout the renderer wants: positions (P, 3) and headings (P, H), view b = p*H + h being the
reference's iteration j = b of the grid / parallel phase.
"""
import numpy as np


def col2x(col, nb_cols, max_meters=10.):
    """x of a grid column, float32 like the reference (sim/_helpers.py:1253-1271)."""
    return np.float32(np.asarray(col) * max_meters) / np.float32(nb_cols)


def row2y(row, nb_rows, max_meters=10.):
    """y of a grid row (sim/_helpers.py:1295-1313)."""
    return np.float32(np.asarray(row) * max_meters) / np.float32(nb_rows)


def ori2yaw(ori, nb_oris, linear=True, degrees=True):
    """Heading of orientation index ``ori`` in [-180, 180) (sim/_helpers.py:1337-1361)."""
    two_pi = np.float32(360. if degrees else (2 * np.pi))
    ori = np.asarray(ori)
    if not linear:
        ori = ((ori / nb_oris - .25) ** 3 + .25) * nb_oris
    return (np.float32(ori) * two_pi / np.float32(nb_oris) + two_pi / 2) % two_pi - two_pi / 2


def grid_views(nb_rows=100, nb_cols=100, nb_orientations=16, z=0.01, max_meters=10.):
    """
    Familiarity-map grid (method="grid", sim/simulation.py:2675-2681): iteration order is
    row-major over (row, col, ori).  Returns pos (rows*cols, 3) float64 and yaw_deg
    (rows*cols, nb_orientations) float64.
    """
    rows, cols = np.meshgrid(np.arange(nb_rows), np.arange(nb_cols), indexing="ij")
    pos = np.empty((nb_rows * nb_cols, 3), dtype=np.float64)
    pos[:, 0] = col2x(cols.ravel(), nb_cols, max_meters)
    pos[:, 1] = row2y(rows.ravel(), nb_rows, max_meters)
    pos[:, 2] = z
    yaw = ori2yaw(np.arange(nb_orientations), nb_orientations, degrees=True).astype(np.float64)
    return pos, np.broadcast_to(yaw, (pos.shape[0], nb_orientations)).copy()


def parallel_views(route, nb_orientations=16, nb_parallel=21, disposition_step=0.02):
    """
    Parallel offset routes (method="parallel", sim/simulation.py:2622-2625, 2682-2701).
    ``route`` is (L, 4) [x, y, z, yaw_deg].  Line k is displaced by
    r_k = step * ((k%2 - (k+1)%2) * ((k+1)//2)) along the normal of (route end - route start);
    iteration order is (k, route step m, ori).  Returns pos (nb_parallel*L, 3), yaw_deg (.., H).
    """
    route = np.asarray(route, dtype=np.float64)
    L = route.shape[0]
    span = route[-1, :2] - route[0, :2]
    unit = span / np.linalg.norm(span)
    normal = np.array([-unit[1], unit[0]])
    k = np.arange(nb_parallel)
    r = disposition_step * ((k % 2 - (k + 1) % 2) * ((k + 1) // 2)).astype(np.float64)
    shift = r[:, None] * normal[None, :]                                   # (K, 2)
    pos = np.empty((nb_parallel, L, 3), dtype=np.float64)
    pos[:, :, 0] = route[None, :, 0] + shift[:, None, 0]
    pos[:, :, 1] = route[None, :, 1] + shift[:, None, 1]
    pos[:, :, 2] = route[None, :, 2]
    d_yaw = ori2yaw(np.arange(nb_orientations), nb_orientations, degrees=True)
    yaw = (route[None, :, 3, None] + d_yaw[None, None, :] + 180) % 360 - 180  # (1, L, H)
    yaw = np.broadcast_to(yaw, (nb_parallel, L, nb_orientations))
    return pos.reshape(-1, 3), yaw.reshape(-1, nb_orientations).astype(np.float64)


def route_views(route):
    """One view per route step (RouteSimulation._step, sim/simulation.py:824-843): pos (L,3), yaw_deg (L,1)."""
    route = np.asarray(route, dtype=np.float64)
    return route[:, :3].copy(), route[:, 3:4].copy()


def shard_positions(nb_positions, rank, world_size):
    """
    Contiguous block of positions for one rank: all headings of a position stay on
    one GPU so its cull / projection is shared.  Returns (start, stop); blocks differ by at most 1.
    """
    base, extra = divmod(nb_positions, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_positions_interleaved(nb_positions, rank, world_size):
    """
    Every world_size-th position from `rank`: the positions of one rank are spread over the whole arena, so that all
    ranks see the same mix of dense and sparse vegetation (contiguous blocks of the heat-map grid differ by +-30 % in
    work, which is what limits strong scaling).  Returns the index array; `unshard_interleaved` undoes it.
    """
    return np.arange(rank, nb_positions, world_size)
