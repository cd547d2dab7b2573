"""
Batched data-collection: what `VisualFamiliarityDataCollectionSimulation` (reference
sim/simulation.py:2546-2722) produces one Python iteration per view, in two render calls (SURVEY.md 8f,
row N2).  The output dictionary keeps the reference's `.npz` schema

    ommatidia_out (L, N, C), xyz_out (L, 4)   -- the outbound route phase (simulation.py:2640-2658)
    ommatidia     (V, N, C), xyz     (V, 4)   -- the grid / parallel-lines phase, iteration order of
                                                 simulation.py:2675-2701;  xyz rows are [x, y, z, yaw_deg]

and the file naming of the example drivers (examples/create_familiarity_map_data.py:25-27), so that
`VisualFamiliarityParallelExplorationSimulation` (simulation.py:2925-2943) can load it.
"""
import os

import numpy as np

from . import views


def dataset_name(nb_scans, nb_rows, nb_cols, ant_no, route_no, world, nb_ommatidia=None, parallel=None):
    """examples/create_familiarity_map_data.py:25-27 / create_parallel_familiarity_lines_data.py naming."""
    base = "dataset-scan%d-" % nb_scans
    base += ("parallel%d-" % parallel) if parallel is not None else ("rows%d-cols%d-" % (nb_rows, nb_cols))
    base += "ant%d-route%d-%s" % (ant_no, route_no, world.__class__.__name__.lower())
    return base + (("-omm%d" % nb_ommatidia) if nb_ommatidia is not None else "")


def collect_familiarity_dataset(route, eye, world, sky=None, method="grid", nb_orientations=16, nb_rows=100,
                                nb_cols=100, nb_parallel=21, disposition_step=0.02, block=None, filename=None):
    """
    route: (L, 4) [x, y, z, yaw_deg]; eye: invertsy_b200.sense.CompoundEye.
    `block=(a, b)` restricts the second phase to positions [a, b) (multi-GPU sharding, all headings of a
    position stay together); the route phase is always rendered in full.
    Returns the stats dict; writes `<filename>.npz` with np.savez_compressed when a filename is given.
    """
    route = np.asarray(route, dtype=np.float64)
    stats = {}
    rpos, ryaw = views.route_views(route)
    stats["ommatidia_out"] = eye.render_batch(world, sky, rpos, ryaw)
    stats["xyz_out"] = np.c_[rpos, ryaw[:, 0]]
    if method == "grid":
        pos, yaw = views.grid_views(nb_rows, nb_cols, nb_orientations, z=route[0, 2])
    elif method == "parallel":
        pos, yaw = views.parallel_views(route, nb_orientations, nb_parallel, disposition_step)
    else:
        raise ValueError("method must be 'grid' or 'parallel'")
    if block is not None:
        pos, yaw = pos[block[0]:block[1]], yaw[block[0]:block[1]]
    stats["ommatidia"] = eye.render_batch(world, sky, pos, yaw)
    stats["xyz"] = np.c_[np.repeat(pos, yaw.shape[1], axis=0), yaw.reshape(-1)]
    if filename is not None:
        path = filename if filename.endswith(".npz") else filename + ".npz"
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        np.savez_compressed(path, **stats)
    return stats
