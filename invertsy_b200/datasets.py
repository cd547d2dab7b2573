"""
Batched data-collection: what `VisualFamiliarityDataCollectionSimulation` (reference
sim/simulation.py:2546-2722) produces one Python iteration per view, rendered in chunks of positions
(SURVEY.md 8f, row N2).  The output keeps the reference's `.npz` schema

    ommatidia_out (L, N, C), xyz_out (L, 4)   -- the outbound route phase (simulation.py:2640-2658)
    ommatidia     (V, N, C), xyz     (V, 4)   -- the grid / parallel-lines phase, iteration order of
                                                 simulation.py:2675-2701;  xyz rows are [x, y, z, yaw_deg]

and the file naming of the example drivers (examples/create_familiarity_map_data.py:25-27), so that
`VisualFamiliarityParallelExplorationSimulation` (simulation.py:2925-2943) can load it.

The photoreceptor transform runs on the device (the render kernel's `resp` output), so only responses cross
PCIe; the second phase is rendered chunk by chunk through `isy_render_host` into one re-used pinned buffer and
streamed into the `.npz` member as it arrives (`NpzStreamWriter`): the heat-map grid of BASELINE.json configs[2]
(1.44 M views x 1000 ommatidia) never needs more host memory than one chunk unless the caller asks to keep it.
"""
import os
import zipfile

import numpy as np

from . import views


def dataset_name(nb_scans, nb_rows, nb_cols, ant_no, route_no, world, nb_ommatidia=None, parallel=None):
    """examples/create_familiarity_map_data.py:25-27 / create_parallel_familiarity_lines_data.py naming."""
    base = "dataset-scan%d-" % nb_scans
    base += ("parallel%d-" % parallel) if parallel is not None else ("rows%d-cols%d-" % (nb_rows, nb_cols))
    base += "ant%d-route%d-%s" % (ant_no, route_no, world.__class__.__name__.lower())
    return base + (("-omm%d" % nb_ommatidia) if nb_ommatidia is not None else "")


class NpzStreamWriter(object):
    """
    Writes an `.npz` (a zip of `.npy` members, what `np.load` reads and `np.savez(_compressed)` writes --
    simulation.py:648) member by member; a member can be streamed in chunks of leading rows, so the whole
    array never has to exist in host memory.
    """

    def __init__(self, path, compressed=True):
        self.path = path if path.endswith(".npz") else path + ".npz"
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        self._zip = zipfile.ZipFile(self.path, "w", zipfile.ZIP_DEFLATED if compressed else zipfile.ZIP_STORED,
                                    allowZip64=True)
        self._member = None

    def add(self, name, array):
        self.begin(name, np.shape(array), np.asarray(array).dtype)
        self.write(np.asarray(array))
        self.end()

    def begin(self, name, shape, dtype):
        assert self._member is None, "finish the open member first"
        self._member = self._zip.open(name + ".npy", "w", force_zip64=True)
        self._left = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._dtype = np.dtype(dtype)
        np.lib.format.write_array_header_2_0(self._member, {"descr": np.lib.format.dtype_to_descr(self._dtype),
                                                            "fortran_order": False, "shape": tuple(int(s) for s in shape)})

    def write(self, chunk):
        chunk = np.ascontiguousarray(chunk, dtype=self._dtype)
        self._left -= chunk.nbytes
        assert self._left >= 0, "more data than the member's shape holds"
        self._member.write(memoryview(chunk).cast("B"))

    def end(self):
        assert self._left == 0, "member is %d bytes short of its shape" % self._left
        self._member.close()
        self._member = None

    def close(self):
        if self._member is not None:
            self.end()
        self._zip.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def collect_familiarity_dataset(route, eye, world, sky=None, method="grid", nb_orientations=16, nb_rows=100,
                                nb_cols=100, nb_parallel=21, disposition_step=0.02, block=None, filename=None,
                                channels=3, chunk_views=32768, keep=True, compressed=True):
    """
    route: (L, 4) [x, y, z, yaw_deg]; eye: invertsy_b200.sense.CompoundEye.
    `block=(a, b)` restricts the second phase to positions [a, b) (multi-GPU sharding, all headings of a
    position stay together); the route phase is always rendered in full.
    channels: 3 = the eye's responses (N, 3) as the reference logs them (simulation.py:2721); 1 = PN inputs (N,)
    (clip(mean over channels), agent.py:744), a third of the bytes.
    chunk_views: views per render call of the second phase; keep=False: `ommatidia` is only written to the file
    (stats["ommatidia"] is None) -- the way to produce data sets larger than host memory.
    Returns the stats dict; writes `<filename>.npz` when a filename is given.
    """
    import torch
    route = np.asarray(route, dtype=np.float64)
    if not keep and filename is None:
        raise ValueError("keep=False needs a filename to stream the views into")
    stats = {}
    rpos, ryaw = views.route_views(route)
    stats["ommatidia_out"] = eye.render_batch(world, sky, rpos, ryaw, channels=channels).copy()
    stats["xyz_out"] = np.c_[rpos, ryaw[:, 0]]
    if method == "grid":
        pos, yaw = views.grid_views(nb_rows, nb_cols, nb_orientations, z=route[0, 2])
    elif method == "parallel":
        pos, yaw = views.parallel_views(route, nb_orientations, nb_parallel, disposition_step)
    else:
        raise ValueError("method must be 'grid' or 'parallel'")
    if block is not None:
        pos, yaw = pos[block[0]:block[1]], yaw[block[0]:block[1]]
    P, H = yaw.shape
    N = eye.nb_ommatidia
    shape = (P * H, N, 3) if channels == 3 else (P * H, N)
    stats["xyz"] = np.c_[np.repeat(pos, H, axis=0), yaw.reshape(-1)]
    writer = NpzStreamWriter(filename, compressed) if filename is not None else None
    if writer is not None:
        writer.add("ommatidia_out", stats["ommatidia_out"])
        writer.add("xyz_out", stats["xyz_out"])
        writer.add("xyz", stats["xyz"])
        writer.begin("ommatidia", shape, np.float32)
    kept = np.empty(shape, dtype=np.float32) if keep else None
    per = max(1, int(chunk_views) // H)                      # positions per chunk
    pinned = torch.empty((min(per, P) * H,) + shape[1:], dtype=torch.float32, pin_memory=torch.cuda.is_available())
    for a in range(0, P, per):
        b = min(P, a + per)
        got = eye.render_batch(world, sky, pos[a:b], yaw[a:b], channels=channels, out=pinned[:(b - a) * H])
        if writer is not None:
            writer.write(got)
        if kept is not None:
            kept[a * H:b * H] = got
    stats["ommatidia"] = kept
    if writer is not None:
        writer.end()
        writer.close()
    return stats
