"""
Multi-GPU rendering: one process per GPU (torchrun), views sharded by *position*.

The path has no exchange step (every view is a pure function of its pose -- reference
sim/simulation.py:2675-2701), so the data path needs no collective: each rank renders a contiguous
block of positions with all their headings (the cull / projection of a position is shared by its
headings, so a position never straddles ranks).  NCCL is used only to gather results -- per-view
outputs or small per-view summaries such as a familiarity heat-map row -- onto one or all ranks.

`render_fn(pos, yaw)` is any callable returning {name: tensor [B_local, ...]} -- normally
``lambda p, y: renderer.render(p, y, outputs=...)`` -- which keeps this module testable on CPU with
the gloo backend.
"""
import torch
import torch.distributed as dist

from .views import shard_positions, shard_positions_interleaved


def bind_to_gpu_numa(device_index):
    """
    Pins this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's `numa_node` and the node's
    `cpulist`), so that pinned host buffers allocated afterwards are first-touched on that node and the D2H copies of
    a rank do not cross the socket interconnect.  Returns the node, or None when the topology is not exposed.
    """
    import os
    try:
        import torch.cuda
        bus = torch.cuda.get_device_properties(device_index)
        pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % pci) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def rank_and_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard(pos, yaw, rank=None, world_size=None, group=None):
    """This rank's block of positions: (pos[a:b], yaw[a:b], (a, b))."""
    if rank is None or world_size is None:
        rank, world_size = rank_and_world(group)
    a, b = shard_positions(len(pos), rank, world_size)
    return pos[a:b], yaw[a:b], (a, b)


def gather_views(local, nb_positions, views_per_position, dst=None, group=None):
    """
    Concatenates per-rank results in position order.  `local` maps names to tensors whose first
    dimension is this rank's view count; shards may differ by one position, so tensors are padded to
    the largest shard for the collective and trimmed afterwards.  dst=None -> every rank gets the
    full result (all_gather); otherwise only rank `dst` does (others get None).
    """
    rank, world = rank_and_world(group)
    if world == 1:
        return dict(local)
    counts = [(shard_positions(nb_positions, r, world)[1] - shard_positions(nb_positions, r, world)[0])
              * views_per_position for r in range(world)]
    biggest = max(counts)
    out = {}
    for name, t in local.items():
        if t.shape[0] != counts[rank]:
            raise ValueError("%s holds %d views, this rank's shard has %d" % (name, t.shape[0], counts[rank]))
        pad = t
        if t.shape[0] < biggest:
            pad = torch.cat([t, t.new_zeros((biggest - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
        pad = pad.contiguous()
        if dst is None:
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad, group=group)
        else:
            parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
            dist.gather(pad, parts, dst=dst, group=group)
        out[name] = None if parts is None else torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return out


def render_sharded(render_fn, pos, yaw, reduce_fn=None, dst=None, group=None):
    """
    Renders this rank's positions and gathers.  `reduce_fn(outputs) -> {name: tensor [B_local, ...]}`
    may shrink what travels (e.g. per-view mean luminance for a heat map) before the gather.
    Returns (gathered dict or None on non-destination ranks, local dict, (a, b) position block).
    """
    p, y, block = shard(pos, yaw, group=group)
    local = render_fn(p, y)          # also for an empty shard: the gather needs every rank's names and shapes
    local = {k: v for k, v in local.items() if torch.is_tensor(v)}
    small = reduce_fn(local) if reduce_fn is not None else local
    views_per_position = yaw.shape[1] if getattr(yaw, "ndim", 1) > 1 else 1
    return gather_views(small, len(pos), views_per_position, dst=dst, group=group), local, block


def all_gather_interleaved(local, nb_positions, views_per_position, out=None, scratch=None, group=None):
    """
    For positions dealt round-robin (`views.shard_positions_interleaved`): `local` holds this rank's views in its own
    order, [ceil(P / world) * H] values (padded to the largest shard; the padding is ignored).  One NCCL all_gather, then
    one on-device transpose puts every view at its global index p * H + h.  Returns the [P * H] tensor (every rank).
    """
    rank, world = rank_and_world(group)
    H = views_per_position
    per = (nb_positions + world - 1) // world
    if local.numel() != per * H:
        raise ValueError("local must hold ceil(P / world) * H = %d values (pad the short shards)" % (per * H))
    if world == 1:
        return local[:nb_positions * H]
    if scratch is None:
        scratch = torch.empty(world * per * H, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(scratch, local, group=group)
    full = scratch.view(world, per, H).permute(1, 0, 2)          # [local index][rank][heading] = global position order
    if out is None:
        out = torch.empty(per * world * H, dtype=local.dtype, device=local.device)
    out.view(per, world, H).copy_(full)
    return out[:nb_positions * H]
