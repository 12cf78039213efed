"""Column sharding across GPUs: every function on the radiation path is column-independent, so the flattened
(lat x lon) axis is cut into contiguous blocks, one per rank, tables are replicated, and the only communication is one
all-gather that reassembles the global flux / heating-rate fields.

Host-side logic only (slicing, packing, the collective); it is backend-agnostic so the N > 1 path is exercised with gloo on
CPU in tests/test_sharding_cpu.py and runs over NCCL in bench.py.
"""
import numpy as np


def shard_bounds(ncol, world):
    """Contiguous blocks of ceil(ncol / world) columns; trailing ranks may be short or empty."""
    per = -(-ncol // world)
    return [(min(r * per, ncol), min((r + 1) * per, ncol)) for r in range(world)]


def column_axis(shape, ncol):
    """Axis that holds the columns in the engine's layouts: (..., ncol) for level-major fields, (nlev, ncol, nband) for the
    band-fastest cloud arrays."""
    if shape and shape[-1] == ncol:
        return len(shape) - 1
    if len(shape) == 3 and shape[1] == ncol:
        return 1
    raise ValueError(f"no column axis of length {ncol} in shape {shape}")


def shard_arrays(arrays, ncol, rank, world):
    """This rank's block of every array (contiguous copies, so each rank's tensors are dense as the engines require)."""
    lo, hi = shard_bounds(ncol, world)[rank]
    out = {}
    for k, a in arrays.items():
        a = np.asarray(a)
        if a.ndim == 0:
            out[k] = a
            continue
        ax = column_axis(a.shape, ncol)
        out[k] = np.ascontiguousarray(np.take(a, np.arange(lo, hi), axis=ax))
    return out


def pack_rows(tensors, names):
    """Stack the (rows, ncol_local) output fields into one (total_rows, ncol_local) tensor for a single collective."""
    import torch
    return torch.cat([tensors[k].reshape(-1, tensors[k].shape[-1]) for k in names], dim=0)


def all_gather_columns(local, names, ncol, group=None):
    """One all-gather of the packed local outputs; returns {name: (rows..., ncol)} global tensors on every rank.
    Ranks with short blocks pad to the common width (all_gather needs equal shapes); the padding is trimmed off."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bounds = shard_bounds(ncol, world)
    per = bounds[0][1] - bounds[0][0]
    packed = pack_rows(local, names)
    rows, nloc = packed.shape
    if nloc < per:
        packed = torch.cat([packed, packed.new_zeros((rows, per - nloc))], dim=1)
    packed = packed.contiguous()
    flat = packed.new_empty((world * rows, per))  # concatenation along dim 0: the form both NCCL and gloo accept
    dist.all_gather_into_tensor(flat, packed, group=group)
    gathered = flat.view(world, rows, per)
    full = torch.cat([gathered[r, :, : hi - lo] for r, (lo, hi) in enumerate(bounds)], dim=1)
    out, r0 = {}, 0
    for k in names:
        shp = tuple(local[k].shape[:-1])
        n = int(np.prod(shp)) if shp else 1
        out[k] = full[r0:r0 + n].reshape(shp + (ncol,))
        r0 += n
    return out


class PackedOutputs:
    """A rank's output fields as row slices of ONE (rows, ncol_local) device buffer, so that the engines write straight into what
    the collective sends: the global flux / heating-rate field is reassembled by a single `all_gather_into_tensor` with no packing
    copy.  `fields` = [(name, rows)]; every rank must hold the same number of columns (the all-gather needs equal shapes).
    Two instances used alternately let the gather of step i overlap the kernels of step i+1 (bench.py)."""

    def __init__(self, fields, ncol_local, world, device="cuda"):
        import torch
        self.fields = list(fields)
        self.rows = sum(n for _, n in self.fields)
        self.ncol_local, self.world = ncol_local, world
        self.buf = torch.empty((self.rows, ncol_local), dtype=torch.float64, device=device)
        self.gathered = torch.empty((world * self.rows, ncol_local), dtype=torch.float64, device=device) if world > 1 else None
        self.views, r0 = {}, 0
        for name, n in self.fields:
            self.views[name] = self.buf[r0:r0 + n]
            r0 += n
        self.pending = None

    def gather_async(self, group=None):
        import torch.distributed as dist
        if self.world > 1:
            self.pending = dist.all_gather_into_tensor(self.gathered, self.buf, group=group, async_op=True)

    def wait(self):
        if self.pending is not None:
            self.pending.wait()
            self.pending = None

    def global_field(self, name):
        """(rows, world * ncol_local) view-free copy of one gathered field (columns in rank order) -- for checks, not the hot path."""
        import torch
        self.wait()
        if self.world == 1:
            return self.views[name]
        g = self.gathered.view(self.world, self.rows, self.ncol_local)
        r0 = 0
        for n, k in self.fields:
            if n == name:
                return torch.cat([g[r, r0:r0 + k] for r in range(self.world)], dim=1)
            r0 += k
        raise KeyError(name)
