"""Multi-GPU: shard independent 15 s cycles across ranks; the only cross-rank step is a host-side gather of records.

A cycle is decoded with no data from any other cycle, so there is no data-path collective and NCCL /
NVLink are not used: one process per GPU decodes its contiguous block of cycles, and the small record arrays are
concatenated on rank 0 with torch.distributed's object gather (any backend; gloo on CPU in the tests).
"""
import numpy as np


def shard_range(n_cycles, rank, world_size):
    """Contiguous block [lo, hi) of cycle indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_cycles, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local_records, local_first_cycle, dist=None, dst=0):
    """Concatenate per-rank record arrays on `dst`, rewriting `cycle` to the global index.  Returns None elsewhere."""
    rec = np.array(local_records, copy=True)
    if len(rec):
        rec["cycle"] += local_first_cycle
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return rec
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(rec, out, dst=dst)
    if dist.get_rank() != dst:
        return None
    out = [r for r in out if len(r)]
    if not out:
        return rec[:0]
    allrec = np.concatenate(out)
    return allrec[np.argsort(allrec["cycle"], kind="stable")]
