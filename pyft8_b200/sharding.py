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


class ShmRecordGather:
    """Single-node record gather without a copy through the process-group transport.

    One process per GPU decodes its block of cycles; the path's only cross-GPU step is "records of all ranks on rank 0".
    `gather_records` pickles every array through gloo (TCP loopback: ~100 ms for 8 x 17 MB), which would dominate a 70 ms
    decode step.  Here every rank owns a POSIX shared-memory segment sized for its worst case (capacity records, two
    alternating slots); `publish` copies the step's records into it (one memcpy, all ranks in parallel), a barrier of the
    process group orders it, and `collect` on rank 0 returns VIEWS of all ranks' segments with `cycle` rewritten to the
    global index by the publishing rank -- the records are in rank 0's address space without a second copy.

    Slots alternate per step, so rank 0 may still be reading step i while the others publish step i + 1.
    """

    HEADER = 64

    def __init__(self, dist, capacity, dtype, tag):
        from multiprocessing import shared_memory
        self.dist, self.dtype, self.capacity = dist, np.dtype(dtype), int(capacity)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.slot_bytes = self.HEADER + self.capacity * self.dtype.itemsize
        self.step = 0
        name = lambda r: f"ft8b200_{tag}_{r}"
        self.mine = shared_memory.SharedMemory(name=name(self.rank), create=True, size=2 * self.slot_bytes)
        dist.barrier()                                   # every segment exists
        self.all = [self.mine if r == self.rank else shared_memory.SharedMemory(name=name(r))
                    for r in range(self.world)] if self.rank == 0 else None
        if self.all:                                     # attached segments belong to their creators: keep Python's resource
            from multiprocessing import resource_tracker  # tracker from unlinking them a second time at exit
            for shm in self.all:
                if shm is not self.mine:
                    try:
                        resource_tracker.unregister(shm._name, "shared_memory")
                    except Exception:
                        pass
        dist.barrier()

    def _views(self, shm, slot):
        off = slot * self.slot_bytes
        n = np.ndarray(1, np.int64, shm.buf, off)
        rec = np.ndarray(self.capacity, self.dtype, shm.buf, off + self.HEADER)
        return n, rec

    def slot_array(self, step=None):
        """Record array (capacity entries) of the slot that step `step` (default: the current one) publishes from: pass it as
        `rec=` to Engine.decode_cycles so the device->host copy lands in the shared segment and `publish_inplace` has nothing
        left to copy."""
        return self._views(self.mine, (self.step if step is None else step) & 1)[1]

    def pin(self):
        """cudaHostRegister this rank's segment (pinned: the record copy-back becomes a straight DMA).  Returns True on success."""
        try:
            import ctypes
            import torch
            addr = ctypes.addressof(ctypes.c_char.from_buffer(self.mine.buf))
            rc = torch.cuda.cudart().cudaHostRegister(addr, 2 * self.slot_bytes, 0)
            self._pinned_addr = addr
            return int(rc) == 0 if not isinstance(rc, tuple) else int(rc[0]) == 0
        except Exception:
            return False

    def publish_inplace(self, n_records, first_cycle):
        """The step's records already sit in slot_array(): publish their count; `first_cycle` travels in the header and is added
        by the reader (collect), so the publishing rank touches no record."""
        n = np.ndarray(2, np.int64, self.mine.buf, (self.step & 1) * self.slot_bytes)
        n[0], n[1] = int(n_records), int(first_cycle)

    def publish(self, records, first_cycle):
        """Copy this rank's records of the step into its segment (cycle -> global index)."""
        n, rec = self._views(self.mine, self.step & 1)
        k = len(records)
        if k > self.capacity:
            raise ValueError("ShmRecordGather: more records than the segment holds")
        rec[:k] = records
        if k and first_cycle:
            rec["cycle"][:k] += first_cycle
        n[0] = k
        np.ndarray(2, np.int64, self.mine.buf, (self.step & 1) * self.slot_bytes)[1] = 0

    def collect(self):
        """Barrier, then on rank 0: list of per-rank record views of this step (rank order = cycle order).  None elsewhere."""
        self.dist.barrier()
        slot = self.step & 1
        self.step += 1
        if self.rank != 0:
            return None
        out = []
        for shm in self.all:
            hdr = np.ndarray(2, np.int64, shm.buf, slot * self.slot_bytes)
            rec = self._views(shm, slot)[1][:int(hdr[0])]
            if hdr[1]:                                   # published in place: the reader applies the cycle offset
                rec["cycle"] += int(hdr[1])
                hdr[1] = 0
            out.append(rec)
        return out

    def close(self):
        self.dist.barrier()
        if getattr(self, "_pinned_addr", None):
            try:
                import torch
                torch.cuda.cudart().cudaHostUnregister(self._pinned_addr)
            except Exception:
                pass
        if self.all:
            for r, shm in enumerate(self.all):
                if shm is not self.mine:
                    shm.close()
        self.mine.close()
        try:
            self.mine.unlink()
        except FileNotFoundError:
            pass
