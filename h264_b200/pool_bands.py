"""Fractal pool matching of ONE picture across the GPUs of a box (BASELINE config 5 "at 1/2/4/8 GPUs").

Range blocks are independent (version1 searches every block of a picture before anything else, V1/src/image.c:1114-1133),
the domain pool is the whole previous reconstructed picture: the ranks split the RANGE block rows into contiguous bands and
every rank holds the full domain plane -- one NCCL broadcast of the plane (2 MB at 1080p; gloo in the CPU tests) from the
rank that owns it, after which each rank builds the pool operands itself (HBM-bound, tens of microseconds) instead of
receiving 64 B per pool entry.  The argmin is per range block, so there is no reduction across ranks; the bands' results
are gathered on one rank only if the caller asks for them.

Host-side orchestration only (torch.distributed + the C ABI); every kernel is the library's."""
import torch
import torch.distributed as dist


def range_band_rows(rank, world, rows8):
    """[first, last) 8x8-block rows of `rank`: as even as possible, the first rows8 % world bands one row taller."""
    base, extra = divmod(rows8, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def replicate_domain(domain_plane, src=0, group=None):
    """the domain plane on every rank (valid on `src` on entry): one broadcast"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(domain_plane, src=src, group=group)
    return domain_plane


class PoolBandSearcher:
    """One rank's band of range-block rows of a range_w x range_h picture against the replicated pool."""

    def __init__(self, range_w, range_h, domain_w, domain_h, pool_size, rank, world, device=0):
        from . import api
        self.rw, self.rh, self.dw, self.dh, self.nd = range_w, range_h, domain_w, domain_h, pool_size
        self.rank, self.world = rank, world
        self.first_row, self.last_row = range_band_rows(rank, world, range_h // 8)
        self.band_h = 8 * (self.last_row - self.first_row)
        self.s = api.PoolSearcher(range_w, self.band_h, domain_w, domain_h, pool_size, device=device) if self.band_h else None
        self.nr = (range_w // 8) * (self.band_h // 8)
        self.first_range = self.first_row * (range_w // 8)

    def search_dev(self, range_plane, domain_plane, out, stream=0, src=0, group=None):
        """range_plane: the full [range_h, range_w] u8 CUDA picture (only this rank's rows are read); domain_plane: [domain_h,
        domain_w] u8 CUDA tensor, valid on rank `src` (broadcast here); out = (dom i32, iso u8, aq i16, beta i16, err i64)
        CUDA tensors of nr entries: this band's results."""
        replicate_domain(domain_plane, src, group)
        if not self.band_h:
            return
        band = range_plane[8 * self.first_row:8 * self.last_row]
        self.s.set_planes_dev(band.data_ptr(), band.stride(0), domain_plane.data_ptr(), domain_plane.stride(0), stream)
        self.s.search_dev(*[t.data_ptr() for t in out], stream)

    def close(self):
        if self.s:
            self.s.close()
