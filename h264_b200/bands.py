"""MB-row bands of ONE picture across the GPUs of a box (BASELINE config 4).

Each rank codes a contiguous band of macroblock rows (one slice per band) and therefore owns the reconstructed
luma rows of its band.  The motion search of the next picture reaches outside the band: an MB row at luma row y
reads reference rows  y + centre - R - 1 - 2 ... y + 15 + centre + R + 1 + 3  (integer window, +-1 pel of
sub-pel refinement, 6-tap filter support: getSubImagesLuma JM/lencod/src/img_luma.c:257-331).  Before every
picture the ranks therefore exchange `halo_rows(R, max_center)` reconstructed rows with their neighbours --
point-to-point sends/receives over NCCL (NVLink), gloo in the CPU tests -- and search their own MB range with
absolute MB indexing (b2me_search_mbs_dev).  There is no other data-path communication: the bands' bitstreams
are independent slices.

Host-side orchestration only (torch.distributed + the C ABI); every kernel is the library's."""
import torch
import torch.distributed as dist


def band_mb_rows(rank, world, mbh):
    """[first, last) macroblock rows of `rank`: as even as possible, the first mbh % world bands one row taller."""
    base, extra = divmod(mbh, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def halo_rows(R, max_center_pel=16):
    """Reference luma rows needed beyond a band edge: search range + the largest |centre| + sub-pel reach (1) +
    6-tap support (3)."""
    return R + max_center_pel + 4


def needed_rows(rank, world, mbh, R, max_center_pel=16):
    """[lo, hi) luma rows of the reference that `rank` reads, clipped to the picture."""
    f, l = band_mb_rows(rank, world, mbh)
    h = halo_rows(R, max_center_pel)
    return max(0, 16 * f - h), min(16 * mbh, 16 * l + h)


def exchange_plan(world, mbh, R, max_center_pel=16):
    """[(src, dst, lo, hi)]: rows [lo, hi) owned by src that dst needs (src != dst)."""
    plan = []
    for dst in range(world):
        nlo, nhi = needed_rows(dst, world, mbh, R, max_center_pel)
        for src in range(world):
            if src == dst:
                continue
            f, l = band_mb_rows(src, world, mbh)
            lo, hi = max(nlo, 16 * f), min(nhi, 16 * l)
            if lo < hi:
                plan.append((src, dst, lo, hi))
    return plan


def exchange_halos(own_band, H, W, R, max_center_pel=16, group=None, out=None):
    """own_band: uint8 [band rows, W] reconstructed luma of this rank's band (CUDA tensor with NCCL, CPU tensor with
    gloo).  Returns a [H, W] picture whose rows needed_rows(rank) are valid (the rest is zero)."""
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)     # one band: nothing to exchange
    mbh = H // 16
    f, l = band_mb_rows(rank, world, mbh)
    assert own_band.shape == (16 * (l - f), W) and own_band.dtype == torch.uint8
    full = out if out is not None else torch.zeros((H, W), dtype=torch.uint8, device=own_band.device)
    full[16 * f:16 * l].copy_(own_band)
    ops, keep = [], []
    for src, dst, lo, hi in exchange_plan(world, mbh, R, max_center_pel):
        if src == rank:
            t = own_band[lo - 16 * f:hi - 16 * f].contiguous()
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, dst, group=group))
        elif dst == rank:
            ops.append(dist.P2POp(dist.irecv, full[lo:hi], src, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full


class BandSearcher:
    """One rank's band of a W x H picture: api.Searcher for the whole picture geometry (absolute MB indexing), of
    which only the band's MB rows are searched."""

    def __init__(self, W, H, nrefs, R, rank, world, device=0, max_center_pel=16):
        from . import api
        self.api = api
        self.W, self.H, self.R, self.nrefs = W, H, R, nrefs
        self.rank, self.world, self.max_center = rank, world, max_center_pel
        self.mbw, self.mbh = W // 16, H // 16
        self.first_row, self.last_row = band_mb_rows(rank, world, self.mbh)
        self.s = api.Searcher(W, H, nrefs, R, device=device)

    @property
    def mb_first(self):
        return self.first_row * self.mbw

    @property
    def mb_count(self):
        return (self.last_row - self.first_row) * self.mbw

    def set_cur_dev(self, cur_full, stream=0):
        self.s.set_cur_dev(cur_full, stream)

    def set_ref_from_band(self, ref_idx, own_band, stream=0, group=None):
        """Halo exchange of the reconstructed band (NCCL), then the sub-pel / search planes of the rows this band reads
        (own rows + halo: b2me_set_ref_rows_dev), in a picture buffer that is allocated once."""
        if getattr(self, "_full", None) is None or self._full.device != own_band.device:
            self._full = torch.zeros((self.H, self.W), dtype=torch.uint8, device=own_band.device)
        full = exchange_halos(own_band, self.H, self.W, self.R, self.max_center, group=group, out=self._full)
        lo, hi = needed_rows(self.rank, self.world, self.mbh, self.R, self.max_center)
        if full.is_cuda:
            self.s.set_ref_rows_dev(ref_idx, full, lo, hi - lo, stream)
        else:
            self.s.set_ref_dev(ref_idx, full, stream)
        return full

    def search(self, pred, center, params, mv_int, cost_int, mv_sub, cost_sub, stream=0, check=True):
        """check: refuse search centres beyond max_center_pel -- the halo was sized for that bound, a centre beyond it would
        read rows that were never exchanged (one device reduction + host read per call; pass False once the caller has
        validated its predictors)."""
        if check and self.mb_count:
            c = center[self.mb_first:self.mb_first + self.mb_count]
            worst = int(c.abs().max().item()) if c.numel() else 0
            if worst > 4 * self.max_center:
                raise ValueError(f"BandSearcher: a search centre of {worst / 4:.2f} pel exceeds max_center_pel = {self.max_center}: "
                                 f"the halo exchange covers R + {self.max_center} + 4 rows only")
        self.s.search_frame_dev(pred, center, params, mv_int, cost_int, mv_sub, cost_sub, stream,
                                mb_first=self.mb_first, mb_count=self.mb_count)
