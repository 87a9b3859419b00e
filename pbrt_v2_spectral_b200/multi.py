"""Multi-GPU host logic: the image is cut into square tiles dealt round-robin to ranks ("tile sets",
SptRenderParams.tile_rank / tile_nranks), every rank renders all samples of its own tiles into a
full-resolution film that is zero elsewhere, and ONE exchange at the end sums the films onto rank 0
(SURVEY.md 8e). The reference itself partitions the image into independent sub-windows
(src/renderers/samplerrenderer.cpp:203-214, src/core/sampler.cpp:47-66); there is no other data-path
collective. With the box filter the tile sets are disjoint; wider filters overlap at tile borders and
the same sum handles them.

`render_fn(params)` renders this rank's tile set into the film tensor; on GPUs it is capi.Scene.render
on a film created over the tensor's own storage (spt_film_create_external), so NCCL reduces the very
buffer K7 accumulated into.

That is the exchange for ranks WITHOUT peer access to one another's memory. On one NVLink host the
exchange is fused into the film kernel instead: the film lives on rank 0, the other ranks
open it (spt_film_ipc_export / spt_film_open_ipc) and `pipelined_frames` below is the frame loop around
it - frames kept in flight over a ring of film buffers, a one-element all-reduce as end-of-frame barrier."""
import copy

import torch
import torch.distributed as dist

TILE = 32


def rank_params(params, rank, world, tile=TILE):
    """This rank's copy of the render parameters."""
    rp = copy.copy(params) if not hasattr(params, "from_buffer_copy") else type(params).from_buffer_copy(bytes(params))
    rp.tile_rank, rp.tile_nranks, rp.tile_size = rank, world, tile
    return rp


def tile_owner(x, y, x_start, y_start, n_tiles_x, world, tile=TILE):
    """Rank that renders sampler pixel (x, y): tiles are numbered row-major and dealt round-robin."""
    return (((y - y_start) // tile) * n_tiles_x + (x - x_start) // tile) % world


def reduce_film(film_t, dst=0):
    """The one exchange of the job: sum of the per-rank films on `dst`."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_t, dst=dst, op=dist.ReduceOp.SUM)
    return film_t


class FilmExchange:
    """The same exchange moving only what can be non-zero: rank r sends the film pixels its tile set can have touched -
    its tiles grown by the filter's reach - and rank 0 ADDS them into its own film. With the box filter that is
    (34/32)^2 / N of the film per rank instead of the whole film through a reduction tree (8 GPUs, 700x700x33 floats:
    9 MB per rank instead of 65 MB), and the result is the same sum (up to the order of fp32 additions where the
    grown tiles of two ranks overlap and both are non-zero).

    film: SptFilmDesc-like (x/y_pixel_start/count, filter_x/ywidth); params: SptRenderParams-like (sample extent).
    Built once per (film, params, world); every rank computes every rank's pixel list (cheap, host side)."""

    def __init__(self, film, params, world, device, tile=TILE):
        import math
        self.world = world
        W, H = film.x_pixel_count, film.y_pixel_count
        xs, ys = film.x_pixel_start, film.y_pixel_start
        # a sample at continuous (ix, iy) reaches film pixels ceil(ix - .5 - w) .. floor(ix - .5 + w) (spectralImage.cpp:88-96)
        hx = int(math.ceil(film.filter_xwidth + 0.5)); hy = int(math.ceil(film.filter_ywidth + 0.5))
        ntx = (params.x_end - params.x_start + tile - 1) // tile
        nty = (params.y_end - params.y_start + tile - 1) // tile
        masks = [torch.zeros((H, W), dtype=torch.bool) for _ in range(world)]
        for t in range(ntx * nty):
            ty, tx = divmod(t, ntx)
            x0 = params.x_start + tx * tile - hx - xs; x1 = params.x_start + (tx + 1) * tile + hx - xs
            y0 = params.y_start + ty * tile - hy - ys; y1 = params.y_start + (ty + 1) * tile + hy - ys
            x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, W), min(y1, H)
            if x1 > x0 and y1 > y0:
                masks[t % world][y0:y1, x0:x1] = True
        idx = [m.reshape(-1).nonzero().reshape(-1) for m in masks]
        self.n = max(int(i.numel()) for i in idx) if idx else 0
        # equal lengths for the gather: padded with pixel 0, whose padded rows are zero
        self.counts = [int(i.numel()) for i in idx]
        self.idx = [torch.cat([i, i.new_zeros(self.n - i.numel())]).to(device) for i in idx]

    def run(self, film_t, rank, dst=0):
        if self.world == 1:
            return film_t
        flat = film_t.view(-1, film_t.shape[-1])
        mine = flat.index_select(0, self.idx[rank])
        if self.counts[rank] < self.n:
            mine[self.counts[rank]:] = 0
        if rank == dst:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.gather(mine, parts, dst=dst)
            for r in range(self.world):
                if r != dst:
                    flat.index_add_(0, self.idx[r], parts[r])
        else:
            dist.gather(mine, None, dst=dst)
        return film_t


def render_distributed(render_fn, film_t, params, rank, world, tile=TILE, exchange=None):
    """Zero the film, render this rank's tile set, exchange. Returns the film tensor (complete on rank 0)."""
    film_t.zero_()
    render_fn(rank_params(params, rank, world, tile))
    if exchange is not None:
        return exchange.run(film_t, rank)
    return reduce_film(film_t)


def pipelined_frames(n_frames, ahead, n_buf, begin, end, clear=None, rank=0, sync_t=None):
    """Frame loop of a host that renders frame after frame, `ahead` frames enqueued beyond the one it waits for
    (spt_render_begin / spt_render_end), on N ranks into a ring of `n_buf` film buffers that live on rank 0.

        begin(k, b)   enqueue this rank's share of frame k into buffer b = k % n_buf   (spt_render_begin)
        end(k)        wait until this rank's frame k has drained                          (spt_render_end)
        clear(b)      rank 0 only: hand frame k's complete film on / zero buffer b        (spt_film_clear_idle)

    N > 1 (torch.distributed initialised, sync_t = a one-element tensor on the ranks' device): frame k's end-of-frame barrier
    is a one-element all-reduce enqueued after end(k) - once it completes, every rank's samples of frame k are in buffer
    k % n_buf. Rank 0 calls clear((k - 1) % n_buf) - complete since barrier k - 1 - BEFORE it enters barrier k, and every
    rank enqueues frame k + ahead + 1, the next frame that writes a buffer cleared in step k (n_buf = ahead + 2), only once
    barrier k has completed. On return every barrier has completed; the last frame's buffer has NOT been cleared.
    N = 1: no barriers, clear is never called (one film may simply accumulate the frames).
    Returns the number of barriers waited for."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    ahead = max(int(ahead), 1)
    if multi and n_buf < ahead + 2:
        raise ValueError("pipelined_frames: %d frames ahead need %d film buffers" % (ahead, ahead + 2))

    def barrier():
        """Enqueue the end-of-frame barrier; returns the call that waits for it."""
        if sync_t.is_cuda:
            dist.all_reduce(sync_t)
            ev = torch.cuda.Event()
            ev.record()
            return ev.synchronize
        return dist.all_reduce(sync_t, async_op=True).wait

    waits = [None] * n_frames
    for j in range(min(ahead, n_frames)):
        begin(j, j % n_buf)
    for k in range(n_frames):
        if k + ahead < n_frames:
            if multi and k >= 1:
                waits[k - 1]()
            begin(k + ahead, (k + ahead) % n_buf)
        end(k)
        if multi:
            if rank == 0 and k >= 1:
                waits[k - 1]()
                clear((k - 1) % n_buf)
            waits[k] = barrier()
    n_waited = 0
    if multi:
        for k in range(max(n_frames - 2, 0), n_frames):
            waits[k]()
        n_waited = n_frames
    return n_waited
