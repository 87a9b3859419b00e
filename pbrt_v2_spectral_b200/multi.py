"""Multi-GPU host logic: the image is cut into square tiles dealt round-robin to ranks ("tile sets",
SptRenderParams.tile_rank / tile_nranks), every rank renders all samples of its own tiles into a
full-resolution film that is zero elsewhere, and ONE exchange at the end sums the films onto rank 0
(SURVEY.md 8e). The reference itself partitions the image into independent sub-windows
(src/renderers/samplerrenderer.cpp:203-214, src/core/sampler.cpp:47-66); there is no other data-path
collective. With the box filter the tile sets are disjoint; wider filters overlap at tile borders and
the same sum handles them.

`render_fn(params)` renders this rank's tile set into the film tensor; on GPUs it is capi.Scene.render
on a film created over the tensor's own storage (spt_film_create_external), so NCCL reduces the very
buffer K7 accumulated into."""
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


def render_distributed(render_fn, film_t, params, rank, world, tile=TILE):
    """Zero the film, render this rank's tile set, reduce. Returns the film tensor (complete on rank 0)."""
    film_t.zero_()
    render_fn(rank_params(params, rank, world, tile))
    return reduce_film(film_t)
