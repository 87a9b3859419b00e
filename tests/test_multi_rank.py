"""N > 1 host logic on CPU: two gloo ranks each render their tile set (with the oracle standing in for
the device kernels, test infrastructure only), the films are summed with dist.reduce, and rank 0 holds
the single-rank image. Mirrors what bench.py --gpus N does over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out_path):
    import oracle_lib as O
    from pbrt_v2_spectral_b200 import ctypes_defs as D, multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lowered, _ = O.load_case(*O.golden_cases(big=False)[0][1:])
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.spp, rp.seed = 4, 11
    fd = lowered.film
    film_t = torch.zeros((fd.y_pixel_count, fd.x_pixel_count, D.NBANDS + 1), dtype=torch.float32)

    def render_fn(p):
        c, w = O.render(lowered, p)
        film_t[..., :D.NBANDS] += torch.from_numpy(c)
        film_t[..., D.NBANDS] += torch.from_numpy(w)

    multi.render_distributed(render_fn, film_t, rp, rank, world, tile=8)
    if rank == 0:
        np.save(out_path, film_t.numpy())
    # the sparse exchange (only the pixels a rank's tiles can reach) must give the same film as the full reduction
    ex = multi.FilmExchange(fd, multi.rank_params(rp, rank, world, 8), world, torch.device("cpu"), tile=8)
    multi.render_distributed(render_fn, film_t, rp, rank, world, tile=8, exchange=ex)
    if rank == 0:
        np.save(out_path + ".sparse.npy", film_t.numpy())
        assert sum(ex.counts) < 2 * film_t.shape[0] * film_t.shape[1]
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reduce_to_the_single_rank_image(tmp_path):
    import oracle_lib as O
    from pbrt_v2_spectral_b200 import ctypes_defs as D
    out = str(tmp_path / "film.npy")
    port = 29700 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    lowered, _ = O.load_case(*O.golden_cases(big=False)[0][1:])
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.spp, rp.seed, rp.tile_size = 4, 11, 8
    c, w = O.render(lowered, rp)
    assert np.array_equal(got[..., D.NBANDS], w)
    assert np.allclose(got[..., :D.NBANDS], c, rtol=1e-6, atol=1e-7)
    sparse = np.load(out + ".sparse.npy")
    assert np.array_equal(sparse[..., D.NBANDS], w)
    assert np.allclose(sparse[..., :D.NBANDS], c, rtol=1e-6, atol=1e-7)


def test_tile_owner_matches_the_render_partition():
    """Every sampler pixel belongs to exactly one rank, and it is the rank multi.tile_owner names."""
    import oracle_lib as O
    from pbrt_v2_spectral_b200 import ctypes_defs as D, multi
    lowered, _ = O.load_case(*O.golden_cases(big=False)[0][1:])
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.spp, rp.seed = 1, 3
    world, tile = 3, 8
    fd = lowered.film
    n_tiles_x = (rp.x_end - rp.x_start + tile - 1) // tile
    owner = np.full((fd.y_pixel_count, fd.x_pixel_count), -1)
    for r in range(world):
        _, w = O.render(lowered, multi.rank_params(rp, r, world, tile))
        ys, xs = np.nonzero(w)
        assert np.all(owner[ys, xs] == -1), "a pixel was rendered by two ranks"
        owner[ys, xs] = r
        for y, x in zip(ys[:200], xs[:200]):
            assert multi.tile_owner(x + fd.x_pixel_start, y + fd.y_pixel_start, rp.x_start, rp.y_start, n_tiles_x, world, tile) == r
    assert np.all(owner >= 0)


def test_film_exchange_covers_a_wide_filter():
    """FilmExchange sends the pixels a rank's tiles can reach: with a filter several pixels wide every pixel a rank's samples
    touched must be in its list, and the lists must stay a small multiple of the film."""
    import ctypes as C
    import oracle_lib as O
    from pbrt_v2_spectral_b200 import ctypes_defs as D, multi
    lowered, _ = O.load_case(*O.golden_cases(big=False)[0][1:])
    fd = D.SptFilmDesc.from_buffer_copy(bytes(lowered.film))
    fd.filter_xwidth = fd.filter_ywidth = 2.0
    fd.filter_inv_xwidth = fd.filter_inv_ywidth = 0.5
    for k in range(256):
        fd.filter_table[k] = 1.0
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params))
    rp.spp, rp.seed = 1, 5
    world, tile = 3, 8
    ex = multi.FilmExchange(fd, multi.rank_params(rp, 0, world, tile), world, torch.device("cpu"), tile=tile)
    npix = fd.x_pixel_count * fd.y_pixel_count
    for r in range(world):
        _, w = O.render(lowered, multi.rank_params(rp, r, world, tile), film=fd)
        touched = set(np.flatnonzero(w.reshape(-1)).tolist())
        mine = set(ex.idx[r][:ex.counts[r]].tolist())
        assert touched <= mine, "rank %d wrote %d pixels outside its exchange list" % (r, len(touched - mine))
    assert sum(ex.counts) < 3 * npix
