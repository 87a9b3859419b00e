"""N > 1 host logic on CPU, two gloo ranks. (1) The exchange for ranks without peer access (multi.render_distributed):
each rank renders its tile set (with the oracle standing in for the device kernels, test infrastructure only), the films
are summed with dist.reduce or the sparse FilmExchange, and rank 0 holds the single-rank image. (2) The frame loop
bench.py --gpus N runs over NCCL around the peer-mapped film (multi.pipelined_frames), on stand-in film buffers."""
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


def _pipeline_worker(rank, world, port, path, ahead, n_frames):
    """Two ranks run multi.pipelined_frames over a ring of stand-in film buffers in shared memory (one cell per buffer and
    rank). begin(k, b) is the EARLIEST moment the device could write frame k into buffer b: the buffer must have been cleared
    of the frame that used it before. end(k) is the LATEST: the rank's contribution lands there. clear(b) runs on rank 0
    behind the frame's barrier: every rank's contribution to exactly that frame must be in the buffer."""
    import random
    import time
    from pbrt_v2_spectral_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_buf = ahead + 2
    buf = np.memmap(path, dtype=np.int64, mode="r+", shape=(n_buf, world))
    rnd = random.Random(17 * rank + ahead)
    in_flight, cleared, max_in_flight = [], [], 0

    def begin(k, b):
        nonlocal max_in_flight
        assert b == k % n_buf
        assert buf[b, rank] == 0, "frame %d enqueued into buffer %d before it was cleared of frame %d" % (k, b, buf[b, rank] - 1)
        in_flight.append(k)
        max_in_flight = max(max_in_flight, len(in_flight))
        time.sleep(rnd.random() * 1e-3)

    def end(k):
        assert in_flight.pop(0) == k                # frames end in the order they were begun
        time.sleep(rnd.random() * (3e-3 if rank else 1e-3))     # rank 1 is the straggler
        buf[k % n_buf, rank] = k + 1

    def clear(b):
        assert rank == 0
        k = len(cleared)                            # buffers are cleared in frame order
        assert b == k % n_buf
        assert np.all(buf[b] == k + 1), "buffer %d handed on with %s, expected every rank's frame %d" % (b, buf[b], k)
        cleared.append(k)
        buf[b] = 0

    sync_t = torch.zeros(1)
    waited = multi.pipelined_frames(n_frames, ahead, n_buf, begin, end, clear=clear, rank=rank, sync_t=sync_t)
    assert waited == n_frames and not in_flight
    assert max_in_flight == min(ahead + 1, n_frames)
    if rank == 0:
        assert cleared == list(range(n_frames - 1))              # the last frame's film is left for the caller
        assert np.all(buf[(n_frames - 1) % n_buf] == n_frames)   # ... complete
    with pytest.raises(ValueError):
        multi.pipelined_frames(4, ahead, ahead + 1, begin, end, clear=clear, rank=rank, sync_t=sync_t)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ahead", [1, 2, 3])
def test_pipelined_frames_never_write_an_uncleared_buffer(tmp_path, ahead):
    """The frame loop bench.py --gpus N runs (frames in flight over a ring of film buffers on rank 0) on two
    gloo ranks with random delays: no frame is enqueued into a buffer that still holds an earlier frame, every frame's film is
    complete when rank 0 hands it on, and `ahead` + 1 frames are in flight."""
    world, n_frames = 2, 24
    path = str(tmp_path / "films.bin")
    np.zeros((ahead + 2, world), dtype=np.int64).tofile(path)
    port = 29950 + (os.getpid() + ahead) % 40
    mp.spawn(_pipeline_worker, args=(world, port, path, ahead, n_frames), nprocs=world, join=True)


def test_pipelined_frames_single_rank_has_no_barriers():
    from pbrt_v2_spectral_b200 import multi
    log = []
    n = multi.pipelined_frames(5, 2, 1, lambda k, b: log.append(("b", k, b)), lambda k: log.append(("e", k)))
    assert n == 0
    assert log == [("b", 0, 0), ("b", 1, 0), ("b", 2, 0), ("e", 0), ("b", 3, 0), ("e", 1), ("b", 4, 0), ("e", 2), ("e", 3), ("e", 4)]
