"""Frames of rank 0's tile set of N GPUs rendered back to back on one GPU: one at a time (spt_render) against pipelined
(spt_render_begin k + D before spt_render_end k, D + 1 films; D = PIPE_DEPTH, default 1 = two frames in flight).
Wall clock per frame over 40 frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pbrt_v2_spectral_b200 import capi, ctypes_defs as D, multi
from pbrt_v2_spectral_b200.scene_io import LoweredScene

root = os.path.join(os.path.dirname(__file__), "..", "..")
lowered = LoweredScene.load(os.path.join(root, "assets", "_lowered", "killeroo_path.spt"))
scene = capi.Scene(lowered)
depth = int(os.environ.get("PIPE_DEPTH", "1"))
films = [capi.Film(lowered.film) for _ in range(depth + 2)]
K = 40
for nranks in [int(x) for x in os.environ.get("PIPE_RANKS", "1,2,4,8").split(",")]:
    rp = D.SptRenderParams.from_buffer_copy(bytes(lowered.params)); rp.seed = 1
    rp = multi.rank_params(rp, 0, nranks); rp.wave_pixels = 0
    for _ in range(3): scene.render(films[0], rp)
    t0 = time.perf_counter(); dev = 0.0
    for k in range(K):
        scene.render(films[k % 3], rp); dev += scene.render_ms()
    t_sync = (time.perf_counter() - t0) / K * 1e3
    for rep in range(2):                       # the first pass sizes the wave buffers of the lanes the pipelined frames use
        t0 = time.perf_counter()
        for k in range(min(depth, K)): scene.render_begin(films[k % len(films)], rp)
        for k in range(K):
            if k + depth < K: scene.render_begin(films[(k + depth) % len(films)], rp)
            scene.render_end()
        t_pipe = (time.perf_counter() - t0) / K * 1e3
    print("N=%d  device time of a frame %.3f ms | wall per frame, one at a time %.3f ms | pipelined (%d in flight) %.3f ms" % (nranks, dev / K, t_sync, depth + 1, t_pipe), flush=True)
