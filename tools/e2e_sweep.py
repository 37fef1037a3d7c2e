#!/usr/bin/env python
"""End-to-end (host buffers) throughput of forward_host vs the number of pipeline chunks."""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 0), strict=True)
m.eval().cuda()
B, L = 256, 64000
x = torch.from_numpy(np.tile(synth.make_mixtures(16, L, 1234), (B // 16, 1))).pin_memory()
kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
for hc in (1, 2, 3, 4, 6, 8):
    m.set_option("host_chunks", hc)
    for _ in range(3):
        m.forward_host(x, kw)
    t0 = time.perf_counter()
    for _ in range(8):
        m.forward_host(x, kw)
    dt = (time.perf_counter() - t0) / 8
    print(f"host_chunks={hc}: {dt * 1e3:.2f} ms per 256 x 4 s -> {B * 4 / dt:.0f} audio-s/s", flush=True)
