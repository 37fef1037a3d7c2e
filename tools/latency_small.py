#!/usr/bin/env python
"""Latency of small batches (BASELINE.json configs[0]-like: one 4 s mixture): device forward, CUDA events, p50 of 50."""
import contextlib, io, os, sys
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
kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
for B, L in ((1, 64000), (8, 64000), (32, 64000), (1, 960000)):
    x = torch.from_numpy(synth.make_mixtures(B, L, 1)).cuda()
    for _ in range(5):
        m(x, kw)
    ts = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m(x, kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"B={B} L={L} ({L / 16000:g} s): p50 {ts[25]:.3f} ms, p99 {ts[49]:.3f} ms -> {B * L / 16000 / (ts[25] * 1e-3):.0f} audio-s/s, launches {m.last_launch_count}", flush=True)
