#!/usr/bin/env python
"""Run-to-run reproducibility per engine mask (bring-up diagnostic)."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 32), strict=True)
m.eval().cuda()
for (B, L) in ((5, 9000), (1, 64000), (16, 64000)):
    x = torch.from_numpy(synth.make_mixtures(B, L, 555)).cuda()
    for eng in (7, 6, 5, 3, 0):
        m.set_engine(eng)
        runs = []
        for _ in range(3):
            out, vad, _ = m(x, {})
            runs.append((out.clone(), vad.clone(), m.masks_b.clone(), m.spectrum.clone()))
        d = lambda i: max((runs[0][i] - runs[k][i]).abs().max().item() for k in (1, 2))
        print(f"B={B} L={L} engine={eng}: out {d(0):.2e} vad {d(1):.2e} logits {d(2):.2e} spectrum {d(3):.2e}", flush=True)
