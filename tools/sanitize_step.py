#!/usr/bin/env python
"""One small forward through every kernel of the headline path (plane layout, CTA-pair dconv, TF32 pair conv1, cluster
residual kernel) for compute-sanitizer: compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 9), strict=True)
m.eval().cuda()
m.set_option("conv1_pair", 2)
kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
x = torch.from_numpy(synth.make_mixtures(3, 40000, 1)).cuda()
out, vad, est = m(x, kw)
torch.cuda.synchronize()
print("launches", m.last_launch_count, "finite", bool(torch.isfinite(out).all()))
