#!/usr/bin/env python
"""One cfg2-sized forward (after warm-up) for ncu: python tools/profile_step.py [B] [L] [warm]"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from septfa_b200 import synth  # noqa: E402
from septfa_b200.model import SeparationModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3
args = synth.CONFIG_WITHOUT_VAD if os.environ.get("PROFILE_CONFIG") == "without_vad" else synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 0), strict=True)
m.eval().cuda()
x = torch.from_numpy(np.tile(synth.make_mixtures(8, L, 1234), ((B + 7) // 8, 1))[:B]).cuda()
kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True) if args.get("final_vad", True) else {}
for _ in range(warm + 1):
    out, vad, est = m(x, kw)
torch.cuda.synchronize()
print("launches per forward:", m.last_launch_count, "finite:", bool(torch.isfinite(out).all()))
