#!/usr/bin/env python
"""Stress check for races in the persistent / cluster kernels: the full-size forward repeated N times under
different co-scheduling conditions (alone, and with a second stream hammering the memory system) must stay
bit-identical - a missed barrier or a stage reused too early shows up as run-to-run differences."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for cfg_name, args in (("with_vad", synth.CONFIG_WITH_VAD), ("without_vad", synth.CONFIG_WITHOUT_VAD)):
    with contextlib.redirect_stdout(io.StringIO()):
        m = SeparationModel(**args)
    m.load_state_dict(synth.make_state_dict(args, 5), strict=True)
    m.eval().cuda()
    for B, L in ((256, 64000), (64, 48000), (7, 150000)):
        x = torch.from_numpy(np.tile(synth.make_mixtures(min(B, 16), L, 321), ((B + 15) // 16, 1))[:B]).cuda()
        ref = None
        side = torch.cuda.Stream()
        junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        worst = 0.0
        for i in range(N):
            if i % 2 == 1:   # disturb: a concurrent memset stream changes the timing of every kernel
                with torch.cuda.stream(side):
                    for _ in range(20):
                        junk.fill_(i)
            out, vad, _ = m(x, {})
            torch.cuda.synchronize()
            cur = (out.clone(), vad.clone() if torch.is_tensor(vad) else None)
            if ref is None:
                ref = cur
            else:
                worst = max(worst, (cur[0] - ref[0]).abs().max().item())
                if cur[1] is not None:
                    worst = max(worst, (cur[1] - ref[1]).abs().max().item())
        print(f"{cfg_name} B={B} L={L}: max run-to-run difference over {N} runs = {worst:.3e}", flush=True)
    del m
