#!/usr/bin/env python
"""Large-batch check (BASELINE.json configs[4], one GPU's shard of 8192 x 4 s at 2 GPUs): 4096 x 4 s in ONE forward -
finite, equal to the same clips run in batches of 256 up to the fp16 operand noise, and its throughput."""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
from conftest import sisdr_db
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = 64000
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 0), strict=True)
m.eval().cuda()
m.materialize.update(estimated_stfts=False, mask_per_speaker=False, spectrum=False, masks_b=False)
x = torch.from_numpy(np.tile(synth.make_mixtures(64, L, 1234), (B // 64, 1))).cuda()
out, vad, _ = m(x, {})
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(out).all()), "launches:", m.last_launch_count, "peak GB:", torch.cuda.max_memory_allocated() / 1e9)
o256, v256, _ = m(x[:256].contiguous(), {})
d = (out[:256] - o256).abs().max().item()
print("first 256 clips vs a 256-batch: |dwav|", d, "|dvad|", (vad[:256] - v256).abs().max().item(), "sisdr", sisdr_db(out[:256].cpu().numpy(), o256.cpu().numpy()))
# tiled copies of the same 64 clips must agree with each other to the noise level
print("clip 0 vs its copy at 64*k: |dwav|", max((out[0] - out[64 * k]).abs().max().item() for k in range(1, B // 64)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m(x, {})
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{B} x 4 s: {ms:.1f} ms per forward -> {B * 4 / (ms * 1e-3):.0f} audio-s/s")
