#!/usr/bin/env python
"""Bring-up: clock64 timeline of CTA 0 of the tensor-core depthwise kernel (build with -DSEPTFA_DM_TIMELINE)."""
import contextlib, ctypes as C, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from septfa_b200 import synth, lib as _lib
from septfa_b200.model import SeparationModel
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 9), strict=True)
m.eval().cuda()
m.materialize.update(estimated_stfts=False, mask_per_speaker=False, spectrum=False, masks_b=False)
B, L = int(os.environ.get("TL_B", 256)), int(os.environ.get("TL_L", 64000))
x = torch.from_numpy(np.tile(synth.make_mixtures(min(B, 8), L, 1), ((B + 7) // 8, 1))[:B]).cuda()
for _ in range(3): m(x, {})
torch.cuda.synchronize()
buf = (C.c_longlong * 640)()
lib = _lib.load()
lib.septfa_debug_dm2_timeline.argtypes = [C.POINTER(C.c_longlong)]
assert lib.septfa_debug_dm2_timeline(buf) == 0
tl = np.array(buf).reshape(10, 64)
t0 = tl[9, 0]
names = ["p_issue", "w_issue", "mini_issue", "main_issue", "d1_ready(w4)", "a2_written(w4)", "epi(d2_full,released,done,-)", "", "w_ready", "start"]
names[7] = "ld_done[0:32] / computed[32:64] (w4)"
names[1] = "mini_issued"; names[8] = "main_issued"
for r in (0, 1, 2, 8, 3, 4, 7, 5, 6):
    row = tl[r]
    print(f"{names[r]:28s}", " ".join(str(int(v - t0)) if v else "-" for v in (row[:64] if r in (6, 7) else row[:32])))
print(f"{'p_full seen by mini issuer':28s}", " ".join(str(int(v - t0)) if v else "-" for v in tl[0][32:64]))
print(f"{'p_peer seen by mini issuer':28s}", " ".join(str(int(v - t0)) if v else "-" for v in tl[1][32:64]))
