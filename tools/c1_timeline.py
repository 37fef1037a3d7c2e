#!/usr/bin/env python
"""Bring-up: globaltimer timeline of CTA 0 / the last CTA of the TMA-fed conv1 kernel (build with -DSEPTFA_C1_TIMELINE)."""
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
x = torch.from_numpy(np.tile(synth.make_mixtures(8, 64000, 1), (32, 1))).cuda()
for _ in range(3): m(x, {})
torch.cuda.synchronize()
buf = (C.c_ulonglong * 128)()
lib = _lib.load()
lib.septfa_debug_c1_timeline.argtypes = [C.POINTER(C.c_ulonglong)]
assert lib.septfa_debug_c1_timeline(buf) == 0
tl = np.array(buf, dtype=np.int64).reshape(2, 64)
t0 = tl[0, 0]
for c in range(2):
    r = tl[c]
    f = lambda v: str(int(v - t0)) if v else "-"
    print(f"cta {'0' if c == 0 else 'last'}: start {f(r[0])} trig {f(r[1])} pdl_done {f(r[2])} sync {f(r[3])} w_ready {[f(v) for v in r[4:8]]}")
    print("   a_ready per chunk:", " ".join(f(v) for v in r[8:24]))
    print("   epi (acc_full, done) per tile:", " ".join(f(v) for v in r[32:40]), " end", f(r[48]))
