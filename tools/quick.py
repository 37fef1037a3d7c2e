#!/usr/bin/env python
"""Bring-up loop: parity of a few small shapes against the oracle, then the per-kernel-class times of a 256 x 4 s forward.
Options: QUICK_OPTS="name=value,name=value" are applied with set_option; QUICK_AB="name" times the option at 0 and 1."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sisdr_db
from oracle import septfa_oracle as O
from septfa_b200 import synth
from septfa_b200.model import SeparationModel

args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 9), strict=True)
m.eval().cuda()
for kv in filter(None, os.environ.get("QUICK_OPTS", "").split(",")):
    k, v = kv.split("="); m.set_option(k, int(v))
W = O.OracleWeights(synth.make_state_dict_numpy(args, 9), args, np.float64)
shapes = ((3, 64000), (2, 40000), (1, 200000), (5, 33000)) if not os.environ.get("QUICK_NOPARITY") else ()
for B, L in shapes:
    xd = synth.make_mixtures(B, L, 4242)
    ref_out, ref_vad, _, _ = O.forward(xd, W, {})
    x = torch.from_numpy(xd).cuda()
    o1, v1, _ = m(x, {}); o2, v2, _ = m(x, {})
    torch.cuda.synchronize()
    tail = (L % 256) if (L % 256) > 200 else 0
    dvo = np.abs(v1.cpu().numpy() - ref_vad).max()
    dw = np.abs(o1.cpu().numpy() - ref_out)[..., :L - tail].max()
    print(f"B={B} L={L}: launches {m.last_launch_count} |dvad| {dvo:.2e} |dwav| {dw:.2e} sisdr {sisdr_db(o1.cpu().numpy(), ref_out):.1f} dB "
          f"finite {bool(torch.isfinite(o1).all())} reproducible {torch.equal(o1, o2) and torch.equal(v1, v2)}", flush=True)
x = torch.from_numpy(np.tile(synth.make_mixtures(8, 64000, 1), (32, 1))).cuda()
kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
if not os.environ.get("QUICK_EXPORTS"):
    m.materialize.update(estimated_stfts=False, mask_per_speaker=False, spectrum=False, masks_b=False)
ab = os.environ.get("QUICK_AB")
for mode in ((0, 1) if ab else (None,)):
    if ab: m.set_option(ab, mode)
    for _ in range(3): m(x, kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m(x, kw)
    e1.record(); torch.cuda.synchronize()
    m.set_profile(True)
    for _ in range(5): m(x, kw)
    prof = m.read_profile(); m.set_profile(False)
    print(f"{ab}={mode}: {e0.elapsed_time(e1) / 10:.3f} ms per 256 x 4 s forward;", {k: round(v[0] / 5, 3) for k, v in prof.items()}, flush=True)
