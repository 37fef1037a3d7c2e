#!/usr/bin/env python
"""Bring-up check of the tensor-core depthwise kernel (dconv_mma.cu): the forward with option "dconv_mma" = 1 against
the CUDA-core depthwise producer (gemm_tc.cu MODE 1, "dconv_mma" = 0) on the same buffers, and both against the oracle."""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sisdr_db
from oracle import septfa_oracle as O
from septfa_b200 import synth
from septfa_b200.model import SeparationModel

args = synth.CONFIG_WITH_VAD
swaps = [int(s) for s in os.environ.get("SWAPS", "0,1").split(",")]
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 9), strict=True)
m.eval().cuda()
W = O.OracleWeights(synth.make_state_dict_numpy(args, 9), args, np.float64)
for B, L in ((1, 33000), (3, 64000), (2, 40000), (1, 200000), (64, 64000)):
    nd = min(B, 4)
    xd = synth.make_mixtures(nd, L, 4242)
    ref_out, ref_vad, _, _ = O.forward(xd, W, {})
    x = torch.from_numpy(np.tile(xd, ((B + nd - 1) // nd, 1))[:B]).cuda()
    m.set_option("dconv_mma", 0)
    o0, v0, _ = m(x, {})
    torch.cuda.synchronize()
    n0 = m.last_launch_count
    print(f"B={B} L={L}: old path launches {n0} |dvad vs oracle| {np.abs(v0.cpu().numpy()[:nd] - ref_vad).max():.2e}", flush=True)
    for swap in swaps:
        m.set_option("dconv_mma", 1)
        m.set_option("dconv_desc_swap", swap)
        o1, v1, _ = m(x, {})
        torch.cuda.synchronize()
        dv = (v1 - v0).abs().max().item(); dw = (o1 - o0).abs().max().item()
        dvo = np.abs(v1.cpu().numpy()[:nd] - ref_vad).max()
        sd = sisdr_db(o1.cpu().numpy()[:nd], ref_out)
        o2, v2, _ = m(x, {})
        rep = torch.equal(o1, o2) and torch.equal(v1, v2)
        print(f"   mma swap={swap}: launches {m.last_launch_count} |dvad vs old| {dv:.2e} |dwav vs old| {dw:.2e} |dvad vs oracle| {dvo:.2e} "
              f"sisdr vs oracle {sd:.1f} dB finite {bool(torch.isfinite(o1).all())} reproducible {rep}", flush=True)
# timing
x = torch.from_numpy(np.tile(synth.make_mixtures(8, 64000, 1), (32, 1))).cuda()
m.materialize.update(estimated_stfts=False, mask_per_speaker=False, spectrum=False, masks_b=False)
for mode in (0, 1):
    m.set_option("dconv_mma", mode); m.set_option("dconv_desc_swap", swaps[0])
    for _ in range(3): m(x, {})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): m(x, {})
    e1.record(); torch.cuda.synchronize()
    m.set_profile(True)
    for _ in range(5): m(x, {})
    prof = m.read_profile(); m.set_profile(False)
    print(f"dconv_mma={mode}: {e0.elapsed_time(e1) / 10:.3f} ms per 256 x 4 s forward;", {k: round(v[0] / 5, 3) for k, v in prof.items()}, flush=True)
