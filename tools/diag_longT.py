#!/usr/bin/env python
"""Diagnostic: VAD-probability error vs the fp64 oracle for long utterances, by engine / precision / shape."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import septfa_oracle as O
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 21), strict=True)
m.eval().cuda()
W = O.OracleWeights(synth.make_state_dict_numpy(args, 21), args, np.float64)
for B, L in ((2, 300000), (1, 300000), (2, 299776), (2, 294912), (2, 295168), (1, 960000)):
    xd = synth.make_mixtures(B, L, 4242)
    taps = {}
    _, ref_vad, _, ex = O.forward(xd, W, {}, taps=taps)
    x = torch.from_numpy(xd).cuda()
    for eng, prec, fr in ((0, 1, 1), (0, 2, 1), (7, 0, 1), (7, 0, 0)):
        m.set_engine(eng); m.set_option("precision", prec); m.set_option("fused_resid", fr)
        out, vad, _ = m(x, {})
        dv = np.abs(vad.cpu().numpy() - ref_vad)
        dl = np.abs(m.masks_b.cpu().numpy() - ex["masks_b"]).max()
        dsp = np.abs(m.spectrum.cpu().numpy() - ex["spectrum"])
        ds = dsp.max()
        sb, sf, st_ = np.unravel_index(dsp.argmax(), dsp.shape)
        if eng == 0 and prec == 1:
            print(f"   spectrum argmax (b={sb}, f={sf}, t={st_}): ours {m.spectrum[sb, sf, st_].item():.4f} ref {ex['spectrum'][sb, sf, st_]:.4f}; ref neighbours {ex['spectrum'][sb, max(sf-1,0):sf+2, st_]}; rel err {ds / np.abs(ex['spectrum']).max():.2e}")
        b, s, t = np.unravel_index(dv.argmax(), dv.shape)
        print(f"B={B} L={L} T={1 + L // 256} engine {eng} prec {prec}: |dvad| {dv.max():.2e} at (b={b}, s={s}, t={t}) |dlogits| {dl:.2e} |dspectrum| {ds:.2e}", flush=True)
m.set_engine(0); m.set_option("precision", 0)
