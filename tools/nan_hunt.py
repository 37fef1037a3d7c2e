import contextlib, io, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
args = synth.CONFIG_WITH_VAD
with contextlib.redirect_stdout(io.StringIO()):
    m = SeparationModel(**args)
m.load_state_dict(synth.make_state_dict(args, 0), strict=True)
m.eval().cuda()
kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
for (B, L, seed) in ((2, 16000, 1234), (1, 16000, 1234), (2, 16128, 1234), (2, 8269, 100), (2, 16000, 77), (3, 16000, 1234), (2, 32000, 1234)):
    x = torch.from_numpy(synth.make_mixtures(B, L, seed)).cuda()
    for eng in (0, 7, 3, 5, 6):
        m.set_engine(eng)
        out, vad, est = m(x, dict(kw))
        torch.cuda.synchronize()
        print(f"B={B} L={L} seed={seed} engine={eng}: nan out={bool(torch.isnan(out).any())} vad={bool(torch.isnan(vad).any())} logits={bool(torch.isnan(m.masks_b).any())} spectrum={bool(torch.isnan(m.spectrum).any())}", flush=True)
