"""Bring-up check: device forward vs forward_host vs stream API at full size, with / without tail balancing."""
import contextlib, io, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from septfa_b200 import synth
from septfa_b200.model import SeparationModel
from conftest import sisdr_db
args = synth.CONFIG_WITH_VAD
kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
xh = torch.from_numpy(synth.make_mixtures(256, 64000, 1234)).pin_memory()
x = xh.cuda()
base = None
for tb in ("0", "1"):
    os.environ["SEPTFA_TAIL_BALANCE"] = tb
    with contextlib.redirect_stdout(io.StringIO()):
        m = SeparationModel(**args)
    m.load_state_dict(synth.make_state_dict(args, 0), strict=True)
    m.eval().cuda()
    o, v, _ = m(x, kw)
    o = o.cpu().numpy(); v = v.cpu().numpy()
    oh, vh = m.forward_host(xh, kw)
    f = m.forward_host_submit(xh, kw, slot=0); os_, vs_ = f.result()
    if base is None:
        base = (o, v)
    for name, (a, b) in {"device": (o, v), "forward_host": (oh.numpy(), vh.numpy()), "stream": (os_.numpy(), vs_.numpy())}.items():
        d = np.abs(a - base[0])
        print(f"tail_balance={tb} {name:13s}: wave max diff vs base {d.max():.3e}  vad(smoothed gate applied) diff {np.abs(b - base[1]).max():.3e}  "
              f"utterances with diff>1e-3: {(d.reshape(256, -1).max(1) > 1e-3).sum()}")
    del m
