#!/usr/bin/env python
"""GPU bring-up diagnostic: run the golden cases through every engine combination and print
per-stage errors (spectrum, logits, VAD, waveform). Not a test; used with gpurun while
developing kernels.  python tools/diag.py [case ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from septfa_b200 import synth  # noqa: E402
from septfa_b200.model import SeparationModel  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sisdr(est, ref):
    est = est.reshape(-1, est.shape[-1]).astype(np.float64)
    ref = ref.reshape(-1, ref.shape[-1]).astype(np.float64)
    a = (est * ref).sum(-1, keepdims=True) / ((ref ** 2).sum(-1, keepdims=True) + 1e-30)
    n = a * ref - est
    return float((10 * np.log10(((a * ref) ** 2).sum(-1) / ((n ** 2).sum(-1) + 1e-30) + 1e-30)).min())


def run_case(name, engines):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(g["meta"]))
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        m = SeparationModel(**meta["args"])
    m.load_state_dict(synth.make_state_dict(meta["args"], meta["weight_seed"]), strict=True)
    m.eval().cuda()
    x = torch.from_numpy(synth.make_mixtures(meta["n"], meta["length"], meta["base_seed"])).cuda()
    st = meta["stride"]
    for eng in engines:
        m.set_engine(eng)
        for i, kw in enumerate(meta["kws"]):
            t0 = time.time()
            try:
                out, vad, est = m(x, dict(kw) if kw else {})
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"{name} engine={eng} kw{i}: FAILED {e}", flush=True)
                return False
            dt = time.time() - t0
            o = out.cpu().numpy()[..., ::st]
            ref = g[f"kw{i}_out"]
            msg = (f"{name} engine={eng} kw{i} ({dt * 1e3:.1f} ms, {m.last_launch_count} launches): "
                   f"wav max|d|={np.abs(o - ref).max():.3e} sisdr={sisdr(o, ref):.1f} dB")
            if torch.is_tensor(vad) and g[f"kw{i}_vad"].size:
                v = vad.cpu().numpy()
                msg += f" vad max|d|={np.abs(v - g[f'kw{i}_vad']).max():.3e}"
            if i == 0 and "logits" in g:
                msg += f" logits max|d|={np.abs(m.masks_b.cpu().numpy() - g['logits']).max():.3e}"
                sp = m.spectrum.cpu().numpy()
                msg += f" spectrum max|d|={np.abs(sp - g['spectrum']).max():.3e} (max {np.abs(g['spectrum']).max():.0f})"
            if f"kw{i}_est" in g:
                e = est.cpu().numpy()
                msg += f" est max|d|={np.abs(e - g[f'kw{i}_est']).max():.3e}"
            print(msg, flush=True)
    return True


if __name__ == "__main__":
    # usage: diag.py ENGINE[,ENGINE...] [case ...]; engine bit mask: 7 = all fp32, 0 = all tcgen05
    engines = [int(e) for e in sys.argv[1].split(",")] if len(sys.argv) > 1 else [7, 0]
    cases = sys.argv[2:] or ["fwd_with_vad_small", "fwd_without_vad_small"]
    print(torch.cuda.get_device_name(0), flush=True)
    for c in cases:
        if not run_case(c, engines):
            sys.exit(1)
