#!/usr/bin/env python
"""Parity report (no assertions): the CUDA path against the fp64 numpy oracle on a sweep of odd shapes, both
configurations, two weight seeds and the three precision modes. Prints one line per case with the error
metrics the GPU tests bound (max |dp| of the VAD probabilities, waveform error / SI-SDR, estimated-STFT
error). Run on a B200:  python tools/parity_report.py [--quick]"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import sisdr_db  # noqa: E402
from oracle import septfa_oracle as O  # noqa: E402
from septfa_b200 import synth  # noqa: E402
from septfa_b200.model import SeparationModel  # noqa: E402

SHAPES = ((1, 257), (1, 32768), (2, 32767), (3, 33000), (9, 64000), (33, 20000), (130, 4000), (2, 300000), (1, 131072),
          (5, 65536), (300, 16000))
QUICK = ((1, 257), (2, 32767), (3, 33000), (9, 64000), (2, 300000))


def main():
    quick = "--quick" in sys.argv
    worst = {}
    for cfg_name, args in (("with_vad", synth.CONFIG_WITH_VAD), ("without_vad", synth.CONFIG_WITHOUT_VAD)):
        for seed in ((9,) if quick else (9, 21)):
            with contextlib.redirect_stdout(io.StringIO()):
                m = SeparationModel(**args)
            m.load_state_dict(synth.make_state_dict(args, seed), strict=True)
            m.eval().cuda()
            W = O.OracleWeights(synth.make_state_dict_numpy(args, seed), args, np.float64)
            for B, L in (QUICK if quick else SHAPES):
                nd = min(B, 8)
                xd = synth.make_mixtures(nd, L, 4242)
                t0 = time.time()
                ref_out, ref_vad, ref_est, _ = O.forward(xd, W, {})
                t_or = time.time() - t0
                x = torch.from_numpy(np.tile(xd, ((B + nd - 1) // nd, 1))[:B]).cuda()
                tail = (L % 256) if (L % 256) > 200 else 0   # ill-conditioned istft tail
                for prec in (1, 0, 2):
                    if prec == 0 and cfg_name == "with_vad":
                        continue   # auto == fast for this configuration
                    m.set_option("precision", prec)
                    out, vad, est = m(x, {})
                    o, v, e = out.cpu().numpy(), vad.cpu().numpy(), est.cpu().numpy()
                    dv = dw = de = der = 0.0
                    sd = 1e9
                    for b in range(B):
                        r = b % nd
                        dv = max(dv, np.abs(v[b] - ref_vad[r]).max())
                        dw = max(dw, np.abs(o[b, :, :L - tail] - ref_out[r, :, :L - tail]).max())
                        de = max(de, np.abs(e[b] - ref_est[r]).max())
                        der = max(der, (np.abs(e[b] - ref_est[r]) / (1.0 + np.abs(ref_est[r]))).max())
                        if b < nd:
                            sd = min(sd, sisdr_db(o[b, :, :L - tail], ref_out[r, :, :L - tail]))
                    key = (cfg_name, prec)
                    worst[key] = max(worst.get(key, 0.0), dv)
                    print(f"{cfg_name} seed {seed} prec {prec} B={B} L={L} T={1 + L // 256}: |dvad| {dv:.2e} |dwav| {dw:.2e} "
                          f"sisdr {sd:.1f} dB |dest| {de:.2e} rel {der:.2e} launches {m.last_launch_count} (oracle {t_or:.1f} s)",
                          flush=True)
            del m
    for k, v in sorted(worst.items()):
        print("worst |dvad|", k, f"{v:.2e}")


if __name__ == "__main__":
    main()
