#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests / golden vectors of its own, so parity is pinned on
outputs of the reference code itself: ``model/model.py::SeparationModel.forward`` and
``model/online_class_unknown_targets.py::OnlineSaving.calc_online``, with seeded random-init
weights from ``septfa_b200.synth`` (the shipped checkpoints are absent) and the seeded synthetic
mixtures of SURVEY.md §8(d). Each case stores the fp32 reference result and the result of the
same module in float64 (for tolerance budgeting). Inputs and weights are NOT stored - they
are regenerated from the recorded seeds.
"""
import os
import sys
import types
import json

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SEPTFA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

# dead imports of the reference that are absent here
sys.modules["turtle"] = types.ModuleType("turtle")
sys.modules["turtle"].forward = None
_mp, _pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
_mp.pyplot = _pp
sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = _mp, _pp

import model.model as RM  # noqa: E402  (the reference)
import model.pit_wrapper as RPW  # noqa: E402
from model.online_class_unknown_targets import OnlineSaving  # noqa: E402

from septfa_b200 import synth  # noqa: E402

KW_GATE = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
KW_SMO = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_unsmo_vad=True, return_smoothed_vad=True,
              threshold_activated_vad=0.45)


def build(args, seed, double=False):
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        m = RM.SeparationModel(**args)
    m.load_state_dict(synth.make_state_dict(args, seed), strict=True)
    m.eval()
    return m.double() if double else m


def run(m, x, kw):
    with torch.no_grad():
        out, vad, est = m(x, dict(kw) if kw else {})
    return out, vad, est


def to_np(t):
    return t.detach().cpu().numpy()


def forward_case(name, args, seed, n, length, base_seed, kws, taps=False, stride=1, store_est=True):
    x = torch.from_numpy(synth.make_mixtures(n, length, base_seed))
    m32, m64 = build(args, seed), build(args, seed, double=True)
    rec = {"meta": json.dumps(dict(args=args, weight_seed=seed, n=n, length=length, base_seed=base_seed,
                                   kws=kws, stride=stride, torch=torch.__version__))}
    for i, kw in enumerate(kws):
        out, vad, est = run(m32, x, kw)
        out64, vad64, est64 = run(m64, x.double(), kw)
        rec[f"kw{i}_out"] = to_np(out)[..., ::stride]
        rec[f"kw{i}_out64"] = to_np(out64).astype(np.float32)[..., ::stride]
        rec[f"kw{i}_vad"] = to_np(vad) if torch.is_tensor(vad) else np.zeros(0)
        rec[f"kw{i}_vad64"] = to_np(vad64) if torch.is_tensor(vad64) else np.zeros(0)
        if store_est:
            rec[f"kw{i}_est"] = to_np(est)
        if i == 0:
            rec["mask"] = to_np(m32.mask_per_speaker) if store_est else np.zeros(0)
            if taps:
                rec["spectrum"] = to_np(m32.spectrum)
                rec["logits"] = to_np(m32.masks_b)
                rec["logits64"] = to_np(m64.masks_b)
                # TCN taps via forward hooks on the fp32 module
                grabbed = {}
                hooks = [m32.TCN.LN.register_forward_hook(lambda mod, i_, o: grabbed.__setitem__("tcn_in", to_np(o)))]
                if args["apply_recursive_ln"]:
                    for b in (0, 1, 23):
                        hooks.append(m32.TCN.ln_second_modules[b].register_forward_hook(
                            lambda mod, i_, o, b=b: grabbed.__setitem__(f"block{b}", to_np(o))))
                run(m32, x, kw)
                for h in hooks:
                    h.remove()
                rec.update(grabbed)
        err = (out.double() - out64).abs().max().item()
        print(f"{name} kw{i}: out {tuple(out.shape)} max|f32-f64|={err:.3e}", flush=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)


def online_case(name, args, seed, length, base_seed, kw, n_streams=1):
    """OnlineSaving.calc_online (unknown targets), one stream at a time (reference semantics)."""
    m32 = build(args, seed)
    sig = []
    for s in range(n_streams):
        x = torch.from_numpy(synth.make_mixtures(1, length, base_seed + s))
        crit = RPW.PITLossWrapper(loss_func=torch.nn.L1Loss(), pit_from="pw_pt")
        o = OnlineSaving(m32, "/tmp/septfa_golden_online", crit)
        o.num_save_samples = 0  # no wav dumps

        class Capture:  # calc_online resets state but keeps online_signal attribute
            pass
        o.calc_online(x, "n", 0, dict(kw))
        sig.append(to_np(o.online_signal)[0])
    rec = {"meta": json.dumps(dict(args=args, weight_seed=seed, length=length, base_seed=base_seed, kw=kw,
                                   n_streams=n_streams, torch=torch.__version__)),
           "online_signal": np.stack(sig)}
    print(f"{name}: online_signal {rec['online_signal'].shape}", flush=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)


def known_targets_case(name, args, seed, length, base_seed, similarity):
    """OnlineSaving of model/online_class_known_targets.py:85-154. The file is stale - it unpacks SIX values from a
    forward that returns three (:105,126) - so the unmodified class is driven through a shim module that pads the
    reference model's 3-tuple to six; everything else (padding, windowing, PIT against the true sources or the L1
    similarity stitch, SI-SDR bookkeeping) is the reference's own code. criterion_separation is test.py's:
    PITLossWrapper(pairwise_neg_sisdr, pit_from='pw_mtx') (test.py:49-56, config_with_vad.json:50-57)."""
    import model.online_class_known_targets as KT
    import model.sdr as SDR
    m32 = build(args, seed)

    class SixTuple(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x):
            out, vad, est = self.inner(x)
            return out, vad, est, None, None, None

    x = torch.from_numpy(synth.make_mixtures(1, length, base_seed))
    rng = np.random.default_rng(base_seed)
    w = torch.from_numpy(rng.uniform(0.3, 0.7, size=(1, 1, length)).astype(np.float32))
    tgt = torch.cat([x[:, None] * w, x[:, None] * (1 - w)], dim=1)        # two "true sources" that sum to the mixture
    crit_sep = RPW.PITLossWrapper(loss_func=SDR.pairwise_neg_sisdr, pit_from="pw_mtx")
    crit_sim = RPW.PITLossWrapper(loss_func=torch.nn.L1Loss(), pit_from="pw_pt") if similarity else None
    o = KT.OnlineSaving(crit_sep, SixTuple(m32), "/tmp/septfa_golden_known", "cpu", crit_sim)
    o.calc_online(x, tgt, "n", 0)
    rec = {"meta": json.dumps(dict(args=args, weight_seed=seed, length=length, base_seed=base_seed, similarity=similarity,
                                   torch=torch.__version__)),
           "online_signal": to_np(o.online_signal), "online_sisdr": np.array(o.online_sisdr),
           "reference_sisdr": np.array(o.reference_sisdr)}
    print(f"{name}: online_signal {rec['online_signal'].shape} online {o.online_sisdr} reference {o.reference_sisdr}", flush=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)


def cli_case(name, args, seed, precision):
    """only_inference.py:27-98 end to end, unmodified, on a synthetic STEREO 16 kHz int16 wav (so the mono selection
    runs; the reference's resampling branch cannot: `resample(torch.tensor(audio))` at :79 yields a Tensor that
    `torch.from_numpy` at :82 rejects with a TypeError) with a synthetic checkpoint: `main(config)` is called with a minimal stand-in
    for parse_config.ConfigParser (which would create run directories and a log file). Stores what the script wrote."""
    import importlib
    import tempfile
    from scipy.io.wavfile import read, write
    sys.argv = ["only_inference.py"]
    OI = importlib.import_module("only_inference")
    tmp = tempfile.mkdtemp(prefix="septfa_cli_golden_")
    wav, ckpt, out_dir = os.path.join(tmp, "mix.wav"), os.path.join(tmp, "model.pth"), os.path.join(tmp, "out")
    write(wav, 16000, synth.make_cli_wav(seed, fs=16000))
    torch.save(synth.make_checkpoint(args, seed), ckpt)

    class Cfg:
        resume = __import__("pathlib").Path(ckpt)
        args = types.SimpleNamespace(save_test_path=out_dir, online=False, path_mix=wav, inference_kw=dict(KW_GATE),
                                     precision_save=precision)

        def get_logger(self, *a, **k):
            import logging
            return logging.getLogger("golden")

        def init_obj(self, key, module, *a, **k):
            import io
            import contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                return getattr(module, "SeparationModel")(**args)

    OI.default_inference_kw = dict(synth.DEFAULT_INFERENCE_KW)      # the module-level dict its __main__ block defines
    import matplotlib.pyplot as plt
    for fn in ("subplots", "savefig", "close", "plot"):            # matplotlib is absent: plotting is a no-op stub
        if not hasattr(plt, fn):
            setattr(plt, fn, lambda *a, **k: (types.SimpleNamespace(), types.SimpleNamespace(
                set_title=lambda *a, **k: None, set_ylabel=lambda *a, **k: None, set_xlabel=lambda *a, **k: None,
                imshow=lambda *a, **k: None, set_xlim=lambda *a, **k: None)) if fn == "subplots" else None)
    figs = types.SimpleNamespace(colorbar=lambda *a, **k: None)
    plt.subplots = lambda *a, **k: (figs, types.SimpleNamespace(
        set_title=lambda *a, **k: None, set_ylabel=lambda *a, **k: None, set_xlabel=lambda *a, **k: None,
        imshow=lambda *a, **k: None, set_xlim=lambda *a, **k: None))
    OI.main(Cfg())
    rec = {"meta": json.dumps(dict(args=args, weight_seed=seed, precision=precision, kw=KW_GATE, torch=torch.__version__))}
    for f in ("Mixed_0.wav", "Speaker_0.wav", "Speaker_1.wav"):
        sr, a = read(os.path.join(out_dir, f))
        assert sr == 16000
        rec[f.replace(".wav", "")] = a
    print(f"{name}: wrote {[ (k, v.shape, v.dtype) for k, v in rec.items() if k != 'meta']}", flush=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)


def main():
    torch.set_num_threads(os.cpu_count())
    W, WO = synth.CONFIG_WITH_VAD, synth.CONFIG_WITHOUT_VAD
    # small, ragged length (L % 256 = 77), both configs, all inference_kw modes, with taps
    forward_case("fwd_with_vad_small", W, 0, 2, 8269, 100, [{}, KW_GATE, KW_SMO], taps=True)
    forward_case("fwd_without_vad_small", WO, 1, 2, 8269, 200, [{}, KW_GATE], taps=True)
    # cfg1: only_inference semantics, 1 x 4 s, with VAD gating (BASELINE.json configs[0])
    forward_case("cfg1_with_vad_4s", W, 2, 1, 64000, 1234, [KW_GATE], store_est=False)
    # 3 s online window length (T=188, ragged tail L % 256 = 128)
    forward_case("fwd_with_vad_3s", W, 3, 1, 48000, 300, [KW_GATE], store_est=False, stride=4)
    # cfg4: long-form 60 s, separation-only config, no inference_kw (strided samples only)
    forward_case("cfg4_without_vad_60s", WO, 4, 1, 960000, 400, [{}], store_est=False, stride=16)
    # online driver: 2 streams of 6 s -> 4 hops each
    online_case("online_with_vad_6s", W, 5, 96000, 500, KW_GATE, n_streams=2)
    # known-targets online driver (through a 6-tuple shim, the reference file is stale): PIT against the true sources,
    # and the L1 similarity stitch
    known_targets_case("online_known_pit_4s", W, 6, 70000, 800, similarity=False)
    known_targets_case("online_known_sim_4s", W, 6, 70000, 800, similarity=True)
    # only_inference.py end to end on a stereo 8 kHz int16 wav + synthetic checkpoint
    cli_case("cli_with_vad_ps32", W, 7, 32)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--new":      # only the cases added in round 2 (the others are unchanged)
        torch.set_num_threads(os.cpu_count())
        W = synth.CONFIG_WITH_VAD
        known_targets_case("online_known_pit_4s", W, 6, 70000, 800, similarity=False)
        known_targets_case("online_known_sim_4s", W, 6, 70000, 800, similarity=True)
        cli_case("cli_with_vad_ps32", W, 7, 32)
    else:
        main()
