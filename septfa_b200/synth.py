"""Seeded synthetic inputs: reference-layout state_dicts and noisy two-speaker mixtures.

The reference's checkpoints (``model_with_vad.pth`` / ``model_without_vad.pth``) are absent
(``/root/reference/.MISSING_LARGE_BLOBS``), so every parity and bench run uses random-init
weights of the same architecture. The generator is pure numpy (PCG64, stable across
platforms) so the build container (where the reference produces golden vectors) and the
GPU box (where the CUDA path is checked against them) see *identical* weights without
shipping a 20 MB checkpoint.

Key families and shapes follow ``model/model.py:210-325,153-171,376-400``.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

CONFIG_WITH_VAD = {  # config_with_vad.json:7-28 ("arch.args")
    "n_fftBins": 512, "BN_dim": 256, "H_dim": 512, "layer": 8, "stack": 3, "kernel": 3,
    "num_spk": 2, "skip": False, "dilated": True, "casual": False, "bool_drop": True,
    "drop_value": 0.05, "weight_norm": True, "final_vad": True, "final_vad_masked_speakers": False,
    "noisy_phase": True, "activity_input_bool": True, "tf_attention": True,
    "apply_recursive_ln": True, "apply_residual_ln": False,
}
CONFIG_WITHOUT_VAD = dict(CONFIG_WITH_VAD, apply_recursive_ln=False, apply_residual_ln=True)  # config_without_vad.json:26-27

DEFAULT_INFERENCE_KW = {  # only_inference.py:102-108
    "filter_signals_by_smo_vad": False, "filter_signals_by_unsmo_vad": False,
    "length_smoothing_filter": 3, "threshold_activated_vad": 0.5, "return_smoothed_vad": False,
}


def _hann(n=512):
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(np.float32)


def make_state_dict_numpy(args, seed=0):
    """Random-init parameters in the reference's state_dict layout (numpy float32).

    Conv weights/biases ~ U(+-1/sqrt(fan_in)) like torch's default init; weight_g = ||v||
    times U(0.8,1.2); PReLU / GroupNorm parameters are moved off their init values so that
    tests are not blind to them."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    C, H = args.get("BN_dim", 256), args.get("H_dim", 512)
    n_bins = args.get("n_fftBins", 512) // 2 + 1
    nspk = args.get("num_spk", 2)
    nblk = args.get("layer", 8) * args.get("stack", 3)

    def u(shape, bound):
        return rng.uniform(-bound, bound, size=shape).astype(np.float32)

    def wn_conv(prefix, co, ci, k):
        bound = 1.0 / np.sqrt(ci * k)
        v = u((co, ci, k), bound)
        sd[prefix + ".bias"] = u((co,), bound)
        norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
        sd[prefix + ".weight_g"] = (norm * rng.uniform(0.8, 1.2, size=norm.shape)).astype(np.float32)
        sd[prefix + ".weight_v"] = v

    def prelu(key):
        sd[key] = rng.uniform(0.1, 0.4, size=(1,)).astype(np.float32)

    def gn(prefix, c):
        sd[prefix + ".weight"] = rng.uniform(0.7, 1.3, size=(c,)).astype(np.float32)
        sd[prefix + ".bias"] = rng.uniform(-0.2, 0.2, size=(c,)).astype(np.float32)

    sd["spec_input.spec.window"] = _hann()
    sd["spec_output.window"] = _hann()
    sd["inv_spec.window"] = _hann()
    gn("TCN.LN", C)
    for i in range(nblk):
        p = f"TCN.TCN.{i}"
        wn_conv(p + ".conv1d", C, C, 1)
        wn_conv(p + ".dconv1d", H, 1, 3)
        wn_conv(p + ".res_out", C, H, 1)
        prelu(p + ".nonlinearity1.weight")
        prelu(p + ".nonlinearity2.weight")
        gn(p + ".reg1", C)
        gn(p + ".reg2", H)
    if args.get("tf_attention", False):
        for i in range(nblk):
            p = f"TCN.time_freq_attnetion.{i}"  # (sic) model/model.py:279
            for n in ("t_1", "t_2"):
                sd[f"{p}.conv1d_{n}.weight"] = u((1, 1, 3), 1.0 / np.sqrt(3.0))
                sd[f"{p}.conv1d_{n}.bias"] = u((1,), 1.0 / np.sqrt(3.0))
            prelu(p + ".prelu_t.weight")
            for n in ("f_1", "f_2"):
                sd[f"{p}.conv1d_{n}.weight"] = u((1, 1, 3), 1.0 / np.sqrt(3.0))
                sd[f"{p}.conv1d_{n}.bias"] = u((1,), 1.0 / np.sqrt(3.0))
            prelu(p + ".prelu_f.weight")
    if args.get("apply_recursive_ln", False):
        for i in range(nblk):
            gn(f"TCN.ln_first_modules.{i}", C)
        for i in range(nblk):
            gn(f"TCN.ln_second_modules.{i}", C)
    if args.get("apply_residual_ln", False):
        for i in range(nblk):
            gn(f"TCN.ln_modules.{i}", C)
    prelu("TCN.output.0.weight")
    gn("TCN.output.1", C)
    wn_conv("TCN.output.2", n_bins * nspk, C, 1)
    if args.get("final_vad", True):
        wn_conv("vad.common.conv1_1", 4, n_bins, 5)
        prelu("vad.common.relu_1.weight")
        gn("vad.common.BN_1", 4)
        wn_conv("vad.output_layer_vad", 1, 4, 3)
    if args.get("activity_input_bool", False):
        sd["activity_input.weight"] = u((1, 1, 3, 3), 1.0 / 3.0)
        sd["activity_input.bias"] = u((1,), 1.0 / 3.0)
        prelu("prelu.weight")
    return sd


def make_state_dict(args, seed=0):
    """Same as :func:`make_state_dict_numpy` but as torch tensors (for ``load_state_dict``)."""
    import torch
    return OrderedDict((k, torch.from_numpy(v.copy())) for k, v in make_state_dict_numpy(args, seed).items())


def make_checkpoint(args, seed=0):
    """Checkpoint dict in the layout of ``only_inference.py:49-55`` / ``base/base_trainer.py:164-171``."""
    return {"arch": "SeparationModel", "epoch": 0, "state_dict": make_state_dict(args, seed),
            "optimizer": "SeparationModel", "monitor_best": 0.0}


def _bandlimited_noise(rng, n, fs=16000, lo=100.0, hi=4000.0):
    spec = np.fft.rfft(rng.standard_normal(n))
    f = np.fft.rfftfreq(n, 1.0 / fs)
    spec[(f < lo) | (f > hi)] = 0
    return np.fft.irfft(spec, n)


def _onoff_envelope(rng, n, fs=16000, activity=0.6):
    env = np.zeros(n)
    t = 0
    on = rng.random() < activity
    while t < n:
        seg = int(rng.uniform(0.2, 1.0) * fs)
        if on:
            env[t:t + seg] = 1.0
        t += seg
        on = rng.random() < (activity if not on else 0.75)
    ramp = int(0.01 * fs)
    k = np.ones(ramp) / ramp
    return np.convolve(env, k, mode="same")


def make_mixture(index, length, base_seed=1234, fs=16000):
    """One synthetic noisy two-speaker mixture: two band-limited (100-4000 Hz)
    Gaussian "speakers" with independent on/off envelopes, SIR U(0,5) dB, white noise at
    SNR U(0,15) dB (create_data/data_conifg_wham.yaml:60-61), then min-max normalised to
    [-0.9, 0.9] exactly as only_inference.py:81. Returns float32 [length]."""
    for attempt in range(16):
        # attempt 0 keeps the original stream; short clips can draw two all-off envelopes (a silent mixture, which
        # min-max normalisation would turn into NaNs) - those are re-drawn from a derived seed
        rng = np.random.default_rng(base_seed + index if attempt == 0 else [base_seed + index, attempt])
        s1 = _bandlimited_noise(rng, length, fs) * _onoff_envelope(rng, length, fs)
        s2 = _bandlimited_noise(rng, length, fs) * _onoff_envelope(rng, length, fs)
        if np.mean((s1 + s2) ** 2) > 1e-8:
            break
    p1 = np.mean(s1 ** 2) + 1e-12
    p2 = np.mean(s2 ** 2) + 1e-12
    sir = rng.uniform(0.0, 5.0)
    s2 = s2 * np.sqrt(p1 / p2 / (10 ** (sir / 10)))
    mix = s1 + s2
    snr = rng.uniform(0.0, 15.0)
    noise = rng.standard_normal(length)
    noise *= np.sqrt(np.mean(mix ** 2) / (10 ** (snr / 10)) / np.mean(noise ** 2))
    mix = (mix + noise).astype(np.float32)
    mix = 1.8 * (mix - mix.min()) / (mix.max() - mix.min()) - 0.9
    return mix.astype(np.float32)


def make_cli_wav(seed, seconds=3.0, fs=8000):
    """A STEREO int16 wav payload [n, 2] at 8 kHz for the command-line tests: channel 0 is a synthetic mixture scaled to
    int16, channel 1 something else entirely (the script must pick channel 0, only_inference.py:70-75)."""
    n = int(seconds * fs)
    a = make_mixture(0, n, 5000 + seed, fs=fs)
    b = make_mixture(1, n, 5000 + seed, fs=fs)
    return np.stack([np.round(a * 30000.0), np.round(b * 12000.0)], axis=1).astype(np.int16)


def make_mixtures(n, length, base_seed=1234, first_index=0):
    """[n, length] float32 batch of :func:`make_mixture` (mixture i uses seed base_seed+first_index+i)."""
    return np.stack([make_mixture(first_index + i, length, base_seed) for i in range(n)])
