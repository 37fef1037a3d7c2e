"""CPU oracle for the Sep-TFAnet-VAD inference forward pass (TEST INFRASTRUCTURE ONLY).

This is a numpy restatement of the reference algorithm. It is the *checker* for the CUDA
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it. The product (``septfa_b200``) never does.

Parity pin: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against outputs of the *reference itself* executed in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference/model/model.py`` and
``model/online_class_unknown_targets.py`` and stores their outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

Third-party arithmetic restated here (absent from /root/reference, unpinned upstream —
semantics taken from the torch 2.11 / torchaudio 2.11 installed in this image):
``torchaudio.transforms.Spectrogram`` / ``InverseSpectrogram`` / ``AmplitudeToDB`` and
``torch.nn.{Conv1d,Conv2d,GroupNorm,PReLU}``, ``torch.nn.utils.weight_norm``.

Every function cites the reference ``file:line`` it follows (paths relative to the
reference repository root).
"""
from __future__ import annotations

import numpy as np

N_FFT = 512
HOP = 256
N_BINS = 257
FS = 16000

DEFAULT_ARGS = {  # model/model.py:362-366
    "n_fftBins": 512, "BN_dim": 256, "H_dim": 512, "layer": 8, "stack": 3, "kernel": 3,
    "num_spk": 2, "skip": False, "dilated": True, "casual": False, "bool_drop": True,
    "drop_value": 0.1, "weight_norm": False, "final_vad": True, "noisy_phase": False,
    "activity_input_bool": False, "tf_attention": False, "apply_recursive_ln": False,
    "apply_residual_ln": False, "final_vad_masked_speakers": False,
}


# --------------------------------------------------------------------------- primitives
def hann_periodic(n=N_FFT, dtype=np.float64):
    """torch.hann_window(n) (periodic) used by model/model.py:19-20,384-387."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def prelu(x, a):
    """nn.PReLU with one scalar slope (model/model.py:93-94,164,187,193,322,400)."""
    a = np.asarray(a, dtype=x.dtype).reshape(())
    return np.where(x >= 0, x, a * x)


def group_norm1(x, gamma, beta, eps):
    """nn.GroupNorm(1, C): statistics over the whole (C, T) plane of each utterance,
    biased variance, per-channel affine (model/model.py:97-98,123-124,165,274,313-319,323).
    x: [B, C, T]."""
    mu = x.mean(axis=(1, 2), keepdims=True)
    var = ((x - mu) ** 2).mean(axis=(1, 2), keepdims=True)
    xn = (x - mu) / np.sqrt(var + x.dtype.type(eps))
    return xn * gamma[None, :, None] + beta[None, :, None]


def fold_weight_norm(g, v):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, norm over all dims but 0
    (model/model.py:104-127,159-163,324)."""
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return (g.astype(np.float64) * v.astype(np.float64) / norm).astype(v.dtype)


def conv1d_pointwise(x, w, b):
    """Conv1d k=1. x [B,Ci,T], w [Co,Ci,1], b [Co] (model/model.py:132,144,357)."""
    return np.einsum("oc,bct->bot", w[:, :, 0], x, optimize=True) + b[None, :, None]


def conv1d_small(x, w, b, pad, dil):
    """Generic small dense Conv1d with zero padding. x [B,Ci,T], w [Co,Ci,K]."""
    B, Ci, T = x.shape
    Co, _, K = w.shape
    xp = np.zeros((B, Ci, T + 2 * pad), dtype=x.dtype)
    xp[:, :, pad:pad + T] = x
    Tout = T + 2 * pad - dil * (K - 1)
    out = np.zeros((B, Co, Tout), dtype=x.dtype)
    for k in range(K):
        out += np.einsum("oc,bct->bot", w[:, :, k], xp[:, :, k * dil:k * dil + Tout], optimize=True)
    if b is not None:
        out += b[None, :, None]
    return out


def depthwise_conv(x, w, b, dil):
    """Conv1d(256->512, k=3, groups=256, dilation=padding=dil): out-channel o reads
    in-channel o//2 (model/model.py:111-114)."""
    B, C, T = x.shape
    Co = w.shape[0]
    mult = Co // C
    xp = np.zeros((B, C, T + 2 * dil), dtype=x.dtype)
    xp[:, :, dil:dil + T] = x
    xr = np.repeat(xp, mult, axis=1)  # [B, Co, T+2d]
    out = np.zeros((B, Co, T), dtype=x.dtype)
    for k in range(3):
        out += w[None, :, 0, k, None] * xr[:, :, k * dil:k * dil + T]
    return out + b[None, :, None]


# --------------------------------------------------------------------------- STFT / iSTFT
def stft(x, dtype=None):
    """transforms.Spectrogram(n_fft=512, hop=256, win=512, hann, power=None) ->
    torch.stft(center=True, pad_mode='reflect', onesided) (model/model.py:19-20,384-385,408-410).
    x [B, L] -> complex [B, 257, T], T = 1 + L // 256. DC row is zeroed (model.py:24,410)."""
    x = np.asarray(x)
    dtype = dtype or x.dtype
    B, L = x.shape
    xp = np.pad(x, ((0, 0), (N_FFT // 2, N_FFT // 2)), mode="reflect")
    T = 1 + L // HOP
    idx = np.arange(T)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = xp[:, idx] * hann_periodic(dtype=dtype)[None, None, :]  # [B, T, 512]
    S = np.fft.rfft(frames.astype(np.float64), axis=-1)  # numpy FFT is float64 internally
    S = np.transpose(S, (0, 2, 1))
    S[:, 0, :] = 0
    ctype = np.complex64 if dtype == np.float32 else np.complex128
    return S.astype(ctype)


def istft(E, length):
    """transforms.InverseSpectrogram -> torch.istft(center=True, length=L)
    (model/model.py:386-387,460). E [B, S, 257, T] complex -> [B, S, L]."""
    B, S, F, T = E.shape
    real_dtype = np.float32 if E.dtype == np.complex64 else np.float64
    w = hann_periodic(dtype=np.float64)
    fr = np.fft.irfft(E.astype(np.complex128), n=N_FFT, axis=2)  # [B,S,512,T]
    fr = fr * w[None, None, :, None]
    n_out = N_FFT + HOP * (T - 1)
    acc = np.zeros((B, S, n_out), dtype=np.float64)
    env = np.zeros(n_out, dtype=np.float64)
    for t in range(T):
        acc[:, :, t * HOP:t * HOP + N_FFT] += fr[:, :, :, t]
        env[t * HOP:t * HOP + N_FFT] += w * w
    out = acc[:, :, HOP:HOP + length] / env[None, None, HOP:HOP + length]
    return out.astype(real_dtype)


# --------------------------------------------------------------------------- weights
class OracleWeights:
    """Folded (weight-norm applied) parameters pulled from a reference-layout state_dict
    (key families: SURVEY.md §8(b); model/model.py:210-325,153-171,398-400)."""

    def __init__(self, state_dict, args, dtype=np.float64):
        self.args = dict(DEFAULT_ARGS)
        self.args.update(args)
        a = self.args
        if not a["weight_norm"] or a["skip"] or a["casual"] or not a["dilated"]:
            raise NotImplementedError("oracle covers the shipped configs (weight_norm, non-causal, no skip)")
        sd = {k: np.asarray(v, dtype=np.float64) if not isinstance(v, np.ndarray) else v.astype(np.float64)
              for k, v in state_dict.items()}
        c = lambda k: sd[k].astype(dtype)  # noqa: E731
        wn = lambda p: fold_weight_norm(sd[p + ".weight_g"], sd[p + ".weight_v"]).astype(dtype)  # noqa: E731
        self.dtype = dtype
        self.nblocks = a["layer"] * a["stack"]
        self.ln_w, self.ln_b = c("TCN.LN.weight"), c("TCN.LN.bias")
        self.blocks = []
        for i in range(self.nblocks):
            p = f"TCN.TCN.{i}."
            blk = dict(
                dil=(i % a["layer"]) % 4 + 1,  # model/model.py:285-293
                w1=wn(p + "conv1d"), b1=c(p + "conv1d.bias"), a1=c(p + "nonlinearity1.weight"),
                g1=c(p + "reg1.weight"), be1=c(p + "reg1.bias"),
                w2=wn(p + "dconv1d"), b2=c(p + "dconv1d.bias"), a2=c(p + "nonlinearity2.weight"),
                g2=c(p + "reg2.weight"), be2=c(p + "reg2.bias"),
                w3=wn(p + "res_out"), b3=c(p + "res_out.bias"),
            )
            if a["tf_attention"]:
                q = f"TCN.time_freq_attnetion.{i}."
                for n in ("t_1", "t_2", "f_1", "f_2"):
                    blk["w" + n], blk["b" + n] = c(q + f"conv1d_{n}.weight"), c(q + f"conv1d_{n}.bias")
                blk["at"], blk["af"] = c(q + "prelu_t.weight"), c(q + "prelu_f.weight")
            if a["apply_recursive_ln"]:
                blk["lf_w"], blk["lf_b"] = c(f"TCN.ln_first_modules.{i}.weight"), c(f"TCN.ln_first_modules.{i}.bias")
                blk["ls_w"], blk["ls_b"] = c(f"TCN.ln_second_modules.{i}.weight"), c(f"TCN.ln_second_modules.{i}.bias")
            elif a["apply_residual_ln"]:
                blk["lm_w"], blk["lm_b"] = c(f"TCN.ln_modules.{i}.weight"), c(f"TCN.ln_modules.{i}.bias")
            self.blocks.append(blk)
        self.out_a = c("TCN.output.0.weight")
        self.out_g, self.out_be = c("TCN.output.1.weight"), c("TCN.output.1.bias")
        self.out_w, self.out_b = wn("TCN.output.2"), c("TCN.output.2.bias")
        if a["final_vad"]:
            self.v_w1, self.v_b1 = wn("vad.common.conv1_1"), c("vad.common.conv1_1.bias")
            self.v_a = c("vad.common.relu_1.weight")
            self.v_g, self.v_be = c("vad.common.BN_1.weight"), c("vad.common.BN_1.bias")
            self.v_w2, self.v_b2 = wn("vad.output_layer_vad"), c("vad.output_layer_vad.bias")
        if a["activity_input_bool"]:
            self.act_w, self.act_b = c("activity_input.weight"), c("activity_input.bias")
            self.act_a = c("prelu.weight")


# --------------------------------------------------------------------------- network pieces
def db_spectrum(S):
    """AmplitudeToDB(stype='power'): 10*log10(clamp(|S|^2, 1e-10)) (model/model.py:382,411-412)."""
    real_dtype = np.float32 if S.dtype == np.complex64 else np.float64
    power = (S.real.astype(real_dtype) ** 2 + S.imag.astype(real_dtype) ** 2)
    return (10.0 * np.log10(np.maximum(power, real_dtype(1e-10)))).astype(real_dtype)


def activity_gate(P, W):
    """spectrum *= PReLU(Conv2d(1->1, 3x3, pad 1)(spectrum)) over the (257, T) plane
    (model/model.py:398-400,414-419). P [B,257,T]."""
    B, F, T = P.shape
    Pp = np.zeros((B, F + 2, T + 2), dtype=P.dtype)
    Pp[:, 1:F + 1, 1:T + 1] = P
    k = W.act_w[0, 0]
    acc = np.zeros_like(P)
    for i in range(3):
        for j in range(3):
            acc += k[i, j] * Pp[:, i:i + F, j:j + T]
    acc += W.act_b[0]
    return P * prelu(acc, W.act_a)


def tf_attention(r, blk):
    """TF_Attention.forward (model/model.py:197-208). r [B,256,T]."""
    def gate(v, w1, b1, w2, b2, a):  # v [B,1,N]
        u1 = conv1d_small(v, w1, b1, pad=1, dil=1)
        u2 = conv1d_small(u1, w2, b2, pad=2, dil=2)
        return sigmoid(prelu(u2, a))
    m_t = r.mean(axis=1, keepdims=True)                      # [B,1,T]
    g_t = gate(m_t, blk["wt_1"], blk["bt_1"], blk["wt_2"], blk["bt_2"], blk["at"])
    m_f = r.mean(axis=2, keepdims=True).transpose(0, 2, 1)   # [B,1,256]
    g_f = gate(m_f, blk["wf_1"], blk["bf_1"], blk["wf_2"], blk["bf_2"], blk["af"]).transpose(0, 2, 1)
    return r * (g_f * g_t)                                   # outer product, model.py:206-207


def depth_conv_block(y, blk):
    """DepthConv1d.forward, weight-norm branch, eval mode (model/model.py:130-149)."""
    h1 = group_norm1(prelu(conv1d_pointwise(y, blk["w1"], blk["b1"]), blk["a1"]), blk["g1"], blk["be1"], 1e-8)
    h2 = group_norm1(prelu(depthwise_conv(h1, blk["w2"], blk["b2"], blk["dil"]), blk["a2"]), blk["g2"], blk["be2"], 1e-8)
    return conv1d_pointwise(h2, blk["w3"], blk["b3"])


def tcn(z0, W, taps=None):
    """TCN.forward (model/model.py:329-358). z0 [B,256,T] -> logits [B,514,T]."""
    a = W.args
    y = group_norm1(z0, W.ln_w, W.ln_b, 1e-8)
    if taps is not None:
        taps["tcn_in"] = y
    for i, blk in enumerate(W.blocks):
        r = depth_conv_block(y, blk)
        if a["tf_attention"]:
            r = tf_attention(r, blk)
        if a["apply_recursive_ln"]:
            y = group_norm1(y + group_norm1(y + r, blk["lf_w"], blk["lf_b"], 1e-5), blk["ls_w"], blk["ls_b"], 1e-5)
        elif a["apply_residual_ln"]:
            y = y + group_norm1(r, blk["lm_w"], blk["lm_b"], 1e-5)
        else:
            y = y + r
        if taps is not None and i in (0, 1, W.nblocks - 1):
            taps[f"block{i}"] = y
    q = group_norm1(prelu(y, W.out_a), W.out_g, W.out_be, 1e-5)
    return conv1d_pointwise(q, W.out_w, W.out_b)


def vad_head(logits_s, W):
    """VAD.forward on one speaker's mask logits [B,257,T] (model/model.py:173-179)."""
    c = conv1d_small(logits_s, W.v_w1, W.v_b1, pad=2, dil=1)
    c = group_norm1(prelu(c, W.v_a), W.v_g, W.v_be, 1e-8)
    return sigmoid(conv1d_small(c, W.v_w2, W.v_b2, pad=1, dil=1))  # [B,1,T]


def smooth_vad(p, thr):
    """Inference-only gating (model/model.py:444-451): threshold (>=), [1,0,1] neighbour
    OR clipped to 1, edge frames copied from the un-smoothed decision. p [B,2,T] ->
    (dcs, sm) both [B,2,T] of {0,1}."""
    dcs = np.where(p >= p.dtype.type(thr), 1.0, 0.0).astype(p.dtype)
    T = p.shape[-1]
    sm = np.zeros_like(dcs)
    if T >= 3:
        sm[..., 1:-1] = np.minimum(dcs[..., :-2] + dcs[..., 2:], 1.0)
    elif T == 2:
        pass  # both frames are edges
    sm[..., 0] = dcs[..., 0]
    sm[..., -1] = dcs[..., -1]
    return dcs, sm


_INFER_KEYS = ("length_smoothing_filter", "threshold_activated_vad", "filter_signals_by_smo_vad",
               "filter_signals_by_unsmo_vad", "return_smoothed_vad")


def forward(x, W, inference_kw=None, taps=None):
    """SeparationModel.forward (model/model.py:402-461).
    Returns (out [B,2,L], vad [B,2,T] or [B,2,1,T], est [B,2,257,T] complex, extras dict)."""
    a = W.args
    x = np.asarray(x, dtype=W.dtype)
    assert x.ndim == 2
    B, L = x.shape
    S = stft(x, dtype=W.dtype)                                     # :408-410
    P = db_spectrum(S)                                             # :411-412
    if a["activity_input_bool"]:
        P = activity_gate(P, W)                                    # :414-419
    logits = tcn(P[:, 1:], W, taps)                                # :421
    T = logits.shape[-1]
    lg = logits.reshape(B, a["num_spk"], N_BINS, T)                # :423
    if a["final_vad"] and not a["final_vad_masked_speakers"]:
        vad = np.concatenate([vad_head(lg[:, s], W) for s in range(a["num_spk"])], axis=1)  # :424-425
    else:
        vad = 0
    M = sigmoid(lg)                                                # :429
    if a["noisy_phase"]:                                           # :430-437
        mag = np.abs(S)[:, None] * M
        est = mag.astype(S.dtype) * np.exp(1j * np.angle(S))[:, None].astype(S.dtype)
    else:
        est = S[:, None] * M                                       # :439
    vad_out = vad
    if inference_kw and a["final_vad"]:                            # :444-457
        for k in _INFER_KEYS:
            inference_kw[k]  # KeyError like the reference when a key is missing
        _, sm = smooth_vad(vad, inference_kw["threshold_activated_vad"])
        if inference_kw["filter_signals_by_smo_vad"] or inference_kw["filter_signals_by_unsmo_vad"]:
            est = sm[:, :, None, :].astype(est.real.dtype) * est
        if inference_kw["return_smoothed_vad"]:
            vad_out = sm[:, :, None, :]
    out = istft(est, L)                                            # :460
    if taps is not None:
        taps.update(spectrum=P, logits=logits, mask=M, stft=S)
    return out, vad_out, est, {"mask_per_speaker": M, "spectrum": P, "masks_b": logits}


# --------------------------------------------------------------------------- online driver
def l1_pit_perm(a, b):
    """PITLossWrapper(L1Loss, 'pw_pt') for n_src=2, one stream
    (model/pit_wrapper.py:149-177,261-312): pw[i,j] = mean|a_i - b_j|; identity vs swap,
    ties -> identity (torch.min first index). a,b [2, n]. Returns perm (2,) ints."""
    pw = np.array([[np.mean(np.abs(a[i] - b[j])) for j in range(2)] for i in range(2)], dtype=a.dtype)
    ident = (pw[0, 0] + pw[1, 1]) / 2
    swap = (pw[1, 0] + pw[0, 1]) / 2
    return np.array([0, 1]) if ident <= swap else np.array([1, 0])


def calc_online(mix, W, inference_kw, fs=FS, max_len=3, save_sec=1, forward_fn=None):
    """OnlineSaving.calc_online, unknown targets (model/online_class_unknown_targets.py:72-105),
    evaluated independently per stream (the reference is only ever called with B=1; see
    SURVEY.md §3.3). mix [S, L]. Returns (online_signal [S,2,hops*fs], perms [hops,S,2])."""
    mix = np.asarray(mix, dtype=W.dtype)
    if mix.shape[-1] < fs * max_len:                                         # :73-74
        mix = np.pad(mix, ((0, 0), (0, fs * max_len - mix.shape[-1])))
    max_indx = int(np.floor((mix.shape[-1] - fs * max_len) / (fs * save_sec)))  # :77
    fwd = forward_fn or (lambda w: forward(w, W, dict(inference_kw) if inference_kw else inference_kw)[0])
    hop = int(np.floor(fs * save_sec))
    online = None
    perms = []
    indx = 0
    while indx <= max_indx:                                                  # :80
        win = mix[:, fs * indx * save_sec: fs * indx * save_sec + max_len * fs]  # :39-41
        pred = fwd(win)
        if indx == 0:
            online = pred[:, :, pred.shape[-1] - hop:]                       # :85-86
        n_on = online.shape[-1]
        start = max(pred.shape[-1] - hop - n_on, 0)                          # python slice clamp, :87
        a = pred[:, :, start: pred.shape[-1] - hop]
        b = online[:, :, max(n_on - (fs * max_len - hop), 0):]               # :88
        perm = np.stack([l1_pit_perm(a[s], b[s]) for s in range(mix.shape[0])])  # :90
        pred = np.stack([pred[s][perm[s]] for s in range(mix.shape[0])])     # :93 reorder_source_mse
        tail = pred[:, :, pred.shape[-1] - hop:]
        online = tail if indx == 0 else np.concatenate([online, tail], axis=-1)  # :28-37
        perms.append(perm)
        indx += 1
    return online, np.stack(perms)


def calc_sisdr(preds, target, zero_mean=True):
    """model/combined_loss.py:16-56 (SI-SDR in dB over the last axis)."""
    preds = np.asarray(preds)
    target = np.asarray(target)
    eps = np.finfo(preds.dtype).eps
    if zero_mean:
        target = target - target.mean(axis=-1, keepdims=True)
        preds = preds - preds.mean(axis=-1, keepdims=True)
    alpha = ((preds * target).sum(-1, keepdims=True) + eps) / ((target ** 2).sum(-1, keepdims=True) + eps)
    ts = alpha * target
    noise = ts - preds
    return 10 * np.log10(((ts ** 2).sum(-1) + eps) / ((noise ** 2).sum(-1) + eps))
