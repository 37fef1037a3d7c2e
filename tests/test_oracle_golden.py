"""CPU: the numpy oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py). This is what pins the oracle."""
import numpy as np
import pytest

from conftest import load_golden, sisdr_db
from oracle import septfa_oracle as O
from septfa_b200 import synth


def _weights(meta, dtype=np.float64):
    return O.OracleWeights(synth.make_state_dict_numpy(meta["args"], meta["weight_seed"]), meta["args"], dtype)


@pytest.mark.parametrize("name", ["fwd_with_vad_small", "fwd_without_vad_small", "fwd_with_vad_3s", "cfg1_with_vad_4s"])
def test_forward_matches_reference(name):
    g, meta = load_golden(name)
    W = _weights(meta)
    x = synth.make_mixtures(meta["n"], meta["length"], meta["base_seed"])
    st = meta["stride"]
    for i, kw in enumerate(meta["kws"]):
        taps = {}
        out, vad, est, extras = O.forward(x, W, dict(kw) if kw else {}, taps)
        # fp64 oracle vs fp64 run of the reference module: only the fp32-rounded window buffer differs
        assert np.abs(out[..., ::st] - g[f"kw{i}_out64"]).max() < 5e-6
        # vs the fp32 reference itself
        assert np.abs(out[..., ::st] - g[f"kw{i}_out"]).max() < 1e-5
        assert sisdr_db(out[..., ::st], g[f"kw{i}_out"]) > 90.0
        if g[f"kw{i}_vad"].size:
            ref_vad = g[f"kw{i}_vad"]
            assert vad.shape == ref_vad.shape  # [B,2,T] or [B,2,1,T] with return_smoothed_vad
            if kw and kw.get("return_smoothed_vad"):
                # decisions may differ only where the fp32 reference probability is within 1e-3 of the threshold
                assert (vad != ref_vad).mean() < 0.02
            else:
                assert np.abs(vad - ref_vad).max() < 2e-5
        if f"kw{i}_est" in g:
            assert np.abs(est - g[f"kw{i}_est"]).max() < 2e-4
        if i == 0 and "logits" in g:
            assert np.abs(taps["logits"] - g["logits"]).max() < 2e-4
            assert np.abs(taps["spectrum"] - g["spectrum"]).max() < 1e-5 * np.abs(g["spectrum"]).max()
            for k in ("tcn_in", "block0", "block1", "block23"):
                if k in g:
                    assert np.abs(taps[k] - g[k]).max() < 2e-3
            assert np.abs(extras["mask_per_speaker"] - g["mask"]).max() < 1e-4


def test_long_form_60s_matches_reference():
    g, meta = load_golden("cfg4_without_vad_60s")
    W = _weights(meta, np.float32)  # fp32 keeps this case to a few seconds
    x = synth.make_mixtures(meta["n"], meta["length"], meta["base_seed"])
    out, vad, _, _ = O.forward(x, W, {})
    st = meta["stride"]
    assert np.abs(out[..., ::st] - g["kw0_out"]).max() < 1e-4
    assert sisdr_db(out[..., ::st], g["kw0_out"]) > 70.0
    assert np.abs(vad - g["kw0_vad"]).max() < 1e-4


def test_online_driver_matches_reference():
    g, meta = load_golden("online_with_vad_6s")
    W = _weights(meta)
    for s in range(meta["n_streams"]):
        x = synth.make_mixtures(1, meta["length"], meta["base_seed"] + s)
        online, perms = O.calc_online(x, W, meta["kw"])
        ref = g["online_signal"][s]
        assert online.shape[1:] == ref.shape
        assert np.abs(online[0] - ref).max() < 1e-5
        assert perms.shape == (ref.shape[-1] // 16000, 1, 2)


def test_smoothing_ignores_filter_length_and_copies_edges():
    # model/model.py:444-451: weights forced to [1,0,1]; edges copied from the thresholded decision
    p = np.array([[[0.9, 0.1, 0.1, 0.6, 0.1, 0.7]]])
    dcs, sm = O.smooth_vad(p, 0.5)
    assert dcs.tolist() == [[[1, 0, 0, 1, 0, 1]]]
    assert sm.tolist() == [[[1, 1, 1, 0, 1, 1]]]


def test_sisdr_known_answer():
    # docstring known answer, model/combined_loss.py:29-33 (torchmetrics' default there is zero_mean=False)
    p, t = np.array([2.5, 0.0, 2.0, 8.0], dtype=np.float32), np.array([3.0, -0.5, 2.0, 7.0], dtype=np.float32)
    assert abs(float(O.calc_sisdr(p, t, zero_mean=False)) - 18.4030) < 1e-3


def test_pit_ties_pick_identity():
    a = np.zeros((2, 10))
    assert O.l1_pit_perm(a, a).tolist() == [0, 1]
    b = np.stack([np.ones(10), np.zeros(10)])
    assert O.l1_pit_perm(b[::-1], b).tolist() == [1, 0]
