"""Evaluation-side companions of the forward pass, on the device.

The reference's ``test.py`` wraps the forward in: a PIT over the pairwise negative SI-SDR matrix
(``loss_separation``: ``PITLossWrapper(pairwise_neg_sisdr, pit_from='pw_mtx')``, test.py:49-56,
model/sdr.py:48-85, model/pit_wrapper.py:100-103), per-speaker SI-SDR and SI-SDR improvement
(test.py:133-143), the mask-based "simple VAD" (Our_utils/utils_test.py:67-72) and the VAD accuracy
(model/metric.py:163-177). These functions give the same quantities for a whole batch from CUDA
tensors: the SI-SDR reductions run in ``septfa_sisdr`` (one pass, five moments per row in double),
the rest are single elementwise / reduction expressions on the device.
"""
from __future__ import annotations

import torch

from .pit import PITLossWrapper, calc_sisdr, reorder_source_mse, sisdr_moments


def pairwise_neg_sisdr(est_targets, targets):
    """``PairwiseNegSDR('sisdr')`` (model/sdr.py:48-85): ``out[b, i, j] = -SI-SDR(est_i, target_j)`` in dB, zero-mean.
    [B, n_src, n] x [B, n_src, n] -> [B, n_src, n_src], CUDA tensors."""
    if targets.shape != est_targets.shape or targets.ndim != 3:
        raise TypeError(f"Inputs must be of shape [batch, n_src, time], got {targets.size()} and {est_targets.size()} instead")
    B, n_src, n = targets.shape
    est = est_targets.detach().float().unsqueeze(2).expand(B, n_src, n_src, n).reshape(-1, n).contiguous()   # est_i
    tgt = targets.detach().float().unsqueeze(1).expand(B, n_src, n_src, n).reshape(-1, n).contiguous()       # target_j
    # the reference's expression (EPS = 1e-8 added to the target energy, to the noise energy and inside the log: pairs
    # that do not correlate saturate at +80 dB) evaluated from the five moments of one device pass
    _, mom = sisdr_moments(est, tgt)
    sp, st, spt, stt, spp = mom.unbind(dim=1)
    eps = 1e-8
    dot = spt - sp * st / n
    tt = stt - st * st / n
    pp = spp - sp * sp / n
    te = tt + eps
    proj2 = dot * dot * tt / (te * te)
    noise2 = (pp - 2.0 * dot * dot / te + proj2).clamp_min(0.0)
    sdr = proj2 / (noise2 + eps)
    return (-10.0 * torch.log10(sdr + eps)).to(est_targets.dtype).view(B, n_src, n_src)


def pit_sisdr(est_targets, targets):
    """The separation criterion of test.py: returns ``(mean loss, reordered estimates, permutation [B, n_src])``."""
    crit = PITLossWrapper(pairwise_neg_sisdr, pit_from="pw_mtx")
    loss, reordered, idx = crit(est_targets, targets, return_est=True, return_incides=True)
    return loss, reordered, idx


def separation_report(mix, est_targets, targets):
    """Per-utterance numbers of test.py:131-143 for a batch: the PIT permutation, SI-SDR per speaker of the reordered
    estimates, SI-SDR of the unprocessed mixture per speaker ("start") and the improvement. All tensors on the device."""
    _, reordered, idx = pit_sisdr(est_targets, targets)
    si_sdr = calc_sisdr(reordered, targets)                                   # [B, n_src]
    start = calc_sisdr(mix.unsqueeze(1).expand_as(targets).contiguous(), targets)
    return {"perm": idx, "si_sdr": si_sdr, "si_sdr_start": start, "si_sdri": si_sdr - start, "reordered": reordered}


def simple_vad_from_masks(masks):
    """``calc_vad`` of Our_utils/utils_test.py:67-72: a frame is active when at least a quarter of the 257 bins have a
    mask value >= 0.5. masks [B, n_spk, 257, T] (``model.mask_per_speaker``) -> int64 [B, n_spk, T]."""
    thr = (masks >= 0.5).to(torch.int64)
    return (thr.sum(dim=2) >= 257 * 0.25).to(torch.int64)


def vad_accuracy(preds, targets):
    """``Accuracy_Vad`` (model/metric.py:163-177) without its in-place thresholding of the caller's tensor: decisions
    are ``p > 0.5``; returns (overall, speaker 0, speaker 1) accuracies as 0-dim tensors."""
    d = (preds > 0.5).to(targets.dtype)
    acc = (d == targets).sum() / targets.numel()
    acc0 = (d[:, 0] == targets[:, 0]).sum() / targets[:, 0].numel()
    acc1 = (d[:, 1] == targets[:, 1]).sum() / targets[:, 1].numel()
    return acc, acc0, acc1


__all__ = ["pairwise_neg_sisdr", "pit_sisdr", "separation_report", "simple_vad_from_masks", "vad_accuracy",
           "reorder_source_mse"]
