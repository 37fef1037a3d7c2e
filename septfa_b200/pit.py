"""Host-side mirror of the PIT / reorder / SI-SDR helpers the online driver uses.

Reference: ``model/pit_wrapper.py:77-177,261-312`` (``PITLossWrapper`` with ``pit_from='pw_pt'``),
``model/combined_loss.py:16-78`` (``calc_sisdr``, ``reorder_source_mse``). Only the path the
online driver exercises is accelerated: ``nn.L1Loss`` + ``pw_pt`` + two sources on CUDA tensors
goes through ``septfa_pit_l1`` (one kernel over all streams). Other loss functions fall back to
the reference's own double loop *of calls to that loss function* (host orchestration only).
"""
from __future__ import annotations

import ctypes as C
from itertools import permutations

import torch

from . import lib as _lib


def reorder_source_mse(preds, batch_indices):
    """model/combined_loss.py:63-78: ``preds[b][batch_indices[b]]`` for every batch item."""
    idx = batch_indices.to(preds.device).long()
    gather_idx = idx.view(idx.shape[0], idx.shape[1], *([1] * (preds.ndim - 2))).expand(-1, -1, *preds.shape[2:])
    return torch.gather(preds, 1, gather_idx)


def sisdr_moments(preds, target):
    """One pass of ``septfa_sisdr`` over rows ``preds[r, :]`` / ``target[r, :]`` (CUDA, float32, contiguous 2-D): returns
    ``(si_sdr_db [rows] float32 with zero_mean=True, moments [rows, 5] float64)`` where the moments are the raw sums
    ``sum p, sum t, sum p t, sum t t, sum p p`` the kernel accumulates in double."""
    rows, n = preds.shape
    out = torch.empty(rows, dtype=torch.float32, device=preds.device)
    scratch = torch.empty(rows * 5, dtype=torch.float64, device=preds.device)
    with torch.cuda.device(preds.device):
        rc = _lib.load().septfa_sisdr(C.c_void_p(preds.data_ptr()), C.c_void_p(target.data_ptr()), rows, n, 1,
                                      C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream(preds.device).cuda_stream))
    if rc != 0:
        raise _lib.SeptfaError(f"septfa_sisdr failed ({rc})")
    return out, scratch.view(rows, 5)


def calc_sisdr(preds, target, zero_mean=True):
    """model/combined_loss.py:16-56 (SI-SDR in dB over the last axis)."""
    if preds.shape != target.shape:
        raise RuntimeError(f"Predictions and targets are expected to have the same shape, pred has shape of "
                           f"{preds.shape} and target has shape of {target.shape}")
    if (preds.is_cuda and target.is_cuda and preds.dtype == torch.float32 and target.dtype == torch.float32
            and preds.ndim >= 1 and 0 < preds.numel() // preds.shape[-1] <= 65535):
        # one fused pass on the device (septfa_sisdr: five moments per row in double), ~1e-3 dB from the element-wise form
        n = preds.shape[-1]
        p2, t2 = preds.detach().reshape(-1, n).contiguous(), target.detach().reshape(-1, n).contiguous()
        out = torch.empty(p2.shape[0], dtype=torch.float32, device=preds.device)
        scratch = torch.empty(p2.shape[0] * 5, dtype=torch.float64, device=preds.device)
        with torch.cuda.device(preds.device):
            rc = _lib.load().septfa_sisdr(C.c_void_p(p2.data_ptr()), C.c_void_p(t2.data_ptr()), p2.shape[0], n, int(bool(zero_mean)),
                                          C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream(preds.device).cuda_stream))
        if rc != 0:
            raise _lib.SeptfaError(f"septfa_sisdr failed ({rc})")
        return out.reshape(preds.shape[:-1])
    # host tensors / other dtypes: the same definition from three inner products per row
    eps = torch.finfo(preds.dtype).eps
    p, t = preds, target
    if zero_mean:
        p, t = p - p.mean(dim=-1, keepdim=True), t - t.mean(dim=-1, keepdim=True)
    scale = ((p * t).sum(dim=-1, keepdim=True) + eps) / ((t * t).sum(dim=-1, keepdim=True) + eps)
    proj = scale * t                                  # projection of the estimate on the target
    ratio = (proj.square().sum(dim=-1) + eps) / ((proj - p).square().sum(dim=-1) + eps)
    return 10 * torch.log10(ratio)


class PITLossWrapper(torch.nn.Module):
    """``PITLossWrapper(loss_func, pit_from)`` (model/pit_wrapper.py:60-147).

    ``per_stream=False`` (default) reproduces the reference bit for bit in its batch quirk:
    ``nn.L1Loss()`` reduces over batch *and* time, so one permutation is chosen jointly for the
    whole batch (model/pit_wrapper.py:173-176). ``per_stream=True`` decides
    every batch item on its own - equal to running the reference once per stream with B = 1,
    which is the only way the reference is ever called.
    """

    def __init__(self, loss_func, pit_from="pw_mtx", perm_reduce=None, per_stream=False, handle_provider=None):
        super().__init__()
        self.loss_func = loss_func
        self.pit_from = pit_from
        self.perm_reduce = perm_reduce
        self.per_stream = per_stream
        self._handle_provider = handle_provider  # callable(device) -> model._Handle (for the CUDA fast path)
        if self.pit_from not in ["pw_mtx", "pw_pt", "perm_avg"]:
            raise ValueError("Unsupported loss function type for now. Expectedone of [`pw_mtx`, `pw_pt`, `perm_avg`]")
        if perm_reduce is not None or pit_from == "perm_avg":
            raise NotImplementedError("septfa_b200.PITLossWrapper covers pw_pt / pw_mtx with the default mean reduce")

    def _fast_l1(self, est, tgt):
        """Pairwise L1 means [B,2,2] via septfa_pit_l1 (device sums / n)."""
        h = self._handle_provider(est.device)
        B, _, n = est.shape
        est = est.contiguous().float()
        tgt = tgt.contiguous().float()
        perm = torch.empty((B, 2), dtype=torch.int32, device=est.device)
        pw = torch.empty((B, 4), dtype=torch.float64, device=est.device)
        stream = C.c_void_p(torch.cuda.current_stream(est.device).cuda_stream)
        with torch.cuda.device(est.device):
            rc = h.lib.septfa_pit_l1(h.ptr, C.c_void_p(est.data_ptr()), C.c_void_p(tgt.data_ptr()), B, n,
                                     C.c_void_p(perm.data_ptr()), C.c_void_p(pw.data_ptr()), stream)
        _lib.check(h.ptr, rc)
        return (pw / n).view(B, 2, 2), perm

    def forward(self, est_targets, targets, target_vad=0, return_est=False, return_incides=False, reduce_kwargs=None,
                **kwargs):
        n_src = targets.shape[1]
        assert n_src < 10, f"Expected source axis along dim 1, found {n_src}"
        fast = (self.pit_from == "pw_pt" and isinstance(self.loss_func, torch.nn.L1Loss) and n_src == 2
                and est_targets.is_cuda and est_targets.ndim == 3 and self._handle_provider is not None
                and self.loss_func.reduction == "mean")
        if fast:
            pw_mean, perm_dev = self._fast_l1(est_targets, targets)  # [B,2,2] per-stream means
            if not self.per_stream:
                # nn.L1Loss batch-mean broadcast into every batch row (reference quirk)
                pw_mean = pw_mean.mean(dim=0, keepdim=True).expand(est_targets.shape[0], -1, -1)
            pw_losses = pw_mean.to(est_targets.dtype)
        elif self.pit_from == "pw_mtx":
            pw_losses = self.loss_func(est_targets, targets, **kwargs)
        else:
            pw_losses = self.get_pw_losses(self.loss_func, est_targets, targets, **kwargs)
            if self.per_stream and pw_losses.ndim == 3 and est_targets.shape[0] > 1:
                pw_losses = torch.stack([self.get_pw_losses(self.loss_func, est_targets[b:b + 1], targets[b:b + 1],
                                                            **kwargs)[0] for b in range(est_targets.shape[0])])
        assert pw_losses.ndim == 3, "Something went wrong with the loss function, please read the docs."
        assert pw_losses.shape[0] == targets.shape[0], "PIT loss needs same batch dim as input"
        min_loss, batch_indices = self.find_best_perm_factorial(pw_losses)
        mean_loss = torch.mean(min_loss)
        if not return_est and not return_incides:
            return mean_loss
        elif not return_est and return_incides:
            return mean_loss, batch_indices
        reordered = reorder_source_mse(est_targets, batch_indices)
        if return_est and return_incides:
            return mean_loss, reordered, batch_indices
        return mean_loss, reordered

    @staticmethod
    def get_pw_losses(loss_func, est_targets, targets, **kwargs):
        """Pairwise loss matrix ``pw[b, i, j] = loss_func(est[:, i], tgt[:, j])`` (semantics of
        model/pit_wrapper.py:149-177, an asteroid-derived MIT-licensed routine): whatever ``loss_func`` returns - a
        scalar for ``nn.L1Loss()`` - is broadcast over the batch axis."""
        n_src = targets.shape[1]
        pw = targets.new_empty(targets.shape[0], n_src, n_src)
        for i in range(n_src):
            for j in range(n_src):
                pw[:, i, j] = loss_func(est_targets[:, i], targets[:, j], **kwargs)
        return pw

    @staticmethod
    def find_best_perm_factorial(pair_wise_losses):
        """Exhaustive search (semantics of model/pit_wrapper.py:261-312 with ``perm_reduce=None``): the score of a
        permutation is the mean over targets ``j`` of ``pw[b, perm[j], j]``; ``torch.min`` keeps the first minimum,
        so ties go to the identity. For two sources this is identity vs swap:
        ``pw[0,0] + pw[1,1] <= pw[1,0] + pw[0,1]``."""
        n_src = pair_wise_losses.shape[-1]
        cand = list(permutations(range(n_src)))
        tgt = torch.arange(n_src, device=pair_wise_losses.device)
        scores = torch.stack([pair_wise_losses[:, list(c), tgt].sum(dim=1) / n_src for c in cand], dim=1)
        best, which = torch.min(scores, dim=1)
        table = torch.tensor(cand, dtype=torch.long, device=pair_wise_losses.device)
        return best, table[which]
