#!/usr/bin/env python
"""Benchmark of the Sep-TFAnet-VAD inference forward pass (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU port of the reference path (oracle)

A "step" is one forward pass over one batch of synthetic noisy two-speaker mixtures
(BASELINE.json configs[1]: config_with_vad.json, 256 x 4 s per GPU, filter_signals_by_smo_vad).
Metric: mixture-seconds processed per second, whole job (all N GPUs). Weak scaling: every rank
runs the same per-GPU batch on its own shard of mixtures; there is no collective on the data
path, torch.distributed only takes the MAX of the per-rank times.

Prints ONE JSON line (rank 0) with the keys the driver contract asks for, plus:
  roofline     - dominant kernel (tcgen05 dconv+res_out GEMM): algorithmic FLOP per launch divided by
                 its CUDA-event duration, against the measured bf16 peak (MEASURED_PEAKS.json)
  kernels      - per-kernel-class device time per step (CUDA events on the launch stream)
  cpu_baseline - the numpy port of the reference path timed on this box's host cores (bounded sample)
  e2e          - the same metric through SeparationModel.forward_host: pinned host buffers, H2D and D2H inside
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 16000
MAC_PER_FRAME = 4899657                     # algorithmic MACs per STFT frame
DCONV_MAC_PER_FRAME = 131072 + 1536         # res_out 512->256 + depthwise k3 (the dominant kernel's share)
CONV1_MAC_PER_FRAME = 65536
DCONV_DRAM_BYTES_NCU = 34972160           # k_tc_gemm<1,1,1>, 256 x 4 s: 33.76 MB read + 1.21 MB written (profiles/r1_ncu_summary.md)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 1400.0, 6650.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            load = [s for s in sm if s >= 0.5 * max(sm)] or sm
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def hbm_kernels(kernels, M, B, L, peak_hbm, exports):
    """Achieved HBM bandwidth of the bandwidth-class kernels: ALGORITHMIC bytes per step (compulsory traffic of the
    layout in DESIGN.md section 2, stated per frame in DESIGN.md section 3) over the class's CUDA-event time."""
    spec = 257 * 8 * M          # complex64 spectrogram
    lg = 2 * 257 * 4 * M        # both speakers' mask logits
    alg = {
        # resid: read stream fp32 + accumulators fp16, write stream fp32 (24 launches)
        "resid": 24 * M * (1024 + 512 + 1024),
        # frontend: read x, write S and the gated dB stream (+ the optional spectrum export: read stream, write [B,257,T])
        "frontend": B * L * 4 + spec + M * 1024 + ((M * 1024 + 257 * 4 * M) if exports else 0),
        # istft: read logits, S and gate, write both waveforms
        "istft": lg + spec + 2 * M * 4 + 2 * B * L * 4,
        # export: read logits and S, write est (complex64), masks and logits in torch layout
        "export": (lg + spec + 2 * spec + lg + lg) if exports else 0,
    }
    out = {}
    for k, nbytes in alg.items():
        ms = kernels.get(k, {}).get("ms_per_step", 0.0)
        if ms <= 0 or nbytes == 0:
            continue
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[k] = {"alg_bytes_per_step": int(nbytes), "ms_per_step": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak_hbm}
    return out


def cpu_reference_throughput(n_mix, length, repeats=1):
    """The oracle (numpy port of the reference path, fp32, BLAS threads = all host cores) on a bounded
    sample of the same workload. Returns (audio-seconds per second, seconds per pass)."""
    import numpy as np
    from oracle import septfa_oracle as O
    from septfa_b200 import synth
    try:  # torchrun exports OMP_NUM_THREADS=1: give the BLAS behind numpy all host cores explicitly
        from threadpoolctl import threadpool_limits
        threadpool_limits(os.cpu_count())
    except Exception:  # noqa: BLE001
        pass
    args = synth.CONFIG_WITH_VAD
    W = O.OracleWeights(synth.make_state_dict_numpy(args, 0), args, np.float32)
    x = synth.make_mixtures(n_mix, length, 1234)
    kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.forward(x, W, dict(kw))
        best = min(best, time.perf_counter() - t0)
    return n_mix * length / FS / best, best


def run_reference(a, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    n_mix = a.ref_sample
    # warm-up + timed steps, each a bounded sample of the workload
    for _ in range(a.warmup):
        cpu_reference_throughput(1, a.length)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_reference_throughput(n_mix, a.length)
    dt = time.perf_counter() - t0
    value = a.steps * n_mix * a.length / FS / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "mixture-seconds processed per second (offline forward)", "value": value,
        "unit": "audio-s/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config_with_vad, {a.batch} x {a.length / FS:g} s mixtures per GPU, "
                               "filter_signals_by_smo_vad", "sample": f"{n_mix} x {a.length / FS:g} s per step"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{n_mix} x {a.length / FS:g} s mixtures per step, {a.steps} steps, numpy fp32 "
                                   "port of model/model.py:402-461 (oracle/septfa_oracle.py)"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="mixtures per GPU per step")
    ap.add_argument("--length", type=int, default=64000, help="samples per mixture (4 s @ 16 kHz)")
    ap.add_argument("--ref-sample", type=int, default=48, help="mixtures per step of the CPU reference arm (~3 s of CPU work)")
    ap.add_argument("--cpu-sample", type=int, default=192, help="mixtures of the cpu_baseline leg (~12 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="skip the optional exports (est/mask/spectrum/logits)")
    ap.add_argument("--online-streams", type=int, default=1024, help="concurrent streams of the online leg (0 = skip)")
    ap.add_argument("--online-hops", type=int, default=12)
    a = ap.parse_args()

    # NCCL logs (its version banner included, which WARN and VERSION both print) go to stdout by default and would
    # precede the JSON line: send them to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if a.impl == "reference":
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[v] = str(os.cpu_count())  # torchrun forces OMP_NUM_THREADS=1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from septfa_b200 import synth
    from septfa_b200.model import SeparationModel
    from septfa_b200.shard import gather_max_time, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    args = synth.CONFIG_WITH_VAD
    with contextlib.redirect_stdout(io.StringIO()):
        model = SeparationModel(**args)
    model.load_state_dict(synth.make_state_dict(args, 0), strict=True)
    model.eval().to(dev)
    if a.lean:
        model.materialize.update(estimated_stfts=False, mask_per_speaker=False, spectrum=False, masks_b=False)
    kw = dict(synth.DEFAULT_INFERENCE_KW, filter_signals_by_smo_vad=True)

    # this rank's shard of the global batch (distinct mixtures per rank; 32 distinct ones tiled)
    B, L = a.batch, a.length
    g0, _ = shard_range(world * B, rank, world)
    n_distinct = min(B, 32)
    base = synth.make_mixtures(n_distinct, L, 1234, first_index=g0)
    x_host = torch.from_numpy(np.tile(base, ((B + n_distinct - 1) // n_distinct, 1))[:B]).pin_memory()
    x_dev = x_host.to(dev)
    T = 1 + L // 256

    def step():
        return model(x_dev, kw)

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): CUDA events on the launch stream, max over ranks
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_dev = gather_max_time(e0.elapsed_time(e1) * 1e-3)
    launches = model.last_launch_count * a.steps

    # ---- per-kernel-class timing (separate pass; events inside forward on the same stream)
    model.set_profile(True)
    for _ in range(a.steps):
        step()
    prof = model.read_profile(dev)
    model.set_profile(False)
    kernels = {k: {"ms_per_step": v[0] / a.steps, "launches_per_step": v[1] / a.steps} for k, v in prof.items()}

    # ---- end to end through the public host API (pinned host buffers, H2D + D2H inside)
    out_h = None
    for _ in range(2):
        out_h = model.forward_host(x_host, kw, device=local_rank)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out_h = model.forward_host(x_host, kw, device=local_rank)
    t_e2e_sync = gather_max_time(time.perf_counter() - t0)
    assert out_h[0].shape == (B, 2, L)

    # ---- the same through the batch-stream API (forward_host_submit / result, two slots): every step still copies
    # its input from pinned host memory and its results back to pinned host memory; successive steps overlap their
    # PCIe copies with each other's kernels, which is how a file loop (only_inference.py:80-100) would call it.
    # The timed region runs from the first submit to the last result (pipeline fill and drain included).
    xs = [x_host, x_host.clone().pin_memory()]
    outs = [torch.empty((B, 2, L), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    vads = [torch.empty((B, 2, T), dtype=torch.float32, pin_memory=True) for _ in range(2)]

    def stream_steps(n):
        pend = [None, None]
        for i in range(n):
            sl = i & 1
            if pend[sl] is not None:
                pend[sl].result()
            pend[sl] = model.forward_host_submit(xs[sl], kw, device=local_rank, slot=sl, out=outs[sl], vad=vads[sl])
        for f in pend:
            if f is not None:
                f.result()

    stream_steps(3)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    stream_steps(a.steps)
    t_e2e = gather_max_time(time.perf_counter() - t0)
    # same results as the synchronous call up to the fp16 operand noise: the two paths tile the batch differently, so a
    # VAD probability within ~4e-4 of the threshold may flip a frame's gate (allowed: |p - thr| < 1e-3) in a few clips
    assert torch.isfinite(outs[0]).all()
    assert ((outs[0] - out_h[0]).abs().amax(dim=(1, 2)) > 1e-3).float().mean().item() < 0.1
    clocks = sampler.stop() if sampler else None   # sampled over the timed, profiled and end-to-end legs

    # ---- online mode (BASELINE.json configs[2]): S concurrent streams, one hop-step = forward on the current 3 s
    # windows + per-stream L1-PIT + reorder + append; latency per hop-step from CUDA events, windows resident
    online = None
    if a.online_streams > 0 and rank == 0:
        import ctypes as C
        from septfa_b200 import lib as _lib
        S = a.online_streams
        h = model._handle(dev)
        st = C.c_void_p()
        _lib.check(h.ptr, h.lib.septfa_online_create(h.ptr, S, C.byref(st)))
        ws = torch.empty(h.lib.septfa_online_workspace_bytes(st), dtype=torch.uint8, device=dev)
        win = torch.from_numpy(np.tile(synth.make_mixtures(16, 48000, 4321), ((S + 15) // 16, 1))[:S]).to(dev)
        emitted = torch.empty((S, 2, 16000), dtype=torch.float32, device=dev)
        perm = torch.empty((S, 2), dtype=torch.int32, device=dev)
        ikw = _lib.InferKw.from_dict(kw)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        times = []
        for i in range(a.online_hops + 3):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            _lib.check(h.ptr, h.lib.septfa_online_step(st, C.c_void_p(win.data_ptr()), C.byref(ikw), C.c_void_p(emitted.data_ptr()),
                                                       C.c_void_p(perm.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), stream))
            s1.record()
            torch.cuda.synchronize()
            if i >= 3:
                times.append(s0.elapsed_time(s1))
        h.lib.septfa_online_destroy(st)
        del ws
        times.sort()
        p50 = times[len(times) // 2]
        p99 = times[min(len(times) - 1, int(0.99 * len(times)))]
        online = {"streams": S, "hops_timed": len(times), "p50_ms_per_hop_step": p50, "p99_ms_per_hop_step": p99,
                  "p50_ms_per_frame": p50 / 62.5, "p99_ms_per_frame": p99 / 62.5,
                  "audio_s_per_s": S * 1.0 / (p50 * 1e-3),
                  "note": "hop-step = forward on S x 3 s windows (T=188) + per-stream L1-PIT + reorder + emit 1 s; "
                          "ms per frame = hop-step / 62.5 frames per hop (the reference has no per-frame entry point)"}

    if rank == 0:
        audio_s = world * B * L / FS * a.steps
        peak_tf, peak_hbm, peak_src = measured_peaks()
        M = B * T
        d_ms, d_n = prof["dconv"]
        dconv_ms = d_ms / max(d_n, 1)
        flop_per_launch = 2.0 * DCONV_MAC_PER_FRAME * M
        achieved = flop_per_launch / (dconv_ms * 1e-3) / 1e12 if dconv_ms > 0 else 0.0
        c_ms, c_n = prof["conv1"]
        conv1_tf = 2.0 * CONV1_MAC_PER_FRAME * M / (c_ms / max(c_n, 1) * 1e-3) / 1e12 if c_ms > 0 else 0.0
        line = {
            "metric": "mixture-seconds processed per second (offline forward)",
            "value": audio_s / t_dev, "unit": "audio-s/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16xf16->f32 (tcgen05), f32 elsewhere", "data": "synthetic",
            "config": {"workload": f"config_with_vad, {B} x {L / FS:g} s mixtures per GPU, filter_signals_by_smo_vad "
                                   "(BASELINE.json configs[1])", "frames_per_step_per_gpu": M,
                       "l2": "working set ~0.6 GB per step >> 126 MB L2 (no explicit flush)",
                       "weights": "seeded random-init, reference layout", "exports": not a.lean},
            "model_flops_utilization": 2.0 * MAC_PER_FRAME * M * world * a.steps / t_dev / 1e12 / (peak_tf * world),
            "roofline": {"kernel": "k_tc_gemm<1> (depthwise conv prologue + res_out 512->256 tcgen05 GEMM)",
                         "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": DCONV_DRAM_BYTES_NCU if (B, L) == (256, 64000) else None,
                         "traffic_source": "profiles/r1_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                           "k_tc_gemm<1,1,1> launch (ncu --set full); the other 31 MB of its 65.8 MB algorithmic "
                                           "fp16 in/out bytes are served by / left in the 126 MB L2",
                         "peak_source": peak_src,
                         "ms_per_launch": dconv_ms, "flop_per_launch": flop_per_launch,
                         "conv1_tflops": conv1_tf},
            "kernels": kernels,
            "hbm_kernels": hbm_kernels(kernels, M, B, L, peak_hbm, not a.lean),
            "e2e": {"value": audio_s / t_e2e, "unit": "audio-s/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                    "d2h_bytes_per_step": int(B * 2 * L * 4 + B * 2 * T * 4), "ms_per_step": 1e3 * t_e2e / a.steps,
                    "api": "SeparationModel.forward_host_submit / HostBatch.result (septfa_forward_host_submit / _wait), "
                           "two batches in flight; pinned host input and output per step; fill and drain inside the timing"},
            "e2e_sync": {"value": audio_s / t_e2e_sync, "unit": "audio-s/s", "ms_per_step": 1e3 * t_e2e_sync / a.steps,
                         "api": "SeparationModel.forward_host (septfa_forward_host): one synchronous call per step, "
                                "two-chunk copy/compute pipeline inside the call"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if online is not None:
            line["online"] = online
        if world == 1 and not a.no_cpu_baseline:
            v, secs = cpu_reference_throughput(a.cpu_sample, L)
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{a.cpu_sample} x {L / FS:g} s mixtures, one pass ({secs:.1f} s), numpy "
                                              "fp32 port of model/model.py:402-461 (oracle/septfa_oracle.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
