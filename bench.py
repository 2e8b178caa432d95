#!/usr/bin/env python
"""Benchmark of the SENSE data-consistency hot path (BASELINE.json configs[1]).

Workload (config.workload = "config_recon.yaml"): unrolled-network inference on a batch of 64
synthetic cine slices per GPU — E=1 map set, 8 coils, 20 frames, 192x160, 12x k-t mask, NUM_UNROLLS=5
(/root/reference/configs/config_recon.yaml:5-17).  One *step* is one pass of the hot path over that
batch: ATy = A^H y, then 5 data-consistency updates x <- x + s (A^H A x - ATy)
(dl_cs/models/unrolled.py:96-104), through the reference-facing API (SenseModel + the PGD loop mirror).
The denoiser between the updates is not part of this path and is not run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = slices/s with inputs resident in HBM; `e2e` = the same through
the same API from pinned host buffers (H2D of y/maps/mask and D2H of x inside the timed region);
`roofline` = the dominant SENSE op (the fused DC step) in algorithmic GB/s against the measured HBM
peak; `ops` gives forward / adjoint / normal too; `cpu_baseline` = the torch-CPU port of the
reference path (oracle/sense_oracle_torch.py) on a bounded sample.  `--impl reference` times only
that CPU port (the Python reference itself cannot travel to the GPU box; kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOAD = dict(B=64, E=1, C=8, T=20, Y=192, X=160, unrolls=5, step=-2.0, accel=12)
METRIC = "cine slices/sec (SENSE data-consistency path of the unrolled recon)"
UNIT = "slices/s"


# --------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes(op: str, B, E, C, T, Y, X, mask_elems) -> int:
    P = B * T * Y * X
    maps = 8 * B * E * C * Y * X
    M = 4 * mask_elems
    if op in ("forward", "adjoint"):
        return 8 * P * (E + C) + maps + M
    if op == "normal":
        return 16 * P * E + maps + M
    if op == "dc_step":
        return 24 * P * E + maps + M
    raise ValueError(op)


def algorithmic_flops(op: str, B, E, C, T, Y, X) -> float:
    n = Y * X
    fft = 5.0 * n * np.log2(n) * B * C * T
    mac = 8.0 * B * T * Y * X * E * C
    return (fft + mac) * (1 if op in ("forward", "adjoint") else 2)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons from NVML while the timed region runs.

    No background thread: NVML queries can hold a driver lock that stalls kernel launches (a 20 ms poller doubled the
    step time in one run), so the host first queues all K timed steps and only then polls NVML, from the launching
    thread, until the closing event has completed - the GPU is still executing the timed steps while it is sampled."""

    _NAMES = (("nvmlClocksThrottleReasonHwSlowdown", 0x8, "hw_slowdown"),
              ("nvmlClocksThrottleReasonHwThermalSlowdown", 0x40, "hw_thermal_slowdown"),
              ("nvmlClocksThrottleReasonSwThermalSlowdown", 0x20, "sw_thermal_slowdown"),
              ("nvmlClocksThrottleReasonSwPowerCap", 0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.sample(record=False)         # NVML's first queries are the slow ones: pay for them before the timed region
        except Exception:
            self.nv = None

    def sample(self, record=True):
        nv = self.nv
        if nv is None:
            return
        try:
            mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            return
        if record:
            self.samples.append(mhz)
            for attr, default, name in self._NAMES:
                if r & getattr(nv, attr, default):
                    self.reasons.add(name)

    def poll_until(self, event, period_s=0.01, max_samples=200):
        """Sample until the CUDA event has completed (at least once)."""
        while True:
            self.sample()
            if event.query() or len(self.samples) >= max_samples:
                return
            time.sleep(period_s)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- synthetic data
def make_mask(B, T, Y, X, accel, seed):
    """0/1 k-t masks [B,1,T,Y,X] from the reference's generator (VDktMaskFunc, dl_cs/mri/subsample.py:65-117, mirrored
    bit-exactly in dl_swin_gan_b200/mri/subsample.py): variable-density ky-t pattern at R = accel, golden-ratio shift per
    phase, partial readout and partial phase encoding 0.25; slice b uses seed + b."""
    from dl_swin_gan_b200.mri.subsample import VDktMaskFunc
    f = VDktMaskFunc((accel, accel), sim_partial_kx=0.25, sim_partial_ky=0.25)
    return np.concatenate([f((1, 1, T, Y, X), seed=seed + b).numpy() for b in range(B)], axis=0)


def device_problem(cfg, dev, seed):
    """maps, mask, y = mask * A_full(x_true) generated on the device (seed 1000 + rank)."""
    from dl_swin_gan_b200.mri import SenseModel
    B, E, C, T, Y, X = (cfg[k] for k in "BECTYX")
    g = torch.Generator(device=dev).manual_seed(seed)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev)) * (0.5 ** 0.5)

    maps = cn(B, E, C, 1, Y, X) / (E * C) ** 0.5
    x_true = cn(B, E, T, Y, X)
    mask = torch.from_numpy(make_mask(B, T, Y, X, cfg["accel"], seed)).to(dev)
    y = SenseModel(maps, weights=mask)(x_true)
    return maps, mask, y


# --------------------------------------------------------------------------- CPU port arm
def cpu_port_rate(cfg, seconds_target=12.0):
    """slices/s of the torch-CPU port of the reference path on a bounded sample of the workload."""
    from oracle import sense_oracle_torch as st
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    E, C, T, Y, X = (cfg[k] for k in "ECTYX")
    g = torch.Generator().manual_seed(1000)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g)) * (0.5 ** 0.5)

    nb = 1
    maps = cn(nb, E, C, 1, Y, X) / (E * C) ** 0.5
    mask = torch.from_numpy(make_mask(nb, T, Y, X, cfg["accel"], 1000))
    y = mask * cn(nb, C, T, Y, X)
    st.dc_path(y, maps, mask, 1, cfg["step"])            # warm-up (thread pool, FFT plans)
    t0 = time.perf_counter()
    n = 0
    while True:
        st.dc_path(y, maps, mask, cfg["unrolls"], cfg["step"])
        n += nb
        el = time.perf_counter() - t0
        if el > seconds_target or n >= 4096:
            break
    return n / el, threads, f"{n} slice(s) of the batch through A^H y + {cfg['unrolls']} DC updates, {el:.1f} s"


def run_reference_arm(args, rank):
    cfg = dict(WORKLOAD)
    if rank != 0:
        return
    rates = []
    for _ in range(max(1, args.warmup)):
        pass                                              # the port warms itself up inside cpu_port_rate
    budget = max(2.0, min(20.0, 60.0 / max(1, args.steps)))
    for _ in range(args.steps):
        rate, threads, sample = cpu_port_rate(cfg, seconds_target=budget)
        rates.append(rate)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cfg["B"] / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (complex64)",
        "data": "synthetic", "config": {"workload": "config_recon.yaml", **cfg},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- window attention (second half of the hot path)
def swin_block_timing(dev, slices=8):
    """forward_part1 of one Swin block (LN, shift, partition, qkv, attention, proj, reverse) for `slices` cine slices at the
    hard-coded denoiser configuration (dim 160, 8 heads, window (7,8,8), feature volume 7x48x40), shifted block."""
    from dl_swin_gan_b200.swin import window_attention as wa
    torch.manual_seed(0)
    D, H, W, C, heads = 7, 48, 40, 160, 8
    blk = wa.SwinTransformerBlock3D(C, heads, (7, 8, 8), (3, 4, 4), qkv_bias=True).to(dev).eval()
    x = torch.randn(slices, D, H, W, C, device=dev)
    ws, ss = wa.get_window_size((D, H, W), (7, 8, 8), (3, 4, 4))
    mask = wa.compute_mask(7, 48, 40, ws, ss, dev)
    with torch.no_grad():
        n0 = _lib_launches()
        ms = time_op(lambda: blk.forward_part1(x, mask), 10, dev)
        launches = (_lib_launches() - n0) // 12
        ms2 = time_op(lambda: blk._part2_tensor_core(x, residual=True), 10, dev)
        ms_block = time_op(lambda: blk(x, mask), 10, dev)
    windows = slices * mask.shape[0]
    useful = windows * (2 * 448 * 160 * 480 + 4 * 448 * 448 * 160 + 2 * 448 * 160 * 160)      # qkv + QK^T/PV + proj
    return {"what": "SwinTransformerBlock3D.forward_part1, shifted, fp16 tensor-core operands", "slices": slices,
            "windows": windows, "ms": ms, "us_per_slice": 1e3 * ms / slices, "useful_TFLOPs": useful / (ms * 1e-3) / 1e12,
            "launches_per_call": launches,
            "part2_mlp_ms": ms2, "part2_useful_TFLOPs": 4.0 * x.numel() * 4 * C / (ms2 * 1e-3) / 1e12,
            "whole_block_ms": ms_block, "whole_block_launches": 5}


def _lib_launches():
    from dl_swin_gan_b200 import _lib
    return int(_lib.load().dlsg_launch_count())


# --------------------------------------------------------------------------- main arm
def time_op(fn, iters, dev):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(dev)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize(dev)
    return s.elapsed_time(e) / iters   # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["B"], help="slices per GPU (default: the config's 64)")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-swin", action="store_true", help="skip the window-attention side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import __graft_entry__ as ge
    if rank == 0 or not os.path.exists(ge.LIB):
        ge.build()
    from dl_swin_gan_b200 import _lib
    from dl_swin_gan_b200.mri import SenseModel
    from dl_swin_gan_b200.unrolled import ProximalGradientDescent
    from dl_swin_gan_b200.sharding import max_over_ranks

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = dict(WORKLOAD)
    cfg["B"] = args.batch
    B, E, C, T, Y, X = (cfg[k] for k in "BECTYX")
    lib = _lib.load()

    maps, mask, y = device_problem(cfg, dev, 1000 + rank)
    A = SenseModel(maps, weights=mask)
    pgd = ProximalGradientDescent(num_unrolls=cfg["unrolls"], step_init=cfg["step"]).to(dev)

    def step():
        with torch.no_grad():
            return pgd(y, A)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    # a fresh process / box needs more than three short steps to reach steady clocks and page everything in: keep warming up
    # (untimed) until 0.5 s have passed; config.warmup_extra_steps reports how many steps that added to the W requested
    clocks = ClockSampler(local_rank)
    # The timed steps are replays of ONE CUDA graph captured from the same public-API calls: a host-side hiccup (GC, CPU
    # throttling of the container; seen as 10-50 ms gaps in eager launches on this pool, tools/stability_probe.py) then cannot
    # starve the GPU inside the timed region.  DLSG_BENCH_EAGER=1 times eager launches instead.
    use_graph = os.environ.get("DLSG_BENCH_EAGER", "0") != "1"
    per_step_launches = None
    if use_graph:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        n_cap = lib.dlsg_launch_count()
        with torch.cuda.graph(graph):
            out = step()
        per_step_launches = int(lib.dlsg_launch_count() - n_cap)
        run_step = graph.replay
    else:
        def run_step():
            nonlocal out
            out = step()
    out = out if use_graph else step()
    extra_warmup = 0
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.5:
        run_step()
        torch.cuda.synchronize(dev)
        extra_warmup += 1
    barrier()
    n0 = lib.dlsg_launch_count()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_ev.record()
    for _ in range(args.steps):
        run_step()
    e_ev.record()
    clocks.poll_until(e_ev)           # all K steps are queued; the GPU is still running them while NVML is read
    barrier()
    launches = args.steps * per_step_launches if use_graph else int(lib.dlsg_launch_count() - n0)
    ms_per_step = max_over_ranks(s_ev.elapsed_time(e_ev), dev) / args.steps      # device-timed, max over ranks
    value = world * B / (ms_per_step * 1e-3)

    # ---- per-op timings for the roofline (CUDA events on the launching stream, batch >> L2)
    hbm_peak, peak_src = measured_peaks()
    x0 = out
    aty = A(y, adjoint=True)
    ops = {}
    iters = 5
    with torch.no_grad():
        t_ms = {
            "forward": time_op(lambda: A(x0), iters, dev),
            "adjoint": time_op(lambda: A(y, adjoint=True), iters, dev),
            "normal": time_op(lambda: A.normal(x0), iters, dev),
            "dc_step": time_op(lambda: A.dc_step(x0, aty, pgd.step_size), iters, dev),
        }
    for name, t in t_ms.items():
        nbytes = algorithmic_bytes(name, B, E, C, T, Y, X, mask.numel())
        ops[name] = {"ms": t, "algorithmic_MB_per_slice": nbytes / B / 1e6, "GBps": nbytes / (t * 1e-3) / 1e9,
                     "frac_hbm": nbytes / (t * 1e-3) / 1e9 / hbm_peak,
                     "TFLOPs_nominal": algorithmic_flops(name, B, E, C, T, Y, X) / (t * 1e-3) / 1e12}
    # the same ops on the dense kernels (what runs when W has no empty ky lines), for comparison
    plan = A._line_plans.get((B, E, C, T, Y, X))
    path = "sampled-lines" if plan is not None else "dense"
    ops_dense = None
    if plan is not None:
        os.environ["DLSG_SAMPLED_LINES"] = "0"
        with torch.no_grad():
            ops_dense = {"forward_ms": time_op(lambda: A(x0), iters, dev),
                         "adjoint_ms": time_op(lambda: A(y, adjoint=True), iters, dev),
                         "dc_step_ms": time_op(lambda: A.dc_step(x0, aty, pgd.step_size), iters, dev)}
        os.environ["DLSG_SAMPLED_LINES"] = "1"
    dom = "dc_step"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_dram_traffic_per_slice.json")
    tkey = "ops_MB_per_slice_sampled_lines" if plan is not None else "ops_MB_per_slice"
    if os.path.exists(tpath) and (E, C, T, Y, X) == (1, 8, 20, 192, 160):
        with open(tpath) as f:
            tj = json.load(f)
        if tkey in tj:
            traffic = tj[tkey][dom] * 1e6 * B      # ncu-measured DRAM bytes per launch chain
    chain = ("cols_fwd_lines + rows_lines<NORMAL> + cols_adj_lines" if plan is not None
             else "rows_fwd + cols<NORMAL> + rows_adj")
    roofline = {"bound": "hbm", "kernel": f"fused DC step chain ({chain})",
                "achieved": ops[dom]["GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": ops[dom]["frac_hbm"],
                "traffic": traffic, "traffic_source": "ncu --set full, profiles/r1_dram_traffic_per_slice.json",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algorithmic_bytes(dom, B, E, C, T, Y, X, mask.numel()),
                # the fused DC step is FP32-issue-bound, not HBM-bound (no k-space in or out, two 2-D transforms per coil image):
                # nominal FFT + coil-MAC flops against the CUDA-core FMA peak (SMs x 128 lanes x 2 x max SM clock), for context
                "fp32_nominal_TFLOPs": ops[dom]["TFLOPs_nominal"],
                "fp32_peak_TFLOPs": torch.cuda.get_device_properties(dev).multi_processor_count * 128 * 2
                                    * (clocks.max_mhz or 1965) * 1e6 / 1e12}

    # ---- end to end from pinned host memory through the same API
    e2e = None
    if not args.skip_e2e:
        hy = torch.empty(y.shape, dtype=y.dtype, pin_memory=True).copy_(y)
        hm = torch.empty(maps.shape, dtype=maps.dtype, pin_memory=True).copy_(maps)
        hw = torch.empty(mask.shape, dtype=mask.dtype, pin_memory=True).copy_(mask)
        hx = torch.empty((B, E, T, Y, X), dtype=torch.complex64, pin_memory=True)
        d2h = hx.numel() * 8

        # The batch goes through the same public API in chunks of 16 slices on two streams, so that the H2D traffic of
        # chunk i+1 and the D2H copy of chunk i-1 overlap the reconstruction of chunk i (slices are independent).
        # The zero-filled k-space stays in pinned host memory: A(y, adjoint=True) reads its sampled lines over PCIe
        # (dlsg_sense_adjoint_lines_gather_c64); maps and mask are copied.  With dense weights it is copied as before.
        host_y = plan is not None and os.environ.get("DLSG_E2E_HOST_Y", "1") != "0"
        n_chunk = int(os.environ.get("DLSG_E2E_CHUNKS", 8))
        if n_chunk < 1 or B % n_chunk:
            n_chunk = 1
        cb = B // n_chunk
        n_streams = int(os.environ.get("DLSG_E2E_STREAMS", 2))
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        xfer = torch.cuda.Stream(device=dev, priority=-1)     # H2D copies, line plan and A^H y (PCIe-bound) ahead of the DC steps
        if host_y:
            sampled_rows = int(plan[0][: B * T].sum().item())                  # sum of n(b,t)
            y_bytes = sampled_rows * C * X * 8
        else:
            y_bytes = hy.numel() * 8
        h2d = y_bytes + hm.numel() * 8 + hw.numel() * 4

        def e2e_step():
            with torch.no_grad():
                for i in range(n_chunk):
                    sl = slice(i * cb, (i + 1) * cb)
                    st = streams[i % n_streams]
                    with torch.cuda.stream(xfer):              # chunk i+1's traffic overlaps chunk i's DC steps
                        dw = hw[sl].to(dev, non_blocking=True)
                        dm = hm[sl].to(dev, non_blocking=True)
                        dy = hy[sl] if host_y else hy[sl].to(dev, non_blocking=True)
                        A_i = SenseModel(dm, weights=dw)
                        aty_i = A_i(dy, adjoint=True)
                        ev = torch.cuda.Event()
                        ev.record(xfer)
                    with torch.cuda.stream(st):
                        st.wait_event(ev)
                        for tns in (dm, dw, aty_i):
                            tns.record_stream(st)
                        res = pgd(None, A_i, ATy=aty_i)
                        hx[sl].copy_(res, non_blocking=True)
                for st in streams:
                    torch.cuda.current_stream(dev).wait_stream(st)

        n_e2e = max(2, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        el = max_over_ranks(time.perf_counter() - t0, dev)
        e2e = {"value": world * B * n_e2e / el, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "pipeline": f"{n_chunk} chunks of {cb} slices: H2D + line plan + A^H y on a high-priority transfer stream, DC steps + D2H on {n_streams} streams",
               "kspace": ("left in pinned host memory, sampled ky lines read in place over PCIe" if host_y
                          else "dense copy to the device")}

    swin = None
    if rank == 0 and not args.skip_swin:
        swin = swin_block_timing(dev)

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        rate, threads, sample = cpu_port_rate(cfg)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (complex64)", "data": "synthetic",
            "config": {"workload": "config_recon.yaml", **cfg, "slices_per_gpu": B,
                       "l2": "working set per step (y 2.5 GB + x/maps/mask) >> 126 MB L2, no flush needed",
                       "workspace_MB": float(os.environ.get("DLSG_WORKSPACE_MB", 4096)),
                       "kernel_path": path, "sampled_ky_lines_max": (plan[1] if plan is not None else Y),
                       "launch": ("one CUDA graph of the step (captured from the public API calls), replayed K times" if use_graph
                                  else "eager launches"),
                       "warmup_extra_steps": extra_warmup},
            "roofline": roofline, "ops": ops, "ops_dense_kernels": ops_dense, "swin_window_attention": swin, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
