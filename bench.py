#!/usr/bin/env python3
"""bench.py -- 1080p frames/s of the MIP mode-decision path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  (N > 1: launched by torch.distributed.run, one rank per GPU; frames are sharded, there is
   no data-path collective -- torch.distributed only carries the barrier and the MAX of the
   per-rank device times)

Workload (BASELINE.json configs[1]): synthetic 1920x1080 10-bit frames, alternative samples,
--FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2.  A step = one pass of the
hot path (filter -> MIP costs -> decisions) over a batch of B distinct frames per GPU.

  value  frames/s with the frame pool already resident in HBM (device-timed, CUDA events)
  e2e    frames/s through the C ABI host path: host frames -> pinned ring -> H2D -> kernels -> D2H of
         the MIP decisions (best mode + its cost for every CU); e2e_costs additionally reads back the
         full int32 cost table (the reference's minSadHad readback, 52.8 MB per 1080p frame)
  roofline      the fused cost kernel against the binding roof: INT32 issue (algorithmic INT32 ops of
                BASELINE.md section 2 / kernel time / measured INT32 peak); the HBM view (algorithmic
                bytes / kernel time / measured copy bandwidth) is nested as roofline.hbm
  cpu_baseline  the CPU oracle (port of the reference algorithm, OpenMP, all host cores) on a
                bounded sample of the same workload (rank 0, N == 1 only)

--impl reference: the reference is OpenCL-only.  The arm runs its UNMODIFIED kernels on one B200 through
NVIDIA's OpenCL driver (oracle/_ref/mipref_ocl, built from /root/reference where it lies) and reports
the CPU port beside it (cpu_baseline); without an OpenCL runtime it falls back to the CPU port alone.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))

W, H = 1920, 1080
FILTER_NAME = "filterFrame_2d_float_5x5_quarterCtu"
FILTER_TYPE, KERNEL_IDX = 8, 2
N_CTUS = 135
COSTS_PER_FRAME_IN = 12_359_520          # (CU, mode) costs of CUs inside a 1080p frame
OPS_PER_FRAME = 1.3148e10                # algorithmic INT32 ops per 1080p frame
ALGO_BYTES_PER_FRAME = 2 * W * H + 4 * COSTS_PER_FRAME_IN   # 53.6 MB: frame in + int32 costs out


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _int32_peak():
    p = os.path.join(ROOT, "profiles", "int32_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["int32_tops"]), d.get("how", "profiles/int32_peak.json")
    return 148 * 128 * 1.965e-3, "nominal 148 SMs x 128 lanes x 1.965 GHz (unmeasured)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def _frame_pool(n: int, seed0: int):
    from mipb200 import frames
    return [frames.natural_frame(W, H, seed0 + i) for i in range(n)]


def _cpu_port_fps(pool, seconds: float, max_frames: int):
    """CPU oracle (port of the reference algorithm, OpenMP on all host cores) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    n = 0
    while n < max_frames and (n == 0 or time.perf_counter() - t0 < seconds):
        O.run_frame(pool[n % len(pool)], FILTER_TYPE, KERNEL_IDX, threads=cores)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} frame(s) of the same workload through oracle/mip_oracle.c (OpenMP, {cores} threads)"}


def _reference_opencl(frame, reps: int):
    """The reference's own unmodified OpenCL kernels on this box's GPU (oracle/_ref/mipref_ocl)."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "mipref_ocl")
    if not os.path.exists(exe):
        return None, "oracle/_ref/mipref_ocl not built (needs /root/reference at build time)"
    with tempfile.TemporaryDirectory() as d:
        fp = os.path.join(d, "f.u16")
        frame.astype("<u2").tofile(fp)
        try:
            r = subprocess.run([exe, fp, str(W), str(H), FILTER_NAME, str(KERNEL_IDX), os.path.join(d, "out"), str(reps)],
                               capture_output=True, text=True, timeout=600)
        except subprocess.TimeoutExpired:
            return None, "mipref_ocl timed out"
    try:
        info = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None, f"mipref_ocl rc={r.returncode}: {r.stderr[-200:]}"
    if "unavailable" in info:
        return None, info["unavailable"]
    return info, None


def run_reference(args, rank: int, world: int) -> None:
    """Reference arm.  The reference is single-device OpenCL with no CPU implementation of its own, so the
    strongest available baseline is used: its unmodified kernels on ONE B200 through NVIDIA's OpenCL driver
    (the number BASELINE.json's ">= 50x" refers to).  Where no OpenCL runtime can be loaded the arm falls back
    to the CPU port.  The CPU port is reported beside it either way (cpu_baseline)."""
    if rank != 0:
        return
    pool = _frame_pool(2, 0)
    cfg = {"workload": f"1920x1080 10-bit natural-like synthetic frames, alternative samples {FILTER_NAME} KernelIdx={KERNEL_IDX}",
           "sample": "one frame per step"}
    cpu = _cpu_port_fps(pool, 10.0, 4)
    info, why = _reference_opencl(pool[0], max(1, min(args.steps, 20)))
    if info is not None:
        frame_bytes = 2 * W * H
        line = {
            "impl": "reference", "metric": "1080p frames/s", "value": info["fps_kernels"], "unit": "frames/s", "n_gpus": 1,
            "steps": info["reps"], "warmup": 1, "ms_per_step": 1e3 / info["fps_kernels"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "reference_kind": "the reference's unmodified OpenCL kernels (intra.cl) on one %s via %s, host = oracle/ocl_ref (full grid for every frame)" % (info["device"], info["opencl_lib"]),
            "kernel_ms": {k: v for k, v in info.items() if k.startswith("ms_")},
            "e2e": {"value": info["fps_e2e"], "unit": "frames/s", "h2d_bytes_per_step": 2 * frame_bytes, "d2h_bytes_per_step": 8 * N_CTUS * 97840},
            "cpu_baseline": cpu, "gpu_launches": 0,
        }
    else:
        line = {
            "impl": "reference", "metric": "1080p frames/s", "value": cpu["value"], "unit": "frames/s", "n_gpus": 1,
            "steps": args.steps, "warmup": 0, "ms_per_step": 1e3 / cpu["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "reference_kind": "CPU port of the reference algorithm (OpenCL unavailable: %s)" % why,
            "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cpu_baseline": cpu, "gpu_launches": 0,
        }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frames per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import mipb200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = args.batch
    dev = torch.device("cuda", local_rank)
    pool_np = _frame_pool(B, 1000 * rank)     # per-GPU work is fixed as N grows: weak scaling
    emit_dec = mipb200.EMIT_DECISIONS
    emit_full = mipb200.EMIT_COSTS | mipb200.EMIT_DECISIONS
    NS = 3                                      # frames in flight, like the engine's host path (3 slots)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    # frames already in HBM; filter -> fused MIP cost kernel -> decisions per frame, frames round-robin
    # over NS streams (independent frames overlap at kernel tails exactly as in the host path)
    engs = [mipb200.Engine(W, H, device=local_rank, filter_type=FILTER_TYPE, kernel_idx=KERNEL_IDX, slots=1, emit=emit_dec) for _ in range(NS)]
    d_pool = torch.from_numpy(np.stack(pool_np).view(np.int16)).to(dev)            # B x H x W, resident in HBM
    d_cost = torch.empty((NS, N_CTUS, mipb200.COSTS_PER_CTU), dtype=torch.int32, device=dev)
    d_bm = torch.empty((NS, N_CTUS, mipb200.CUS_PER_CTU), dtype=torch.uint8, device=dev)
    d_bc = torch.empty((NS, N_CTUS, mipb200.CUS_PER_CTU), dtype=torch.int32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]   # real (non-default) streams
    torch.cuda.synchronize()

    def step_device():
        for i in range(B):
            k = i % NS
            engs[k].run_device(d_pool[i].data_ptr(), d_cost[k].data_ptr(), d_best_mode=d_bm[k].data_ptr(),
                               d_best_cost=d_bc[k].data_ptr(), stream=streams[k].cuda_stream)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sum(e.kernel_launches() for e in engs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(streams[0])
    for s_ in streams[1:]:
        s_.wait_stream(streams[0])
    for _ in range(args.steps):
        step_device()
    for s_ in streams[1:]:
        streams[0].wait_stream(s_)
    ev1.record(streams[0])
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = sum(e.kernel_launches() for e in engs) - l0
    clocks = sampler.stop()
    for e in engs:
        e.close()

    # dominant kernel in isolation (roofline numerator): the fused cost kernel back to back on one stream
    eng_k = mipb200.Engine(W, H, device=local_rank, filter_type=0, slots=1, emit=mipb200.EMIT_COSTS)
    sp = streams[0].cuda_stream
    for _ in range(3):
        eng_k.run_device(d_pool[0].data_ptr(), d_cost[0].data_ptr(), stream=sp)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(16, min(96, args.steps * B))
    k0.record(streams[0])
    for i in range(reps):
        eng_k.run_device(d_pool[i % B].data_ptr(), d_cost[i % NS].data_ptr(), stream=sp)
    k1.record(streams[0])
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / reps
    eng_k.close()

    # ---------------- end to end through the host API (`e2e`, `e2e_costs`) ----------------
    # the application's frames live in page-locked host memory (as the contract asks): the engine DMAs them in place
    pool_pin = torch.empty((B, H, W), dtype=torch.int16, pin_memory=True)
    pool_pin.numpy()[...] = np.stack(pool_np).view(np.int16)
    pool_host = [pool_pin[i].numpy().view(np.uint16) for i in range(B)]

    def step_host(e, touch_costs):
        sub = got = 0
        checksum = 0
        while got < B:
            while sub < B and e.in_flight() < NS:
                e.submit(pool_host[sub], poc=sub)    # async H2D from pinned memory + fused kernel + async D2H
                sub += 1
            r = e.collect()                          # waits for this frame's results to be resident on the host
            checksum += int(r.best_cost[0, 0]) + int(r.best_mode[-1, -1])
            if touch_costs:
                checksum += int(r.cost[-1, -1])
            got += 1
        return checksum

    def time_host(emit, touch_costs):
        e = mipb200.Engine(W, H, device=local_rank, filter_type=FILTER_TYPE, kernel_idx=KERNEL_IDX, slots=NS, emit=emit)
        for _ in range(2):
            step_host(e, touch_costs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host(e, touch_costs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e.close()
        return dt

    e2e_steps = max(2, min(args.steps, 8))
    e2e_dec_s = time_host(emit_dec, False)
    e2e_s = time_host(emit_full, True)

    # ---------------- aggregate over ranks (MAX of times) ----------------
    times = torch.tensor([dev_ms, e2e_s * 1e3, e2e_dec_s * 1e3, kernel_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_dec_ms, kernel_ms = times.tolist()
    total_frames = B * args.steps * world
    value = total_frames / (dev_ms * 1e-3)
    e2e_fps = B * e2e_steps * world / (e2e_ms * 1e-3)
    e2e_dec_fps = B * e2e_steps * world / (e2e_dec_ms * 1e-3)

    if rank == 0:
        hbm_peak, peak_src = _peaks()
        int32_peak, int32_src = _int32_peak()
        achieved_gbs = ALGO_BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "cost_kernel_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        cpu = _cpu_port_fps(pool_np, 12.0, 8) if (world == 1 and not args.no_cpu_baseline) else None
        frame_bytes = 2 * W * H
        d2h_frame = 4 * N_CTUS * mipb200.COSTS_PER_CTU + 5 * N_CTUS * mipb200.CUS_PER_CTU
        line = {
            "metric": "1080p frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"1920x1080 10-bit natural-like synthetic frames, alternative samples {FILTER_NAME} KernelIdx={KERNEL_IDX}; "
                                   f"batch of {B} distinct frames per GPU per step; filter + MIP costs (97840 per CTU) + decisions",
                       "frames_per_step_per_gpu": B, "sharding": f"frames over {world} GPU(s), no collective",
                       "l2": f"input pool {(B * frame_bytes) >> 20} MiB + 3 rotating 50 MiB cost tables exceed the 126 MiB L2 (no flush needed)"},
            "e2e": {"value": e2e_dec_fps, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes,
                    "d2h_bytes_per_step": B * 5 * N_CTUS * mipb200.CUS_PER_CTU,
                    "result": "MIP decisions: best_mode u8 + best_cost i32 for each of the 5380 CUs of every CTU (pinned host memory)"},
            "e2e_costs": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes, "d2h_bytes_per_step": B * d2h_frame,
                          "result": "decisions + the full int32 cost table (the reference's minSadHad readback, 52.8 MB per frame)"},
            "gpu_launches": launches,
            "clocks": clocks,
            # BASELINE.json's metric asks for the fraction of the SLOWER of the INT32-issue and HBM rooflines.  This path is
            # INT32-issue bound (50-90x further from the HBM roof), so `roofline` is the INT32 view and the HBM view rides
            # inside it.  frac: the kernel alone, launches back to back on one stream (each launch pays its own ramp-up
            # and tail); frac_timed_region: the same launches inside the timed region, where frames overlap on the slot
            # streams (timed-region time / launches = what one launch costs the GPU in steady state).
            "roofline": {"bound": "int32", "kernel": "mip_cost_kernel",
                         "achieved": OPS_PER_FRAME / (kernel_ms * 1e-3) / 1e12, "peak": int32_peak, "unit": "Tops/s",
                         "frac": OPS_PER_FRAME / (kernel_ms * 1e-3) / 1e12 / int32_peak,
                         "frac_timed_region": OPS_PER_FRAME * value / world / 1e12 / int32_peak,
                         "traffic": traffic, "ops_per_launch": OPS_PER_FRAME, "peak_source": int32_src,
                         "kernel_ms_per_frame": kernel_ms, "ms_per_launch_timed_region": dev_ms / max(1, launches),
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                                 "algorithmic_bytes_per_launch": ALGO_BYTES_PER_FRAME, "peak_source": peak_src}},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
