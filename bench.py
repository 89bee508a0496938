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
         full int32 cost table (the reference's minSadHad readback, 52.8 MB per 1080p frame) and is
         repeated as e2e_like_for_like: the result the reference arm's e2e returns
         e2e_costs_compact: the same table in the lossless compact transport (71 % of the bytes)
  sizes  the same three numbers for BASELINE configs 4 (3840x2160, original samples) and 5 (7680x4320,
         alternative samples) at the current GPU count
  shard_check   N > 1: every rank runs its poc % N share of a fixed 16-frame set, the decision hashes are
                gathered and rank 0 compares them with its own unsharded run
  roofline      the fused cost kernel, in the bench configuration, against the binding roof: INT32 issue
                (algorithmic INT32 ops of BASELINE.md section 2 / average launch duration over the timed region /
                measured INT32 peak); roofline.lone_frame: the same for one frame at a time; the HBM view
                (algorithmic bytes / the same duration / measured copy bandwidth) is nested as roofline.hbm
  cpu_baseline  the CPU oracle (port of the reference algorithm, OpenMP, all host cores) on a
                bounded sample of the same workload (rank 0, N == 1 only)

--impl reference: the reference is OpenCL-only.  The arm runs its UNMODIFIED kernels on one B200 through
NVIDIA's OpenCL driver (oracle/_ref/mipref_ocl, built from /root/reference where it lies), one frame per
step, --steps timed steps after --warmup untimed ones, and reports the CPU port beside it (cpu_baseline);
without an OpenCL runtime it falls back to the CPU port alone.
"""
from __future__ import annotations

import argparse
import hashlib
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
COSTS_PER_CTU, CUS_PER_CTU = 97840, 5380
# the one workload string both arms print (BASELINE.json configs[1])
WORKLOAD = f"1920x1080 10-bit natural-like synthetic frames, alternative samples {FILTER_NAME} KernelIdx={KERNEL_IDX}"
# per frame: CTUs, in-frame (CU, mode) costs, algorithmic INT32 ops (SURVEY.md 8(d) / BASELINE.md section 2)
GEOM = {
    (1920, 1080): dict(n_ctus=135, costs_in=12_359_520, ops=1.3148e10),
    (3840, 2160): dict(n_ctus=510, costs_in=49_494_960, ops=5.2766e10),
    (7680, 4320): dict(n_ctus=2040, costs_in=198_125_280, ops=2.1155e11),
}
NS = 3                                      # frames in flight, like the engine's host path (3 slots)


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


def _traffic():
    """Steady-state DRAM bytes per launch of the cost kernel from the committed ncu range capture (None if absent)."""
    for name in ("r02_cost_kernel_traffic.json", "cost_kernel_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p)).get("dram_bytes_per_launch"), "profiles/" + name
    return None, None


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


def _frame_pool(n: int, seed0: int, w: int = W, h: int = H):
    """n distinct natural-like frames.  2160p / 4320p frames are mosaics of 1080p ones (2x2 / 4x4 tiles, every tile a
    different frame): the arithmetic of the path is data independent, so this only saves generation time."""
    import numpy as np
    from mipb200 import frames
    if (w, h) == (W, H):
        return [frames.natural_frame(W, H, seed0 + i) for i in range(n)]
    t = w // W
    base = [frames.natural_frame(W, H, seed0 + i) for i in range(t * t + n - 1)]
    return [np.ascontiguousarray(np.block([[base[i + r * t + c] for c in range(t)] for r in range(t)])) for i in range(n)]


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


def _reference_opencl(frame, reps: int, warmup: int):
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
            r = subprocess.run([exe, fp, str(W), str(H), FILTER_NAME, str(KERNEL_IDX), os.path.join(d, "out"), str(reps), str(warmup)],
                               capture_output=True, text=True, timeout=900)
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
    to the CPU port.  The CPU port is reported beside it either way (cpu_baseline).  A step is ONE frame
    (23 ms of kernels): --steps timed steps after --warmup untimed ones, exactly as asked."""
    if rank != 0:
        return
    pool = _frame_pool(2, 0)
    n_ctus = GEOM[(W, H)]["n_ctus"]
    cfg = {"workload": WORKLOAD, "sample": "one frame per step (the reference host processes frames one at a time)"}
    cpu = _cpu_port_fps(pool, 10.0, 4)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    info, why = _reference_opencl(pool[0], steps, warmup)
    if info is not None:
        frame_bytes = 2 * W * H
        line = {
            "impl": "reference", "metric": "1080p frames/s", "value": info["fps_kernels"], "unit": "frames/s", "n_gpus": 1,
            "steps": info["reps"], "warmup": info.get("warmup", warmup), "ms_per_step": 1e3 / info["fps_kernels"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "reference_kind": "the reference's unmodified OpenCL kernels (intra.cl) on one %s via %s, host = oracle/ocl_ref (full grid for every frame)" % (info["device"], info["opencl_lib"]),
            "kernel_ms": {k: v for k, v in info.items() if k.startswith("ms_")},
            # e2e: the stock host's behaviour -- blocking write, kernels, blocking read of the `long` table into pageable memory
            "e2e": {"value": info["fps_e2e"], "unit": "frames/s", "h2d_bytes_per_step": 2 * frame_bytes, "d2h_bytes_per_step": 8 * n_ctus * COSTS_PER_CTU,
                    "result": "the full minSadHad table as 64-bit integers (main_aux_functions.h:585-630)"},
            # the host as it is meant to run: next upload and previous read-back overlapping the kernels (main.cpp:886-898)
            "e2e_overlapped": {"value": info.get("fps_overlapped"), "unit": "frames/s"},
            "cpu_baseline": cpu, "gpu_launches": 0,
        }
    else:
        line = {
            "impl": "reference", "metric": "1080p frames/s", "value": cpu["value"], "unit": "frames/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / cpu["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
            "reference_kind": "CPU port of the reference algorithm (OpenCL unavailable: %s)" % why,
            "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cpu_baseline": cpu, "gpu_launches": 0,
        }
    print(json.dumps(line), flush=True)


class Workload:
    """One (size, filter) configuration on this rank's GPU: device-resident and host-path timing."""

    def __init__(self, mip, torch, np, dev_index, w, h, ft, kidx, batch, seed0, barrier):
        self.mip, self.torch, self.np, self.dev_index = mip, torch, np, dev_index
        self.w, self.h, self.ft, self.kidx, self.B, self.barrier = w, h, ft, kidx, batch, barrier
        self.n_ctus = GEOM[(w, h)]["n_ctus"]
        self.pool_np = _frame_pool(batch, seed0, w, h)
        self.dev = torch.device("cuda", dev_index)
        self.d_pool = torch.from_numpy(np.stack(self.pool_np).view(np.int16)).to(self.dev)          # B x H x W, resident in HBM
        self.d_cost = torch.empty((NS, self.n_ctus, COSTS_PER_CTU), dtype=torch.int32, device=self.dev)
        self.d_bm = torch.empty((NS, self.n_ctus, CUS_PER_CTU), dtype=torch.uint8, device=self.dev)
        self.d_bc = torch.empty((NS, self.n_ctus, CUS_PER_CTU), dtype=torch.int32, device=self.dev)
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(NS)]   # real (non-default) streams

    def engine(self, slots, emit):
        return self.mip.Engine(self.w, self.h, device=self.dev_index, filter_type=self.ft, kernel_idx=self.kidx, slots=slots, emit=emit)

    def device_resident(self, steps, warmup, sampler=None):
        """frames already in HBM; filter -> fused MIP cost kernel (costs + decisions) per frame, frames round-robin over NS
        streams (independent frames overlap at kernel tails exactly as in the host path).  -> (ms, launches)"""
        torch, mip = self.torch, self.mip
        engs = [self.engine(1, mip.EMIT_DECISIONS) for _ in range(NS)]

        def step():
            for i in range(self.B):
                k = i % NS
                engs[k].run_device(self.d_pool[i].data_ptr(), self.d_cost[k].data_ptr(), d_best_mode=self.d_bm[k].data_ptr(),
                                   d_best_cost=self.d_bc[k].data_ptr(), stream=self.streams[k].cuda_stream)

        for _ in range(warmup):
            step()
        self.barrier()
        if sampler:
            sampler.start()
        l0 = sum(e.kernel_launches() for e in engs)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.streams[0])
        for s_ in self.streams[1:]:
            s_.wait_stream(self.streams[0])
        for _ in range(steps):
            step()
        for s_ in self.streams[1:]:
            self.streams[0].wait_stream(s_)
        ev1.record(self.streams[0])
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = sum(e.kernel_launches() for e in engs) - l0
        for e in engs:
            e.close()
        return ms, launches

    def kernel_alone(self, reps):
        """the fused kernel of THIS configuration (filter, costs + decisions), one frame at a time: launches back to back on
        one stream, the engine told that nothing overlaps (MIPB200_LAUNCH_LATENCY) -> ms per launch"""
        torch, mip = self.torch, self.mip
        eng = self.engine(1, mip.EMIT_DECISIONS)
        eng.set_launch_mode(mip.LAUNCH_LATENCY)      # one frame at a time on the GPU: the split the engine uses for a lone frame
        sp = self.streams[0].cuda_stream

        def go(i):
            eng.run_device(self.d_pool[i % self.B].data_ptr(), self.d_cost[i % NS].data_ptr(), d_best_mode=self.d_bm[i % NS].data_ptr(),
                           d_best_cost=self.d_bc[i % NS].data_ptr(), stream=sp)

        for i in range(3):
            go(i)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(self.streams[0])
        for i in range(reps):
            go(i)
        k1.record(self.streams[0])
        torch.cuda.synchronize()
        eng.close()
        return k0.elapsed_time(k1) / reps

    def host_path(self, emit, touch_costs, steps):
        """the application's frames live in page-locked host memory (as the contract asks): the engine DMAs them in place,
        runs the fused kernel and copies the results back; a frame counts when collect() has returned it.  -> seconds"""
        torch, np = self.torch, self.np
        if not hasattr(self, "pool_host"):
            pin = torch.empty((self.B, self.h, self.w), dtype=torch.int16, pin_memory=True)
            pin.numpy()[...] = np.stack(self.pool_np).view(np.int16)
            self._pin = pin
            self.pool_host = [pin[i].numpy().view(np.uint16) for i in range(self.B)]
        e = self.engine(NS, emit)

        def run(n):
            # one pipelined pass over n frames (n / B steps): NS frames in flight, no drain between the steps of a run, like
            # the device-resident measurement, whose steps are not separated by a synchronisation either
            sub = got = 0
            checksum = 0
            while got < n:
                while sub < n and e.in_flight() < NS:
                    e.submit(self.pool_host[sub % self.B], poc=sub)    # async H2D from pinned memory + fused kernel + async D2H
                    sub += 1
                r = e.collect()                               # waits for this frame's results to be resident on the host
                checksum += int(r.best_cost[0, 0]) + int(r.best_mode[-1, -1])
                if touch_costs:
                    checksum += int(r.cost[-1, -1])
                got += 1
            return checksum

        run(2 * self.B)
        self.barrier()
        t0 = time.perf_counter()
        run(steps * self.B)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e.close()
        return dt

    def free(self):
        del self.d_pool, self.d_cost, self.d_bm, self.d_bc
        if hasattr(self, "_pin"):
            del self.pool_host, self._pin
        self.torch.cuda.empty_cache()


def _shard_check(mip, np, dist, rank, world, dev_index):
    """Results must not depend on the number of GPUs: a fixed 16-frame set, every rank runs its poc % world share through
    the host path and hashes the decisions; rank 0 also runs all 16 alone and compares.  With one GPU two engines stand in
    for two ranks."""
    from mipb200 import frames, shard
    n = 16
    fs = [frames.natural_frame(W, H, 9000 + i) if i % 4 else frames.noise_frame(W, H, 9000 + i) for i in range(n)]

    def run(pocs):
        out = {}
        with mip.Engine(W, H, device=dev_index, filter_type=FILTER_TYPE, kernel_idx=KERNEL_IDX, slots=NS, emit=mip.EMIT_DECISIONS) as eng:
            shard.run_pipelined(eng, fs, pocs, lambda poc, r: out.__setitem__(poc, hashlib.sha256(r.best_mode.tobytes() + r.best_cost.tobytes()).hexdigest()[:16]))
        return out

    if world == 1:
        parts, how = [run(shard.frames_for_rank(n, g, 2)) for g in range(2)], "2 engines on one GPU, frames poc % 2"
    else:
        mine = run(shard.frames_for_rank(n, rank, world))
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        how = f"{world} ranks, frames poc % {world}, hashes all-gathered"
    if rank != 0:
        return None
    whole = run(list(range(n)))
    try:
        merged = shard.merge_in_poc_order(parts, n)
    except ValueError as ex:
        return {"status": f"FAILED: {ex}", "how": how}
    bad = [i for i in range(n) if merged[i] != whole[i]]
    return {"status": "ok" if not bad else f"FAILED: frames {bad} differ", "frames": n, "how": how}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="1080p frames per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sizes", action="store_true", help="skip the 2160p / 4320p configurations")
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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=torch.device("cuda", local_rank))
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    emit_dec = mipb200.EMIT_DECISIONS
    emit_full = mipb200.EMIT_COSTS | mipb200.EMIT_DECISIONS
    B = args.batch
    e2e_steps = max(2, min(args.steps, 8))

    # ---------------- the headline configuration: 1080p, alternative samples ----------------
    wl = Workload(mipb200, torch, np, local_rank, W, H, FILTER_TYPE, KERNEL_IDX, B, 1000 * rank, barrier)   # per-GPU work fixed as N grows: weak scaling
    sampler = ClockSampler(local_rank)
    dev_ms, launches = wl.device_resident(args.steps, args.warmup, sampler)
    clocks = sampler.stop()
    kernel_ms = wl.kernel_alone(max(16, min(96, args.steps * B)))
    e2e_dec_s = wl.host_path(emit_dec, False, e2e_steps)
    e2e_s = wl.host_path(emit_full, True, e2e_steps)
    e2e_cmp_s = wl.host_path(mipb200.EMIT_COSTS_COMPACT | mipb200.EMIT_DECISIONS, False, e2e_steps)
    pool_np = wl.pool_np
    wl.free()

    # what the box's host link gives a plain read-back of the same bytes (one frame's table + decisions, device -> page-locked
    # host memory, back to back on one stream, all ranks at once): the ceiling of `e2e_costs` on THIS box, whatever the kernel
    d2h_frame_bytes = 4 * GEOM[(W, H)]["n_ctus"] * COSTS_PER_CTU + 5 * GEOM[(W, H)]["n_ctus"] * CUS_PER_CTU
    link_src = torch.empty(d2h_frame_bytes, dtype=torch.uint8, device="cuda")
    link_dst = [torch.empty(d2h_frame_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    link_n = 24
    for i in range(4):
        link_dst[i & 1].copy_(link_src, non_blocking=True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(link_n):
        link_dst[i & 1].copy_(link_src, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    link_ms = ev0.elapsed_time(ev1)
    del link_src, link_dst
    dev_ms, e2e_ms, e2e_dec_ms, kernel_ms, e2e_cmp_ms, link_ms = max_over_ranks([dev_ms, e2e_s * 1e3, e2e_dec_s * 1e3, kernel_ms, e2e_cmp_s * 1e3, link_ms])
    link_fps = link_n * world / (link_ms * 1e-3)
    value = B * args.steps * world / (dev_ms * 1e-3)
    e2e_fps = B * e2e_steps * world / (e2e_ms * 1e-3)
    e2e_dec_fps = B * e2e_steps * world / (e2e_dec_ms * 1e-3)
    e2e_cmp_fps = B * e2e_steps * world / (e2e_cmp_ms * 1e-3)

    # ---------------- BASELINE configs 4 and 5 at this GPU count ----------------
    int32_peak, int32_src = _int32_peak()
    sizes = []
    if not args.no_sizes:
        for name, w, h, ft, kidx, b in (("BASELINE config 4: 3840x2160 10-bit, original samples", 3840, 2160, 0, 0, 8),
                                        (f"BASELINE config 5: 7680x4320, alternative samples {FILTER_NAME} KernelIdx={KERNEL_IDX}", 7680, 4320, FILTER_TYPE, KERNEL_IDX, 4)):
            s_steps = max(2, min(args.steps, 6))
            w2 = Workload(mipb200, torch, np, local_rank, w, h, ft, kidx, b, 2000 + 1000 * rank, barrier)
            ms, _ = w2.device_resident(s_steps, 3)
            host_s = w2.host_path(emit_dec, False, s_steps)
            w2.free()
            ms, host_ms = max_over_ranks([ms, host_s * 1e3])
            g = GEOM[(w, h)]
            fps = b * s_steps * world / (ms * 1e-3)
            sizes.append({"workload": name, "frames_per_step_per_gpu": b, "steps": s_steps, "value": fps, "unit": "frames/s",
                          "ms_per_frame_per_gpu": ms / (b * s_steps),
                          "e2e": {"value": b * s_steps * world / (host_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": b * 2 * w * h,
                                  "d2h_bytes_per_step": b * 5 * g["n_ctus"] * CUS_PER_CTU, "result": "MIP decisions"},
                          "frac_timed_region": g["ops"] * fps / world / 1e12 / int32_peak})

    shard = _shard_check(mipb200, np, dist, rank, world, local_rank)

    if rank == 0:
        hbm_peak, peak_src = _peaks()
        g = GEOM[(W, H)]
        algo_bytes = 2 * W * H + 4 * g["costs_in"]            # 53.6 MB: frame in + int32 costs out
        traffic, traffic_src = _traffic()
        cpu = _cpu_port_fps(pool_np, 12.0, 8) if (world == 1 and not args.no_cpu_baseline) else None
        frame_bytes = 2 * W * H
        d2h_frame = 4 * g["n_ctus"] * COSTS_PER_CTU + 5 * g["n_ctus"] * CUS_PER_CTU
        e2e_costs = {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes, "d2h_bytes_per_step": B * d2h_frame,
                     "result": "decisions + the full int32 cost table (the reference's minSadHad readback, 52.8 MB per frame)",
                     "d2h_link": {"what": "plain cudaMemcpyAsync of one frame's results (device -> page-locked host), back to back, measured in this run on this box: the ceiling of this number",
                                  "gbs_per_gpu": d2h_frame_bytes * link_n / (link_ms * 1e-3) / 1e9, "ceiling": link_fps, "unit": "frames/s", "frac": e2e_fps / link_fps}}
        line = {
            "metric": "1080p frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch": f"{B} distinct frames per GPU per step; filter + MIP costs (97840 per CTU) + decisions",
                       "frames_per_step_per_gpu": B, "sharding": f"frames over {world} GPU(s), no collective",
                       "l2": f"input pool {(B * frame_bytes) >> 20} MiB + 3 rotating 50 MiB cost tables exceed the 126 MiB L2 (no flush needed)"},
            "e2e": {"value": e2e_dec_fps, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes,
                    "d2h_bytes_per_step": B * 5 * g["n_ctus"] * CUS_PER_CTU,
                    "result": "MIP decisions: best_mode u8 + best_cost i32 for each of the 5380 CUs of every CTU (pinned host memory)"},
            "e2e_costs": e2e_costs,
            # full table against full table: what the reference arm's e2e returns (there as 64-bit integers)
            "e2e_like_for_like": dict(e2e_costs, compare_with="the reference arm's e2e / e2e_overlapped (full minSadHad table on the host)"),
            # the same table in the compact transport (uint16 for CUs of <= 32 samples, lossless): what the link allows then
            "e2e_costs_compact": {"value": e2e_cmp_fps, "unit": "frames/s", "h2d_bytes_per_step": B * frame_bytes,
                                  "d2h_bytes_per_step": B * (mipb200.COMPACT_BYTES_PER_CTU * g["n_ctus"] + 5 * g["n_ctus"] * CUS_PER_CTU),
                                  "result": "decisions + the compact cost table (MIPB200_EMIT_COSTS_COMPACT: 276 672 instead of 391 360 bytes per CTU; mipb200_expand_costs() restores int32)"},
            "sizes": sizes,
            "shard_check": shard,
            "gpu_launches": launches,
            "clocks": clocks,
            # BASELINE.json's metric asks for the fraction of the SLOWER of the INT32-issue and HBM rooflines.  This path is
            # INT32-issue bound (50-90x further from the HBM roof), so `roofline` is the INT32 view and the HBM view rides
            # inside it.  As the bench contract asks, `achieved` = algorithmic ops per launch / the kernel's average launch
            # duration over the timed region (CUDA events on the launching streams: timed-region time / launches -- the frames
            # overlap on the slot streams, this is what one launch costs the GPU).  `lone_frame` is the stricter figure: the
            # same kernel, same configuration, one frame at a time (launches back to back on one stream, the engine's
            # lone-frame block split), each launch paying its own ramp-up and tail.
            "roofline": {"bound": "int32", "kernel": "mip_cost_kernel",
                         "achieved": g["ops"] * value / world / 1e12, "peak": int32_peak, "unit": "Tops/s",
                         "frac": g["ops"] * value / world / 1e12 / int32_peak,
                         "frac_timed_region": g["ops"] * value / world / 1e12 / int32_peak,
                         "ms_per_launch_timed_region": dev_ms / max(1, launches),
                         "lone_frame": {"achieved": g["ops"] / (kernel_ms * 1e-3) / 1e12, "frac": g["ops"] / (kernel_ms * 1e-3) / 1e12 / int32_peak,
                                        "kernel_ms_per_frame": kernel_ms},
                         "traffic": traffic, "traffic_source": traffic_src, "ops_per_launch": g["ops"], "peak_source": int32_src,
                         "kernel_config": f"{FILTER_NAME} KernelIdx={KERNEL_IDX}, costs + decisions",
                         "hbm": {"achieved": algo_bytes * value / world / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": algo_bytes * value / world / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src}},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
