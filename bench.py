#!/usr/bin/env python
"""Benchmark of the fast WordPiece encode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload en|ru|ja|zh|adv]

One step = one pass of the encode path over one batch of synthetic text.  At
N = 1 the batch is BASELINE.json configs[1]: 1 GiB of English-like text with the
29k-entry bert-shaped vocabulary, resident in HBM before the timed region.  At
N > 1 (launched by torchrun, one rank per GPU) every rank encodes its own 1 GiB
block-aligned shard of one corpus (the configs[3] shape: contiguous byte ranges
cut at whitespace, replicated vocabulary, no data-path collective) => weak scaling.

Prints ONE JSON line (rank 0).  `value` is whole-job GB/s of text with device-
resident input; `e2e` is the same metric through the host-buffer C-ABI call
(wp_encode_into: pinned host text -> H2D -> kernel -> D2H ids); `roofline` is
(text bytes + 4 x ids) / kernel time against the measured HBM copy peak;
`cpu_baseline` is the reference's own CPU fast path on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

GIB = 1 << 30
MIB = 1 << 20
METRIC = "GB/s text tokenized"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="en", choices=["en", "ru", "ja", "zh", "adv"])
    ap.add_argument("--mib", type=int, default=1024, help="text MiB per GPU (default 1024 = BASELINE configs[1])")
    ap.add_argument("--cpu-sample-mib", type=int, default=256, help="sample size for the CPU reference timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(args) -> str:
    size = f"{args.mib // 1024}GiB" if args.mib % 1024 == 0 else f"{args.mib}MiB"
    vocab = {"en": "bert-cased-29k", "adv": "long-m100"}.get(args.workload, "mbert-120k")
    return f"{args.workload}-{size}-per-gpu/{vocab}"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_from_profile():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def cpu_reference_timing(text: np.ndarray, vocab, sample_mib: int, repeats: int = 1):
    """Time the reference's own CPU fast path (oracle/_ref, else the C port) on a bounded sample."""
    from _oracle import Oracle, Ref

    n = min(text.size, sample_mib * MIB)
    sample = text[:n]
    # cut at the last whitespace so the sample is a valid shard
    tail = sample[-4096:].tobytes()
    cut = max(tail.rfind(b" "), tail.rfind(b"\n"))
    if cut >= 0:
        sample = sample[: n - len(tail) + cut + 1]
    b = sample.tobytes()
    if Ref.available():
        cores = os.cpu_count() or 1
        Ref.lib(cores)  # the pool size is frozen by the first call (utils.cpp:25-28)
        best, n_ids = None, 0
        for _ in range(repeats):
            ids, sec = Ref.encode(b, vocab, "fast", return_seconds=True)
            n_ids = int(ids.size)
            best = sec if best is None else min(best, sec)
        return {"kind": "reference", "cores": int(Ref.pool_threads), "seconds": best, "bytes": len(b), "ids": n_ids,
                "sample": f"first {len(b) / MIB:.0f} MiB of the workload, reference fast::encode(text, vocab) in-process "
                          f"(oracle/_ref, -O3 -flto -march=x86-64-v3), {Ref.pool_threads} pool threads, "
                          f"includes its per-call vocab parse + hash-map build"}
    o = Oracle(vocab)
    t0 = time.perf_counter()
    ids = o.encode(b)
    sec = time.perf_counter() - t0
    return {"kind": "port", "cores": 1, "seconds": sec, "bytes": len(b), "ids": int(ids.size),
            "sample": f"first {len(b) / MIB:.0f} MiB of the workload, oracle/wp_oracle.c (scalar port, 1 thread)"}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU fast path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from wordpiece_b200 import synth

    g = synth.generator(args.workload)
    sample_mib = min(args.cpu_sample_mib, args.mib)
    text = g.generate(sample_mib * MIB, seed=2)
    vocab = g.spec.vocab
    times, info = [], None
    for i in range(args.warmup + args.steps):
        info = cpu_reference_timing(text, vocab, sample_mib)
        if i >= args.warmup:
            times.append(info["seconds"])
    total = sum(times)
    gbs = info["bytes"] * len(times) / total / 1e9
    line = {
        "impl": "reference",
        "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "step": f"one reference fast::encode call over a {sample_mib} MiB sample",
                   "timing": "steady_clock inside the reference process (host only; no GPU involved)"},
        "tokens_per_s": info["ids"] * len(times) / total,
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 from here on (NCCL prints its
    version banner there from C code, libraries may log) goes to stderr; emit() writes to the saved fd."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import wordpiece_b200
    from wordpiece_b200 import synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path to measure)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if "WP_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["WP_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- workload: this rank's shard of the corpus (block-aligned => cut at whitespace)
    n_bytes = args.mib * MIB
    g = synth.generator(args.workload)
    vocab_tokens = g.spec.vocab
    blocks_per_rank = args.mib
    seed = 2 if world == 1 else 4
    threads = max(1, (os.cpu_count() or 8) // max(1, min(world, 8)))
    h_text = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    g.fill(h_text.numpy(), seed=seed, first_block=rank * blocks_per_rank, n_threads=threads)
    d_text = h_text.to(dev, non_blocking=False)

    vocab = wordpiece_b200.Vocab(vocab_tokens, device=local_rank)
    cap = n_bytes // 2 + 4096
    d_ids = torch.empty(cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)

    # one checked pass (count; also the first warm-up)
    _, n_ids = vocab.encode_device(d_text, d_ids)
    stats = vocab.stats()
    assert n_ids <= cap, "id buffer too small for this workload"

    for _ in range(max(args.warmup, 3) - 1):
        vocab.encode_device_async(d_text, d_ids, d_cnt)
    barrier()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = wordpiece_b200.kernel_launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        vocab.encode_device_async(d_text, d_ids, d_cnt)
    e1.record()
    barrier()
    sampler.stop()
    launches = int(sum_over_ranks(float(wordpiece_b200.kernel_launch_count() - launches0)))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    assert int(d_cnt.item()) == n_ids

    total_bytes = sum_over_ranks(float(n_bytes))
    total_ids = sum_over_ranks(float(n_ids))
    value = total_bytes / (ms_step * 1e-3) / 1e9
    tokens_per_s = total_ids / (ms_step * 1e-3)

    # ---- roofline.  Algorithmic bytes = text + 4 x ids, this rank.  The path is three kernels
    # per 64 MiB range; their device times come from CUDA events the library records on the launching
    # stream around each launch (wp_set_kernel_timing), in a separate short pass so that the timed loop
    # above carries no extra events.  The roofline object is for the DOMINANT kernel, as the contract asks;
    # `whole_path_frac` is the same bytes over the whole step.
    peak, peak_src = measured_peak()
    algo_bytes = float(n_bytes) + 4.0 * float(n_ids)
    my_ms = e0.elapsed_time(e1) / args.steps
    names = ["wp_split_kernel", "wp_match_kernel", "wp_scatter_kernel"]
    vocab.set_kernel_timing(True)
    k_ms = [0.0, 0.0, 0.0]
    prof_steps = 3
    n_ranges = 1
    for _ in range(prof_steps):
        vocab.encode_device_async(d_text, d_ids, d_cnt)
        ms, n_ranges = vocab.last_kernel_ms()
        k_ms = [a + b for a, b in zip(k_ms, ms)]
    vocab.set_kernel_timing(False)
    k_ms = [x / prof_steps for x in k_ms]
    dom = int(np.argmax(k_ms))
    achieved = algo_bytes / (k_ms[dom] * 1e-3) / 1e9
    prof = traffic_from_profile()
    traffic = None
    if prof and names[dom] in prof.get("per_kernel", {}):
        traffic = prof["per_kernel"][names[dom]].get("dram_bytes")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": names[dom],
                "algo_bytes_per_launch": algo_bytes / max(n_ranges, 1), "launches_per_step": n_ranges,
                "kernel_ms_per_launch": k_ms[dom] / max(n_ranges, 1),
                "kernel_ms_per_step": dict(zip(names, k_ms)), "step_ms": my_ms,
                "whole_path_achieved": algo_bytes / (my_ms * 1e-3) / 1e9,
                "whole_path_frac": algo_bytes / (my_ms * 1e-3) / 1e9 / peak,
                "peak_source": peak_src,
                "note": "achieved = (text bytes + 4 x ids) of one range / device time of the dominant kernel for that "
                        "range (CUDA events on the launching stream); a step = launches_per_step x (scratch memset + "
                        "K1 + K2 + K3)"}
    if prof:
        # the capture is one launch over a FULL 64 MiB range (a step also has a 2 MiB, an 8 MiB and a tail range)
        roofline["traffic_launch_algo_bytes"] = algo_bytes * min(1.0, 64.0 * MIB / n_bytes)
        roofline["traffic_source"] = prof.get("source")

    # ---- e2e: the host-buffer C-ABI call (pinned host text in, host ids out), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        h_ids = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        h_ids_np = h_ids.numpy()
        h_text_np = h_text.numpy()
        e2e_steps = max(2, min(args.steps, 5))
        vocab.encode_into(h_text_np, h_ids_np)  # warm-up: sizes the library's device staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            got = vocab.encode_into(h_text_np, h_ids_np)
        torch.cuda.synchronize(dev)
        sec = time.perf_counter() - t0
        barrier()
        assert got == n_ids
        sec = max_over_ranks(sec)
        e2e = {"value": total_bytes * e2e_steps / sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n_bytes),
               "d2h_bytes_per_step": int(4 * n_ids), "steps": e2e_steps,
               "call": "wp_encode_into (C ABI, include/wordpiece_b200.h) with pinned host text and id buffers",
               "timing": "host wall clock around synchronous calls, max over ranks"}
        # the ids that came back are the ones the device-resident path produced
        assert np.array_equal(h_ids_np[:n_ids], d_ids[:n_ids].cpu().numpy())

    # ---- CPU baseline: the reference's own fast path on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        info = cpu_reference_timing(h_text.numpy(), vocab_tokens, min(args.cpu_sample_mib, args.mib))
        cpu = {"value": info["bytes"] / info["seconds"] / 1e9, "unit": UNIT, "cores": info["cores"],
               "kind": info["kind"], "sample": info["sample"], "tokens_per_s": info["ids"] / info["seconds"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": workload_name(args),
                "text_bytes_per_gpu": n_bytes, "ids_per_gpu": n_ids, "vocab_entries": len(vocab_tokens),
                "sharding": "none (1 GPU)" if world == 1 else f"{world} contiguous block-aligned shards of one corpus, "
                                                              f"cut at whitespace, replicated vocab, no collective",
                "l2": "input (1 GiB) and output (>1 GB) are far larger than the 126 MB L2; no flush needed",
                "tiles_per_gpu": stats.n_tiles, "dirty_tiles": stats.dirty_tiles, "long_segments": stats.long_segments,
            },
            "tokens_per_s": tokens_per_s,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
