#!/usr/bin/env python3
"""bench.py -- decoded info Gbit/s of the B200 min-sum LDPC decoder (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg5|cfg1]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's CPU decode on the host cores

A step = one pass of the decode hot path over one batch of synthetic BPSK-AWGN channel values.
Default workload `cfg2` = BASELINE configs[1]: Test.cpp's code (802.16e rate 3/4B, N=576, K=432),
65,536 codewords per GPU, 40-iteration cap, sigma = 1.0 (snr arg 0 dB) so that no word converges:
a fixed 40 iterations under the reference's own early-termination rule.  Independent codewords are
sharded over GPUs with no collective (weak scaling: every GPU decodes its own 65,536 words).

Prints ONE JSON line (rank 0).  `value`: device-resident decode, CUDA events, max over ranks.
`e2e`: the same work through the host-buffer C-ABI call (pinned host LLRs in, bits out).
`roofline`: the binding on-chip (shared-memory) roofline from the algorithmic message bytes,
with the HBM figures beside it.  `cpu_baseline`: the oracle restatement of Coder::decodeCPU timed on
the host cores on a bounded sample of the very same floats (and used to spot-check parity).
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "decoded_info_gbit_per_s"
UNIT = "Gbit/s"

WORKLOADS = {
    # name: (description, sigma, cap, default codewords per GPU)
    "cfg1": ("wimax_3_4b_n576_4096cw", 1.0, 40, 4096),
    "cfg2": ("wimax_3_4b_n576_65536cw", 1.0, 40, 65536),
    "cfg3": ("regular_3_6_n8192", 1.0, 40, 131072),
    "cfg5": ("ira_n64800_cap50", 1.0, 50, 4736),
}


def make_code(workload: str):
    import myldpccppapi_b200 as m

    if workload in ("cfg1", "cfg2"):
        N, K = 576, 432
        rp, ci, M = m.wimax_csr(K, N, m.rate_3_4_b)
        return M, N, K, rp, ci
    if workload == "cfg3":
        return m.codes.regular_code()
    if workload == "cfg5":
        return m.codes.ira_code()
    raise SystemExit("unknown workload " + workload)


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed regions run."""

    BITS = {0x1: "gpu_idle", 0x2: "app_clocks", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
            0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake", 0x100: "display_clock"}

    def __init__(self, uuid: str):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.BITS.items():
                        if r & bit and name != "gpu_idle":
                            self.reasons.add(name)
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                except Exception:
                    pass
            time.sleep(0.02)

    def start(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def region(self, on: bool):
        (self._active.set if on else self._active.clear)()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_decode_timed(code, cap, llr_np, target_s: float = 12.0, literal: bool = True):
    """Time the oracle (reference-semantics CPU decode) on as many of the given words as fit in
    about `target_s` seconds with all host threads.  Returns (words, seconds, outputs)."""
    import oracle

    M, N, K, rp, ci = code
    o = oracle.Oracle(M, N, K, rp, ci, times=cap)
    thr = host_threads()
    probe = min(len(llr_np), max(thr * 4, 32))
    t0 = time.perf_counter()
    o.decode(llr_np[:probe], threads=thr, literal=literal, want_post=False, want_hard=False)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(len(llr_np), max(probe, target_s / dt * probe)))
    n = max(thr, n // thr * thr)
    t0 = time.perf_counter()
    out = o.decode(llr_np[:n], threads=thr, literal=literal, want_post=False, want_hard=False)
    return n, time.perf_counter() - t0, out, thr


def run_reference(args, rank: int, world: int) -> None:
    """--impl reference: the reference's CPU decode of the same workload on the host cores."""
    if rank != 0:
        return
    code = make_code(args.workload)
    M, N, K, rp, ci = code
    desc, sigma, cap, ncw_default = WORKLOADS[args.workload]
    sigma = args.sigma if args.sigma is not None else sigma
    thr = host_threads()
    rng = np.random.default_rng(12345)
    # bounded sample: sized from a probe so that steps+warmup stay within a few minutes
    import oracle
    o = oracle.Oracle(M, N, K, rp, ci, times=cap)
    probe_n = max(thr * 2, 16)
    y = (1.0 + sigma * rng.standard_normal((probe_n, N), dtype=np.float32)).astype(np.float32)
    t0 = time.perf_counter()
    o.decode(y, threads=thr, literal=True, want_post=False, want_hard=False)
    rate = probe_n / max(time.perf_counter() - t0, 1e-6)
    budget_s = 150.0 / max(args.steps + args.warmup, 1)
    sample = int(max(thr, min(ncw_default, rate * min(budget_s, 8.0)) // thr * thr))
    y = (1.0 + sigma * rng.standard_normal((sample, N), dtype=np.float32)).astype(np.float32)
    kind, runner = "port", None
    try:
        from oracle import ref as oref
        if oref.available("O2") and args.workload in ("cfg1", "cfg2"):
            kind = "reference"
            runner = lambda: oref.decode_cpu_parallel(K, N, 4, y, cap, thr, opt="O2")  # noqa: E731
    except Exception:
        runner = None
    if runner is None:
        runner = lambda: o.decode(y, threads=thr, literal=True, want_post=False, want_hard=False)  # noqa: E731
    for _ in range(args.warmup):
        runner()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        runner()
    dt = time.perf_counter() - t0
    value = sample * K * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic BPSK-AWGN (all-zero codeword + noise), seeded",
        "config": {"workload": desc, "code": "N=%d K=%d nnz=%d" % (N, K, int(rp[-1])), "sigma": sigma, "max_iter": cap,
                   "early_termination": True, "codewords_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": kind,
                         "sample": "%d codewords per step on %d host threads (%s)" % (
                             sample, thr, "oracle/_ref: the reference's own Coder::decodeCPU compiled -O2 (its Makefile uses -O0), one Coder per thread" if kind == "reference"
                             else "oracle port of Coder::decodeCPU, literal O(dc^2) check loop, gcc -O2")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--ncw", type=int, default=None, help="codewords per GPU (default: the workload's)")
    ap.add_argument("--sigma", type=float, default=None)
    ap.add_argument("--max-iter", type=int, default=None, help="override the workload's iteration cap")
    ap.add_argument("--path", type=int, default=-1, help="force a kernel path (see ldpc_b200.h)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import myldpccppapi_b200 as m

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decoder has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

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

    desc, sigma, cap, ncw = WORKLOADS[args.workload]
    sigma = args.sigma if args.sigma is not None else sigma
    ncw = args.ncw or ncw
    cap = args.max_iter or cap
    code = make_code(args.workload)
    M, N, K, rp, ci = code
    nnz = int(rp[-1])
    dec = m.Decoder(M, N, K, rp, ci, device=local_rank, max_iter=cap, early_termination=True)
    if args.path >= 0:
        dec.set_path(args.path)
    info = dec.info()

    # every rank's shard: its own seeded slice of the synthetic stream (weak scaling)
    llr = m.synth_llr(ncw, N, sigma, seed=0x4C445043 + rank, device=local_rank)
    out = {}
    for _ in range(max(args.warmup, 1)):
        dec.decode_device(llr, out=out)
    torch.cuda.synchronize()

    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", ""))
    sampler.start()

    # ---- device-resident timing: CUDA events on the launching (current) stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = dec.launches
    barrier()
    torch.cuda.synchronize()
    sampler.region(True)
    ev0.record()
    for _ in range(args.steps):
        dec.decode_device(llr, out=out)
    ev1.record()
    torch.cuda.synchronize()
    sampler.region(False)
    barrier()
    launches = dec.launches - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    total_cw = sum_over_ranks(float(ncw))
    value = total_cw * K / (ms_step * 1e-3) / 1e9
    iters_dev = out["iters"]
    mean_iters = float(iters_dev.float().mean().item())

    # ---- end to end through the host-buffer C-ABI call (pinned host memory in, bits out)
    e2e = None
    numa = None
    if not args.no_e2e:
        # pinned staging buffers on the NUMA node of this rank's GPU (placement only; affinity restored below)
        saved_affinity = os.sched_getaffinity(0)
        numa = m.shard.bind_to_gpu_numa_node(local_rank)
        h_llr = torch.empty((ncw, N), dtype=torch.float32, pin_memory=True)
        h_llr.copy_(llr)
        h_out = {"info": torch.empty((ncw, dec.KB), dtype=torch.uint8, pin_memory=True),
                 "iters": torch.empty((ncw,), dtype=torch.int32, pin_memory=True)}
        torch.cuda.synchronize()
        os.sched_setaffinity(0, saved_affinity)
        for _ in range(max(1, min(args.warmup, 3))):
            dec.decode_host(h_llr, out=h_out)
        barrier()
        torch.cuda.synchronize()
        sampler.region(True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dec.decode_host(h_llr, out=h_out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sampler.region(False)
        barrier()
        dt = max_over_ranks(dt)
        e2e = {"value": total_cw * K * args.steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(total_cw) * N * 4, "d2h_bytes_per_step": int(total_cw) * (dec.KB + 4),
               "ms_per_step": dt / args.steps * 1e3, "api": "ldpc_b200_decode_host (pinned host buffers; " + ("one persistent launch fed by a copy stream)" if info["path_name"] == "qc" else "3-stream pipeline)")}
        same = bool(torch.equal(h_out["info"], out["info"].cpu()) and torch.equal(h_out["iters"], out["iters"].cpu()))
        e2e["matches_device_path"] = same
        e2e["host_numa"] = numa
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (one decode launch per step)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    import ctypes as C
    smem_gbs = C.c_double(0.0)
    m.lib.check(m.load().ldpc_b200_probe_smem_bandwidth(local_rank, C.byref(smem_gbs)))
    per_gpu_cw_s = ncw / (ms_step * 1e-3)
    b_hbm = 4 * N + (K + 7) // 8 + 1                       # channel values in, info bits + count out
    b_msg = mean_iters * (16 * nnz + 8 * N)                # on-chip message bytes per word
    onchip = info["path_name"] in ("lane_smem", "lane16", "group", "cluster", "qc")
    roof_smem = {"bound": "smem", "achieved": per_gpu_cw_s * b_msg / 1e9, "peak": smem_gbs.value, "unit": "GB/s",
                 "peak_source": "measured live: ldpc_b200_probe_smem_bandwidth (LDS.128 stream on all SMs)"}
    roof_smem["frac"] = roof_smem["achieved"] / roof_smem["peak"] if roof_smem["peak"] else None
    roof_hbm = {"bound": "hbm", "achieved": per_gpu_cw_s * (b_hbm if onchip else b_msg) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "peak_source": hbm_src, "traffic": None}
    roof_hbm["frac"] = roof_hbm["achieved"] / hbm_peak
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch)
    try:
        tr = json.loads((ROOT / "profiles" / ("ncu_traffic_%s.json" % args.workload)).read_text())
        if tr.get("codewords_per_launch") == ncw:
            roof_hbm["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roof_hbm["traffic_source"] = tr["source"]
            roof_hbm["algorithmic_bytes_per_launch"] = (b_hbm if onchip else b_msg) * ncw
    except Exception:
        pass
    roofline = dict(roof_smem if onchip else roof_hbm)
    roofline.update({
        "kernel": {"qc": "ldpc_ms_qc_kernel", "group": "ldpc_ms_group_kernel", "cluster": "ldpc_ms_cluster_kernel", "lane16": "ldpc_ms_lane16_kernel",
                   "lane_smem": "ldpc_ms_lane_kernel<true>", "stream": "ldpc_ms_stream_kernel"}.get(info["path_name"], "ldpc_ms_lane_kernel<false>"),
        "launch_ms": ms_step, "traffic": roof_hbm.get("traffic"),
        "algorithmic_bytes_per_codeword": {"hbm": b_hbm, "messages": b_msg, "mean_iterations": mean_iters},
        "hbm": roof_hbm, "smem": roof_smem,
        "note": ("messages stay in shared memory: HBM carries only channel values and bits, so shared-memory "
                 "bandwidth is the binding roofline" if onchip else
                 "messages live in a global workspace: HBM/L2 bandwidth is the binding roofline"),
    })

    # ---- CPU baseline: the oracle on a bounded sample of the same floats (+ parity spot check)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sample_max = min(ncw, 65536)  # cpu_decode_timed sizes the sample for ~12 s of host work
        y = llr[:sample_max].cpu().numpy()
        n, secs, ref, thr = cpu_decode_timed(code, cap, y)
        ok = bool(np.array_equal(ref[0], out["info"][:n].cpu().numpy()) and np.array_equal(ref[1], out["iters"][:n].cpu().numpy()))
        cpu = {"value": n * K / secs / 1e9, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": "first %d codewords of the GPU batch, oracle (literal Coder::decodeCPU restatement, gcc -O2) on %d threads, %.1f s" % (n, thr, secs),
               "codewords_per_s": n / secs, "gpu_matches_oracle_on_sample": ok}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic BPSK-AWGN channel values (all-zero codeword + seeded noise, generated on device)",
        "config": {"workload": desc, "code": "N=%d K=%d M=%d nnz=%d" % (N, K, M, nnz), "codewords_per_gpu": ncw,
                   "sigma": sigma, "max_iter": cap, "early_termination": True, "mean_iterations": mean_iters,
                   "parallelism": "codeword sharding x%d, no collective" % world,
                   "l2_policy": "inputs larger than L2 (%.0f MB of channel values per step)" % (ncw * N * 4 / 1e6),
                   "kernel_path": info["path_name"], "threads_per_cta": info["threads_per_cta"], "ctas": info["ctas"],
                   "smem_bytes_per_cta": info["smem_bytes"]},
        "codewords_per_s": total_cw / (ms_step * 1e-3),
        "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
