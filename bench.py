#!/usr/bin/env python3
"""bench.py -- decoded info Gbit/s of the B200 LDPC decoder (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg5|cfg1|cfg4]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's CPU decode on the host cores

A step = one pass of the decode hot path over one batch of synthetic BPSK-AWGN channel values.
Headline workload `cfg2` = BASELINE configs[1]: Test.cpp's code (802.16e rate 3/4B, N=576, K=432),
65,536 codewords per GPU, 40-iteration cap, sigma = 1.0 (snr arg 0 dB) so that no word converges:
a fixed 40 iterations under the reference's own early-termination rule.  Independent codewords are
sharded over GPUs with no collective (weak scaling: every GPU decodes its own 65,536 words).

Prints ONE JSON line (rank 0):
  value        device-resident decode, CUDA events on the launching stream, max over ranks
  e2e          the same work through the host-buffer C-ABI call (pinned host values in, bits out)
  e2e_plugin   the same through the drop-in C++ Coder::decode (libmyldpc_b200.so) with malloc'd buffers: the call
               a user of the reference makes -- first call and steady state apart
  roofline     the binding roofline of the dominant kernel from the algorithmic bytes
  sustained    the headline loop held for >= 2 s (clocks and power under load)
  workloads    the other BASELINE configs measured in the same process: cfg3 (regular (3,6), N=8192), cfg5 (IRA N=64800,
               cap 50), cfg4 (Test.cpp's code at Eb/N0 = 3.5 dB with early termination) -- value, roofline, clocks,
               oracle spot check each
  parity_all_ranks   every rank decodes the head of its shard with the oracle; AND over ranks
  cpu_baseline the oracle restatement of Coder::decodeCPU on the host cores (rank 0, N=1), bounded sample
  setdevices   (N > 1) one process driving all N GPUs through Coder::setDevices
"""
from __future__ import annotations

import argparse
import json
import math
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "decoded_info_gbit_per_s"
UNIT = "Gbit/s"


def sigma_from_ebn0(ebn0_db: float, rate: float) -> float:
    return float(math.sqrt(1.0 / (2.0 * rate * 10.0 ** (ebn0_db / 10.0))))


WORKLOADS = {
    # name: (description, sigma, cap, default codewords per GPU)
    "cfg1": ("wimax_3_4b_n576_4096cw", 1.0, 40, 4096),
    "cfg2": ("wimax_3_4b_n576_65536cw", 1.0, 40, 65536),
    "cfg3": ("regular_3_6_n8192", 1.0, 40, 131072),
    "cfg4": ("wimax_3_4b_n576_ebn0_3.5dB_early_termination", sigma_from_ebn0(3.5, 0.75), 40, 65536),
    "cfg5": ("ira_n64800_cap50", 1.0, 50, 4736),
    # two more points of the sweep and of the reference's code family, measured as sub-records only
    "cfg4_4dB": ("wimax_3_4b_n576_ebn0_4.0dB_early_termination", sigma_from_ebn0(4.0, 0.75), 40, 65536),
    "family_z60": ("wimax_3_4b_n1440_z60_no_compiled_profile", 1.0, 40, 32768),
}
# (steps, warmup, oracle spot-check words) of the sub-records in `workloads`
PLACE_EFFORT = {"cfg3": 48}   # LDPC_B200_PLACE_EFFORT of the decoder's setup (see measure())
SUB = {"cfg3": (5, 3, 32), "cfg5": (3, 3, 2), "cfg4": (200, 3, 512), "cfg4_4dB": (200, 3, 512), "family_z60": (10, 3, 128)}


def make_code(workload: str):
    """(M, N, K, row_ptr, col_idx).  Host-side only: the 802.16e expansion comes from the oracle package here so that
    the reference arm never loads the product's CUDA library (the product's expansion is checked equal in tests)."""
    if workload in ("cfg1", "cfg2", "cfg4", "cfg4_4dB"):
        import oracle
        N, K = 576, 432
        rp, ci, M = oracle.wimax_H(N, "3/4B")
        return M, N, K, rp, ci
    if workload == "family_z60":   # a block size Coder::initCheckMatrix accepts that has no compiled lockstep profile
        import oracle
        N, K = 1440, 1080
        rp, ci, M = oracle.wimax_H(N, "3/4B")
        return M, N, K, rp, ci
    from myldpccppapi_b200 import codes
    if workload == "cfg3":
        return codes.regular_code()
    if workload == "cfg5":
        return codes.ira_code()
    raise SystemExit("unknown workload " + workload)


def common_config(workload: str, code, sigma: float, cap: int, ncw: int) -> dict:
    """The part of `config` both arms print identically."""
    M, N, K, rp, ci = code
    return {"workload": WORKLOADS[workload][0], "code": "N=%d K=%d M=%d nnz=%d" % (N, K, M, int(rp[-1])),
            "codewords_per_gpu": ncw, "sigma": round(sigma, 6), "max_iter": cap, "early_termination": True}


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed regions run."""

    BITS = {0x1: "gpu_idle", 0x2: "app_clocks", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
            0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake", 0x100: "display_clock"}

    def __init__(self, uuid: str):
        self.max_mhz = None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thr = None
        self.reset()
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def reset(self):
        self.samples, self.reasons, self.power = [], set(), []

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

    def summary(self, reset: bool = False):
        if not self.samples:
            s = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        else:
            s = {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                 "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}
        if reset:
            self.reset()
        return s


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
    """--impl reference: the reference's CPU decode of the same workload on the host cores.  Loads only oracle/
    (liboracle.so, oracle/_ref): none of the product's native code is on this arm."""
    if rank != 0:
        return
    code = make_code(args.workload)
    M, N, K, rp, ci = code
    desc, sigma, cap, ncw_default = WORKLOADS[args.workload]
    sigma = args.sigma if args.sigma is not None else sigma
    ncw_default = args.ncw or ncw_default
    cap = args.max_iter or cap
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
        if oref.available("O2") and args.workload in ("cfg1", "cfg2", "cfg4"):
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
        "config": common_config(args.workload, code, sigma, cap, ncw_default),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": kind,
                         "sample": "%d codewords of the workload per step on %d host threads (%s)" % (
                             sample, thr, "oracle/_ref: the reference's own Coder::decodeCPU compiled -O2 (its Makefile uses -O0), one Coder per thread" if kind == "reference"
                             else "oracle port of Coder::decodeCPU, literal O(dc^2) check loop, gcc -O2")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_record(line)


_RECORD_FD = None


def emit_record(line: dict) -> None:
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _RECORD_FD is None:
        os.write(1, data)
    else:
        os.write(_RECORD_FD, data)


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
    ap.add_argument("--algorithm", type=int, default=0, help="0 min-sum, 1 sum-product, 2 layered (LDPC_B200_ALG_*)")
    ap.add_argument("--option", action="append", default=[], help="name=value for ldpc_b200_set_option (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no sustained / workloads / plugin / setdevices legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)
    # stdout carries exactly ONE line, the JSON record: whatever libraries print there while the bench runs (NCCL's version
    # banner, for one) goes to stderr; the record is written to the saved descriptor at the end
    global _RECORD_FD
    sys.stdout.flush()
    _RECORD_FD = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C

    import torch
    import torch.distributed as dist

    import myldpccppapi_b200 as m
    import oracle

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decoder has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # CPU-side barrier (gloo) for the leg in which rank 0 alone drives every GPU: ranks parked in an NCCL barrier would
    # keep a spinning kernel on the very GPUs rank 0 is timing
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce(x, dist.ReduceOp.MAX)   # noqa: E731
    min_over_ranks = lambda x: reduce(x, dist.ReduceOp.MIN)   # noqa: E731
    sum_over_ranks = lambda x: reduce(x, dist.ReduceOp.SUM)   # noqa: E731

    def gather(x: float):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        outl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outl, t)
        return [float(v.item()) for v in outl]

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    smem_gbs = C.c_double(0.0)
    m.lib.check(m.load().ldpc_b200_probe_smem_bandwidth(local_rank, C.byref(smem_gbs)))

    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", ""))
    sampler.start()

    KERNELS = {"qc": "ldpc_ms_qc_kernel", "group": "ldpc_ms_group_kernel", "cluster": "ldpc_ms_cluster_kernel", "lane16": "ldpc_ms_lane16_kernel",
               "lane_smem": "ldpc_ms_lane_kernel<true>", "stream": "ldpc_ms_stream_kernel", "lane_global": "ldpc_ms_lane_kernel<false>"}

    def roofline_of(info, code, ncw, ms_step, mean_iters, workload):
        """algorithmic bytes per codeword x codewords per launch / launch time, against the measured peak of
        the binding resource (shared memory when the messages are on chip, HBM when they live in the global workspace)."""
        M, N, K, rp, ci = code
        nnz = int(rp[-1])
        per_gpu_cw_s = ncw / (ms_step * 1e-3)
        b_hbm = 4 * N + (K + 7) // 8 + 1
        b_msg = mean_iters * (16 * nnz + 8 * N)
        onchip = info["path_name"] in ("lane_smem", "lane16", "group", "cluster", "qc") or args.algorithm == 2
        roof_smem = {"bound": "smem", "achieved": per_gpu_cw_s * b_msg / 1e9, "peak": smem_gbs.value, "unit": "GB/s",
                     "peak_source": "measured live: ldpc_b200_probe_smem_bandwidth (LDS.128 stream on all SMs)"}
        roof_smem["frac"] = roof_smem["achieved"] / roof_smem["peak"] if roof_smem["peak"] else None
        roof_hbm = {"bound": "hbm", "achieved": per_gpu_cw_s * (b_hbm if onchip else b_msg) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "peak_source": hbm_src, "traffic": None}
        roof_hbm["frac"] = roof_hbm["achieved"] / hbm_peak
        try:  # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch)
            tr = json.loads((ROOT / "profiles" / ("ncu_traffic_%s.json" % workload)).read_text())
            if tr.get("codewords_per_launch") == ncw:
                roof_hbm["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roof_hbm["traffic_source"] = tr["source"]
                roof_hbm["algorithmic_bytes_per_launch"] = (b_hbm if onchip else b_msg) * ncw
        except Exception:
            pass
        r = dict(roof_smem if onchip else roof_hbm)
        kname = KERNELS.get(info["path_name"], "?")
        if info["path_name"] == "qc":
            kname = {0: "ldpc_ms_qc_kernel", 1: "ldpc_ms_qcw_kernel", 2: "ldpc_ms_qc_ring_kernel", 3: "ldpc_ms_qcm_kernel", 4: "ldpc_ms_qcm_multi_kernel"}.get(info.get("kernel_variant", 0), kname)
        r.update({"kernel": kname if args.algorithm == 0 else {1: "ldpc_sp_group_kernel", 2: "ldpc_tdmp_group_kernel"}[args.algorithm],
                  "launch_ms": ms_step, "traffic": roof_hbm.get("traffic"),
                  "algorithmic_bytes_per_codeword": {"hbm": b_hbm, "messages": b_msg, "mean_iterations": mean_iters},
                  "hbm": roof_hbm, "smem": roof_smem,
                  "note": ("messages stay in shared memory: HBM carries only channel values and bits, so shared-memory "
                           "bandwidth is the binding roofline" if onchip else
                           "messages live in a global workspace: HBM/L2 bandwidth is the binding roofline")})
        return r

    def spot_check(code, cap, llr, out, nwords, alg=0):
        """This rank's first `nwords` words through the oracle on the very same floats; AND over ranks."""
        M, N, K, rp, ci = code
        n = int(min(nwords, llr.shape[0]))
        y = llr[:n].cpu().numpy()
        o = oracle.Oracle(M, N, K, rp, ci, times=cap)
        if alg == 1:
            r = oracle.decode_sp(o, y)
        elif alg == 2:
            r = oracle.decode_tdmp(o, y, N // 24)
        else:
            r = o.decode(y, literal=False, want_post=False, want_hard=False)
        ok = bool(np.array_equal(r[0], out["info"][:n].cpu().numpy()) and np.array_equal(r[1], out["iters"][:n].cpu().numpy()))
        return min_over_ranks(1.0 if ok else 0.0) == 1.0, n

    def measure(workload, ncw, sigma, cap, steps, warmup, forced_path=-1, alg=0, hold_s=0.0):
        """Device-resident decode of this rank's shard of `workload`: CUDA events on the launching stream, barrier and
        synchronize on both sides, max over ranks."""
        code = make_code(workload)
        M, N, K, rp, ci = code
        # cfg3 (32 different checks per warp instruction): the bank-placement search of the setup runs at effort 48 instead
        # of the library default 12 -- 7.5 s instead of 2.8 s of one-off setup per code for +2.2 % decode throughput
        # (DESIGN.md section 4); the setup time is reported beside the number
        effort = PLACE_EFFORT.get(workload)
        if effort and "LDPC_B200_PLACE_EFFORT" not in os.environ:
            os.environ["LDPC_B200_PLACE_EFFORT"] = str(effort)
        else:
            effort = None
        t_setup = time.perf_counter()
        dec = m.Decoder(M, N, K, rp, ci, device=local_rank, max_iter=cap, early_termination=True)
        t_setup = time.perf_counter() - t_setup
        if effort:
            del os.environ["LDPC_B200_PLACE_EFFORT"]
        if workload in ("cfg1", "cfg2", "cfg4", "cfg4_4dB", "family_z60"):
            dec.set_layer_height(N // 24)
        if forced_path >= 0:
            dec.set_path(forced_path)
        if alg:
            dec.set_algorithm(alg)
        for kv in args.option:
            name, _, val = kv.partition("=")
            dec.set_option(name, int(val))
        info = dec.info()
        llr = m.synth_llr(ncw, N, sigma, seed=0x4C445043 + rank, device=local_rank)  # every rank: its own seeded shard
        out = {}
        t_first = time.perf_counter()
        dec.decode_device(llr[: min(ncw, 64)])   # (the tables of the kernel path are built and uploaded at the first launch)
        torch.cuda.synchronize()
        t_setup += time.perf_counter() - t_first
        for _ in range(max(warmup, 3)):
            dec.decode_device(llr, out=out)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = dec.launches
        barrier()
        torch.cuda.synchronize()
        sampler.summary(reset=True)
        sampler.region(True)
        ev0.record()
        for _ in range(steps):
            dec.decode_device(llr, out=out)
        ev1.record()
        torch.cuda.synchronize()
        sampler.region(False)
        barrier()
        ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        info = dec.info()   # (after the launches: kernel_variant says which quasi-cyclic kernel the handle settled on)
        res = {"dec": dec, "info": info, "code": code, "llr": llr, "out": out, "ms_step": ms_step, "setup_s": t_setup, "place_effort": effort,
               "launches": dec.launches - launches0, "clocks": sampler.summary(reset=True),
               "mean_iters": float(out["iters"].float().mean().item()), "total_cw": sum_over_ranks(float(ncw))}
        if hold_s > 0:  # the same loop held for hold_s seconds: sustained clocks and power
            n_hold = max(steps, int(math.ceil(hold_s * 1e3 / ms_step)))
            barrier()
            torch.cuda.synchronize()
            sampler.region(True)
            ev0.record()
            for _ in range(n_hold):
                dec.decode_device(llr, out=out)
            ev1.record()
            torch.cuda.synchronize()
            sampler.region(False)
            barrier()
            ms_hold = max_over_ranks(ev0.elapsed_time(ev1)) / n_hold
            res["sustained"] = {"steps": n_hold, "seconds": ms_hold * n_hold * 1e-3, "ms_per_step": ms_hold,
                                "value": res["total_cw"] * K / (ms_hold * 1e-3) / 1e9, "unit": UNIT,
                                "clocks": sampler.summary(reset=True)}
        return res

    def packed_legs(pdec, pllr, pncw, pN, pK, steps, want_f32=False):
        """End to end through ldpc_b200_decode_host_packed: the same channel values as float16 and as int8 (x 1/8) in pinned
        host buffers, widened on the device -- half / a quarter of the PCIe bytes.  (Additive API: the reference's is
        fp32; the quantised values are a different input, so the iteration counts move a little.)"""
        res = {}
        legs = [("f16", pllr.to(torch.float16).cpu().pin_memory(), 1.0, 2),
                ("i8", torch.clamp(torch.round(pllr * 8.0), -127, 127).to(torch.int8).cpu().pin_memory(), 0.125, 1)]
        if want_f32:
            legs.insert(0, ("f32", pllr.cpu().pin_memory(), 1.0, 4))
        h_o = {"info": torch.empty((pncw, pdec.KB), dtype=torch.uint8, pin_memory=True),
               "iters": torch.empty((pncw,), dtype=torch.int32, pin_memory=True)}
        for name, buf, scale, esz in legs:
            call = (lambda: pdec.decode_host(buf, out=h_o)) if name == "f32" else (lambda: pdec.decode_host_packed(buf, scale=scale, out=h_o))
            for _ in range(3):
                call()
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                call()
            torch.cuda.synchronize()
            dtp = max_over_ranks(time.perf_counter() - t0)
            barrier()
            res[name] = {"value": sum_over_ranks(float(pncw)) * pK * steps / dtp / 1e9, "unit": UNIT, "ms_per_step": dtp / steps * 1e3,
                         "h2d_bytes_per_step": int(sum_over_ranks(float(pncw))) * pN * esz, "mean_iterations": float(h_o["iters"].float().mean().item())}
            del buf
        return res

    # ---- headline: device-resident decode of the chosen workload
    desc, sigma, cap, ncw = WORKLOADS[args.workload]
    sigma = args.sigma if args.sigma is not None else sigma
    ncw = args.ncw or ncw
    cap = args.max_iter or cap
    R = measure(args.workload, ncw, sigma, cap, args.steps, args.warmup, forced_path=args.path, alg=args.algorithm,
                hold_s=0.0 if args.no_extras else 2.2)
    dec, info, code, llr, out = R["dec"], R["info"], R["code"], R["llr"], R["out"]
    M, N, K, rp, ci = code
    nnz = int(rp[-1])
    ms_step, total_cw, mean_iters, launches = R["ms_step"], R["total_cw"], R["mean_iters"], R["launches"]
    value = total_cw * K / (ms_step * 1e-3) / 1e9
    parity_ok, parity_n = spot_check(code, cap, llr, out, 256 if N <= 8192 else 2, alg=args.algorithm)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory in, bits out)
    e2e = None
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
        for _ in range(3):
            dec.decode_host(h_llr, out=h_out)
        dec.timing(reset=True)
        barrier()
        torch.cuda.synchronize()
        sampler.region(True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dec.decode_host(h_llr, out=h_out)
        torch.cuda.synchronize()
        dt_rank = time.perf_counter() - t0
        sampler.region(False)
        barrier()
        dt = max_over_ranks(dt_rank)
        tm = dec.timing()
        h2d_rank = ncw * N * 4 * args.steps / max(tm["h2d_s"], 1e-9) / 1e9
        e2e = {"value": total_cw * K * args.steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(total_cw) * N * 4, "d2h_bytes_per_step": int(total_cw) * (dec.KB + 4),
               "ms_per_step": dt / args.steps * 1e3,
               "api": "ldpc_b200_decode_host (pinned host buffers; " + ("one persistent launch fed by a copy stream)" if info["path_name"] == "qc" and args.algorithm == 0 else "3-stream pipeline)"),
               "matches_device_path": bool(torch.equal(h_out["info"], out["info"].cpu()) and torch.equal(h_out["iters"], out["iters"].cpu())),
               "host_numa": numa,
               # where the end-to-end time goes, per rank: the library's own phase timers (CUDA events on its streams)
               "phase_ms_per_step": {k: tm[k] / args.steps * 1e3 for k in ("wall_s", "h2d_s", "kernel_s", "d2h_s")},
               "h2d_gbs_per_gpu": [round(v, 2) for v in gather(h2d_rank)],
               "h2d_gbs_aggregate": round(sum_over_ranks(ncw * N * 4 * args.steps) / dt / 1e9, 2),
               "ms_per_step_per_rank": [round(v / args.steps * 1e3, 3) for v in gather(dt_rank)]}
        # the host's copy ceiling: the same pinned buffer copied to the device by every rank at once, nothing else running
        d_tmp = torch.empty_like(h_llr, device=dev)
        for _ in range(2):
            d_tmp.copy_(h_llr, non_blocking=True)
        barrier()
        torch.cuda.synchronize()
        cev0, cev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cev0.record()
        for _ in range(10):
            d_tmp.copy_(h_llr, non_blocking=True)
        cev1.record()
        torch.cuda.synchronize()
        barrier()
        ceil_rank = ncw * N * 4 * 10 / (cev0.elapsed_time(cev1) * 1e-3) / 1e9
        ceil = gather(ceil_rank)
        e2e["h2d_ceiling_gbs_per_gpu"] = [round(v, 2) for v in ceil]
        e2e["h2d_ceiling_note"] = ("plain cudaMemcpyAsync of the same pinned buffer on all %d ranks at once: what the host can feed; "
                                   "the end-to-end step cannot be shorter than bytes / this rate" % world)
        e2e["h2d_bound_ms_per_step"] = round(max(ncw * N * 4 / (v * 1e9) for v in ceil) * 1e3, 3)
        del h_llr, d_tmp
        if not args.no_extras and args.algorithm == 0:
            try:
                e2e["packed_input"] = packed_legs(dec, llr, ncw, N, K, max(5, args.steps // 2))
            except Exception as e:
                e2e["packed_input"] = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- the call a user of the reference makes: Coder::decode of the drop-in C++ class, malloc'd buffers
    plugin = None
    if not (args.no_extras or args.no_e2e) and args.workload in ("cfg1", "cfg2", "cfg4") and args.algorithm == 0:
        src_len = ncw * (K // 8)
        post = np.empty(ncw * N, dtype=np.float32)          # pageable (malloc/mmap'd) memory, as Test.cpp:39-41
        post[:] = llr.cpu().numpy().reshape(-1)
        src = np.zeros(src_len + 1, dtype=np.uint8)
        coder = m.Coder(K, N, m.rate_3_4_b, device=local_rank)
        coder.setMaxIter(cap)
        t0 = time.perf_counter()
        coder.forDecoder(ncw)                                # Test.cpp:47-48, 62: setup is outside the reference's timed region
        coder.addDecodeType(m.DecodeMS)
        setup_s = time.perf_counter() - t0
        barrier()
        t0 = time.perf_counter()
        coder.decode(post, src, src_len, m.DecodeMS)
        first_s = max_over_ranks(time.perf_counter() - t0)
        steps_p = max(5, args.steps // 2)
        barrier()
        sampler.region(True)
        t0 = time.perf_counter()
        for _ in range(steps_p):
            coder.decode(post, src, src_len, m.DecodeMS)
        dt_p = max_over_ranks(time.perf_counter() - t0)
        sampler.region(False)
        barrier()
        # the same buffer page-locked on first sight (Coder::setRegisterHostBuffers): no staging copy on the host afterwards
        coder2 = m.Coder(K, N, m.rate_3_4_b, device=local_rank)
        coder2.setMaxIter(cap)
        coder2.setRegisterHostBuffers(True)
        coder2.forDecoder(ncw)
        coder2.addDecodeType(m.DecodeMS)
        barrier()
        t0 = time.perf_counter()
        coder2.decode(post, src, src_len, m.DecodeMS)
        first_reg_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps_p):
            coder2.decode(post, src, src_len, m.DecodeMS)
        dt_r = max_over_ranks(time.perf_counter() - t0)
        barrier()
        registered = {"api": "the same call after Coder::setRegisterHostBuffers(true): the malloc'd postCode is page-locked at the first decode",
                      "value": total_cw * K * steps_p / dt_r / 1e9, "unit": UNIT, "ms_per_step": dt_r / steps_p * 1e3,
                      "first_call_ms": first_reg_s * 1e3,
                      "bytes_match_device_path": bool(np.array_equal(src[:src_len], out["info"].cpu().numpy().reshape(-1)))}
        coder2.close()
        plugin = {"api": "Coder::decode(postCode, srcCode, srcLength, DecodeMS) of libmyldpc_b200.so, malloc'd host buffers, setup (forDecoder + addDecodeType) outside the timed region as in Test.cpp",
                  "registered": registered,
                  "value": total_cw * K * steps_p / dt_p / 1e9, "unit": UNIT, "ms_per_step": dt_p / steps_p * 1e3, "steps": steps_p,
                  "first_call_ms": first_s * 1e3, "setup_ms": setup_s * 1e3,
                  "phase_ms_last_call": {k: v * 1e3 for k, v in coder.lastStepTimes().items()},
                  "bytes_match_device_path": bool(np.array_equal(src[:src_len], out["info"].cpu().numpy().reshape(-1)))}
        coder.close()
        del post

    # ---- the other BASELINE configs, same process, every N
    workloads = {}
    if not args.no_extras and args.workload == "cfg2" and args.algorithm == 0 and args.path < 0:
        del llr
        R["llr"] = None
        torch.cuda.empty_cache()
        for wl in ("cfg4", "cfg4_4dB", "family_z60", "cfg3", "cfg5"):
            wdesc, wsigma, wcap, wncw = WORKLOADS[wl]
            wsteps, wwarm, wcheck = SUB[wl]
            try:
                W = measure(wl, wncw, wsigma, wcap, wsteps, wwarm)
                wK = W["code"][2]
                ok, nchk = spot_check(W["code"], wcap, W["llr"], W["out"], wcheck)
                workloads[wl] = {"config": common_config(wl, W["code"], wsigma, wcap, wncw), "steps": wsteps, "warmup": max(wwarm, 3),
                                 "value": W["total_cw"] * wK / (W["ms_step"] * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": W["ms_step"],
                                 "mean_iterations": W["mean_iters"], "kernel_path": W["info"]["path_name"],
                                 "setup_s": round(W["setup_s"], 2), "place_effort": W["place_effort"],
                                 "roofline": {k: v for k, v in roofline_of(W["info"], W["code"], wncw, W["ms_step"], W["mean_iters"], wl).items()
                                              if k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "launch_ms", "traffic", "algorithmic_bytes_per_codeword")},
                                 "clocks": W["clocks"], "gpu_launches": W["launches"],
                                 "gpu_matches_oracle_all_ranks": ok, "oracle_words_per_rank": nchk}
                if wl == "cfg4" and not args.no_e2e:   # the regime in which the PCIe copy, not the kernel, bounds a host call
                    workloads[wl]["e2e_by_input_format"] = packed_legs(W["dec"], W["llr"], wncw, W["code"][1], wK, 10, want_f32=True)
                W["dec"].close()
                del W
            except Exception as e:  # a sub-record must never take the headline down
                workloads[wl] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.empty_cache()

    # ---- one process, all N GPUs through Coder::setDevices (rank 0 drives; the other ranks wait)
    setdev = None
    if world > 1 and not (args.no_extras or args.no_e2e) and args.workload in ("cfg1", "cfg2", "cfg4") and args.algorithm == 0:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                tot = ncw * world
                src_len = tot * (K // 8)
                rng = np.random.default_rng(7)
                post = (1.0 + sigma * rng.standard_normal(tot * N, dtype=np.float32)).astype(np.float32)
                src = np.zeros(src_len + 1, dtype=np.uint8)
                coder = m.Coder(K, N, m.rate_3_4_b)
                coder.setMaxIter(cap)
                coder.setDevices(list(range(world)))
                coder.forDecoder(tot)
                coder.addDecodeType(m.DecodeMS)
                t0 = time.perf_counter()
                coder.decode(post, src, src_len, m.DecodeMS)
                first_s = time.perf_counter() - t0
                t0 = time.perf_counter()
                for _ in range(5):
                    coder.decode(post, src, src_len, m.DecodeMS)
                dts = (time.perf_counter() - t0) / 5
                chk = oracle.Oracle(M, N, K, rp, ci, times=cap).decode(post.reshape(tot, N)[:: max(1, tot // 256)][:256], literal=False, want_post=False, want_hard=False)
                got = src[:src_len].reshape(tot, K // 8)[:: max(1, tot // 256)][:256]
                setdev = {"api": "one process, Coder::setDevices(0..%d), one host thread per GPU, malloc'd buffers" % (world - 1),
                          "codewords": tot, "value": tot * K / dts / 1e9, "unit": UNIT, "ms_per_step": dts * 1e3, "first_call_ms": first_s * 1e3,
                          "gpu_matches_oracle_on_sample": bool(np.array_equal(got, chk[0]))}
                coder.close()
                # the same with the caller's buffer page-locked at its first decode: every GPU DMAs its shard straight out of it
                src2 = np.zeros(src_len + 1, dtype=np.uint8)
                coder = m.Coder(K, N, m.rate_3_4_b)
                coder.setMaxIter(cap)
                coder.setDevices(list(range(world)))
                coder.setRegisterHostBuffers(True)
                coder.forDecoder(tot)
                coder.addDecodeType(m.DecodeMS)
                t0 = time.perf_counter()
                coder.decode(post, src2, src_len, m.DecodeMS)
                first_r = time.perf_counter() - t0
                t0 = time.perf_counter()
                for _ in range(5):
                    coder.decode(post, src2, src_len, m.DecodeMS)
                dtr = (time.perf_counter() - t0) / 5
                setdev["registered"] = {"api": "the same after Coder::setRegisterHostBuffers(true)", "value": tot * K / dtr / 1e9, "unit": UNIT,
                                        "ms_per_step": dtr * 1e3, "first_call_ms": first_r * 1e3,
                                        "bytes_match_unregistered": bool(np.array_equal(src, src2))}
                coder.close()
            except Exception as e:
                setdev = {"error": "%s: %s" % (type(e).__name__, e)}
        dist.barrier(group=cpu_group)
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roofline = roofline_of(info, code, ncw, ms_step, mean_iters, args.workload)

    # ---- CPU baseline: the oracle on a bounded sample of the same floats (+ parity check of that sample)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        y = m.synth_llr(min(ncw, 65536), N, sigma, seed=0x4C445043 + rank, device=local_rank).cpu().numpy()
        n, secs, ref, thr = cpu_decode_timed(code, cap, y)
        ok = bool(np.array_equal(ref[0], out["info"][:n].cpu().numpy()) and np.array_equal(ref[1], out["iters"][:n].cpu().numpy())) if args.algorithm == 0 else None
        cpu = {"value": n * K / secs / 1e9, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": "first %d codewords of the GPU batch, oracle (literal Coder::decodeCPU restatement, gcc -O2) on %d threads, %.1f s" % (n, thr, secs),
               "codewords_per_s": n / secs, "gpu_matches_oracle_on_sample": ok}

    topo = None
    if not args.no_extras:
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[:world + 12]
        except Exception:
            topo = None

    config = common_config(args.workload, code, sigma, cap, ncw)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic BPSK-AWGN channel values (all-zero codeword + seeded noise, generated on device)",
        "config": config,
        "details": {"mean_iterations": mean_iters, "parallelism": "codeword sharding x%d, no collective" % world,
                    "l2_policy": "inputs larger than L2 (%.0f MB of channel values per step)" % (ncw * N * 4 / 1e6),
                    "kernel_path": info["path_name"], "threads_per_cta": info["threads_per_cta"], "ctas": info["ctas"],
                    "smem_bytes_per_cta": info["smem_bytes"], "algorithm": args.algorithm},
        "codewords_per_s": total_cw / (ms_step * 1e-3),
        "e2e": e2e, "e2e_plugin": plugin, "gpu_launches": launches, "clocks": R["clocks"], "roofline": roofline,
        "sustained": R.get("sustained"), "parity_all_ranks": {"ok": parity_ok, "words_per_rank": parity_n, "ranks": world,
                                                              "what": "info bytes and iteration counts of each rank's first words against the oracle on the same floats"},
        "workloads": workloads or None, "setdevices": setdev, "cpu_baseline": cpu, "pcie_topology": topo,
    }
    emit_record(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
