#!/usr/bin/env python
"""bench.py -- SA build MB/s (1 GiB input) on N B200s; queries/sec as a secondary figure.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload rep_1G|rand_256M|acgt_4M|...] [--no-queries] [--no-cpu-baseline]

One "step" = one full suffix-array construction of the workload (default: BASELINE config
"SA of 1 GiB highly repetitive text", the 1 GiB input the metric is quoted on).
  value      whole-job MB/s (10^6 input bytes / s), text already resident in HBM, device-timed
  e2e        same metric through the reference-facing call gsa_divsufsort() with HOST buffers
             (pinned), host->device and device->host copies inside the timed region
  roofline   dominant kernel k_radix_pass: 24 B moved per element and launch (12 read + 12
             written), duration from CUDA events around every launch inside the timed region
  cpu_baseline  the reference's own C libdivsufsort (oracle/_ref) on the box's host cores, on
             a bounded sample of the same workload (N=1, rank 0 only)
N > 1 (torchrun): sacapart's model -- every rank builds the SA of its own 1 GiB partition,
no data-path collective (scaling "weak"); time is the max over ranks.
--impl reference: the reference's CPU implementation timed on the same metric (bounded sample).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MB = 1e6
CPU_SAMPLE_BYTES = 32 << 20  # bounded sample of the workload for the CPU arms (~5-8 s / build)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(name: str, rank: int = 0) -> np.ndarray:
    from stringsearch_b200 import synth

    if name == "rep_1G":
        return synth.repetitive(1 << 30, 3 + 1000 * rank)
    if name == "rand_256M":
        return synth.random_bytes(1 << 28, 2 + 1000 * rank)
    if name == "acgt_4M":
        return synth.acgt(4 << 20, 1 + 1000 * rank)
    if name == "acgt_512M":
        return synth.acgt(536870913, 4 + 1000 * rank)
    if name.startswith("rep_") and name.endswith("M"):
        return synth.repetitive(int(name[4:-1]) << 20, 3 + 1000 * rank)
    if name.startswith("rand_") and name.endswith("M"):
        return synth.random_bytes(int(name[5:-1]) << 20, 2 + 1000 * rank)
    if name.startswith("acgt_") and name.endswith("M"):
        return synth.acgt(int(name[5:-1]) << 20, 1 + 1000 * rank)
    raise SystemExit(f"unknown workload {name}")


WORKLOAD_DESC = {
    "rep_1G": "SA of 1 GiB period-1000 text with 1e-3 byte mutations (BASELINE config 2, many doubling rounds)",
    "rand_256M": "SA of 256 MiB uniform-random bytes (BASELINE config 1)",
    "acgt_4M": "SA of 4 MiB ACGT (BASELINE config 0)",
}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation (oracle/_ref = its vendored C libdivsufsort)
# ------------------------------------------------------------------------------------------
def cpu_reference_build(sample_views, threads: int):
    """Build the SA of every sample with the reference C divsufsort, `threads` samples at a time
    (the sacapart / rayon model).  Returns wall seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle

    ref = oracle.ref(ndebug=True)  # -DNDEBUG build: the faster of the two
    outs = [np.empty(v.size, dtype=np.int32) for v in sample_views]

    def one(i):
        v = sample_views[i]
        rc = ref.lib.divsufsort(v.ctypes.data_as(C.POINTER(C.c_uint8)), outs[i].ctypes.data_as(C.POINTER(C.c_int32)), v.size)
        assert rc == 0

    t0 = time.perf_counter()
    if threads <= 1 or len(sample_views) == 1:
        for i in range(len(sample_views)):
            one(i)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(len(sample_views))))
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm
    from oracle import oracle

    oracle.build()
    n_gpus = args.gpus
    samples = [np.ascontiguousarray(make_workload(args.workload, r)[:CPU_SAMPLE_BYTES]) for r in range(n_gpus)]
    threads = min(n_gpus, os.cpu_count() or 1)
    for _ in range(args.warmup):
        cpu_reference_build(samples[:1], 1) if n_gpus == 1 else cpu_reference_build(samples, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_reference_build(samples, threads)
    total_bytes = sum(s.size for s in samples) * args.steps
    value = total_bytes / t / MB
    sample_desc = f"first {samples[0].size >> 20} MiB of each rank's {args.workload} input ({len(samples)} sample(s), {threads} thread(s))"
    out = {
        "impl": "reference", "metric": "SA build MB/s", "value": value, "unit": "MB/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 text / i32 indices", "data": "synthetic",
        "config": {"workload": args.workload, "desc": WORKLOAD_DESC.get(args.workload, args.workload),
                   "bytes_per_gpu": int(make_workload_size(args.workload)), "sample": sample_desc},
        "cpu_baseline": {"value": value, "unit": "MB/s", "cores": threads, "kind": "reference",
                         "sample": sample_desc + "; libdivsufsort C from crates/cdivsufsort/c-sources, gcc -O3 -DNDEBUG"},
        "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)
    return 0


def make_workload_size(name: str) -> int:
    if name == "rep_1G":
        return 1 << 30
    if name == "rand_256M":
        return 1 << 28
    if name == "acgt_512M":
        return 536870913
    return int(name.split("_")[1][:-1]) << 20


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from stringsearch_b200 import _native as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; stringsearch_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) out of it
        os.environ["NCCL_DEBUG"] = os.environ.get("GSA_NCCL_DEBUG", "NONE")  # (NCCL prints the banner at VERSION and at WARN)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_host = make_workload(args.workload, rank)
    n = int(t_host.size)
    log(f"[rank {rank}] workload {args.workload}: n={n}")
    d_t = torch.from_numpy(t_host).to(dev)
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    ws_bytes = N.lib.gsa_build_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)  # workspace is part of the resident state
    stream = torch.cuda.current_stream(dev).cuda_stream
    stats = N.BuildStats()

    def step():
        rc = N.lib.gsa_build_device(d_t.data_ptr(), d_sa.data_ptr(), n, ws.data_ptr(), ws_bytes, stream, C.byref(stats))
        if rc != 0:
            raise RuntimeError(f"gsa_build_device rc={rc}: {N.last_error()}")

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms = pass_elems = pass_launches = launches = 0
    rounds_log = None
    ev0.record()
    for _ in range(args.steps):
        step()
        pass_ms += stats.ms_radix_passes
        pass_elems += stats.radix_pass_elements
        pass_launches += stats.radix_pass_launches
        launches += stats.kernel_launches
        rounds_log = stats.rounds_list()
        alg_bytes = stats.algorithmic_bytes()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * n * args.steps / (ms_max / 1e3) / MB

    # correctness gate inside the bench: O(n) GPU sufcheck of the last SA
    bad = C.c_int64(-1)
    rc = N.lib.gsa_sufcheck_device(d_t.data_ptr(), d_sa.data_ptr(), n, stream, C.byref(bad))
    if rc != 0:
        raise RuntimeError(f"bench: the SA failed sufcheck (rc={rc}, slot {bad.value})")

    # ---- secondary figure: LCP array of the resident text + SA (lcp.cu) -----------------------
    lcp_info = None
    if args.queries and world == 1:
        try:
            d_lcp = torch.empty(n, dtype=torch.int32, device=dev)

            def lcp_step():
                rc = N.lib.gsa_lcp_device(d_t.data_ptr(), d_sa.data_ptr(), d_lcp.data_ptr(), n, ws.data_ptr(), ws_bytes, stream)
                if rc != 0:
                    raise RuntimeError(f"gsa_lcp_device rc={rc}: {N.last_error()}")

            lcp_step()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            for _ in range(3):
                lcp_step()
            l1.record()
            torch.cuda.synchronize()
            lcp_ms = l0.elapsed_time(l1) / 3
            lcp_info = {"ms": lcp_ms, "value": n / (lcp_ms / 1e3) / MB, "unit": "MB/s of text",
                        "max_lcp": int(d_lcp.max().item()), "mean_lcp": float(d_lcp.double().mean().item()),
                        "api": "gsa_lcp_device on the resident text + SA (irreducible-PLCP scheme, lcp.cu)"}
            del d_lcp
        except Exception as e:  # secondary figure: never lose the headline line
            lcp_info = {"error": repr(e)}

    # ---- end to end through the reference-facing call (host pointers) -----------------------
    del ws
    torch.cuda.empty_cache()
    e2e = None
    if args.e2e:
        e2e = run_e2e(args, N, torch, dist, dev, local_rank, world, t_host, d_sa, n, barrier)

    # ---- queries at N > 1: the un-partitioned index replicated on every GPU, needles split across ranks ----
    queries_multi = None
    if args.queries and world > 1:
        try:
            queries_multi = bench_queries_replicated(dev, local_rank, rank, world, dist)
        except Exception as e:  # secondary figure: never lose the headline line
            queries_multi = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    achieved = (pass_elems * 24) / (pass_ms / 1e3) / 1e9 if pass_ms > 0 else 0.0
    out = {
        "metric": "SA build MB/s", "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 keys / u32 indices", "data": "synthetic",
        "config": {"workload": args.workload, "desc": WORKLOAD_DESC.get(args.workload, args.workload),
                   "bytes_per_gpu": n, "l2": "inputs larger than L2 (text + sort state of ~65 bytes per text byte per GPU)",
                   "parallelism": f"{world} independent partition(s), one per GPU (sacapart model)"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "kernel": "k_radix_pass (onesweep LSD pass, u64 key + u32 value)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": 24.3 * pass_elems / max(1, pass_launches), "traffic_source": "ncu --set full: dram read+write = 24.3 B/element (profiles/r1/v5_pass_rep1G.details.txt)", "launches": int(pass_launches), "avg_launch_ms": pass_ms / max(1, pass_launches),
            "algorithmic_bytes_per_launch": 24.0 * pass_elems / max(1, pass_launches), "algorithmic_bytes_formula": "24 B x elements (8+4 read, 8+4 written)",
            "share_of_step": pass_ms / ms if ms > 0 else None,
            "whole_build": {"algorithmic_bytes": int(alg_bytes), "achieved": alg_bytes * args.steps / (ms / 1e3) / 1e9,
                            "frac_of_peak": alg_bytes * args.steps / (ms / 1e3) / 1e9 / peak,
                            "formula": "round0 n(41+24p0) + sum_k (52 L_k + 24 p_k S_k + 32 B_k), S_k <= L_k suffixes actually sorted, B_k suffixes refined in the bag (SURVEY.md 8(d))"},
        },
        "rounds": rounds_log,
    }
    if args.cpu_baseline:
        sample = np.ascontiguousarray(t_host[:CPU_SAMPLE_BYTES])
        secs = cpu_reference_build([sample], 1)
        out["cpu_baseline"] = {"value": sample.size / secs / MB, "unit": "MB/s", "cores": 1, "kind": "reference",
                               "sample": f"first {sample.size >> 20} MiB of the same input, reference C libdivsufsort "
                                         f"(oracle/_ref, gcc -O3 -DNDEBUG), 1 thread, {secs:.1f} s; host has {os.cpu_count()} cpus"}
    if lcp_info is not None:
        out["lcp"] = lcp_info
    if queries_multi is not None:
        out["queries"] = queries_multi
    if args.queries and world == 1:
        try:
            out["queries"] = bench_queries(dev, local_rank)
        except Exception as e:  # secondary figure: never lose the headline line
            out["queries"] = {"error": repr(e)}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(args, N, torch, dist, dev, local_rank, world, t_host, d_sa, n, barrier):
    pin_t = N.PinnedBuffer(n)
    pin_sa = N.PinnedBuffer(4 * n)
    pin_t.array[:] = t_host
    sa_view = pin_sa.view(np.int32, n)
    e2e_steps = max(1, min(args.steps, 3))
    est = N.BuildStats()

    def e2e_step():
        rc = N.lib.gsa_divsufsort_ex(pin_t.array.ctypes.data, sa_view.ctypes.data, n, local_rank, C.byref(est))
        if rc != 0:
            raise RuntimeError(f"gsa_divsufsort_ex rc={rc}: {N.last_error()}")

    e2e_step()  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t_e.item()) / MB
    e2e_same = bool((torch.from_numpy(sa_view[: 1 << 20].copy()).to(dev) == d_sa[: 1 << 20]).all().item())
    if not e2e_same:
        raise RuntimeError("bench: e2e SA differs from the device-resident SA")
    e2e = {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": 4 * n,
           "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "ms_h2d": est.ms_h2d, "ms_d2h": est.ms_d2h,
           "ms_build": est.ms_total, "api": "gsa_divsufsort_ex(host T, host SA) with pinned host buffers; device scratch block cached between calls"}
    pin_t.free()
    pin_sa.free()
    return e2e


def bench_queries(dev, device_index):
    """BASELINE config 4 shape: 10M patterns x 32 B against the SA of a 1 GiB ACGT text
    (1 GPU here; the multi-GPU fan-out is exercised by tests).  Device-timed, patterns
    resident; plus the host-pointer call; plus the CPU oracle on a sample with all cores."""
    import torch
    from oracle import oracle
    from stringsearch_b200 import _native as N
    from stringsearch_b200 import synth

    n, Q, m = 1 << 30, 10_000_000, 32
    t = synth.acgt(n, 5)
    h = C.c_void_p()
    rc = N.lib.gsa_index_create(t.ctypes.data, n, device_index, C.byref(h), None)
    if rc != 0:
        raise RuntimeError(f"gsa_index_create rc={rc}: {N.last_error()}")
    try:
        flat, off = synth.patterns_from_text(t, Q, m, 6)
        d_p = torch.from_numpy(flat).to(dev)
        d_o = torch.from_numpy(off.astype(np.int64)).to(dev)
        d_s = torch.empty(Q, dtype=torch.int64, device=dev)
        d_l = torch.empty(Q, dtype=torch.int32, device=dev)
        d_left = torch.empty(Q, dtype=torch.int32, device=dev)
        d_cnt = torch.empty(Q, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        res = {"text": "1 GiB ACGT (seed 5)", "patterns": Q, "pattern_len": m}
        for name, fn in (("longest_substring_match", lambda: N.lib.gsa_lsm_device(h, d_p.data_ptr(), d_o.data_ptr(), Q, m, 0, 0, d_s.data_ptr(), d_l.data_ptr(), stream)),
                         ("search_all", lambda: N.lib.gsa_search_all_device(h, d_p.data_ptr(), d_o.data_ptr(), Q, m, d_left.data_ptr(), d_cnt.data_ptr(), stream))):
            assert fn() == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                assert fn() == 0
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            steps = 31 if name == "longest_substring_match" else 60
            res[name] = {"queries_per_s": Q / (ms / 1e3), "ms": ms,
                         "algorithmic_GBps": Q * steps * (4 + m) / (ms / 1e3) / 1e9,
                         "sector_GBps": Q * steps * (32 + 64) / (ms / 1e3) / 1e9}
        # host-pointer call, pinned buffers (H2D of 320 MB patterns + 80 MB offsets, D2H of results inside)
        pb = [N.PinnedBuffer(flat.nbytes), N.PinnedBuffer(off.nbytes), N.PinnedBuffer(8 * Q), N.PinnedBuffer(4 * Q)]
        pb[0].array[:] = flat
        pb[1].view(np.uint64, Q + 1)[:] = off
        st = pb[2].view(np.uint64, Q)
        ln = pb[3].view(np.uint32, Q)
        for _ in range(2):  # first call warms the allocator
            t0 = time.perf_counter()
            rc = N.lib.gsa_lsm_batch(h, pb[0].array.ctypes.data, pb[1].array.ctypes.data, Q, st.ctypes.data, ln.ctypes.data)
            dt = time.perf_counter() - t0
            assert rc == 0
        res["longest_substring_match"]["e2e_queries_per_s"] = Q / dt
        res["longest_substring_match"]["e2e_bytes"] = {"h2d": int(flat.nbytes + off.nbytes), "d2h": 12 * Q}
        st, ln = st.copy(), ln.copy()
        for b in pb:
            b.free()
        # CPU: oracle port of sacabase::longest_substring_match, OpenMP over patterns, sample of the batch
        port = oracle.port()
        sa = np.empty(n, dtype=np.int32)
        assert N.lib.gsa_index_sa(h, sa.ctypes.data) == 0
        qs = 400_000
        sub = (flat[: qs * m], off[: qs + 1])
        t0 = time.perf_counter()
        cs, cl = port.lsm_batch(t, sa, sub, threads=0)
        secs = time.perf_counter() - t0
        res["cpu_baseline"] = {"queries_per_s": qs / secs, "cores": port.max_threads(), "kind": "port",
                               "sample": f"first {qs} patterns, oracle longest_substring_match, OpenMP"}
        assert (cs == st[:qs]).all() and (cl == ln[:qs]).all(), "bench: GPU query results differ from the oracle"
        res["hit_fraction"] = float((ln == m).mean())
        return res
    finally:
        N.lib.gsa_index_destroy(h)


def bench_queries_replicated(dev, device_index, rank, world, dist):
    """BASELINE config 4 read literally ("1 GiB SA, N GPUs"): every rank builds the same 1 GiB ACGT
    index, answers its 1/N of the 10M x 32 B needles, and the answers are all-gathered (NCCL)
    inside the timed region.  queries/s = all needles / max over ranks of the device time."""
    import torch
    from stringsearch_b200 import _native as N
    from stringsearch_b200 import synth

    n, Q, m = 1 << 30, 10_000_000, 32
    t = synth.acgt(n, 5)
    h = C.c_void_p()
    rc = N.lib.gsa_index_create(t.ctypes.data, n, device_index, C.byref(h), None)
    if rc != 0:
        raise RuntimeError(f"gsa_index_create rc={rc}: {N.last_error()}")
    try:
        flat, _ = synth.patterns_from_text(t, Q, m, 6)
        per = (Q + world - 1) // world
        lo, hi = min(Q, rank * per), min(Q, (rank + 1) * per)
        d_p = torch.from_numpy(flat[lo * m:hi * m].copy()).to(dev)
        d_o = (torch.arange(hi - lo + 1, dtype=torch.int64, device=dev) * m)
        d_s = torch.zeros(per, dtype=torch.int64, device=dev)
        d_l = torch.zeros(per, dtype=torch.int32, device=dev)
        g_s = torch.empty(world * per, dtype=torch.int64, device=dev)
        g_l = torch.empty(world * per, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream

        d_left = torch.zeros(per, dtype=torch.int32, device=dev)
        d_cnt = torch.zeros(per, dtype=torch.int32, device=dev)
        g_left = torch.empty(world * per, dtype=torch.int32, device=dev)
        g_cnt = torch.empty(world * per, dtype=torch.int32, device=dev)

        def step_lsm():
            rc = N.lib.gsa_lsm_device(h, d_p.data_ptr(), d_o.data_ptr(), hi - lo, m, 0, 0, d_s.data_ptr(), d_l.data_ptr(), stream)
            if rc != 0:
                raise RuntimeError(f"gsa_lsm_device rc={rc}: {N.last_error()}")
            dist.all_gather_into_tensor(g_s, d_s)
            dist.all_gather_into_tensor(g_l, d_l)

        def step_all():
            rc = N.lib.gsa_search_all_device(h, d_p.data_ptr(), d_o.data_ptr(), hi - lo, m, d_left.data_ptr(), d_cnt.data_ptr(), stream)
            if rc != 0:
                raise RuntimeError(f"gsa_search_all_device rc={rc}: {N.last_error()}")
            dist.all_gather_into_tensor(g_left, d_left)
            dist.all_gather_into_tensor(g_cnt, d_cnt)

        res = {"text": "1 GiB ACGT (seed 5), replicated on every GPU", "patterns": Q, "pattern_len": m, "gpus": world}
        for name, step in (("longest_substring_match", step_lsm), ("search_all", step_all)):
            step()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            dist.barrier()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[name] = {"queries_per_s": Q / (float(ms.item()) / 1e3), "ms": float(ms.item()),
                         "includes": "all-gather of the two result arrays over NCCL"}
        res["hit_fraction"] = float((g_l[:Q] == m).double().mean().item())
        return res
    finally:
        N.lib.gsa_index_destroy(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rep_1G")
    ap.add_argument("--no-queries", dest="queries", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false", help="profiling runs only: skip the host-pointer leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if args.warmup < 3:
        log("bench.py: warm-up raised to 3 (timing rules)")
        args.warmup = 3
    return run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
