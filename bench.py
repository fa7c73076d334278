#!/usr/bin/env python
"""bench.py -- SA build MB/s (1 GiB input) on N B200s; partitioned index and queries/sec next to it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload rep_1G|rand_256M|acgt_4M|part_4G|...] [--no-queries] [--no-part]
                    [--no-extras] [--no-cpu-baseline] [--no-e2e]

One "step" = one full suffix-array construction of the workload.  Default workload: `rep_1G`,
BASELINE.json configs[2] "SA of 1 GiB highly repetitive text" -- the 1 GiB input the metric is quoted on.
  value      whole-job MB/s (10^6 input bytes / s), text already resident in HBM, device-timed
  e2e        same metric through the reference-facing call gsa_divsufsort() with HOST buffers
             (pinned), host->device and device->host copies inside the timed region; the host SA is
             compared slot by slot with the device-resident one
  roofline   dominant kernel k_radix_pass: 24 B moved per element and launch (12 read + 12
             written), duration from CUDA events around every launch inside the timed region;
             whole_build = the same for the whole step with the byte count of DESIGN.md section 2
  cpu_baseline  the reference's own C libdivsufsort (oracle/_ref) on the box's host cores, on
             a bounded sample of the same workload (N=1, rank 0 only)
Further keys of the same JSON line (each measured through the repo's public API):
  rand_256M, acgt_4M   the other two single-text BASELINE configs, device-timed + host-to-host
  part_4G    BASELINE configs[3]: ONE 4 GiB ACGT text, PartitionedSuffixArray with 8 partitions of
             536 870 913 bytes, partition i on rank i % N: build (max over ranks) and the fan-out query
             (pattern broadcast -> per-shard answers -> NCCL all-gather -> device merge) all timed
  queries    BASELINE configs[4]: 10 M x 32-byte patterns against a 1 GiB ACGT index: the kernels with
             resident patterns, the host-pointer calls, and ReplicatedSuffixArray over N ranks; every
             answer of the batch is compared with the CPU oracle
  parity_gate  (N > 1) DistributedPartitionedSuffixArray and ReplicatedSuffixArray answers over NCCL
             against the CPU oracle on >= 100 k needles; a mismatch makes the job exit non-zero
N > 1 (torchrun): sacapart's model -- every rank builds the SA of its own 1 GiB partition, no
data-path collective (scaling "weak"); time is the max over ranks.
--impl reference: the reference's CPU implementation (its vendored C libdivsufsort compiled into
oracle/_ref) timed on the same metric and config on the host cores; this process never loads libgsa.so.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
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
CPU_SAMPLE_BYTES = 256 << 20   # bounded sample of a 1 GiB workload for our arm's cpu_baseline (one build, ~20 s)
REF_SAMPLE_BYTES = 128 << 20   # ... and for each of the K timed steps of --impl reference (~10 s per step)
CPU_WARMUP_BYTES = 32 << 20    # CPU warm-up steps only page the library and buffers in
PART_N, PART_P = 1 << 32, 8    # BASELINE configs[3]
QUERY_N, QUERY_Q, QUERY_M = 1 << 30, 10_000_000, 32  # BASELINE configs[4]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_synth():
    """stringsearch_b200/synth.py is pure numpy; it is loaded by path so that the reference arm does not
    import the package (which would map libgsa.so into the CPU process)."""
    spec = importlib.util.spec_from_file_location("gsa_synth", os.path.join(ROOT, "stringsearch_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synth = load_synth()


def make_workload(name: str, rank: int = 0) -> np.ndarray:
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


def workload_size(name: str) -> int:
    if name == "rep_1G":
        return 1 << 30
    if name == "rand_256M":
        return 1 << 28
    if name == "acgt_512M":
        return 536870913
    if name == "part_4G":
        return PART_N
    return int(name.split("_")[1][:-1]) << 20


# BASELINE.json `configs` is a 0-based list: [0] divsuftest 4 MiB ACGT, [1] 256 MiB random, [2] 1 GiB repetitive,
# [3] 4 GiB partitioned, [4] 10 M patterns against a 1 GiB SA.
WORKLOAD_DESC = {
    "rep_1G": "SA of 1 GiB period-1000 text with 1e-3 byte mutations (BASELINE configs[2], many doubling rounds)",
    "rand_256M": "SA of 256 MiB uniform-random bytes (BASELINE configs[1])",
    "acgt_4M": "SA of 4 MiB ACGT (BASELINE configs[0], the divsuftest case)",
    "part_4G": "PartitionedSuffixArray of one 4 GiB ACGT text, 8 partitions of 536870913 bytes, partition i on rank i % N (BASELINE configs[3])",
}


def config_of(workload: str, world: int) -> dict:
    """The `config` object -- the same for both arms (the reference arm runs on our arm's config)."""
    if workload == "part_4G":
        return {"workload": workload, "desc": WORKLOAD_DESC[workload], "bytes_total": PART_N, "partitions": PART_P,
                "l2": "inputs larger than L2 (512 MiB of text + ~33 GB of sort state per shard build)",
                "parallelism": f"{PART_P} partitions over {world} GPU(s), no collective in the build; query = broadcast + all-gather"}
    n = workload_size(workload)
    return {"workload": workload, "desc": WORKLOAD_DESC.get(workload, workload), "bytes_per_gpu": n,
            "l2": "inputs larger than L2 (text + sort state of ~65 bytes per text byte per GPU)",
            "parallelism": f"{world} independent partition(s), one per GPU (sacapart model)"}


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


def pass_traffic_per_element():
    """DRAM bytes per element and launch of k_radix_pass from the newest committed `ncu --set full` capture
    (profiles/*/pass_traffic.json, written by tools/ncu_summary.py --traffic); None if there is none."""
    best = None
    prof = os.path.join(ROOT, "profiles")
    for d in sorted(os.listdir(prof)) if os.path.isdir(prof) else []:
        p = os.path.join(prof, d, "pass_traffic.json")
        if os.path.exists(p):
            try:
                j = json.load(open(p))
                best = (float(j["dram_bytes_per_element"]), f"{os.path.relpath(p, ROOT)}: {j.get('source', '')}")
            except Exception:
                pass
    return best


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
    cores = os.cpu_count() or 1
    t_start = time.perf_counter()
    if args.workload == "part_4G":
        # sacapart on the CPU: the 8 chunks of the one 4 GiB text on min(8, cores) threads (rayon's par_chunks),
        # each step a bounded sample: the first CPU_SAMPLE_BYTES / 4 of every chunk
        text = synth.acgt(PART_N, 4)
        ps = PART_N // PART_P + 1
        per = min(ps, REF_SAMPLE_BYTES // 4)
        full = [text[i * ps:min(PART_N, (i + 1) * ps)] for i in range(PART_P)]
        samples = [np.ascontiguousarray(c[:per]) for c in full]
        threads = min(PART_P, cores)
        sample_desc = (f"first {per >> 20} MiB of each of the {PART_P} partitions of the 4 GiB text, {threads} thread(s) "
                       f"(rayon par_chunks model, crates/sacapart/src/lib.rs:45-49)")
    else:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(min(n_gpus, cores)) as ex:  # the texts of the N ranks (numpy releases the GIL in its generators)
            full = list(ex.map(lambda r: make_workload(args.workload, r), range(n_gpus)))
        per = min(full[0].size, REF_SAMPLE_BYTES)
        samples = [np.ascontiguousarray(t[:per]) for t in full]
        threads = min(n_gpus, cores)
        sample_desc = (f"first {per >> 20} MiB of each rank's {args.workload} input" if per < full[0].size
                       else f"the full {args.workload} input of each rank") + f" ({len(samples)} text(s), {threads} thread(s))"
    warm = [np.ascontiguousarray(s[:CPU_WARMUP_BYTES]) for s in samples]
    for _ in range(args.warmup):
        cpu_reference_build(warm, threads)
    # the full-size config once (no sampling; N = 1 only, ~2 min): shows what the per-step sample hides
    full_once = None
    if args.full_once and n_gpus == 1 and per < full[0].size:
        secs = cpu_reference_build([np.ascontiguousarray(x) for x in full], threads)
        full_once = {"bytes": int(sum(x.size for x in full)), "seconds": secs, "value": sum(x.size for x in full) / secs / MB,
                     "unit": "MB/s", "cores": threads,
                     "what": "every text of the config at full size, built once (not part of the K timed steps)"}
    t = 0.0
    for _ in range(args.steps):
        t += cpu_reference_build(samples, threads)
    total_bytes = sum(s.size for s in samples) * args.steps
    value = total_bytes / t / MB
    lib_desc = "libdivsufsort C from crates/cdivsufsort/c-sources (oracle/_ref, gcc -O3 -DNDEBUG)"
    out = {
        "impl": "reference", "metric": "SA build MB/s", "value": value, "unit": "MB/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "part_4G" else "weak", "vs_baseline": None,
        "dtype": "u8 text / i32 indices", "data": "synthetic",
        "config": config_of(args.workload, n_gpus),
        "cpu_baseline": {"value": value, "unit": "MB/s", "cores": threads, "kind": "reference",
                         "sample": f"{sample_desc}; {lib_desc}; warm-up steps on {CPU_WARMUP_BYTES >> 20} MiB slices; host has {cores} cpus"},
        "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if full_once is not None:
        out["full_config_once"] = full_once
        out["cpu_baseline"]["sample"] += f"; the un-sampled config built once: {full_once['value']:.2f} MB/s (full_config_once)"
    print("\n" + json.dumps(out), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def bind_near_gpu(gpu_index: int) -> dict:
    """Multi-GPU runs: pin this process (and therefore the first-touch placement of its page-locked host buffers) to
    the CPUs NVML reports as local to its GPU.  With N ranks copying 5 GiB per step each, buffers on the far socket cross
    the inter-socket link as well as the PCIe switch.  What a deployment does with numactl; harmless when NVML is absent."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return {"bound": False, "why": "NVML reports no narrower CPU set", "cpus": len(allowed)}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": len(cpus), "first_cpu": min(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": f"{type(e).__name__}: {e}"[:120]}


class Ctx:
    """Per-process state shared by the legs of the bench."""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        from stringsearch_b200 import _native as N

        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        self.torch, self.dist, self.N = torch, dist, N
        self.dev = torch.device("cuda", local_rank)
        self.numa = bind_near_gpu(local_rank) if world > 1 else {"bound": False, "why": "one process, all cores"}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        if self.world == 1:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream


def whole_build_bytes(rounds) -> int:
    """Algorithmic bytes of one build from its per-round log ("whole build"):
    round 0: n (41 + 24 p0); round k >= 1: 8 L_k (the two label reads of every walked suffix -- an inert
    member of a huge group costs nothing else) + (44 + 24 p_k) S_k (a sorted suffix: key + suffix written,
    histogram, p_k passes, slot + rebuild reads, label / SA writes) + 32 B_k (a bag entry: suffix, slot,
    label gather, key half out and in, suffix + slot out, SA)."""
    total = 0
    for i, r in enumerate(rounds):
        if i == 0:
            total += r["live"] * (41 + 24 * r["passes"])
        else:
            total += 8 * r["live"] + (44 + 24 * r["passes"]) * r["sorted"] + 32 * r["bag"]
    return total


def timed_builds(cx: Ctx, d_t, d_sa, n, ws, ws_bytes, warmup, steps, sample_clocks=False):
    """`warmup` untimed + `steps` timed gsa_build_device calls.  -> dict(ms, ms_max, pass_*, launches, rounds, clocks)."""
    torch, N = cx.torch, cx.N
    stats = N.BuildStats()
    stream = cx.stream()

    def step():
        rc = N.lib.gsa_build_device(d_t.data_ptr(), d_sa.data_ptr(), n, ws.data_ptr() if ws is not None else None, ws_bytes, stream, C.byref(stats))
        if rc != 0:
            raise RuntimeError(f"gsa_build_device rc={rc}: {N.last_error()}")

    for _ in range(warmup):
        step()
    sampler = ClockSampler(cx.local_rank) if sample_clocks else None
    cx.barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r = dict(pass_ms=0.0, pass_elems=0, pass_bytes=0, pass_launches=0, launches=0, rounds=None)
    ev0.record()
    for _ in range(steps):
        step()
        r["pass_ms"] += stats.ms_radix_passes
        r["pass_elems"] += stats.radix_pass_elements
        r["pass_bytes"] += stats.radix_pass_bytes
        r["pass_launches"] += stats.radix_pass_launches
        r["launches"] += stats.kernel_launches
        r["rounds"] = stats.rounds_list()
    ev1.record()
    cx.barrier()
    r["clocks"] = sampler.stop() if sampler else None
    r["ms"] = ev0.elapsed_time(ev1)
    r["ms_max"] = cx.max_over_ranks(r["ms"])
    return r


def sufcheck_or_die(cx: Ctx, d_t, d_sa, n, what):
    bad = C.c_int64(-1)
    rc = cx.N.lib.gsa_sufcheck_device(d_t.data_ptr(), d_sa.data_ptr(), n, cx.stream(), C.byref(bad))
    if rc != 0:
        raise RuntimeError(f"bench: the SA of {what} failed sufcheck (rc={rc}, slot {bad.value})")


def run_e2e(cx: Ctx, t_host, d_sa, n, steps):
    """gsa_divsufsort_ex(host T, host SA), pinned buffers; the WHOLE host SA is compared with the device one."""
    torch, N = cx.torch, cx.N
    pin_t = N.PinnedBuffer(n)
    pin_sa = N.PinnedBuffer(4 * n)
    pin_t.array[:] = t_host
    sa_view = pin_sa.view(np.int32, n)
    est = N.BuildStats()

    def e2e_step():
        rc = N.lib.gsa_divsufsort_ex(pin_t.array.ctypes.data, sa_view.ctypes.data, n, cx.local_rank, C.byref(est))
        if rc != 0:
            raise RuntimeError(f"gsa_divsufsort_ex rc={rc}: {N.last_error()}")

    t0 = time.perf_counter()
    e2e_step()  # first call: allocates the cached device scratch block (reported as the cold time)
    cold_s = time.perf_counter() - t0
    e2e_step()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = cx.max_over_ranks(time.perf_counter() - t0)
    same = True
    chunk = 1 << 28
    for lo in range(0, n, chunk):  # every slot, through the device
        hi = min(n, lo + chunk)
        same = same and bool(torch.equal(torch.from_numpy(sa_view[lo:hi]).to(cx.dev), d_sa[lo:hi]))
    if not same:
        raise RuntimeError("bench: the SA returned by gsa_divsufsort_ex differs from the device-resident SA")
    e2e = {"value": cx.world * n * steps / e2e_s / MB, "unit": "MB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": 4 * n,
           "steps": steps, "ms_per_step": e2e_s / steps * 1e3, "ms_h2d": est.ms_h2d, "ms_d2h": est.ms_d2h, "ms_build": est.ms_total,
           "ms_first_call_cold": cold_s * 1e3, "checked": "all n slots equal the device-resident SA",
           "api": "gsa_divsufsort_ex(host T, host SA) with pinned host buffers; device scratch block cached between calls "
                  "(ms_first_call_cold = the first call of the process, which allocates it)"}
    pin_t.free()
    pin_sa.free()
    N.lib.gsa_release_cached_memory()  # the next leg allocates its own state
    return e2e


def small_config(cx: Ctx, name: str, warmup=3, steps=5):
    """One of the smaller single-text configs: device-timed build + host-pointer call + sufcheck."""
    torch, N = cx.torch, cx.N
    t = make_workload(name, cx.rank)
    n = int(t.size)
    d_t = torch.from_numpy(t).to(cx.dev)
    d_sa = torch.empty(n, dtype=torch.int32, device=cx.dev)
    ws_bytes = N.lib.gsa_build_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cx.dev)
    r = timed_builds(cx, d_t, d_sa, n, ws, ws_bytes, warmup, steps)
    sufcheck_or_die(cx, d_t, d_sa, n, name)
    del ws
    torch.cuda.empty_cache()
    e2e = run_e2e(cx, t, d_sa, n, steps)
    return {"desc": WORKLOAD_DESC.get(name, name), "bytes_per_gpu": n, "value": cx.world * n * steps / (r["ms_max"] / 1e3) / MB,
            "unit": "MB/s", "ms_per_step": r["ms_max"] / steps, "steps": steps, "warmup": warmup,
            "rounds": [(x["depth"], x["live"], x["sorted"], x["passes"]) for x in r["rounds"]],
            "e2e": {k: e2e[k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "checked": "gsa_sufcheck_device on the last SA; host SA == device SA (all slots)"}


# ---- BASELINE configs[3]: one 4 GiB text, 8 partitions over the N ranks ----------------------------
def bench_part_4g(cx: Ctx, n=PART_N, P=PART_P, Q=2_000_000, m=32, steps=3):
    torch, N = cx.torch, cx.N
    from stringsearch_b200 import sacapart

    pin = N.PinnedBuffer(n)  # the caller's text, page-locked (a pageable numpy text is staged by the driver at ~11 GB/s)
    rng = np.random.default_rng(4)
    step_b = 1 << 28
    for lo in range(0, n, step_b):  # == synth.acgt(n, 4) (Generator.integers draws in order), without the 4 GiB temporaries
        pin.array[lo:lo + step_b] = synth._ACGT[rng.integers(0, 4, min(step_b, n - lo), dtype=np.uint8)]
    text = pin.array
    res = {"desc": WORKLOAD_DESC["part_4G"], "bytes_total": n, "partitions": P, "gpus": cx.world,
           "api": "sacapart.DistributedPartitionedSuffixArray(text, 8, device): gsa_index_create_shard per local partition; "
                  "longest_substring_match_device: broadcast + gsa_lsm_device per shard + all_gather + gsa_lsm_reduce_device"}
    psa = None
    walls, devs = [], []
    for it in range(1 + steps):  # first build is the warm-up (allocator, module load)
        if psa is not None:
            psa.close()
        cx.barrier()
        t0 = time.perf_counter()
        psa = sacapart.DistributedPartitionedSuffixArray(text, P, cx.local_rank)
        torch.cuda.synchronize()
        wall = cx.max_over_ranks(time.perf_counter() - t0)
        dev_ms = cx.max_over_ranks(sum(psa.build_ms))
        if it > 0:
            walls.append(wall)
            devs.append(dev_ms)
    assert psa.num_partitions() == P and psa._ps == n // P + 1
    res["build"] = {"value": n / (sum(walls) / len(walls)) / MB, "unit": "MB/s (host text -> resident shards, max over ranks)",
                    "ms_per_step": sum(walls) / len(walls) * 1e3, "steps": steps,
                    "device_only": {"value": n / (sum(devs) / len(devs) / 1e3) / MB, "ms_per_step": sum(devs) / len(devs),
                                    "what": "sum of gsa_build_stats.ms_total over the rank's shards, max over ranks"},
                    "h2d_bytes_per_step": n + (P - 1) * 4096, "partitions_per_rank": len(psa.local_partitions())}
    # every local shard passes the O(n) sufcheck
    ok = True
    for _, _, h in psa._shards:
        bad = C.c_int64(-1)
        ok = ok and N.lib.gsa_index_verify(h, C.byref(bad)) == 0
    if not cx.all_ok(ok):
        raise RuntimeError("bench: a part_4G shard failed sufcheck")
    # ---- fan-out query: needles cut from the text (anywhere, partition boundaries included) and random ones
    flat = off = None
    t_pat = t_off = None
    if cx.rank == 0:
        qrng = np.random.default_rng(44)
        o = qrng.integers(0, n - m, Q // 2)
        flat = np.empty((Q, m), dtype=np.uint8)
        idx = o[:, None] + np.arange(m)[None, :]
        flat[0::2] = text[idx]
        flat[1::2] = synth._ACGT[qrng.integers(0, 4, (Q - Q // 2, m), dtype=np.uint8)]
        ps = n // P + 1
        k = min(Q // 2, 7 * 64)  # plus needles that straddle every partition boundary
        bnd = np.array([(1 + j % 7) * ps - 1 - (j // 7) % (m - 1) for j in range(k)], dtype=np.int64)
        flat[0:2 * k:2] = text[bnd[:, None] + np.arange(m)[None, :]]
        o[:k] = bnd
        flat = flat.reshape(-1)
        off = np.arange(Q + 1, dtype=np.uint64) * np.uint64(m)
        t_pat = torch.from_numpy(flat).to(cx.dev)
        t_off = torch.from_numpy(off.astype(np.int64)).to(cx.dev)
    for _ in range(2):  # warm-up (NCCL channels, halo, prefix-bucket tables of the shards)
        psa.longest_substring_match_device(t_pat, t_off)
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d_s, d_l = psa.longest_substring_match_device(t_pat, t_off)
    e1.record()
    cx.barrier()
    q_ms = cx.max_over_ranks(e0.elapsed_time(e1) / steps)
    cx.barrier()
    t0 = time.perf_counter()
    hs, hl = psa.longest_substring_match_batch((flat, off) if cx.rank == 0 else None)  # host needles -> host answers
    q_wall = cx.max_over_ranks(time.perf_counter() - t0)
    res["query"] = {"patterns": Q, "pattern_len": m, "queries_per_s": Q / (q_ms / 1e3), "ms": q_ms,
                    "includes": "pattern broadcast, 8 per-shard searches, all-gather of (start, len), device merge" if cx.world > 1
                                else "8 per-shard searches on one GPU (world 1: no collective)",
                    "e2e_queries_per_s": Q / q_wall, "e2e_bytes": {"h2d": Q * m + 8 * (Q + 1), "d2h": 12 * Q}}
    # properties that hold whatever the tie-breaks: the reported match is real; a needle cut from the text is found in
    # full -- the halo rule (lib.rs:77-84) extends a match that touches the end of a shard, so that also holds
    # for a needle straddling a boundary provided its part in front of the boundary is unique in that shard (>= 24 bytes of
    # ACGT: 4^24 >> 512 Mi positions, so no second occurrence; shorter heads may legitimately yield shorter matches)
    if cx.rank == 0:
        assert (hs.astype(np.int64) == d_s.cpu().numpy()).all() and (hl.astype(np.int32) == d_l.cpu().numpy()).all()
        sub = np.arange(0, Q, max(1, Q // 200_000))
        st, ln = hs[sub].astype(np.int64), hl[sub].astype(np.int64)
        pat = flat.reshape(Q, m)[sub]
        got = text[np.minimum(st[:, None] + np.arange(m)[None, :], n - 1)]
        real = ((got == pat) | (np.arange(m)[None, :] >= ln[:, None])).all()
        hits = sub[sub % 2 == 0]
        inside = (o[hits // 2] % ps) <= ps - m  # the needle lies inside one partition
        full = hl[hits] == m
        head = ps - (o[hits // 2] % ps)  # bytes in front of the boundary for a straddling needle
        ok = bool(real) and bool(full[inside].all()) and bool(full[(~inside) & (head >= 24)].all())
        res["query"]["checked"] = (f"{sub.size} answers: text[start:start+len] == needle[:len]; {int(inside.sum())} in-partition needles and "
                                   f"{int(((~inside) & (head >= 24)).sum())} boundary-straddling ones found in full; device == host API results")
        res["query"]["hit_fraction"] = float((hl == m).mean())
    else:
        ok = True
    if not cx.all_ok(ok):
        raise RuntimeError("bench: part_4G query answers violate the match properties")
    psa.close()
    pin.free()
    N.lib.gsa_release_cached_memory()
    torch.cuda.empty_cache()
    return res


# ---- BASELINE configs[4]: 10 M x 32-byte patterns against a 1 GiB index ------------------------------
def bench_queries(cx: Ctx, n=QUERY_N, Q=QUERY_Q, m=QUERY_M, steps=3):
    """Kernels with resident patterns, the host-pointer calls (1 GPU), and sacapart.ReplicatedSuffixArray
    over the N ranks (broadcast + all-gather inside the timed region); on rank 0 EVERY answer of both
    searches is compared with the CPU oracle (OpenMP port of sacabase / sa_search)."""
    torch, N = cx.torch, cx.N
    from oracle import oracle
    from stringsearch_b200 import sacapart

    t = synth.acgt(n, 5)
    rsa = sacapart.ReplicatedSuffixArray(t, cx.local_rank)  # every rank builds its own copy (deterministic)
    h = rsa._h
    res = {"text": f"{n >> 20} MiB ACGT (seed 5)", "patterns": Q, "pattern_len": m, "gpus": cx.world,
           "api": "sacapart.ReplicatedSuffixArray.query_device: rank 0 sends every rank its 1/N of the needles (header broadcast + grouped NCCL send/recv), rank r answers them "
                  "(gsa_lsm_device / gsa_search_all_device), one all-gather per result array"}
    flat = off = t_pat = t_off = None
    if cx.rank == 0:
        flat, off = synth.patterns_from_text(t, Q, m, 6)
        t_pat = torch.from_numpy(flat).to(cx.dev)
        t_off = torch.from_numpy(off.astype(np.int64)).to(cx.dev)
    out = {}
    for what, key, nsteps in (("lsm", "longest_substring_match", 31), ("search_all", "search_all", 60)):
        for _ in range(2):  # warm-up (the first call also builds the prefix-bucket table of the index)
            rsa.query_device(t_pat, t_off, what)
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            a, b = rsa.query_device(t_pat, t_off, what)
        e1.record()
        cx.barrier()
        ms = cx.max_over_ranks(e0.elapsed_time(e1) / steps)
        out[what] = (a, b)
        res[key] = {"queries_per_s": Q / (ms / 1e3), "ms": ms,
                    "includes": ("header broadcast, point-to-point scatter of the needle slices, all-gather of the two result arrays over NCCL" if cx.world > 1 else
                                 "ReplicatedSuffixArray.query_device at world 1: header / max-length reductions and result allocation around the kernel")}
        if cx.world == 1:
            # the kernel alone: patterns, offsets and result arrays resident, one launch per step
            d_a = torch.empty(Q, dtype=torch.int64 if what == "lsm" else torch.int32, device=cx.dev)
            d_b = torch.empty(Q, dtype=torch.int32, device=cx.dev)

            def launch():
                if what == "lsm":
                    rc = N.lib.gsa_lsm_device(h, t_pat.data_ptr(), t_off.data_ptr(), Q, m, 0, 0, d_a.data_ptr(), d_b.data_ptr(), cx.stream())
                else:
                    rc = N.lib.gsa_search_all_device(h, t_pat.data_ptr(), t_off.data_ptr(), Q, m, d_a.data_ptr(), d_b.data_ptr(), cx.stream())
                if rc != 0:
                    raise RuntimeError(f"search kernel rc={rc}: {N.last_error()}")

            launch()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                launch()
            e1.record()
            torch.cuda.synchronize()
            kms = e0.elapsed_time(e1) / steps
            res[key]["kernel"] = {"queries_per_s": Q / (kms / 1e3), "ms": kms,
                                  "algorithmic_GBps": Q * nsteps * (4 + m) / (kms / 1e3) / 1e9,
                                  "sector_GBps": Q * nsteps * (32 + 64) / (kms / 1e3) / 1e9,
                                  "bytes_model": f"reference walk: {nsteps} steps x (4 B SA entry + {m} B text) per query (SURVEY 8(d)); sector = 32 + 64 B per step",
                                  "what": "gsa_lsm_device / gsa_search_all_device, everything resident"}
            assert torch.equal(d_a, a) and torch.equal(d_b, b)
    if cx.rank == 0:
        # host-pointer calls, pinned buffers (H2D of 320 MB patterns + 80 MB offsets, D2H of results inside)
        pb = [N.PinnedBuffer(flat.nbytes), N.PinnedBuffer(off.nbytes), N.PinnedBuffer(8 * Q), N.PinnedBuffer(4 * Q), N.PinnedBuffer(4 * Q)]
        pb[0].array[:] = flat
        pb[1].view(np.uint64, Q + 1)[:] = off
        st, ln = pb[2].view(np.uint64, Q), pb[3].view(np.uint32, Q)
        left, cnt = pb[3].view(np.int32, Q), pb[4].view(np.int32, Q)
        dts = []
        for _ in range(4):  # first call warms the allocator; the median of the other three is reported
            t0 = time.perf_counter()
            rc = N.lib.gsa_lsm_batch(h, pb[0].array.ctypes.data, pb[1].array.ctypes.data, Q, st.ctypes.data, ln.ctypes.data)
            dts.append(time.perf_counter() - t0)
            assert rc == 0
        dt = sorted(dts[1:])[1]
        res["longest_substring_match"]["e2e_queries_per_s"] = Q / dt
        res["longest_substring_match"]["e2e_bytes"] = {"h2d": int(flat.nbytes + off.nbytes), "d2h": 12 * Q}
        h_st, h_ln = st.copy(), ln.copy()
        dts = []
        for _ in range(4):
            t0 = time.perf_counter()
            rc = N.lib.gsa_search_all_batch(h, pb[0].array.ctypes.data, pb[1].array.ctypes.data, Q, left.ctypes.data, cnt.ctypes.data)
            dts.append(time.perf_counter() - t0)
            assert rc == 0
        dt = sorted(dts[1:])[1]
        res["search_all"]["e2e_queries_per_s"] = Q / dt
        res["search_all"]["e2e_bytes"] = {"h2d": int(flat.nbytes + off.nbytes), "d2h": 8 * Q}
        h_left, h_cnt = left.copy(), cnt.copy()
        for b in pb:
            b.free()
        # CPU: the oracle port (sacabase::longest_substring_match and sa_search), OpenMP over ALL patterns
        port = oracle.port()
        sa = np.empty(n, dtype=np.int32)
        assert N.lib.gsa_index_sa(h, sa.ctypes.data) == 0
        nthr = max(1, (os.cpu_count() or 1) // cx.world)  # (torchrun sets OMP_NUM_THREADS=1: ask for the cores explicitly)
        t0 = time.perf_counter()
        cs, cl = port.lsm_batch(t, sa, (flat, off), threads=nthr)
        secs_lsm = time.perf_counter() - t0
        t0 = time.perf_counter()
        c_left, c_cnt = port.search_all_batch(t, sa, (flat, off), threads=nthr)
        secs_all = time.perf_counter() - t0
        res["cpu_baseline"] = {"longest_substring_match_queries_per_s": Q / secs_lsm, "search_all_queries_per_s": Q / secs_all,
                               "cores": nthr, "kind": "port",
                               "sample": f"all {Q} patterns, oracle longest_substring_match / sa_search, OpenMP"}
        d_s, d_l = out["lsm"]
        d_left, d_cnt = out["search_all"]
        ok = ((cs == h_st).all() and (cl == h_ln).all() and (c_left == h_left).all() and (c_cnt == h_cnt).all()
              and (d_s.cpu().numpy().astype(np.uint64) == cs).all() and (d_l.cpu().numpy().astype(np.uint32) == cl).all()
              and (d_left.cpu().numpy() == c_left).all() and (d_cnt.cpu().numpy() == c_cnt).all())
        res["checked"] = f"all {Q} answers of both searches, host-pointer calls and ReplicatedSuffixArray, equal the CPU oracle's"
        res["hit_fraction"] = float((h_ln == m).mean())
    else:
        ok = True
    rsa.close()
    N.lib.gsa_release_cached_memory()
    torch.cuda.empty_cache()
    if not cx.all_ok(bool(ok)):
        raise RuntimeError("bench: GPU query results differ from the oracle")
    return res


# ---- N > 1: answers over NCCL against the CPU oracle -----------------------------------------------
def parity_gate(cx: Ctx, n=48 << 20, Q=120_000):
    """DistributedPartitionedSuffixArray (P = 8 and P = 5) and ReplicatedSuffixArray on a 48 MiB ACGT text over all
    ranks: every answer must equal the CPU oracle's (reference C divsufsort per partition + the port of
    crates/sacapart/src/lib.rs:69-97 / sacabase / sa_search).  Raises on mismatch -> non-zero exit."""
    torch = cx.torch
    from oracle import oracle
    from stringsearch_b200 import sacapart

    t = synth.acgt(n, 77)
    rng = np.random.default_rng(78)
    lens = rng.integers(1, 48, Q)
    starts = rng.integers(0, n - 64, Q)
    needles = []
    for j in range(Q):
        b = t[starts[j]:starts[j] + lens[j]].copy()
        if j % 3 == 1:
            b[rng.integers(0, b.size)] = 65 + j % 20  # a byte that may not occur in the text
        elif j % 3 == 2:
            b = synth._ACGT[rng.integers(0, 4, b.size)]
        needles.append(b.tobytes())
    # needles across the partition boundaries of the P = 8 plan
    ps8 = n // 8 + 1
    for i in range(1, 8):
        for k in (1, 5, 20, 40):
            needles.append(t[i * ps8 - k:i * ps8 - k + 44].tobytes())
    report = {"text": f"{n >> 20} MiB ACGT (seed 77)", "needles": len(needles)}
    ok = True
    port = oracle.port() if cx.rank == 0 else None
    ref = oracle.ref(ndebug=True) if cx.rank == 0 else None
    nthr = max(1, (os.cpu_count() or 1) // cx.world)  # (torchrun sets OMP_NUM_THREADS=1)
    for P in (8, 5):
        psa = sacapart.DistributedPartitionedSuffixArray(t, P, cx.local_rank)
        s, l = psa.longest_substring_match_batch(needles if cx.rank == 0 else None)
        psa.close()
        if cx.rank == 0:
            ps, sas = port.part_build(t, P, builder=ref.sa_build)
            es, el = port.part_lsm_batch(t, ps, sas, needles, threads=nthr)
            good = bool((s == es).all() and (l == el).all())
            report[f"partitioned_P{P}"] = "equal" if good else f"{int(((s != es) | (l != el)).sum())} answers differ"
            ok = ok and good
    rsa = sacapart.ReplicatedSuffixArray(t, cx.local_rank)
    s, l = rsa.longest_substring_match_batch(needles if cx.rank == 0 else None)
    left, cnt = rsa.search_all_batch(needles if cx.rank == 0 else None)
    rsa.close()
    if cx.rank == 0:
        sa = ref.sa_build(t)
        es, el = port.lsm_batch(t, sa, needles, threads=nthr)
        e_left, e_cnt = port.search_all_batch(t, sa, needles, threads=nthr)
        good = bool((s == es).all() and (l == el).all() and (left == e_left).all() and (cnt == e_cnt).all())
        report["replicated"] = "equal" if good else "answers differ"
        ok = ok and good
    torch.cuda.empty_cache()
    if not cx.all_ok(ok):
        raise RuntimeError(f"bench: multi-GPU parity gate failed: {report}")
    report["result"] = "every answer over NCCL equals the CPU oracle's"
    return report


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; stringsearch_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    cx = Ctx(args, rank, local_rank, world)
    N, dev = cx.N, cx.dev
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's log (kept on: it shows the ranks and the NVLink / NVLS transport)
        # goes to stderr unless the caller routed it elsewhere
        os.environ["NCCL_DEBUG"] = os.environ.get("GSA_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        if rank == 0:
            log(f"[bench] torch.distributed backend=nccl nranks={dist.get_world_size()} NCCL {'.'.join(map(str, torch.cuda.nccl.version()))}")

    if args.workload == "part_4G":
        return run_part_main(cx)

    t_host = make_workload(args.workload, rank)
    n = int(t_host.size)
    log(f"[rank {rank}] workload {args.workload}: n={n}")
    d_t = torch.from_numpy(t_host).to(dev)
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    ws_bytes = N.lib.gsa_build_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)  # workspace is part of the resident state
    r = timed_builds(cx, d_t, d_sa, n, ws, ws_bytes, args.warmup, args.steps, sample_clocks=True)
    ms, ms_max = r["ms"], r["ms_max"]
    value = world * n * args.steps / (ms_max / 1e3) / MB
    sufcheck_or_die(cx, d_t, d_sa, n, args.workload)  # correctness gate: O(n) GPU sufcheck of the last SA

    extras = {}

    def leg(name, fn, *a, **kw):
        """A secondary figure must never lose the headline line -- except a parity failure (RuntimeError('bench: ...'))."""
        try:
            extras[name] = fn(*a, **kw)
        except RuntimeError as e:
            if str(e).startswith("bench:"):
                raise
            extras[name] = {"error": repr(e)}
        except Exception as e:
            extras[name] = {"error": repr(e)}
        # all ranks stay in step even if one of them failed a leg
        cx.barrier()

    # ---- secondary figure: LCP array of the resident text + SA (lcp.cu) -----------------------
    def lcp_leg():
        d_lcp = torch.empty(n, dtype=torch.int32, device=dev)

        def lcp_step():
            rc = N.lib.gsa_lcp_device(d_t.data_ptr(), d_sa.data_ptr(), d_lcp.data_ptr(), n, ws.data_ptr(), ws_bytes, cx.stream())
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
        return {"ms": lcp_ms, "value": n / (lcp_ms / 1e3) / MB, "unit": "MB/s of text",
                "max_lcp": int(d_lcp.max().item()), "mean_lcp": float(d_lcp.double().mean().item()),
                "api": "gsa_lcp_device on the resident text + SA (irreducible-PLCP scheme, lcp.cu)"}

    if args.extras and world == 1:
        leg("lcp", lcp_leg)

    # ---- end to end through the reference-facing call (host pointers) -----------------------
    del ws
    torch.cuda.empty_cache()
    e2e = run_e2e(cx, t_host, d_sa, n, max(1, min(args.steps, 3))) if args.e2e else None
    del d_t, d_sa
    torch.cuda.empty_cache()

    cpu_baseline = None
    if args.cpu_baseline and rank == 0 and world == 1:
        sample = np.ascontiguousarray(t_host[:CPU_SAMPLE_BYTES])
        secs = cpu_reference_build([sample], 1)
        cpu_baseline = {"value": sample.size / secs / MB, "unit": "MB/s", "cores": 1, "kind": "reference",
                        "sample": f"first {sample.size >> 20} MiB of the same input, reference C libdivsufsort "
                                  f"(oracle/_ref, gcc -O3 -DNDEBUG), 1 thread, {secs:.1f} s; host has {os.cpu_count()} cpus"}
    del t_host

    if args.extras:
        for name in ("rand_256M", "acgt_4M"):
            if name != args.workload:
                leg(name, small_config, cx, name)
    if args.part:
        leg("part_4G", bench_part_4g, cx)
    if args.queries:
        leg("queries", bench_queries, cx)
    if world > 1 and args.gate:
        extras["parity_gate"] = parity_gate(cx)  # raises on mismatch: the job must fail

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    pass_ms, pass_elems, pass_launches = r["pass_ms"], r["pass_elems"], r["pass_launches"]
    pass_bytes = r["pass_bytes"]  # 12 B read + 12 B written per element and launch (less for the key-generating pass)
    achieved = pass_bytes / (pass_ms / 1e3) / 1e9 if pass_ms > 0 else 0.0
    traffic = pass_traffic_per_element()
    alg_bytes = whole_build_bytes(r["rounds"])
    out = {
        "metric": "SA build MB/s", "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 keys / u32 indices", "data": "synthetic",
        "config": config_of(args.workload, world),
        "clocks": r["clocks"],
        "host_binding": cx.numa,
        "e2e": e2e,
        "gpu_launches": int(r["launches"]),
        "roofline": {
            "bound": "hbm", "kernel": "k_radix_pass (onesweep LSD pass, u64 key + u32 value)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": traffic[0] * pass_bytes / 24.0 / max(1, pass_launches) if traffic else None,
            "traffic_source": traffic[1] if traffic else None,
            "launches": int(pass_launches), "avg_launch_ms": pass_ms / max(1, pass_launches),
            "algorithmic_bytes_per_launch": pass_bytes / max(1, pass_launches),
            "algorithmic_bytes_formula": "24 B x elements (8+4 read, 8+4 written); the key-generating first pass of round 0 reads "
                                         "bits_per_symbol / 8 B of packed text per element instead of a pair (gsa_build_stats.radix_pass_bytes)",
            "share_of_step": pass_ms / ms if ms > 0 else None,
            "whole_build": {"algorithmic_bytes": int(alg_bytes), "achieved": alg_bytes * args.steps / (ms / 1e3) / 1e9,
                            "frac_of_peak": alg_bytes * args.steps / (ms / 1e3) / 1e9 / peak,
                            "frac_of_8TBps": alg_bytes * args.steps / (ms / 1e3) / 1e9 / 8000.0,
                            "formula": "round0 n(41+24p0) + sum_k (8 L_k + (44+24 p_k) S_k + 32 B_k): L_k suffixes walked (two label reads each), "
                                       "S_k of them sorted, B_k refined in the bag (DESIGN.md section 2)"},
        },
        "rounds": r["rounds"],
    }
    if cpu_baseline is not None:
        out["cpu_baseline"] = cpu_baseline
    out.update(extras)
    print("\n" + json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_part_main(cx: Ctx):
    """--workload part_4G as the main line: value = build MB/s of the whole 4 GiB text (strong scaling over N)."""
    args = cx.args
    sampler = ClockSampler(cx.local_rank)
    sampler.start()
    res = bench_part_4g(cx, steps=max(1, args.steps))
    clocks = sampler.stop()
    gate = parity_gate(cx) if cx.world > 1 and args.gate else None
    if cx.rank == 0:
        b = res["build"]
        out = {"metric": "SA build MB/s", "value": b["device_only"]["value"], "unit": "MB/s", "n_gpus": cx.world, "steps": b["steps"],
               "warmup": 1, "ms_per_step": b["device_only"]["ms_per_step"], "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "u64 keys / u32 indices", "data": "synthetic", "config": config_of("part_4G", cx.world),
               "clocks": clocks, "gpu_launches": None,
               "e2e": {"value": b["value"], "unit": "MB/s", "h2d_bytes_per_step": b["h2d_bytes_per_step"], "d2h_bytes_per_step": 0,
                       "ms_per_step": b["ms_per_step"], "api": res["api"]},
               "part_4G": res}
        if gate:
            out["parity_gate"] = gate
        print("\n" + json.dumps(out), flush=True)
    if cx.world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rep_1G")
    ap.add_argument("--no-queries", dest="queries", action="store_false")
    ap.add_argument("--no-part", dest="part", action="store_false", help="skip the part_4G leg (BASELINE configs[3])")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip rand_256M / acgt_4M / LCP")
    ap.add_argument("--no-gate", dest="gate", action="store_false", help="skip the multi-GPU parity gate (N > 1)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-full-once", dest="full_once", action="store_false", help="reference arm: skip the un-sampled full-size build")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false", help="profiling runs only: skip the host-pointer leg")
    ap.add_argument("--only-build", action="store_true", help="profiling runs: no e2e / extras / part / queries / gate / cpu baseline")
    args = ap.parse_args()
    if args.only_build:
        args.e2e = args.extras = args.part = args.queries = args.gate = args.cpu_baseline = False
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
