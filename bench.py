#!/usr/bin/env python
"""bench.py -- embed+extract throughput of the Gaussian-Shading codec hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, libgswm.so)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # CPU arm: the reference's algorithm on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1, one rank per GPU

One "step" = one pass of the hot path over one batch: embed B latents (K2) and extract B latents
(K3) -- BASELINE config[1] + config[2]: B = 4096 SD-2.1 latents (4x64x64), 256-bit message, shared
default key/nonce, message 'lthero', extraction input = embedded latents + sigma*N(0,1), sigma = 0.325.
`value` = pairs (latent embedded AND latent decoded) per second with inputs resident in HBM;
`e2e` = the same through the host-buffer C ABI (gswm_pipe_*), PCIe copies inside the timed region.
Weak scaling: every rank processes its own B latents (global latent index = rank * B + b).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "a-watermark-for-diffusion-models_b200"))
sys.path.insert(0, ROOT)

METRIC = "embed+extract latents/s (4x64x64, 256-bit); bit-exact decode"
UNIT = "latents/s"
SHAPES = {"sd21": (4, 64, 64), "sdxl": (4, 128, 128)}
SIGMA = 0.325
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback
NCU_SUMMARY = {"embed_kernel": "r02h_embed_shared_ncu_summary.json", "extract_kernel": "r02h_extract_f32_ncu_summary.json"}
E2E_CHUNK = 4096              # latents per pipe chunk.  The kernels take ~0.1 ms of a ~5.5 ms PCIe-bound step, so there is nothing to
                              # gain from overlapping them with the copies, while every extra copy costs ~65 us when both PCIe
                              # directions are busy (tools/e2ebench.py: 7.4 / 6.5 / 6.0 / 5.65 ms at 64 / 256 / 1024 / 4096 latents)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="latents per GPU per step")
    ap.add_argument("--shape", default="sd21", choices=sorted(SHAPES))
    ap.add_argument("--msg-bits", type=int, default=256)
    ap.add_argument("--per-latent-keys", action="store_true", help="BASELINE config[4]: distinct key/nonce/message per latent")
    ap.add_argument("--total-latents", type=int, default=None,
                    help="strong scaling: this many latents in total, sharded over the ranks (BASELINE configs[3]: --shape sdxl "
                         "--total-latents 65536; configs[4]: --per-latent-keys --total-latents 1048576)")
    ap.add_argument("--chunk-latents", type=int, default=None, help="latents per kernel launch when a shard is processed in chunks")
    ap.add_argument("--sustain-seconds", type=float, default=1.2, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--no-issue-rates", action="store_true")
    ap.add_argument("--collective", default="auto", choices=["auto", "gswm", "nccl"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline sample duration")
    ap.add_argument("--ref-step-seconds", type=float, default=None, help="--impl reference: CPU seconds per step (default: sized so the run ends within ~2 minutes)")
    return ap.parse_args()


def workload_name(args):
    c, h, w = SHAPES[args.shape]
    keys = "per-latent key/nonce/message" if args.per_latent_keys else "shared default key/nonce, message 'lthero'"
    if getattr(args, "total_latents", None):
        which = "configs[4]" if args.per_latent_keys else "configs[3]"
        return (f"BASELINE {which}: {args.total_latents} latents ({c}x{h}x{w}) in total, sharded over the GPUs: every GPU embeds its shard "
                f"(per-sample noise) and extracts its shard of noisy latents (sigma={SIGMA}; every message must decode exactly), "
                f"{args.msg_bits}-bit message, {keys}")
    return (f"BASELINE configs[1]+[2]: embed {args.batch} latents ({c}x{h}x{w}, per-sample noise) + extract {args.batch} "
            f"noisy latents (sigma={SIGMA}), {args.msg_bits}-bit message, {keys}")


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(job):
    """One worker: `pairs` embed+extract pairs of the oracle port (numpy/scipy/cryptography), returns seconds."""
    from oracle import gs_oracle as O
    pairs, n, msg_bits, seed = job
    key, nonce = bytes.fromhex(O.DEFAULT_KEY_HEX), bytes.fromhex(O.DEFAULT_NONCE_HEX)
    rs = np.random.RandomState(seed)
    noise = np.float32(SIGMA) * rs.standard_normal(n).astype(np.float32)   # synthetic perturbation, made once
    ks = O.chacha20_keystream_lib                                           # the reference's own OpenSSL call
    ok = 0
    t0 = time.perf_counter()
    for _ in range(pairs):
        u = rs.uniform(size=n)                                             # gs_insert.py:62
        z = O.embed("lthero", key, nonce, u, msg_bits, keystream=ks).astype(np.float32)   # gs_insert.py:23-66 + .float()
        bits = O.recover_message_bits(z + noise, key, nonce, msg_bits, keystream=ks)       # extract.py:72-101
        ok += int(O.bits_to_bytes(bits)[:6] == b"lthero")
    return time.perf_counter() - t0, ok


def cpu_pool(cores):
    """One worker process per host core, started (and warmed: imports done) once, so that a step times the reference's
    arithmetic and not process start-up."""
    import multiprocessing as mp

    if cores == 1:
        return None
    pool = mp.get_context("fork").Pool(cores)
    pool.map(_cpu_worker, [(1, 16384, 256, i) for i in range(cores)])
    return pool


def cpu_pairs_per_second(n_elems, msg_bits, pairs_per_worker, cores, pool=None):
    """All `cores` workers run `pairs_per_worker` pairs concurrently; returns (pairs/s, total pairs, all decoded ok)."""
    jobs = [(pairs_per_worker, n_elems, msg_bits, 1000 + i) for i in range(cores)]
    own = pool is None and cores > 1
    if own:
        pool = cpu_pool(cores)
    t0 = time.perf_counter()
    res = [_cpu_worker(jobs[0])] if cores == 1 else pool.map(_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
    total = pairs_per_worker * cores
    return total / wall, total, sum(r[1] for r in res) == total


def calibrate_cpu(n_elems, msg_bits):
    """Seconds per pair on one core, measured after a warm-up call (the first call pays the scipy / cryptography imports)."""
    _cpu_worker((2, n_elems, msg_bits, 1))
    t, _ = _cpu_worker((16, n_elems, msg_bits, 1))
    return t / 16


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = int(np.prod(SHAPES[args.shape]))
    cores = len(os.sched_getaffinity(0))
    per_pair = calibrate_cpu(n, args.msg_bits)
    # each step: a bounded sample so that warmup + steps stays within ~2 minutes
    budget_per_step = args.ref_step_seconds or min(6.0, 100.0 / max(1, args.steps + args.warmup))
    ppw = max(2, int(budget_per_step / per_pair))
    pool = cpu_pool(cores)
    for _ in range(args.warmup):
        cpu_pairs_per_second(n, args.msg_bits, max(1, ppw // 4), cores, pool)
    t0 = time.perf_counter()
    total = 0
    ok = True
    for _ in range(args.steps):
        _, tp, o = cpu_pairs_per_second(n, args.msg_bits, ppw, cores, pool)
        total += tp
        ok &= o
    wall = time.perf_counter() - t0
    if pool is not None:
        pool.close()
    value = total / wall
    sample = f"{ppw * cores} embed+extract pairs per step ({ppw} per core x {cores} processes), vectorised numpy/scipy/cryptography port of the reference"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample, "decoded_ok": bool(ok)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def wait_ready(self, timeout=5.0):
        """nvidia-smi takes a few hundred ms to print its first line: block until it is streaming, so that the timed
        region (tens of ms) is sampled from its first millisecond."""
        t0 = time.perf_counter()
        while self.proc is not None and time.perf_counter() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return True
            except OSError:
                pass
            time.sleep(0.02)
        return False

    def stop(self, window=None):
        """Median SM clock / throttle reasons over the samples whose timestamp lies inside `window` (datetime pair: the
        timed region and the per-kernel bursts after it); all samples if none falls inside."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        import datetime as dt
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = dt.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                    rows.append((ts, float(f[1]), float(f[2]), float(f[3]), [nm for nm, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
                except ValueError:
                    continue
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        self.rows = rows
        return self.window_stats(window)

    def window_stats(self, window):
        rows = getattr(self, "rows", None) or []
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if window and window[0] <= r[0] <= window[1]]
        near = []
        if window and not inside:                      # a window shorter than the sampling period: the two samples nearest to it
            mid = window[0] + (window[1] - window[0]) / 2
            near = sorted(rows, key=lambda r: abs((r[0] - mid).total_seconds()))[:2]
        use = inside or near or rows
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": float(max(r[2] for r in use)),
                "reasons": sorted({nm for r in use for nm in r[4]}), "samples": len(use), "samples_total": len(rows),
                "window": "timed region + per-kernel bursts" if inside else
                          ("the two samples nearest to the timed region (it is shorter than the sampling period)" if near else "whole run"),
                "power_w_max": max(r[3] for r in use)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram bytes per launch from the committed ncu capture, if profiles/ncu_traffic.json has it."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:  # noqa: BLE001
        return None


def gpu_local_cpus(torch, index):
    """CPUs of the NUMA node the GPU's PCIe slot hangs off, intersected with this process's affinity (empty set if unknown)."""
    try:
        pr = torch.cuda.get_device_properties(index)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return set()
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus & os.sched_getaffinity(0)
    except Exception:  # noqa: BLE001
        return set()


def ncu_pipe_summary(kernel):
    """Issue / pipe utilisation of `kernel` from the committed ncu --set full capture (profiles/), for the roofline block:
    the embed kernel is bound by the FMA pipes, not by HBM, and these are the numbers that say so."""
    try:
        with open(os.path.join(ROOT, "profiles", NCU_SUMMARY[kernel])) as f:
            k = json.load(f)["kernels"][0]
        pick = {"issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "fma_heavy_pipe_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
                "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
                "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "duration_us_under_ncu": "gpu__time_duration.sum"}
        out = {a: round(float(k[b]["value"]), 2) for a, b in pick.items()}
        out["source"] = f"profiles/{NCU_SUMMARY[kernel]} (cold, serialised launch)"
        return out
    except Exception:  # noqa: BLE001
        return None


def gpu_local_cpus_nvml(index):
    """CPU affinity of GPU `index` as NVML reports it (the ideal CPUs for its PCIe root), intersected with this process's
    affinity; empty set if unknown.  /sys's numa_node reads -1 inside these containers, NVML still knows."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        return cpus & os.sched_getaffinity(0)
    except Exception:  # noqa: BLE001
        return set()


def verbatim_reference_note():
    """BASELINE config[0] timed on the unmodified reference in the build container (tools/verbatim_reference_timing.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_verbatim_reference.json")) as f:
            v = json.load(f)
        return ("the UNMODIFIED reference (gs_insert.py:8-75 + extract.py:72-110), 1 core, build container, median of %d: %.2f s embed + "
                "%.2f s extract per latent = %.3f pairs/s (profiles/r02_verbatim_reference.json)"
                % (v["repeats"], v["embed_s_median"], v["extract_s_median"], v["pairs_per_s"]))
    except Exception:  # noqa: BLE001
        return None


def embed_issue_profile():
    """Executed warp instructions per latent element of the embed kernel, from the committed ncu capture of this build
    (smsp__inst_executed.sum / elements): a property of the binary, not of the run."""
    for name in ("r02h_embed_shared_ncu_summary.json", "r01h_embed_ncu_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                k = json.load(f)["kernels"][0]
            inst = float(k["smsp__inst_executed.sum"]["value"])
            return inst / (4096 * 16384), f"profiles/{name}"
        except Exception:  # noqa: BLE001
            continue
    return None, None


def run_gpu_arm(args):
    # stdout carries the one JSON line only: libraries that printf to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import ctypes as C
    import datetime as _dt

    import torch
    import torch.distributed as dist

    import gswm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: gswm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        gswm.build()                                  # no-op when libgswm.so is up to date; there is no fallback path
    if world > 1:
        dist.barrier()
    lib = gswm._lib.lib()
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get('BENCH_NO_SAMPLER'):
        sampler.start()                               # started early: it needs a few hundred ms before its first line

    from gswm.codec import _DeviceJob
    shape = SHAPES[args.shape]
    n = int(np.prod(shape))
    L = args.msg_bits
    strong = args.total_latents is not None
    if strong:                                         # BASELINE configs[3] / [4]: a fixed total, sharded over the ranks
        lo, hi = gswm.sharding.shard_range(args.total_latents, rank, world)
        B, first = hi - lo, lo
    else:                                              # weak scaling: every rank its own B latents
        B, first = args.batch, rank * args.batch
    # Both sides of a step stay RESIDENT when they fit (2 x B x n x 4 bytes <= 150 GB of the 180 GB): one launch per kernel, the
    # two co-scheduled, exactly like the default workload.  A shard that does not fit is processed as a round trip in chunks.
    fits = 2 * B * n * 4 <= 150e9 and not args.chunk_latents
    chunk = B if fits else min(B, args.chunk_latents or max(1, (1 << 30) // (n * 4)))   # latents per launch
    pipeline = chunk < B
    n_chunks = (B + chunk - 1) // chunk
    if args.per_latent_keys:
        rs = np.random.RandomState(2025)               # keys / nonces / messages from RandomState(2025).bytes
        tot = args.total_latents if strong else B * world
        allk = np.frombuffer(rs.bytes(32 * tot), np.uint8).reshape(tot, 32)
        alln = np.frombuffer(rs.bytes(16 * tot), np.uint8).reshape(tot, 16)
        allm = np.frombuffer(rs.bytes((L // 8) * tot), np.uint8).reshape(tot, L // 8)
        km = gswm.KeyMaterial.make(allk[first:first + B], alln[first:first + B], allm[first:first + B], L)
    else:
        km = gswm.KeyMaterial.make(bytes.fromhex(gswm.DEFAULT_KEY_HEX), bytes.fromhex(gswm.DEFAULT_NONCE_HEX),
                                   gswm.pad_message("lthero", L // 8), L)
    seed = 0x5EED
    NC = gswm._lib.N_COUNTERS

    # ---- resident buffers -------------------------------------------------------------------------------------------
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    xstream = torch.cuda.Stream(dev)
    xp = xstream.cuda_stream
    dj_all = _DeviceJob(km, B, n, dev)                  # key material on the device (all of this rank's rows)
    row_k, row_n, row_m = (32, 16, L // 8) if km.per_latent else (0, 0, 0)

    def chunk_job(c):
        """ctypes job for chunk c of this rank's shard (per-latent rows are offset into the resident key arrays)."""
        f0 = c * chunk
        nb = min(chunk, B - f0)
        j = dj_all.job
        return gswm._lib.Job(nb, n, L, j.flags, j.keys + f0 * row_k, j.nonces + f0 * row_n, j.msgs + f0 * row_m), f0, nb

    jobs = [chunk_job(c) for c in range(n_chunks)]
    msgs = torch.empty((chunk, L // 8), dtype=torch.uint8, device=dev)
    matched = torch.empty((chunk,), dtype=torch.int32, device=dev)
    counters = torch.zeros((NC,), dtype=torch.int64, device=dev)
    reduced = torch.zeros((NC,), dtype=torch.int64, device=dev)
    rendezvous = torch.zeros((1,), dtype=torch.int64, device=dev)
    zbuf = [torch.empty((chunk, *shape), dtype=torch.float32, device=dev) for _ in range(2 if pipeline else 1)]
    z = zbuf[0]
    if not pipeline:
        # BASELINE configs[1]+[2]: the extract side reads its own resident batch of noisy latents (sigma = 0.325)
        gswm._lib.check(lib.gswm_embed(C.byref(jobs[0][0]), seed, 0, first, z.data_ptr(), sp), "gswm_embed")
        g = torch.Generator(dev).manual_seed(99 + rank)
        z_noisy = z.clone()
        flat = z_noisy.view(-1)
        for o in range(0, flat.numel(), 1 << 28):                  # 1 GiB slices: no batch-sized temporaries (the batch may be 69 GB)
            sl = flat[o:o + (1 << 28)]
            sl.add_(torch.randn(sl.shape, device=dev, generator=g), alpha=SIGMA)
    else:
        z_noisy = None

    # ---- the path's one collective: the sum of the bit-match counters over the ranks ------------------------------------
    comm, collective = None, "none (1 GPU)"
    if world > 1:
        collective = "NCCL all-reduce through torch.distributed (c10d)"
        if args.collective != "nccl":
            try:
                comm = gswm.Comm(dev)
                collective = ("gswm_comm: fused into the last extract launch of the timed region (gswm_extract_allreduce) -- "
                              "per-rank mailboxes in HBM, peers store into them over NVLink (CUDA IPC mapping), no host round trip")
            except Exception as e:  # noqa: BLE001
                if args.collective == "gswm":
                    raise
                collective += f" [gswm_comm unavailable: {e}]"

    def launch_embed(c, buf, st):
        job, f0, _ = jobs[c]
        gswm._lib.check(lib.gswm_embed(C.byref(job), seed, 0, first + f0, buf.data_ptr(), st), "gswm_embed")

    def launch_extract(c, src, st, fused=False):
        job = jobs[c][0]
        if fused:
            gswm._lib.check(lib.gswm_extract_allreduce(C.byref(job), src.data_ptr(), 0, msgs.data_ptr(), None, matched.data_ptr(),
                                                       None, counters.data_ptr(), comm.handle, reduced.data_ptr(), st),
                            "gswm_extract_allreduce")
        else:
            gswm._lib.check(lib.gswm_extract(C.byref(job), src.data_ptr(), 0, msgs.data_ptr(), None, matched.data_ptr(), None,
                                             counters.data_ptr(), st), "gswm_extract")

    # The two halves of a step are bound by different resources (embed: the SMs' FMA pipes, extract: HBM), so they are
    # launched on two streams and share the SMs.  Default workload: they are independent (embed writes z, extract reads
    # z_noisy).  Sharded-total workloads (configs[3], [4]): a ROUND TRIP in chunks -- extract decodes the chunk embed just
    # wrote (two buffers: embed of chunk k+1 overlaps extract of chunk k), so every message must come back exactly.
    ev_e = [torch.cuda.Event() for _ in range(2)]
    ev_x = [torch.cuda.Event() for _ in range(2)]

    def step(serial=False, last=False):
        if not pipeline:
            launch_embed(0, z, sp)
            launch_extract(0, z_noisy, sp if serial else xp, fused=last and comm is not None)
            return
        serial = True          # measured: the chunked round trip gains nothing from two streams (embed of chunk c + 1 is resident
                               # before extract of chunk c becomes eligible; tools/benchq.py sweep, profiles/r02_roundtrip_schedules.txt)
        for c in range(n_chunks):
            b_ = c & 1
            if c >= 2 and not serial:
                stream.wait_event(ev_x[b_])                    # the buffer's previous reader is done
            launch_embed(c, zbuf[b_], sp)
            if serial:
                launch_extract(c, zbuf[b_], sp, fused=last and c == n_chunks - 1 and comm is not None)
            else:
                ev_e[b_].record(stream)
                xstream.wait_event(ev_e[b_])
                launch_extract(c, zbuf[b_], xp, fused=last and c == n_chunks - 1 and comm is not None)
                ev_x[b_].record(xstream)

    def timed(n_steps, serial=False, reduce=True):
        """Device time of n_steps steps: fork the extract stream off the launch stream, join it back before the end event.
        The path's only collective -- the FINAL sum of the bit-match counters the extract kernels have been accumulating
        (north_star) -- is part of the timed region: fused into the last extract launch (gswm_comm), or one c10d all-reduce."""
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        joined = torch.cuda.Event()
        if comm is not None and reduce:
            # device-side rendezvous AHEAD of the start event: the host barrier releases the ranks tens of microseconds apart, and
            # with a collective at the end of the region every rank would be charged the latest rank's late start
            comm.allreduce_counters(rendezvous)
        t0.record(stream)
        xstream.wait_event(t0)
        for i in range(n_steps):
            step(serial, last=reduce and world > 1 and i == n_steps - 1)
        joined.record(xstream)
        stream.wait_event(joined)
        if world > 1 and reduce and comm is None:
            reduced.copy_(counters)
            dist.all_reduce(reduced, op=dist.ReduceOp.SUM)
        t1.record(stream)
        return t0, t1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(3, args.warmup)
    timed(warm)                                        # warm-up steps (and the communicator's first exchange)
    barrier()
    if rank == 0:
        sampler.wait_ready()
    launches0 = gswm.launch_count()
    load_begin = _dt.datetime.now()
    barrier()
    t_begin, t_end = timed(args.steps)
    barrier()
    launches = gswm.launch_count() - launches0
    total_ms = t_begin.elapsed_time(t_end)
    n_steps_total = warm + args.steps
    torch.cuda.synchronize(dev)
    final = (reduced if world > 1 else counters).cpu().numpy().tolist()   # accumulated over every step so far
    tot_lat = (args.total_latents if strong else B * world) * n_steps_total
    # every message of every step decodes exactly (sigma = 0.325 / noise-free round trip), nothing was rejected
    exact = final[2] == final[3] == tot_lat and final[0] == final[1] == tot_lat * L and final[4] == final[5] == 0
    if comm is not None and comm.status() != 0:
        exact = False

    # Per-kernel durations for the roofline block: the same launches, each kernel back to back `inst_steps` times
    # between one pair of events (so no event record or dependent-launch gap sits inside the measured interval).
    inst_steps = max(1, min(args.steps, 200) // n_chunks) if pipeline else min(args.steps, 200)
    barrier()
    s_begin, s_end = timed(max(1, min(args.steps, 200 // n_chunks if pipeline else 200)), serial=True, reduce=False)
    barrier()
    serial_ms = s_begin.elapsed_time(s_end) / max(1, min(args.steps, 200 // n_chunks if pipeline else 200))

    def burst(launch):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(inst_steps):
            launch()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) / inst_steps

    src = z_noisy if not pipeline else zbuf[0]
    if pipeline:
        launch_embed(0, zbuf[0], sp)
    burst_sleep = float(os.environ.get("BENCH_BURST_SLEEP_MS", "0")) * 1e-3
    n_bursts = int(os.environ.get("BENCH_BURSTS", "5"))   # the GPU toggles between two states ~4 % apart on a 100 ms scale
                                                          # (profiles/r02_burst_states.txt): five bursts, best and median reported

    def bursts(launch):
        out = []
        for _ in range(n_bursts):
            if burst_sleep:
                time.sleep(burst_sleep)
            out.append(burst(launch))
        return out

    eb, xb = bursts(lambda: launch_embed(0, zbuf[0], sp)), bursts(lambda: launch_extract(0, src, sp))
    if os.environ.get("BENCH_BURST_TRACE") and rank == 0:
        print("bursts embed", [round(x * 1e3, 2) for x in eb], "extract", [round(x * 1e3, 2) for x in xb], file=sys.stderr)
    embed_bursts, extract_bursts = sorted(eb), sorted(xb)
    embed_ms, extract_ms = embed_bursts[0], extract_bursts[0]
    embed_med, extract_med = embed_bursts[len(eb) // 2], extract_bursts[len(xb) // 2]
    burst_latents = jobs[0][2]
    kernels_end = _dt.datetime.now()

    # ---- issue-rate denominators, same run, same GPU ---------------------------------------------------------
    issue = None
    if rank == 0 and not args.no_issue_rates:
        issue = {}
        for name, kind in gswm._lib.ISSUE_KINDS.items():
            r_, g_ = C.c_double(), C.c_double()
            if lib.gswm_debug_issue_rate(kind, C.byref(r_), C.byref(g_)) == 0:
                issue[name] = {"warp_inst_per_clk_per_smsp": round(r_.value, 4), "sm_ghz": round(g_.value, 3)}
    # ---- sustained leg: >= --sustain-seconds of back-to-back steps, whatever --steps was --------------------------------
    sustained = None
    if args.sustain_seconds > 0:
        per_step = total_ms / args.steps
        n_sus = int(min(200000, max(args.steps, args.sustain_seconds * 1e3 / max(per_step, 1e-3))))
        barrier()
        sus_begin = _dt.datetime.now()
        s0, s1 = timed(n_sus, reduce=False)
        barrier()
        sus_end = _dt.datetime.now()
        sus_ms = s0.elapsed_time(s1) / n_sus
        sustained = {"steps": n_sus, "ms_per_step": sus_ms, "window": (sus_begin, sus_end)}

    barrier()
    clocks = sampler.stop((load_begin, kernels_end)) if rank == 0 else None     # timed region + per-kernel bursts; the sustained leg has its own window
    if rank == 0 and sustained is not None:
        sc = sampler.window_stats(sustained.pop("window"))
        sustained.update({"sm_mhz": sc.get("sm_mhz"), "reasons": sc.get("reasons"), "power_w_max": sc.get("power_w_max"),
                          "clock_samples": sc.get("samples")})
    elif sustained is not None:
        sustained.pop("window")
    tm = torch.tensor([total_ms, embed_ms, extract_ms, serial_ms, sustained["ms_per_step"] if sustained else 0.0],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, embed_ms, extract_ms, serial_ms, sus_ms = tm.cpu().tolist()
    ms_per_step = total_ms / args.steps
    lat_per_step = args.total_latents if strong else B * world
    value = lat_per_step / (ms_per_step * 1e-3)
    if sustained is not None:
        sustained["ms_per_step"] = sus_ms
        sustained["value"] = lat_per_step / (sus_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI pipe (PCIe inside the timed region) -------------------------------
    # Two pipes (one per direction) driven from two host threads: the embed side's D2H and the extract side's H2D
    # use the two directions of the PCIe link at the same time (ctypes releases the GIL during the calls).
    # host buffers and the pipe's staging live on the GPU's own NUMA node (first touch happens under this affinity)
    affinity0 = os.sched_getaffinity(0)
    local_cpus = gpu_local_cpus(torch, local) or gpu_local_cpus_nvml(local)
    if local_cpus:
        os.sched_setaffinity(0, local_cpus)
    Be = min(B, max(1, (1 << 30) // (n * 4)))          # e2e batch: the rank's shard, bounded to 1 GiB per direction
    kme = km if not km.per_latent else gswm.KeyMaterial.make(km.keys[:Be], km.nonces[:Be], km.msgs[:Be], L)
    pipe_e = gswm.HostPipe(local, max_elems=n, chunk_latents=min(E2E_CHUNK, Be))
    pipe_x = gswm.HostPipe(local, max_elems=n, chunk_latents=min(E2E_CHUNK, Be))
    h_out = torch.empty((Be, *shape), dtype=torch.float32).pin_memory()
    if pipeline:
        launch_embed(0, zbuf[0], sp)
        torch.cuda.synchronize(dev)
        h_in = zbuf[0][:Be].cpu().pin_memory()
    else:
        h_in = z_noisy[:Be].cpu().pin_memory()
    e2e_steps = max(1, args.e2e_steps)
    box = {}

    from concurrent.futures import ThreadPoolExecutor
    embed_side = ThreadPoolExecutor(max_workers=1)      # one long-lived host thread for the D2H direction (no thread start per step)

    def e2e_step():
        fut = embed_side.submit(pipe_e.embed, h_out, kme, seed, 0, first)
        box["x"] = pipe_x.extract(h_in, kme)
        fut.result()
        return box["x"]

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, _, _, h_cnt, h_flags = e2e_step()
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_checks = {"all_messages_exact": int(h_cnt[2]) == Be and int(h_cnt[0]) == Be * L, "none_rejected": not bool(h_flags.any())}
    # the embedded latents that came back over PCIe must be the ones the resident path produced
    chk = torch.empty((min(8, Be), *shape), dtype=torch.float32, device=dev)
    cj = gswm._lib.Job(chk.shape[0], n, L, dj_all.job.flags, dj_all.job.keys, dj_all.job.nonces, dj_all.job.msgs)
    gswm._lib.check(lib.gswm_embed(C.byref(cj), seed, 0, first, chk.data_ptr(), sp), "gswm_embed")
    e2e_checks["embedded_latents_equal_resident_path"] = bool(torch.equal(h_out[:chk.shape[0]], chk.cpu()))
    # the box's ceiling for exactly this traffic: plain pinned cudaMemcpyAsync of the same bytes, both directions at once,
    # every rank at the same time -- no kernels, no pipe
    d_a = torch.empty((Be, n), dtype=torch.float32, device=dev)
    d_b = torch.empty((Be, n), dtype=torch.float32, device=dev)
    cs1, cs2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy_both():
        with torch.cuda.stream(cs1):
            d_a.copy_(h_in.reshape(Be, n), non_blocking=True)
        with torch.cuda.stream(cs2):
            h_out.reshape(Be, n).copy_(d_b, non_blocking=True)

    copy_both()
    barrier()
    tc = time.perf_counter()
    for _ in range(e2e_steps):
        copy_both()
    torch.cuda.synchronize(dev)
    ceil_s = time.perf_counter() - tc
    te = torch.tensor([e2e_s, ceil_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, ceil_s = te.cpu().tolist()
    e2e_value = Be * world * e2e_steps / e2e_s
    e2e_ok = all(e2e_checks.values())
    key_bytes = kme.keys.nbytes + kme.nonces.nbytes + (kme.msgs.nbytes if kme.msgs is not None else 0)
    h2d = Be * n * 4 + key_bytes * 2                  # latents in + key material once per pipe call
    d2h = Be * n * 4 + Be * (L // 8) + Be * 4 + Be + 8 * NC
    embed_side.shutdown()
    pipe_e.close()
    pipe_x.close()
    os.sched_setaffinity(0, affinity0)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if comm is not None:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    lat_bytes = burst_latents * n * 4                  # algorithmic bytes of ONE launch of either kernel (4 B per element)
    step_bytes = 2 * B * n * 4                         # ... and of one step on one GPU
    k_embed = {"ms": embed_ms, "ms_median": embed_med, "GBps": lat_bytes / (embed_ms * 1e-3) / 1e9, "algorithmic_bytes": lat_bytes}
    k_extract = {"ms": extract_ms, "ms_median": extract_med, "GBps": lat_bytes / (extract_ms * 1e-3) / 1e9, "algorithmic_bytes": lat_bytes}
    dom_name, dom = ("embed_kernel", k_embed) if embed_ms >= extract_ms else ("extract_kernel", k_extract)
    # issue view of the embed kernel: executed warp instructions (ncu, per element) x elements / time, against the measured
    # one-instruction-per-clock peak and against the pipe model (sum over instruction classes of count / measured class rate)
    ipe, ipe_src = embed_issue_profile()
    issue_view = None
    if issue and ipe and "FFMA(imm)" in issue:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # clock: the one the microbenchmark measured IN its kernel (clock64 ticks / elapsed time), right after the bursts.  nvidia-smi
        # keeps reporting 1965 MHz while the SMs really run at 1.83 .. 1.92 GHz, and the embed kernel's time follows the real clock
        # to the percent (tools/clock_vs_embed.py: time x clock = 101.8 k cycles per launch, profiles/r02_clock_vs_embed.json)
        ghz = issue["FFMA(imm)"]["sm_ghz"]
        peak_wi = issue["FFMA(imm)"]["warp_inst_per_clk_per_smsp"] * 4 * sms * ghz * 1e9      # warp instructions / s, whole GPU
        ach_wi = ipe * burst_latents * n / (embed_ms * 1e-3)
        issue_view = {"achieved": ach_wi, "peak": peak_wi, "unit": "warp-inst/s", "frac": ach_wi / peak_wi,
                      "inst_per_element": round(ipe, 3), "inst_source": ipe_src,
                      "peak_source": "gswm_debug_issue_rate(FFMA imm) in this run: %.3f warp-inst/clk/SMSP x 4 x %d SMs x %.3f GHz (SM clock measured in the microbenchmark kernel)"
                                     % (issue["FFMA(imm)"]["warp_inst_per_clk_per_smsp"], sms, ghz)}
    hbm_frac = dom["GBps"] / peak
    bound = "hbm"
    if dom_name == "embed_kernel" and issue_view and issue_view["frac"] > hbm_frac:
        bound = "issue"
    roofline = {"bound": bound, "kernel": dom_name, "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                "frac": hbm_frac, "traffic": ncu_traffic(dom_name), "peak_source": peak_src,
                "issue": issue_view, "issue_rates": issue,
                "kernels": {"embed_kernel": dict(k_embed, frac=k_embed["GBps"] / peak, ncu=ncu_pipe_summary("embed_kernel")),
                            "extract_kernel": dict(k_extract, frac=k_extract["GBps"] / peak, ncu=ncu_pipe_summary("extract_kernel"))},
                # the whole co-scheduled step against the same peak: both kernels' algorithmic bytes / step time
                "step": {"ms": ms_per_step, "GBps": step_bytes / (ms_per_step * 1e-3) / 1e9, "algorithmic_bytes": step_bytes,
                         "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak, "serial_ms": serial_ms}}
    if sustained is not None:
        sustained["GBps"] = step_bytes / (sustained["ms_per_step"] * 1e-3) / 1e9
        sustained["frac"] = sustained["GBps"] / peak

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        per_pair = calibrate_cpu(n, L)
        ppw = max(2, int(args.cpu_seconds / per_pair))
        v, tp, ok = cpu_pairs_per_second(n, L, ppw, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{tp} embed+extract pairs ({ppw} per core x {cores} processes) of the same latent shape, vectorised "
                         f"numpy/scipy/cryptography port of the reference (oracle/gs_oracle.py); decoded_ok={ok}",
               "note": verbatim_reference_note()}

    c, h, w = shape
    ceil_gbs = Be * n * 4 * e2e_steps / ceil_s / 1e9
    e2e_gbs = Be * n * 4 * e2e_steps / e2e_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "latents_per_gpu": B, "latent_shape": [c, h, w], "msg_bits": L,
                   "latents_per_launch": burst_latents, "launches_per_step_per_kernel": n_chunks,
                   "l2": "inputs larger than L2 (%d MB streamed per step vs 126 MB L2)" % (step_bytes // 1000000) if step_bytes > 2 * 126e6 else
                         "WARNING: working set fits L2", "timing": "CUDA events on the launch stream (the extract stream is forked after the start event and joined before the end event), max over ranks; per-kernel durations: best (ms) and median (ms_median) of 5 bursts of %d back-to-back launches, each burst between one event pair" % inst_steps,
                   "schedule": ("chunked round trip on one stream: extract decodes the chunk embed just wrote (noise-free)"
                                if pipeline else
                                "embed and extract of a step run on two CUDA streams and share the SMs (FMA-bound embed next to HBM-bound extract); roofline.step.serial_ms is the same step on one stream"),
                   "collective": collective,
                   "uniform_source": "Philox4x32-%d, 23 bits per element, outermost cell refined to 51 bits (uniforms v4)" % lib.gswm_philox_rounds(),
                   "decode_exact": bool(exact), "counters": final},
        "roofline": roofline, "sustained": sustained, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "decode_exact": bool(e2e_ok), "checks": e2e_checks, "latents_per_gpu": Be,
                "path": "gswm_pipe_embed -> pinned host fp32 and pinned host fp32 -> gswm_pipe_extract, two host threads (one per PCIe direction), %d-latent chunks, 2 slots each" % min(E2E_CHUNK, Be),
                "GBps_each_way_per_gpu": e2e_gbs,
                "pcie_ceiling_GBps_each_way_per_gpu": ceil_gbs,
                "pcie_ceiling": "plain pinned cudaMemcpyAsync of the same bytes, H2D and D2H at once, all %d ranks at the same time, this run" % world,
                "pcie_frac": e2e_gbs / ceil_gbs,
                "numa_local_cpus": len(local_cpus)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
