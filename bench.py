#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configuration.

metric   : decoded samples/s (a sample = one channel-sample), whole job, with PCM GB/s and the HBM roofline beside it
workload : configs[1] -- 1 h 24-bit stereo 96 kHz FLAC, blocksize 4096, LPC order <= 12, adaptive mid/side, decoded to
           one PCM buffer per GPU.  Synthetic: 60 s of seeded integer PCM encoded by corpus/bncorpus.c, tiled x60 at the
           frame level (renumbered, CRC-8/CRC-16 recomputed) -- the construction SURVEY.md 8(d) prescribes.
step     : one pass of the whole pipeline (frame scan + CRC + subframe parse + residual decode + restore + interleave)
           over one 1 h stream per GPU.
value    : inputs already resident in HBM, output left in HBM (bnflac_open_device + bnflac_decode_device).
e2e      : the same stream through the reference-facing call with HOST buffers (bnflac_open_memory + bnflac_decode_all):
           pinned host FLAC bytes -> device, decode, PCM -> pinned host, every step.
N > 1    : one process per GPU (torchrun), weak scaling: every rank decodes its own 1 h stream (sharded by file, no
           collective on the data path); `--scaling strong` instead shards ONE 1 h stream by frame ranges.
--impl reference : the reference's CPU decoder on the host cores (oracle/_ref = the reference's LibFlac.dll hosted by
           oracle/refdll when it runs on this box, else the C port oracle/flac_oracle.c), bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

UNIQUE_SECONDS = 60
WORKLOAD = "cfg2: 1 h 24-bit stereo 96 kHz FLAC, bs 4096, LPC<=12, adaptive mid/side (60 s unique synthetic audio tiled x60)"


def make_stream(seconds: int):
    import pycorpus
    tiles = max(1, seconds // UNIQUE_SECONDS)
    uniq = min(seconds, UNIQUE_SECONDS)
    t0 = time.time()
    s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=uniq, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=tiles, seed=2026)
    s.gen_seconds = time.time() - t0
    return s


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  Sampled in-process through
    NVML every 4 ms (the recipe's `nvidia-smi -lms` needs ~100 ms to produce its first row: a short timed region would end
    up with no sample at all); nvidia-smi is the fallback when pynvml is missing."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.rows = []            # (time, sm_mhz, sm_max_mhz, set(reasons))
        self.proc = None
        self.nvml = None
        self.index = index
        self.stopping = False
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index))
            self._sample_nvml()                      # fails here rather than in the thread if the queries are unsupported
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _sample_nvml(self):
        nv, h = self.nvml
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(h))
        self.rows.append((time.time(), sm, mx, {n for n, b in self.BITS.items() if mask & b}))

    def _poll(self):
        while not self.stopping:
            try:
                self._sample_nvml()
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            try:
                row = (time.time(), float(f[0]), float(f[1]),
                       {n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]) if v.lower().startswith("active")})
            except Exception:
                continue
            self.rows.append(row)

    def stop(self, t0, t1):
        if not self.nvml and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        self.stopping = True
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.15] or list(self.rows)
        sm = [r[1] for r in rows]; mx = [r[2] for r in rows]
        reasons = set().union(*[r[3] for r in rows]) if rows else set()
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ---------------------------------------------------------------------------------------------- CPU baselines
def ref_runnable():
    exe = os.path.join(ROOT, "oracle", "_ref", "refflac")
    dll = os.path.join(ROOT, "oracle", "_ref", "LibFlac.dll")
    if not (os.path.exists(exe) and os.path.exists(dll)):
        return None
    try:
        r = subprocess.run([exe, "probe"], capture_output=True, timeout=10)
        if r.returncode not in (2, 3):   # prints usage (2) when the 32-bit binary executes at all
            return None
    except Exception:
        return None
    return exe, dll


def cpu_decode_rate(sample_flac: bytes, nsamples_all_ch: int, threads: int, reps: int):
    """Decode `sample_flac` `reps` times on each of `threads` host threads; returns (samples/s aggregate, kind, seconds)."""
    ref = ref_runnable()
    if ref:
        exe, dll = ref
        path = os.path.join("/tmp", f"bnflac_cpu_sample_{os.getpid()}.flac")
        with open(path, "wb") as f:
            f.write(sample_flac)
        t0 = time.time()
        procs = [subprocess.Popen([exe, "bench", dll, path, str(reps)], stdout=subprocess.PIPE, text=True) for _ in range(threads)]
        outs = [p.communicate()[0] for p in procs]
        dt = time.time() - t0
        os.unlink(path)
        # per-process decode time as the process itself clocks it (excludes exec + file read)
        per = []
        for o in outs:
            us = [int(l.split("ms=")[1].split("us")[0]) for l in o.splitlines() if l.startswith("ms=")]
            if len(us) == reps:
                per.append(sum(us) / 1e6)
        if len(per) == threads:
            return threads * reps * nsamples_all_ch / max(per), "reference", dt
    import pyoracle
    L = pyoracle.lib()
    need = L.fo_decode(sample_flac, len(sample_flac), None, 0, None, 0, None, None, None, 0, None)
    import ctypes as C
    bufs = [C.create_string_buffer(int(need) + 1) for _ in range(threads)]

    def work(i):
        for _ in range(reps):
            L.fo_decode(sample_flac, len(sample_flac), bufs[i], int(need), None, 0, None, None, None, 0, None)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.time()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    return threads * reps * nsamples_all_ch / dt, "port", dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    s = make_stream(min(args.seconds, UNIQUE_SECONDS))
    n_all = s.total_samples * s.channels
    # calibrate: one step is a bounded sample of the 1 h workload on every host thread -- at most ~8 s of CPU work, and
    # short enough that the whole --steps K --warmup W run stays within ~2.5 minutes whatever K is
    _, _, dt1 = cpu_decode_rate(s.flac, n_all, cores, 1)
    per_step = min(8.0, 150.0 / max(1, args.steps + (1 if args.warmup else 0)))
    reps = max(1, min(200, int(per_step / max(dt1, 1e-3))))
    for _ in range(1 if args.warmup else 0):
        cpu_decode_rate(s.flac, n_all, cores, 1)
    rates, kind, secs = [], "port", 0.0
    for _ in range(args.steps):
        r, kind, dt = cpu_decode_rate(s.flac, n_all, cores, reps)
        rates.append(r); secs += dt
    value = statistics.mean(rates)
    sample = f"{reps} x {UNIQUE_SECONDS} s tile of the cfg2 stream per thread, {cores} threads, per step"
    line = {"metric": "decoded samples/s", "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64",
            "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "pcm_gbps": value * 3 / 1e9}
    print(json.dumps(line), flush=True)


class StdoutToStderr:
    """File descriptor 1 points at stderr inside the block (native libraries that printf are covered too)."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def bind_to_gpu_cpus(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers of the e2e leg
    (first touched by this process) and the copies that read/write them stay on the GPU's NUMA node.  With one process
    per GPU (torchrun) every rank otherwise shares whatever node the launcher happened to start on."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from birdnest.audio_b200 import _abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or _abi.lib().bnflac_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the decode engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with StdoutToStderr():        # NCCL prints its version banner on stdout at communicator creation: stdout carries ONE JSON line
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm, op=dist.ReduceOp.MAX)
            dist.barrier()
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s = make_stream(args.seconds)
    numa = bind_to_gpu_cpus(local)            # after the (multi-threaded) stream generator, before any pinned allocation
    flac_len = len(s.flac)
    strong = args.scaling == "strong" and world > 1
    shard = dict(shard_index=rank, shard_count=world) if strong else {}
    total_samples_all = s.total_samples * s.channels          # per stream
    job_samples = total_samples_all if strong else total_samples_all * world
    pcm_bytes_stream = total_samples_all * 3

    # ---- value: device-resident in, device-resident out -------------------------------------------------
    host_in = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).pin_memory()
    d_in = torch.zeros(flac_len + 256, dtype=torch.uint8, device=dev)
    d_in[:flac_len].copy_(host_in, non_blocking=True)
    d_out = torch.empty(pcm_bytes_stream + 256, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    h = _abi.open_device(d_in.data_ptr(), flac_len, s.flac[:1 << 20], device=local, stream=stream.cuda_stream, keep=d_in, **shard)
    torch.cuda.synchronize()
    written = 0
    for _ in range(args.warmup):
        _, written = h.decode_device(d_out.data_ptr(), d_out.numel())
    # correctness gate on the benchmark's own output (not timed): md5 of the decoded PCM == STREAMINFO md5
    if not strong and args.verify:
        import hashlib
        got = hashlib.md5(d_out[:written].cpu().numpy().tobytes()).digest()
        if got != s.md5:
            raise SystemExit("bench.py: decoded PCM md5 != STREAMINFO md5 -- refusing to report a number")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _abi.lib().bnflac_kernel_launches()
    stage = {"scan": 0.0, "crc": 0.0, "link": 0.0, "parse": 0.0, "decode": 0.0, "total": 0.0}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        _, written = h.decode_device(d_out.data_ptr(), d_out.numel())
        t = h.timing()
        for k in stage:
            stage[k] += getattr(t, k)
    e1.record(stream)
    barrier()
    t1 = time.time()
    launches = _abi.lib().bnflac_kernel_launches() - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / args.steps
    value = job_samples / (ms_per_step / 1e3)
    for k in stage:
        stage[k] /= args.steps
    h.close()

    # ---- e2e: host FLAC bytes -> host PCM through the reference-facing call, copies inside the timed region ----
    host_out = torch.empty(pcm_bytes_stream + 256, dtype=torch.uint8).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    if args.no_e2e:
        e2e_steps = 0

    def e2e_step():
        with _abi.open_memory(host_in, device=local, flags=_abi.OPT_BORROW_INPUT, **shard) as hh:
            return hh.decode_all(host_out)
    nbytes = 0
    for _ in range(2 if e2e_steps else 0):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        nbytes = e2e_step()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    e2e_ms = (w1 - w0) * 1e3 / max(1, e2e_steps) if e2e_steps else float("inf")
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = job_samples / (e2e_ms / 1e3)
    h2d = flac_len if not strong else flac_len // world
    d2h = nbytes

    if rank == 0:
        peak, peak_src = peaks()
        # roofline of the dominant kernel: algorithmic bytes = compressed bytes in + packed PCM bytes out,
        # the units one launch processes = the whole stream (one k_decode launch per step)
        units_bytes = (flac_len + pcm_bytes_stream) if not strong else (flac_len + pcm_bytes_stream) / world
        dom = max(("decode", "parse", "crc", "scan"), key=lambda k: stage[k])
        dom_ms = stage[dom]
        achieved = units_bytes / (dom_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(dom)
        line = {
            "metric": "decoded samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "int32/int64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "seconds_per_stream": args.seconds, "streams": 1 if strong else world,
                       "frames_per_stream": len(s.frame_bs), "compressed_bytes": flac_len, "pcm_bytes": pcm_bytes_stream,
                       "parallelism": f"{'frame-range' if strong else 'file'} shards x{world}, no collective",
                       "host_cpus_bound_to_gpu": numa,
                       "l2": "inputs (1.1 GB) and outputs (2.1 GB) exceed the 126 MB L2; no flush needed"},
            "pcm_gbps": value * 3 / 1e9,
            "pipeline_hbm_frac": (units_bytes * (1 if strong else 1)) / (ms_per_step / 1e3) / 1e9 / peak,
            "stage_ms": stage,
            "roofline": {"bound": "hbm", "kernel": {"decode": "k_decode", "parse": "k_parse", "crc": "k_crc", "scan": "k_scan"}[dom],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": units_bytes},
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "bnflac_open_memory(pinned host FLAC, BORROW_INPUT) + bnflac_decode_all(pinned host PCM): pipelined sub-shards, PCIe both ways inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            try:
                os.sched_setaffinity(0, range(cores))      # undo bind_to_gpu_cpus: the CPU baseline gets every host thread
            except OSError:
                pass
            uniq = make_stream(min(args.seconds, UNIQUE_SECONDS)) if args.seconds > UNIQUE_SECONDS else s
            n_all = uniq.total_samples * uniq.channels
            _, _, dt1 = cpu_decode_rate(uniq.flac, n_all, cores, 1)
            reps = max(4, min(400, int(15.0 / max(dt1, 1e-3))))          # ~15 s of CPU work per host thread
            r, kind, dt = cpu_decode_rate(uniq.flac, n_all, cores, reps)
            line["cpu_baseline"] = {"value": r, "unit": "samples/s", "cores": cores, "kind": kind,
                                    "sample": f"{reps} x {min(args.seconds, UNIQUE_SECONDS)} s tile of the cfg2 stream on each of {cores} threads ({dt:.1f} s wall)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=int, default=3600, help="stream length (default: the named 1 h configuration)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-verify", dest="verify", action="store_false")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
