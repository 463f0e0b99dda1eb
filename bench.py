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
WORKLOAD = "cfg2: 1 h 24-bit stereo 96 kHz FLAC, bs 4096, LPC<=12, adaptive mid/side (59.99 s of unique synthetic audio tiled x60 = 84,360 frames)"


def base_config(args, world, strong=False):
    """the workload description both arms print (the reference arm decodes a bounded sample of the same stream shape)"""
    return {"workload": WORKLOAD, "seconds_per_stream": args.seconds, "streams": 1 if strong else world,
            "parallelism": f"{'frame-range' if strong else 'file'} shards x{world}, no collective",
            "l2": "inputs (1.1 GB) and outputs (2.1 GB) exceed the 126 MB L2; no flush needed"}


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
            "data": "synthetic", "impl": "reference", "config": base_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
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


# ---------------------------------------------------------------------------------------------- corpora of the other configurations
def cfg1_kwargs(scale):
    return dict(ch=2, bps=16, sr=44100, seconds=60, bs=4096, lpc=8, maxpo=5, tile=max(1, int(60 * scale)), seed=2026)


def cfg3_kwargs(scale):       # 600 s: 586 tiles of 12 frames (1.024 s each)
    return dict(ch=8, bps=24, sr=192000, samples=16384 * 12, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, tile=max(1, int(586 * scale)), seed=5)


def cfg5_formats(scale):
    """BASELINE configs[4] as SURVEY.md 8(d) spells it out: 10 h in six formats, each a unique pool tiled at the frame level."""
    t = lambda n: max(1, int(n * scale))
    return [
        ("s16_44k_2ch_lpc8", "3 h 16-bit stereo 44.1 kHz, bs 4096, LPC<=8", dict(ch=2, bps=16, sr=44100, seconds=60, bs=4096, lpc=8, maxpo=5, tile=t(180), seed=31)),
        ("s16_48k_mono_fixed", "2 h 16-bit mono 48 kHz, bs 1152, FIXED", dict(ch=1, bps=16, sr=48000, seconds=60, bs=1152, lpc=0, tile=t(120), seed=32)),
        ("s24_96k_2ch_lpc12", "2 h 24-bit stereo 96 kHz, bs 4096, LPC<=12", dict(ch=2, bps=24, sr=96000, seconds=60, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=t(120), seed=33)),
        ("s24_48k_6ch_special", "1 h 24-bit 6 ch 48 kHz, bs 1152, CONSTANT / VERBATIM / wasted-bits segments", dict(ch=6, bps=24, sr=48000, seconds=30, bs=1152, lpc=8, kind=1, period=1152, tile=t(120), seed=34)),
        ("s24_192k_8ch_lpc32", "1 h 24-bit 8 ch 192 kHz, bs 16384, LPC 32, Rice2 partition order 8", dict(ch=8, bps=24, sr=192000, samples=16384 * 12, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, tile=t(3516), seed=35)),
        ("s16_44k_2ch_var", "1 h 16-bit stereo 44.1 kHz, variable blocksize (0xFFF9)", dict(ch=2, bps=16, sr=44100, seconds=60, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304), tile=t(60), seed=36)),
    ]


def cfg4_pool():
    import pycorpus
    pool = []
    for i in range(200):
        kw = dict(ch=1 + (i & 1), bps=16, sr=44100, seconds=0.5 + (i * 37 % 26) / 10.0, lpc=0 if i % 4 < 2 else 8, seed=1000 + i)
        if i % 10 == 3:
            kw["var"] = (4096, 1152, 4080, 720, 16, 192, 2304)
        else:
            kw["bs"] = (576, 1152, 2304, 4096, 4608)[i % 5]
        pool.append(pycorpus.make(**kw))
    return pool


CFG4_TOTAL_CLIPS = 100_000


def cfg4_clip(pool, k):
    return pool[(7 * k) % len(pool)]


def verify_periodic(torch, d_out, written, tile_dev, start):
    """the decoded PCM of a tiled stream is the unique tile's PCM repeated: d_out[0:written] against the tile from offset `start`"""
    L = tile_dev.numel()
    pos = 0
    while pos < written:
        off = (start + pos) % L
        n = min(L - off, written - pos)
        if not torch.equal(d_out[pos:pos + n], tile_dev[off:off + n]):
            return False
        pos += n
    return True


def device_pass(ctx, flac, out_cap, steps, warmup, shard=None):
    """Device-resident decode of `flac` (host bytes / numpy view; uploaded once, outside the timed region) `steps` times.
    -> dict(ms per step on this rank, stage ms, written bytes, d_out)"""
    torch, _abi, dev, local, stream = ctx
    d_out = torch.empty(int(out_cap) + 256, dtype=torch.uint8, device=dev)
    kw = dict(shard_index=shard[0], shard_count=shard[1]) if shard else {}
    with _abi.open_memory(flac, device=local, stream=stream.cuda_stream, flags=_abi.OPT_BORROW_INPUT, **kw) as h:
        written = 0
        for _ in range(max(1, warmup)):
            _, written = h.decode_device(d_out.data_ptr(), d_out.numel())
        stage = {"scan": 0.0, "crc": 0.0, "link": 0.0, "parse": 0.0, "decode": 0.0, "total": 0.0}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            _, written = h.decode_device(d_out.data_ptr(), d_out.numel())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ssteps = max(1, min(steps, 3))                     # stage breakdown from further steps, read outside the timed region
        for _ in range(ssteps):
            h.decode_device(d_out.data_ptr(), d_out.numel())
            t = h.timing()
            for k in stage:
                stage[k] += getattr(t, k)
        for k in stage:
            stage[k] = round(stage[k] / ssteps, 4)
    return {"ms": ms, "stage_ms": stage, "written": int(written), "d_out": d_out}


def config_entry(ctx, name, kwargs, steps, peak, keep=None):
    """one BASELINE configuration, single GPU, device-resident: generate, decode, check against the unique tile's PCM"""
    import pycorpus
    torch, _abi, dev, local, stream = ctx
    t0 = time.time()
    s = pycorpus.make(md5=False, view=True, **kwargs)
    gen_s = time.time() - t0
    B = (s.bps + 7) // 8
    n_all = s.total_samples * s.channels
    r = device_pass(ctx, s.flac, n_all * B, steps, 2)
    tile_dev = torch.frombuffer(bytearray(s.pcm), dtype=torch.uint8).to(dev)
    ok = r["written"] == n_all * B and verify_periodic(torch, r["d_out"], r["written"], tile_dev, 0)
    if not ok:
        raise SystemExit(f"bench.py: {name}: decoded PCM differs from the generator's PCM -- refusing to report a number")
    alg = len(s.flac) + n_all * B
    e = {"workload": name, "samples": n_all, "frames": len(s.frame_bs), "compressed_bytes": int(len(s.flac)), "pcm_bytes": n_all * B,
         "ms_per_step": round(r["ms"], 4), "samples_per_s": n_all / (r["ms"] / 1e3), "pcm_gbps": n_all * B / (r["ms"] / 1e3) / 1e9,
         "algorithmic_gbps": alg / (r["ms"] / 1e3) / 1e9, "hbm_frac": alg / (r["ms"] / 1e3) / 1e9 / peak, "stage_ms": r["stage_ms"],
         "steps": steps, "pcm_check": "equal to the generator's PCM, every tile", "corpus_gen_s": round(gen_s, 1)}
    del r, tile_dev
    if keep is not None:
        keep.append(s)              # the caller decodes it again (and frees it)
    else:
        s.free()
    torch.cuda.empty_cache()
    _abi.lib().bnflac_trim_pools()             # the engine caches device blocks between handles: give this stream's back before the next one
    return e


def batch_entry(ctx, pool, ks, steps):
    """cfg4: the clips with indices `ks` of the 100,000-clip batch, host memory -> device PCM through bnflac_decode_batch"""
    torch, _abi, dev, local, stream = ctx
    clips = [cfg4_clip(pool, k) for k in ks]
    blobs = [c.flac for c in clips]
    n_all = sum(c.total_samples * c.channels for c in clips)
    out = torch.empty(n_all * 2 + 256, dtype=torch.uint8, device=dev)
    res = None
    t0 = time.perf_counter()
    table = _abi.BatchTable(blobs)         # the bnflac_span array: built once per batch (Python fills ctypes structs at ~4 us each; a C# / C++ caller has the array)
    table_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(2):
        n, res = _abi.decode_batch(table, device=local, dst=out, dst_is_device=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        n, res = _abi.decode_batch(table, device=local, dst=out, dst_is_device=True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    bad = sum(1 for r, c in zip(res, clips) if r.status != 0 or r.pcm_bytes != len(c.pcm))
    for k in list(range(0, len(clips), max(1, len(clips) // 64))):       # spot check of the PCM itself
        r = res[k]
        if bytes(out[r.pcm_offset:r.pcm_offset + r.pcm_bytes].cpu().numpy()) != clips[k].pcm:
            bad += 1
    if bad or n != n_all * 2:
        raise SystemExit("bench.py: cfg4 batch: decoded clips differ from the generator's PCM -- refusing to report a number")
    return {"clips": len(clips), "samples": n_all, "compressed_bytes": sum(len(b) for b in blobs), "pcm_bytes": n_all * 2, "ms_per_step": round(ms, 3),
            "samples_per_s": n_all / (ms / 1e3), "steps": steps, "span_table_python_ms": round(table_ms, 1),
            "path": "bnflac_decode_batch: clips in (pageable) host memory -> gathered + uploaded -> one pass per format group -> PCM left on the device; wall clock of the C-ABI call, "
                    "staging and H2D included; the bnflac_span table is built once per batch outside the timed calls (span_table_python_ms: ctypes marshalling in Python, not part of the C ABI)"}


def stream_surface_entry(flac, n_all, local):
    """e2e through the drop-in surface itself: the C++ mirror of FLACDecoder (csrc/flac_decoder.hpp) over an istream that hands out
    <= 16 KiB per Read, drained with Read(buf, 0, 81920) into pageable memory -- OpenALDemo/Program.cs:26-38's loop -- in its own
    process (birdnest/audio_b200/flacdecoder_demo --bench), timed inside that process."""
    exe = os.path.join(ROOT, "birdnest", "audio_b200", "flacdecoder_demo")
    if not os.path.exists(exe):
        return {"unavailable": "flacdecoder_demo not built (make host)"}
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize > len(flac) + (1 << 28) else "/tmp"
    path = os.path.join(d, f"bnflac_bench_{os.getpid()}.flac")
    try:
        with open(path, "wb") as f:
            f.write(flac)
        runs = {}
        for mode in ("--bench", "--bench-drain"):
            r = subprocess.run([exe, mode, path, str(local), "3"], capture_output=True, text=True, timeout=600)
            reps = []
            for line in r.stdout.splitlines():
                f = line.split()
                if f and f[0] == "rep":
                    reps.append({f[i]: float(f[i + 1]) for i in range(2, len(f) - 1, 2)})
            if r.returncode != 0 or not reps:
                return {"unavailable": (r.stderr or r.stdout)[-200:]}
            runs[mode] = reps
    finally:
        if os.path.exists(path):
            os.unlink(path)
    best = min(runs["--bench"], key=lambda x: x["total_ms"])
    drain = min(runs["--bench-drain"], key=lambda x: x["total_ms"])
    return {"value": n_all / (best["total_ms"] / 1e3), "unit": "samples/s", "ms_total": best["total_ms"],
            "first_read_ms": min(x["first_read_ms"] for x in runs["--bench"][1:] or runs["--bench"]),
            "reads": int(best["reads"]), "pcm_bytes": int(best["bytes"]), "reps_ms": [round(x["total_ms"], 1) for x in runs["--bench"]],
            "drain_only": {"value": n_all / (drain["total_ms"] / 1e3), "ms_total": drain["total_ms"], "reps_ms": [round(x["total_ms"], 1) for x in runs["--bench-drain"]],
                           "what": "the same Read(buf,0,81920) loop with the buffer reused instead of appended to a MemoryStream: decoder + Stream buffering alone"},
            "path": "flacdecoder_demo --bench: new FLACDecoder(istream) [BNFLAC_OPT_LAZY_PULL, <= 16 KiB per stream read] + Read(buf,0,81920) until 0, every buffer appended to a growing pageable vector (Stream.CopyTo(MemoryStream), Program.cs:33); file in tmpfs; first repetition includes CUDA context creation",
            "where_the_time_goes": "single host thread: ~0.3 s pulling 1.2 GB through 16 KiB stream reads, ~0.1 s staging pageable uploads, ~0.25 s copying 2.07 GB of PCM out of pinned memory in 80 KB pieces; the rest of ms_total is the caller's MemoryStream (growth copies + first-touch page faults of ~4 GB); GPU work is ~3 ms per 64 MiB sub-shard, downloads are never waited for (BNFLAC_TRACE=1 prints this split)"}


def corpus_job(ctx, maps, info, shard, steps, prio=False):
    """the streams of a corpus decoded CONCURRENTLY on one GPU (one CUDA stream and one host thread each: a pass below one wave of
    CTAs leaves most of the GPU idle, and the passes of different streams fill it; different handles may be used from different
    threads, include/bnflac.h) -> (ms per step on this rank by CUDA events, output tensors, bytes written)"""
    from concurrent.futures import ThreadPoolExecutor
    torch, _abi, dev, local, stream = ctx
    nf = len(maps)
    # prio: the stream with the most samples (the 8-channel one: a handful of warps walking long subframes) on a high-priority CUDA
    # stream.  Measured with tools/strong_probe.py (one rank's eighth of the corpus on one GPU): 7.58 ms without, 7.68 ms with -- CTA
    # placement is not what the longest pass waits for; off.
    big = max(range(nf), key=lambda fi: info[fi][1]) if prio else -1
    streams = [torch.cuda.Stream(device=dev, priority=-1 if fi == big else 0) for fi in range(nf)]
    hs, outs = [], []
    for fi in range(nf):
        flen, n_all, B, _ = info[fi]
        cap = n_all * B // (shard[1] if shard else 1) + (64 << 20)
        outs.append(torch.empty(cap + 256, dtype=torch.uint8, device=dev))
        kw = dict(shard_index=shard[0], shard_count=shard[1]) if shard else {}
        hs.append(_abi.open_memory(maps[fi], device=local, stream=streams[fi].cuda_stream, flags=_abi.OPT_BORROW_INPUT, **kw))
    written = [0] * nf

    def one(fi):
        _, written[fi] = hs[fi].decode_device(outs[fi].data_ptr(), outs[fi].numel())
    with ThreadPoolExecutor(nf) as pool_ex:
        def step():
            list(pool_ex.map(one, range(nf)))
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        for _ in range(steps):
            step()
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    for hh in hs:
        hh.close()
    return ms, outs, written


def encode_entry(ctx, s, label, reps, steps=3, **enc):
    """the GPU encoder.  `reps` copies of the unique PCM tile of stream `s`, device-resident in, device-resident stream out
    (bnflac_encode_device, MD5 off: it is serial host work); the stream is then decoded by the GPU decoder and compared with the PCM
    on the device."""
    torch, _abi, dev, local, stream = ctx
    tile = torch.frombuffer(bytearray(s.pcm), dtype=torch.uint8).to(dev)
    d_pcm = tile.repeat(reps)
    n = d_pcm.numel()
    B = (s.bps + 7) // 8
    o = _abi.enc_opts(s.sample_rate, s.channels, s.bps, flags=_abi.ENC_NO_MD5 | enc.pop("flags", 0), device=local, **enc)
    cap = _abi.encode_bound(n, o)
    d_flac = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
    best = None
    for _ in range(steps + 1):
        w, st = _abi.encode_device(d_pcm.data_ptr(), n, o, d_flac.data_ptr(), cap)
        if best is None or st.total_ms < best.total_ms:
            best = st
    hdr = bytes(d_flac[:4096].cpu().numpy())
    d_out = torch.empty(n + 256, dtype=torch.uint8, device=dev)
    h = _abi.open_device(d_flac.data_ptr(), w, hdr, device=local, stream=stream.cuda_stream, keep=d_flac)
    _, written = h.decode_device(d_out.data_ptr(), d_out.numel())
    h.close()
    torch.cuda.synchronize()
    if written != n or not torch.equal(d_out[:n], d_pcm):
        raise SystemExit("bench.py: encode: the GPU decoder does not return the PCM the GPU encoder was given -- refusing to report a number")
    samples = n // B
    return {"workload": f"{label}: {n / 1e6:.0f} MB of PCM, device PCM -> device FLAC stream",
            "samples": samples, "frames": int(best.frames), "ms": round(best.total_ms, 3), "plan_ms": round(best.plan_ms, 3), "write_ms": round(best.write_ms, 3),
            "samples_per_s": samples / (best.total_ms / 1e3), "pcm_gbps": n / (best.total_ms / 1e3) / 1e9, "compressed_bytes": int(w), "ratio": w / n,
            "ratio_cpu_corpus_encoder": len(s.flac) / max(1, s.tiles) / len(s.pcm),
            "check": "stream decoded by the GPU decoder == the input PCM (device compare); tests/test_encode_gpu.py decodes such streams with the oracle and the reference DLL"}


DECODE_KERNEL_SOURCES = ("bnflac_dev.h", "kernels.cu", "kernels_common.cuh", "kernels_decode.cuh", "kernels_decode_narrow.cu", "kernels_decode_wide.cu")


def kernels_fingerprint():
    """Hash of the sources the DECODE kernels are compiled from (what their DRAM traffic depends on; the host runtime and the encoder
    are other translation units)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "birdnest", "audio_b200", "csrc")
    for name in DECODE_KERNEL_SOURCES:
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


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

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def gather_ints(x):
        if world == 1:
            return [int(x)]
        tt = [torch.zeros(1, device=dev, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(tt, torch.tensor([int(x)], device=dev, dtype=torch.int64))
        return [int(t.item()) for t in tt]

    s = make_stream(args.seconds)
    numa = bind_to_gpu_cpus(local)            # after the (multi-threaded) stream generator, before any pinned allocation
    flac_len = len(s.flac)
    strong_main = args.scaling == "strong" and world > 1
    shard = dict(shard_index=rank, shard_count=world) if strong_main else {}
    total_samples_all = s.total_samples * s.channels          # per stream
    job_samples = total_samples_all if strong_main else total_samples_all * world
    pcm_bytes_stream = total_samples_all * 3

    # ---- value: device-resident in, device-resident out -------------------------------------------------
    host_in = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).pin_memory()
    d_in = torch.zeros(flac_len + 256, dtype=torch.uint8, device=dev)
    d_in[:flac_len].copy_(host_in, non_blocking=True)
    d_out = torch.empty(pcm_bytes_stream + 256, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    ctx = (torch, _abi, dev, local, stream)
    h = _abi.open_device(d_in.data_ptr(), flac_len, s.flac[:1 << 20], device=local, stream=stream.cuda_stream, keep=d_in, **shard)
    torch.cuda.synchronize()
    written = 0
    for _ in range(args.warmup):
        _, written = h.decode_device(d_out.data_ptr(), d_out.numel())
    # correctness gate on the benchmark's own output (not timed): md5 of the decoded PCM == STREAMINFO md5
    if not strong_main and args.verify:
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
    e1.record(stream)
    barrier()
    t1 = time.time()
    launches = _abi.lib().bnflac_kernel_launches() - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = job_samples / (ms_per_step / 1e3)
    # stage breakdown: the library's own CUDA events of further steps, read outside the timed region (reading six event pairs through
    # ctypes after every step kept the GPU idle for ~20 us of each 2.4 ms step)
    stage_steps = max(1, min(args.steps, 10))
    for _ in range(stage_steps):
        h.decode_device(d_out.data_ptr(), d_out.numel())
        t = h.timing()
        for k in stage:
            stage[k] += getattr(t, k)
    for k in stage:
        stage[k] /= stage_steps
    h.close()
    del d_out, d_in

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
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = job_samples / (e2e_ms / 1e3)
    h2d = flac_len if not strong_main else flac_len // world
    d2h = nbytes
    del host_out
    peak, peak_src = peaks()

    # ---- the other BASELINE configurations (single GPU: the N = 1 run carries them) ----------------------------------
    extra = {}
    small = args.corpus_scale
    pool = None
    if not args.no_configs and world == 1:
        cfgs = {}
        cfgs["cfg1"] = config_entry(ctx, "cfg1 shape: 1 h 16-bit stereo 44.1 kHz, bs 4096, LPC<=8 (configs[0] is the reference's CPU case; GPU run of the same shape)", cfg1_kwargs(small), 10, peak)
        cfgs["cfg3"] = config_entry(ctx, "cfg3: 600 s 24-bit 8 ch 192 kHz, bs 16384, LPC 32, Rice2 partition order 8", cfg3_kwargs(small), 5, peak)
        fm, kept = {}, []
        for key, label, kw in cfg5_formats(small):
            fm[key] = config_entry(ctx, label, kw, 5, peak, keep=kept)
        tot_s = sum(f["samples"] for f in fm.values()); tot_ms = sum(f["ms_per_step"] for f in fm.values())
        tot_alg = sum(f["compressed_bytes"] + f["pcm_bytes"] for f in fm.values())
        cfgs["cfg5"] = {"workload": "cfg5: 10 h mixed-format corpus, one pass per format back to back on one GPU", "formats": fm, "samples": tot_s, "ms_per_step": round(tot_ms, 3),
                        "samples_per_s": tot_s / (tot_ms / 1e3), "algorithmic_gbps": tot_alg / (tot_ms / 1e3) / 1e9, "hbm_frac": tot_alg / (tot_ms / 1e3) / 1e9 / peak}
        pool = cfg4_pool()
        share = max(1, int(CFG4_TOTAL_CLIPS / 8 * small))
        cfgs["cfg4_share"] = batch_entry(ctx, pool, range(0, share * 8, 8), 3)
        cfgs["cfg4_share"]["workload"] = f"cfg4: {share} clips = one GPU's share (1/8) of the 100,000-clip batch, 16-bit mono/stereo 0.5-3 s, mixed FIXED/LPC and blocksizes"
        extra["configs"] = cfgs
        # the same corpus as ONE job (what `strong` shards over N GPUs): the six streams decoded concurrently on this GPU
        info5 = [(len(g.flac), g.total_samples * g.channels, (g.bps + 7) // 8, len(g.pcm)) for g in kept]
        job_ms, outs, wr = corpus_job(ctx, [g.flac for g in kept], info5, None, 5)
        for g, o, w in zip(kept, outs, wr):
            if w != g.total_samples * g.channels * ((g.bps + 7) // 8) or not verify_periodic(torch, o, w, torch.frombuffer(bytearray(g.pcm), dtype=torch.uint8).to(dev), 0):
                raise SystemExit("bench.py: strong (N = 1): decoded PCM differs from the generator's PCM -- refusing to report a number")
        del outs
        for g in kept:
            g.free()
        kept.clear()
        torch.cuda.empty_cache()
        _abi.lib().bnflac_trim_pools()
        extra["strong"] = {"job": "cfg5: 10 h mixed-format corpus (six streams) as one job on one GPU, the six streams decoded concurrently (one CUDA stream + one host thread per stream); "
                                  "N > 1 shards every stream by frame ranges over the ranks", "n_gpus": 1, "samples": tot_s, "ms_per_step": round(job_ms, 3),
                           "samples_per_s": tot_s / (job_ms / 1e3), "hbm_frac": tot_alg / (job_ms / 1e3) / 1e9 / peak, "one_stream_at_a_time_ms": cfgs["cfg5"]["ms_per_step"]}
        extra["by_file"] = dict(cfgs["cfg4_share"], n_gpus=1, note="N = 1: one GPU's 1/8 share of the batch (the whole 100,000 clips are split over the ranks when N > 1)")
        if not args.no_e2e:
            extra["e2e_stream"] = stream_surface_entry(s.flac, total_samples_all, local)
        import pycorpus
        g3 = pycorpus.make(**dict(cfg3_kwargs(small), tile=1))
        extra["encode"] = {
            "cfg2_shape": encode_entry(ctx, s, "8 x 60 s 24-bit stereo 96 kHz, blocksize 4096, LPC <= 12, partition order <= 6, adaptive mid/side", 8,
                                       blocksize=4096, max_lpc_order=12, max_partition_order=6),
            "cfg3_shape": encode_entry(ctx, g3, "100 x 12 frames 24-bit 8 ch 192 kHz, blocksize 16384, LPC 32 (fixed order), Rice2 partition order 8", 100,
                                       blocksize=16384, max_lpc_order=32, min_partition_order=8, max_partition_order=8, flags=_abi.ENC_FIXED_ORDER)}
        g3.free()

    # ---- N > 1: the partitions BASELINE names -- one corpus by frame ranges, the clip batch by file ---------------------
    if not args.no_configs and world > 1:
        import numpy as np
        import pycorpus
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        formats = cfg5_formats(small)
        nf = len(formats)
        # rank 0 generates the corpus once; every rank maps it (a rank only ever reads its own byte range of each stream)
        meta = torch.zeros(nf, 4, dtype=torch.int64, device=dev)
        tiles = []
        for fi, (key, label, kw) in enumerate(formats):
            if rank == 0:
                g = pycorpus.make(md5=False, view=True, **kw)
                with open(os.path.join(shm, f"bnflac_bench_{key}.flac"), "wb") as f:
                    f.write(g.flac)
                meta[fi] = torch.tensor([len(g.flac), g.total_samples * g.channels, (g.bps + 7) // 8, len(g.pcm)], dtype=torch.int64)
                tiles.append(g.pcm)
                g.free()
        barrier()
        dist.broadcast(meta, 0)
        info = [tuple(int(x) for x in meta[fi].tolist()) for fi in range(nf)]         # (compressed bytes, samples, bytes per sample, tile PCM bytes)
        tile_dev = []
        for fi in range(nf):
            t = torch.empty(info[fi][3], dtype=torch.uint8, device=dev)
            if rank == 0:
                t.copy_(torch.frombuffer(bytearray(tiles[fi]), dtype=torch.uint8))
            dist.broadcast(t, 0)
            tile_dev.append(t)
        maps = [np.memmap(os.path.join(shm, f"bnflac_bench_{key}.flac"), dtype=np.uint8, mode="r") for key, _, _ in formats]
        # the same job on ONE GPU, in the same run (the other ranks wait): what the N-GPU time is compared with
        n1_ms, n1_ok = 0.0, True
        if rank == 0:
            n1_ms, outs, written = corpus_job(ctx, maps, info, None, 3)
            for fi in range(nf):
                n1_ok = n1_ok and written[fi] == info[fi][1] * info[fi][2] and verify_periodic(torch, outs[fi], written[fi], tile_dev[fi], 0)
            del outs
            torch.cuda.empty_cache()
            _abi.lib().bnflac_trim_pools()
        barrier()
        ms_n, outs, written = corpus_job(ctx, maps, info, (rank, world), 5)
        ms_n = max_over_ranks(ms_n)
        ok = True
        sizes_all = []
        for fi in range(nf):
            sizes = gather_ints(written[fi])
            sizes_all.append(sizes)
            ok = ok and sum(sizes) == info[fi][1] * info[fi][2] and verify_periodic(torch, outs[fi], written[fi], tile_dev[fi], sum(sizes[:rank]))
        # host-side gather (the only exchange there is): every rank's PCM slices to pinned host memory, timed on the device
        total_w = sum(written)
        gbytes = min(max(written), 2 << 30)
        hb = torch.empty(gbytes, dtype=torch.uint8).pin_memory()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record(stream)
        copied = 0
        for fi in range(nf):
            n = min(written[fi], gbytes)
            hb[:n].copy_(outs[fi][:n], non_blocking=True)
            copied += n
        g1.record(stream)
        torch.cuda.synchronize()
        gather_ms = max_over_ranks(g0.elapsed_time(g1) * (total_w / max(1, copied)))
        okall = max_over_ranks(0.0 if ok else 1.0) == 0.0
        del outs, hb, maps, tile_dev
        torch.cuda.empty_cache()
        _abi.lib().bnflac_trim_pools()
        barrier()
        if rank == 0:
            for key, _, _ in formats:
                os.unlink(os.path.join(shm, f"bnflac_bench_{key}.flac"))
            if not okall or not n1_ok:
                raise SystemExit("bench.py: strong: sharded PCM differs from the generator's PCM -- refusing to report a number")
            tot_s = sum(i[1] for i in info)
            tot_alg = sum(i[0] + i[1] * i[2] for i in info)
            extra["strong"] = {"job": "cfg5: 10 h mixed-format corpus (six streams), every stream sharded by frame ranges over the ranks (bnflac_opts.shard_index / shard_count), no collective; "
                                      "a rank decodes its six shards concurrently (one CUDA stream + one host thread per stream)",
                               "n_gpus": world, "samples": tot_s, "ms_per_step": round(ms_n, 3), "samples_per_s": tot_s / (ms_n / 1e3),
                               "n1_ms_per_step": round(n1_ms, 3), "speedup_vs_n1": n1_ms / ms_n, "efficiency_vs_n1": n1_ms / ms_n / world,
                               "hbm_frac_per_gpu": tot_alg / world / (ms_n / 1e3) / 1e9 / peak, "host_gather_ms": round(gather_ms, 1),
                               "formats": {key: {"workload": label, "samples": info[fi][1], "compressed_bytes": info[fi][0], "pcm_bytes": info[fi][1] * info[fi][2], "shard_pcm_bytes": sizes_all[fi]}
                                           for fi, (key, label, _) in enumerate(formats)},
                               "timing": "CUDA events on this rank's streams around 5 steps (device-resident shards, uploaded before), max over ranks; n1 = the same six streams whole, "
                                         "decoded the same concurrent way on rank 0 alone in the same run",
                               "pcm_check": "every rank's slice of every stream equal to the generator's PCM at its offset; slices add up to the whole"}
        # by file: the 100,000 clips dealt round-robin to the ranks
        pool = cfg4_pool()
        nclips = max(world, int(CFG4_TOTAL_CLIPS * small))
        barrier()
        be = batch_entry(ctx, pool, range(rank, nclips, world), 3)
        ms_b = max_over_ranks(be["ms_per_step"])
        tot_samples = sum(gather_ints(be["samples"]))
        if rank == 0:
            extra["by_file"] = {"job": f"cfg4: {nclips} clips dealt round-robin to {world} ranks, each rank one bnflac_decode_batch call per step (host clips -> device PCM), no collective",
                                "n_gpus": world, "clips": nclips, "samples": tot_samples, "ms_per_step": round(ms_b, 3), "samples_per_s": tot_samples / (ms_b / 1e3),
                                "rank0": be, "timing": "wall clock per rank around 3 calls (gather to pinned staging + H2D + kernels), max over ranks"}

    if rank == 0:
        # roofline of the dominant kernel: algorithmic bytes = compressed bytes in + packed PCM bytes out,
        # the units one launch processes = the whole stream (one k_decode launch per step)
        units_bytes = (flac_len + pcm_bytes_stream) if not strong_main else (flac_len + pcm_bytes_stream) / world
        dom = max(("decode", "parse", "crc", "scan"), key=lambda k: stage[k])
        dom_ms = stage[dom]
        achieved = units_bytes / (dom_ms / 1e3) / 1e9
        # DRAM traffic of that kernel: from one `ncu --set full` capture (never measured inside a bench run), valid only for the
        # build it was taken from -- profiles/traffic.json names the capture and carries the fingerprint of the kernel sources
        traffic, traffic_note = None, "no capture on record"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("kernels_fingerprint") == kernels_fingerprint():
                traffic, traffic_note = tj.get(dom), f"{tj.get('source')} (same kernel sources: {tj.get('kernels_fingerprint')})"
            else:
                traffic_note = f"{tj.get('source')} was taken from other kernel sources ({tj.get('kernels_fingerprint')} != {kernels_fingerprint()}): not quoted"
        line = {
            "metric": "decoded samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong_main else "weak", "vs_baseline": None,
            "dtype": "int32/int64", "data": "synthetic",
            "config": base_config(args, world, strong_main),
            "workload_detail": {"frames_per_stream": len(s.frame_bs), "compressed_bytes": flac_len, "pcm_bytes": pcm_bytes_stream, "host_cpus_bound_to_gpu": numa,
                                "stream_note": "60 tiles of 1406 frames (59.99 s each: whole frames of 4096 samples) = 84,360 frames, 3599.4 s; this encoder's ratio is 0.589 (the reference encoder's on the same shape: 0.528)"},
            "pcm_gbps": value * 3 / 1e9,
            "pipeline_hbm_frac": units_bytes / (ms_per_step / 1e3) / 1e9 / peak,
            "pipeline_hbm_frac_of_nominal_8000_gbps": units_bytes / (ms_per_step / 1e3) / 1e9 / 8000.0,      # both denominators
            "stage_ms": stage,
            "roofline": {"bound": "hbm", "kernel": {"decode": "k_decode", "parse": "k_parse", "crc": "k_crc", "scan": "k_scan"}[dom],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": units_bytes},
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "bnflac_open_memory(pinned host FLAC, BORROW_INPUT) + bnflac_decode_all(pinned host PCM): pipelined sub-shards, PCIe both ways inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line.update(extra)
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
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end legs (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="only the cfg2 line: skip the other BASELINE configurations / the strong and by-file partitions")
    ap.add_argument("--corpus-scale", type=float, default=1.0, help="shrink the corpora of --configs (tile counts, clip count) by this factor (smoke runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
