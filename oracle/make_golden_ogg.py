#!/usr/bin/env python
"""Golden records for the Ogg FLAC container layer, produced by the REFERENCE's own decoder binary.

The reference's LibFlac.dll (libFLAC 1.2.1) carries libogg and exports FLAC__stream_decoder_init_ogg_stream, although the
reference's C# never binds it.  `oracle/_ref/refflac decogg` drives it with FLACDecoder's callbacks (read callback with
libFLAC's own end-of-stream convention: with the C#'s "short read = END_OF_STREAM" the DLL's Ogg layer drops the frames of
the last read).  This script muxes small synthetic streams into Ogg pages (tests/oggmux.py), runs the DLL on them here, in
the build container, and commits the pages + what the DLL produced: tests/golden/ogg_*.oga and tests/golden/golden_ogg.json.
Run from the repo root:  python oracle/make_golden_ogg.py     (needs oracle/_ref, i.e. `make ref`)"""
import hashlib, json, os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import pycorpus
from oggmux import mux, native_packets, page

REF, DLL = os.path.join(ROOT, "oracle", "_ref", "refflac"), os.path.join(ROOT, "oracle", "_ref", "LibFlac.dll")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_dll(blob):
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.oga"), os.path.join(d, "out.pcm")
        open(fi, "wb").write(blob)
        out = subprocess.run([REF, "decogg", DLL, fi, fo], capture_output=True, text=True, check=True).stdout
        pcm = open(fo, "rb").read()
    kv = dict(x.split("=") for x in out.splitlines()[0].split())
    errors = [int(l.split("=")[1].split()[0]) for l in out.splitlines()[1:] if l.startswith("error[")]
    return {"frames": int(kv["frames"]), "bytes": len(pcm), "pcm_md5": hashlib.md5(pcm).hexdigest(), "state": int(kv["state"]), "errors": errors,
            "channels": int(kv["ch"]), "bps": int(kv["bps"]), "sample_rate": int(kv["sr"]), "total_samples": int(kv["total"]), "si_md5": kv["si_md5"]}


def one_frame_per_page(s, serial=7):
    headers, frames = native_packets(s)
    lac = lambda b: [255] * (len(b) // 255) + [len(b) % 255]
    pages = [page(serial, 0, 2, 0, lac(headers[0]), headers[0])]
    for hp in headers[1:]:
        pages.append(page(serial, len(pages), 0, 0, lac(hp), hp))
    first_audio = len(pages)
    for i, f in enumerate(frames):
        pages.append(page(serial, len(pages), 4 if i + 1 == len(frames) else 0, i, lac(f), f))
    return pages, first_audio


var = pycorpus.make(ch=2, bps=16, sr=44100, seconds=1.1, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304))
mono = pycorpus.make(ch=1, bps=16, sr=44100, seconds=0.7, bs=576, lpc=0)
s24 = pycorpus.make(ch=2, bps=24, sr=96000, seconds=0.25, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, seed=52)
cases = {
    "ogg_var_small_pages_foreign_stream": b"".join(mux(var, random.Random(1), max_segs=17, other_serial=0x77)),
    "ogg_mono_one_segment_pages": b"".join(mux(mono, random.Random(2), max_segs=1)),
    "ogg_24bit_full_pages": b"".join(mux(s24, None, max_segs=255)),
}
pages, fa = one_frame_per_page(mono)
dmg = list(pages)
b = bytearray(dmg[fa + 5]); b[-3] ^= 1; dmg[fa + 5] = bytes(b)          # page CRC mismatch
del dmg[fa + 2]                                                          # page missing
dmg.insert(fa + 9, b"Ogg but not a page, then OggS\x01 and junk" + bytes(40))
cases["ogg_mono_lost_corrupt_pages_junk"] = b"".join(dmg)
native_md5 = {"ogg_var_small_pages_foreign_stream": var, "ogg_mono_one_segment_pages": mono, "ogg_24bit_full_pages": s24}
rec = {"note": "produced by oracle/make_golden_ogg.py running the reference LibFlac.dll (libFLAC 1.2.1 + libogg) through FLAC__stream_decoder_init_ogg_stream in the build container"}
for name, blob in cases.items():
    open(os.path.join(GOLD, name + ".oga"), "wb").write(blob)
    r = run_dll(blob)
    if name in native_md5:                      # intact layouts: the DLL must reproduce the native stream's PCM
        assert r["pcm_md5"] == hashlib.md5(native_md5[name].pcm).hexdigest() == r["si_md5"], name
    rec[name] = r
    print(name, len(blob), r)
json.dump(rec, open(os.path.join(GOLD, "golden_ogg.json"), "w"), indent=1, sort_keys=True)
