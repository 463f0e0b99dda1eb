#!/usr/bin/env python
"""Generates tests/golden/ by RUNNING THE REFERENCE'S OWN DECODER (LibFlac.dll hosted by oracle/refdll) in this
container.  /root/reference does not exist on the GPU box, so what the reference produced is committed here as small
fixtures, together with this script:

  tests/golden/ref_*.flac        streams ENCODED by the reference DLL's encoder (FLAC__stream_encoder_*) from synthetic PCM
  tests/golden/golden.json       for every fixture and every synthetic test case (tests/conftest.py CASES):
                                   md5 of the PCM the reference DLL decodes, its frame count, final state, error callbacks
                                 plus fault-injection cases (bit flips) with the reference's observed behaviour.

Run:  python oracle/make_golden.py        (needs oracle/_ref/refflac + LibFlac.dll: `make ref`)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "corpus"), os.path.join(ROOT, "oracle")]
REF = os.path.join(ROOT, "oracle", "_ref", "refflac")
DLL = os.path.join(ROOT, "oracle", "_ref", "LibFlac.dll")
GOLD = os.path.join(ROOT, "tests", "golden")


def ref_decode(flac: bytes):
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "i.flac"), os.path.join(d, "o.pcm")
        open(fi, "wb").write(flac)
        r = subprocess.run([REF, "dec", DLL, fi, fo], capture_output=True, text=True)
        if r.returncode != 0:
            return {"failed": r.stdout.strip() + r.stderr.strip(), "rc": r.returncode}
        pcm = open(fo, "rb").read() if os.path.exists(fo) else b""
        head = r.stdout.splitlines()[0]
        kv = dict(t.split("=") for t in head.split())
        errs = [(int(l.split("=")[1].split()[0]), int(l.split("state=")[1])) for l in r.stdout.splitlines()[1:] if l.startswith("error[")]
        return {"pcm": pcm, "frames": int(kv["frames"]), "state": int(kv["state"]), "errors": [e[0] for e in errs], "error_states": [e[1] for e in errs],
                "sr": int(kv["sr"]), "ch": int(kv["ch"]), "bps": int(kv["bps"]), "total": int(kv["total"]), "si_md5": kv["si_md5"]}


def ref_encode(pcm: bytes, ch, bps, sr, bs, lpc, minpo, maxpo, ms, ex):
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "i.pcm"), os.path.join(d, "o.flac")
        open(fi, "wb").write(pcm)
        subprocess.check_call([REF, "enc", DLL, str(ch), str(bps), str(sr), str(bs), str(lpc), str(minpo), str(maxpo), str(ms), str(ex), fi, fo],
                              stdout=subprocess.DEVNULL)
        return open(fo, "rb").read()


def _block(btype: int, payload: bytes, last: bool = False) -> bytes:
    return bytes([(0x80 if last else 0) | btype]) + len(payload).to_bytes(3, "big") + payload


def metadata_variants(flac: bytes):
    """The same audio behind different container prefixes (LibFLACSharp.cs:270-280 lists the block types; the
    decoder answers STREAMINFO only, FLACDecoder.cs:431-473, and must skip the rest).  Returns {name: bytes}."""
    assert flac[:4] == b"fLaC" and (flac[4] & 0x7f) == 0
    si = flac[8:8 + 34]
    pos, last = 4, False
    while not last:                                   # first frame = after the last metadata block of the original
        last = bool(flac[pos] & 0x80)
        pos += 4 + int.from_bytes(flac[pos + 1:pos + 4], "big")
    frames = flac[pos:]
    vendor = b"reference libFLAC 1.2.1 20070917"
    vorbis = len(vendor).to_bytes(4, "little") + vendor + (2).to_bytes(4, "little")
    for c in (b"TITLE=synthetic", b"ARTIST=bnflac tests"):
        vorbis += len(c).to_bytes(4, "little") + c
    seek = b"".join(((i * 4096).to_bytes(8, "big") + (i * 9000).to_bytes(8, "big") + (4096).to_bytes(2, "big")) for i in range(3))
    seek += b"\xff" * 8 + bytes(10)                   # one placeholder point
    cue = bytes(128) + (0).to_bytes(8, "big") + bytes([0]) + bytes(258) + bytes([1])          # catalog, lead-in, flags, reserved, 1 track
    cue += (0).to_bytes(8, "big") + bytes([170]) + bytes(12) + bytes([0]) + bytes(13) + bytes([0])   # lead-out track, no indices
    mime, desc, data = b"image/png", b"cover", bytes(range(256)) * 3 + b"\xff\xf8\xc9\x18\x00\x00"      # a fake sync code inside
    pic = (3).to_bytes(4, "big") + len(mime).to_bytes(4, "big") + mime + len(desc).to_bytes(4, "big") + desc
    pic += (16).to_bytes(4, "big") * 2 + (24).to_bytes(4, "big") + (0).to_bytes(4, "big") + len(data).to_bytes(4, "big") + data
    app = b"bnfl" + b"\xff\xf8" * 40                    # sync-code look-alikes inside an APPLICATION block
    out = {}
    out["meta_all_block_types"] = (b"fLaC" + _block(0, si) + _block(2, app) + _block(3, seek) + _block(4, vorbis) + _block(5, cue) +
                                   _block(6, pic) + _block(1, bytes(77)) + _block(30, b"reserved type") + _block(1, bytes(5), last=True) + frames)
    id3 = b"ID3\x03\x00\x00" + bytes([0, 0, 2, 44]) + (b"TIT2" + bytes(296))                  # syncsafe size 300
    out["meta_id3v2_prefix"] = id3 + flac
    out["meta_big_padding"] = b"fLaC" + _block(0, si) + _block(1, bytes(70000), last=True) + frames   # > the 16 KiB read buffer of FLACDecoder.cs:19-24
    si_unknown = si[:13] + bytes([si[13] & 0xF0]) + bytes(4) + bytes(16)                         # total_samples = 0, md5 = 0
    si_unknown = si_unknown[:4] + bytes(6) + si_unknown[10:]                                       # min/max frame size unknown
    out["meta_unknown_length_no_md5"] = b"fLaC" + _block(0, si_unknown, last=True) + frames
    return out


def metadata_cases(out):
    """4. container / metadata breadth: what the reference decoder does with every block type, an ID3v2 tag, a large PADDING
    block and a STREAMINFO that does not know the length -- fixtures committed as tests/golden/meta_*.flac."""
    import pycorpus
    import pyoracle
    base = pycorpus.make(ch=2, bps=16, sr=44100, seconds=0.6, bs=1152, lpc=8, seed=404)
    out["metadata"] = {}
    ok = True
    for name, blob in sorted(metadata_variants(base.flac).items()):
        r = ref_decode(blob)
        o_pcm, o_n, _, o_err = pyoracle.decode(blob)
        same = r.get("pcm") == o_pcm == base.pcm and r.get("errors") == o_err == [] and r.get("frames") == o_n
        ok &= same
        open(os.path.join(GOLD, name + ".flac"), "wb").write(blob)
        out["metadata"][name] = {"ref_pcm_md5": hashlib.md5(r.get("pcm", b"")).hexdigest(), "frames": r.get("frames"), "state": r.get("state"), "errors": r.get("errors"),
                                 "sr": r.get("sr"), "ch": r.get("ch"), "bps": r.get("bps"), "total": r.get("total"), "si_md5": r.get("si_md5"),
                                 "bytes": len(r.get("pcm", b"")), "oracle_equal": bool(same)}
        print(("OK  " if same else "FAIL"), name, len(blob), r.get("frames"), r.get("errors"), r.get("total"), r.get("failed"))
    return ok


def main():
    if "--only-metadata" in sys.argv:
        out = json.load(open(os.path.join(GOLD, "golden.json")))
        ok = metadata_cases(out)
        json.dump(out, open(os.path.join(GOLD, "golden.json"), "w"), indent=1, sort_keys=True)
        print("metadata pinned" if ok else "MISMATCHES PRESENT")
        return 0 if ok else 1
    import pycorpus
    import pyoracle
    from conftest import CASES
    os.makedirs(GOLD, exist_ok=True)
    out = {"note": "produced by oracle/make_golden.py running the reference LibFlac.dll (libFLAC 1.2.1) in the build container",
           "cases": {}, "fixtures": {}, "faults": {}}
    ok = True
    # 1. every synthetic case of the test-suite: reference decode == oracle decode == synthesis PCM
    for name, kw in sorted(CASES.items()):
        s = pycorpus.make(**kw)
        r = ref_decode(s.flac)
        o_pcm, o_n, _, o_err = pyoracle.decode(s.flac)
        same = ("pcm" in r) and r["pcm"] == o_pcm == s.pcm * s.tiles and r["errors"] == o_err and r["frames"] == o_n
        ok &= same
        out["cases"][name] = {"ref_pcm_md5": hashlib.md5(r.get("pcm", b"")).hexdigest(), "frames": r.get("frames"), "state": r.get("state"),
                              "errors": r.get("errors"), "flac_md5": hashlib.md5(s.flac).hexdigest(), "oracle_equal": bool(same)}
        print(("OK  " if same else "FAIL"), name, r.get("frames"), r.get("errors"))
    # 2. streams encoded by the REFERENCE encoder (different model/partition choices than corpus/bncorpus.c)
    fixtures = {
        "ref_16bit_stereo_lpc8": dict(ch=2, bps=16, sr=44100, samples=4096 * 5 + 123, bs=4096, lpc=8, minpo=0, maxpo=5, ms=1, ex=0, noise=6),
        "ref_24bit_stereo_lpc12": dict(ch=2, bps=24, sr=96000, samples=4096 * 4 + 7, bs=4096, lpc=12, minpo=0, maxpo=6, ms=1, ex=1, noise=12),
        "ref_24bit_8ch_lpc32_po8": dict(ch=8, bps=24, sr=192000, samples=16384, bs=16384, lpc=32, minpo=8, maxpo=8, ms=0, ex=0, noise=19),
        "ref_16bit_mono_fixed": dict(ch=1, bps=16, sr=44100, samples=576 * 20 + 5, bs=576, lpc=0, minpo=0, maxpo=4, ms=0, ex=0, noise=6, kind=1),
        "ref_24bit_6ch_special": dict(ch=6, bps=24, sr=48000, samples=1152 * 18, bs=1152, lpc=8, minpo=0, maxpo=4, ms=0, ex=0, noise=10, kind=1),
        "ref_8bit_stereo": dict(ch=2, bps=8, sr=22050, samples=1024 * 9 + 100, bs=1024, lpc=6, minpo=0, maxpo=3, ms=1, ex=0, noise=3),
    }
    import ctypes as C
    for name, f in fixtures.items():
        n, ch, bps = f["samples"], f["ch"], f["bps"]
        pcm32 = (C.c_int32 * (n * ch))()
        pycorpus.lib().bnc_synth(pcm32, n, ch, bps, f["sr"], f["noise"], f.get("kind", 0), f["bs"], 77)
        B = (bps + 7) // 8
        packed = C.create_string_buffer(n * ch * B + 1)
        pycorpus.lib().bnc_pack_pcm(pcm32, n * ch, bps, packed)
        pcm = packed.raw[:n * ch * B]
        flac = ref_encode(pcm, ch, bps, f["sr"], f["bs"], f["lpc"], f["minpo"], f["maxpo"], f["ms"], f["ex"])
        r = ref_decode(flac)
        o_pcm, o_n, _, o_err = pyoracle.decode(flac)
        same = r["pcm"] == o_pcm == pcm and not r["errors"] and not o_err
        ok &= same
        open(os.path.join(GOLD, name + ".flac"), "wb").write(flac)
        out["fixtures"][name] = {"pcm_md5": hashlib.md5(pcm).hexdigest(), "frames": r["frames"], "bytes": len(pcm), "si_md5": r["si_md5"],
                                 "channels": ch, "bps": bps, "sample_rate": f["sr"], "total_samples": n, "oracle_equal": bool(same)}
        print(("OK  " if same else "FAIL"), name, len(flac), r["frames"], r["si_md5"] == hashlib.md5(pcm).hexdigest())
    # 3. fault injection: what the reference does with damaged streams
    base = pycorpus.make(**CASES["cfg1_16bit_stereo_lpc8"])
    flips = {"payload_bit_mid_frame": (base.frame_off[3] + (base.frame_off[4] - base.frame_off[3]) // 2, 0x10),
             "payload_bit_near_end": (base.frame_off[7] - 5, 0x01),
             "crc16_byte": (base.frame_off[5] - 1, 0x80),
             "header_blocksize_bits": (base.frame_off[2] + 2, 0x40),
             "sync_byte": (base.frame_off[6], 0x08),
             "last_frame_payload": (base.frame_off[-2] + 40, 0x04)}
    for name, (pos, mask) in flips.items():
        b = bytearray(base.flac)
        b[pos] ^= mask
        r = ref_decode(bytes(b))
        o_pcm, o_n, _, o_err = pyoracle.decode(bytes(b))
        same = r.get("pcm") == o_pcm and r.get("errors") == o_err
        out["faults"][name] = {"case": "cfg1_16bit_stereo_lpc8", "pos": pos, "mask": mask, "ref_pcm_md5": hashlib.md5(r.get("pcm", b"")).hexdigest(),
                               "ref_bytes": len(r.get("pcm", b"")), "frames": r.get("frames"), "errors": r.get("errors"), "error_states": r.get("error_states"),
                               "state": r.get("state"), "oracle_equal": bool(same)}
        print(("OK  " if same else "DIFF"), "fault", name, r.get("frames"), r.get("errors"), r.get("error_states"), "oracle:", o_n, o_err, len(o_pcm), len(r.get("pcm", b"")))
    ok &= metadata_cases(out)
    json.dump(out, open(os.path.join(GOLD, "golden.json"), "w"), indent=1, sort_keys=True)
    print("all pinned" if ok else "MISMATCHES PRESENT")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
