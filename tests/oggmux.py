"""A small Ogg FLAC muxer for the tests (pages, lacing, page CRC-32) -- test infrastructure, independent of the library's
de-pager: the CRC goes through zlib's reflected CRC on bit-reversed bytes, not through a table of the Ogg polynomial."""
import binascii
import struct

_REV8 = bytes(int(f"{i:08b}"[::-1], 2) for i in range(256))


def ogg_crc(data: bytes) -> int:
    """CRC-32 with polynomial 0x04C11DB7, no reflection, initial value 0, no final XOR -- computed with zlib's reflected
    CRC on bit-reversed bytes (an independent route from the table-driven one in the library)."""
    r = binascii.crc32(data.translate(_REV8), 0xFFFFFFFF) ^ 0xFFFFFFFF
    return int(f"{r:032b}"[::-1], 2)


def page(serial, seq, flags, granule, lacing, body):
    hdr = b"OggS\0" + bytes([flags]) + struct.pack("<qIII", granule, serial, seq, 0) + bytes([len(lacing)]) + bytes(lacing)
    crc = ogg_crc(hdr + body)
    return hdr[:22] + struct.pack("<I", crc) + hdr[26:] + body


def native_packets(s):
    """Ogg FLAC packets of a native stream: header packet, one packet per further metadata block, one per frame."""
    flac = s.flac
    assert flac[:4] == b"fLaC"
    blocks, pos, last = [], 4, False
    while not last:
        last = bool(flac[pos] & 0x80)
        n = int.from_bytes(flac[pos + 1:pos + 4], "big")
        blocks.append(flac[pos:pos + 4 + n])
        pos += 4 + n
    assert pos == s.frame_off[0]
    offs = list(s.frame_off)                          # frame starts + the end of the stream
    assert offs[-1] == len(flac)
    frames = [flac[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    first = b"\x7fFLAC\x01\x00" + struct.pack(">H", len(blocks) - 1) + b"fLaC" + blocks[0]
    return [first] + blocks[1:], frames


def mux(s, rng=None, max_segs=255, serial=0x1234, other_serial=None, split_pages=True):
    """Pages as an Ogg muxer would write them: the header packet alone on the first page, the other header packets on the
    next, then audio packets packed into pages of up to `max_segs` segments, packets spanning pages where they do not fit."""
    headers, frames = native_packets(s)
    out, seq = [], 0
    out.append(page(serial, seq, 2, 0, [len(headers[0])], headers[0])); seq += 1
    if other_serial is not None:                      # a second logical stream multiplexed in (its pages must be ignored)
        out.append(page(other_serial, 0, 2, 0, [8], b"\x01garbage"))
    lacing, body, pending = [], bytearray(), []

    def flush(cont, last=False):
        nonlocal lacing, body, seq
        if not lacing and not last:
            return
        out.append(page(serial, seq, (1 if cont else 0) | (4 if last else 0), seq, lacing, bytes(body)))
        seq += 1
        if other_serial is not None and seq % 5 == 0:
            out.append(page(other_serial, seq // 5, 0, 0, [255, 3], bytes(258)))
        lacing, body = [], bytearray()

    cont = False
    def add_packet(pkt):
        nonlocal cont, lacing, body
        segs = [255] * (len(pkt) // 255) + [len(pkt) % 255]       # a multiple of 255 ends with a 0 lacing value
        at = 0
        for L in segs:
            if len(lacing) >= max_segs:
                flush(cont)
                cont = at > 0                                   # the next page continues this packet
            lacing.append(L)
            body += pkt[at:at + L]
            at += L
        if rng is not None and split_pages and rng.random() < 0.3:
            flush(cont); cont = False

    for hp in headers[1:]:
        add_packet(hp)
    flush(cont); cont = False
    for f in frames:
        add_packet(f)
    flush(cont, last=True)
    return out
