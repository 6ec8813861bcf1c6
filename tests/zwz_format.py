"""`.zwz` record container (compression.cpp:73-104 writer, decompression.cpp:65-92 reader) — test-side helper.

record := i32 total_size (= 4 + path_len + 4 + 1 + payload_len)   i32 path_len   path   i32 sequence_id
          u8 is_last_chunk   payload   [32 hex chars of MD5 iff is_last_chunk; not counted in total_size]
Native (little) endian, no header/footer/index.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional


@dataclass
class Record:
    path: str
    seq: int
    last: bool
    payload: bytes
    md5: Optional[bytes]  # 32 hex bytes when last


def parse(buf: bytes) -> List[Record]:
    out: List[Record] = []
    o = 0
    n = len(buf)
    while o + 4 <= n:
        (total,) = struct.unpack_from("<i", buf, o)
        (plen,) = struct.unpack_from("<i", buf, o + 4)
        path = buf[o + 8:o + 8 + plen].decode()
        seq, last = struct.unpack_from("<iB", buf, o + 8 + plen)
        clen = total - (4 + plen + 4 + 1)
        p0 = o + 8 + plen + 5
        payload = buf[p0:p0 + clen]
        o = p0 + clen
        md5 = None
        if last:
            md5 = buf[o:o + 32]
            o += 32
        out.append(Record(path, seq, bool(last), payload, md5))
    return out


def serialize(recs: List[Record]) -> bytes:
    parts = []
    for r in recs:
        p = r.path.encode()
        total = 4 + len(p) + 4 + 1 + len(r.payload)
        parts.append(struct.pack("<ii", total, len(p)) + p + struct.pack("<iB", r.seq, 1 if r.last else 0) + r.payload)
        if r.last:
            assert r.md5 is not None and len(r.md5) == 32
            parts.append(r.md5)
    return b"".join(parts)
