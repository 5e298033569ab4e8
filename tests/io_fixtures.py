"""Writers for the file formats the reference reads (test fixtures for include/trgt_denovo_b200_io.hpp), written from
the SAM/BAM, BGZF, VCF and faidx specifications -- independent of the C++ readers they exercise.  No htslib here:
these are the only files of those formats the environment can produce."""
import struct
import zlib


def bgzf_bytes(data: bytes, block=0xff00) -> bytes:
    """data as a series of BGZF blocks (gzip members with the BC extra field) + the empty EOF block."""
    out = bytearray()
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        if i is not None and not chunk:
            continue
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(comp) + 8 - 1
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize)
        out += comp + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    return bytes(out)


def write_fasta(path, contigs, width=60):
    """contigs: list of (name, sequence str).  Writes path and path + '.fai'."""
    fai = []
    with open(path, "wb") as o:
        off = 0
        for name, seq in contigs:
            hdr = f">{name} synthetic\n".encode()
            o.write(hdr)
            off += len(hdr)
            fai.append(f"{name}\t{len(seq)}\t{off}\t{width}\t{width + 1}\n")
            for i in range(0, len(seq), width):
                line = seq[i:i + width].encode() + b"\n"
                o.write(line)
                off += len(line)
    with open(path + ".fai", "w") as o:
        o.write("".join(fai))


def write_text(path, text: str, bgzf: bool):
    with open(path, "wb") as o:
        o.write(bgzf_bytes(text.encode()) if bgzf else text.encode())


_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _tag(tag: str, ty: str, value) -> bytes:
    t = tag.encode() + ty.encode()
    if ty in "ZH":
        return t + value.encode() + b"\0"
    if ty == "B":  # value = (subtype, list)
        sub, vals = value
        fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
        return t + sub.encode() + struct.pack("<i", len(vals)) + b"".join(struct.pack("<" + fmt, v) for v in vals)
    return t + struct.pack("<" + {"A": "c", "c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[ty],
                           value.encode() if ty == "A" else value)


def bam_record(name: str, ref_id: int, pos: int, seq: str, tags) -> bytes:
    """tags: list of (tag, type, value).  One soft-clip-free match CIGAR, quality 0xff (absent)."""
    l_seq = len(seq)
    packed = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        packed[i // 2] |= _CODE.get(c, 15) << (0 if i & 1 else 4)
    rn = name.encode() + b"\0"
    cigar = struct.pack("<I", (l_seq << 4) | 0) if l_seq else b""
    body = struct.pack("<iiBBHHHiiii", ref_id, pos, len(rn), 60, 4680, 1 if l_seq else 0, 0, l_seq, -1, -1, 0)
    body += rn + cigar + bytes(packed) + b"\xff" * l_seq + b"".join(_tag(*t) for t in tags)
    return struct.pack("<i", len(body)) + body


def write_bam(path, header_text: str, refs, records):
    """refs: list of (name, length); records: iterable of bam_record() bytes."""
    data = bytearray(b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text.encode() + struct.pack("<i", len(refs)))
    for name, ln in refs:
        data += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    for r in records:
        data += r
    with open(path, "wb") as o:
        o.write(bgzf_bytes(bytes(data)))
