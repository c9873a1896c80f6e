"""CPU: the block FASTQ reader (kart_b200/host/read_input.cpp: newline index + parallel entry parsing) yields exactly the batches of the
entry-at-a-time reader that mirrors the reference's GetNextEntry/GetNextChunk (src/GetData.cpp:51-143), on well-formed and on odd inputs."""
import gzip
import itertools
import os
import random
import subprocess

import pytest

import parity_util as pu

HERE = os.path.join(pu.ROOT, "tests", "host_io")


@pytest.fixture(scope="module")
def io_check():
    exe = os.path.join(HERE, "io_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "io_check.cpp"), os.path.join(pu.ROOT, "kart_b200", "host", "read_input.cpp"), "-lz", "-lpthread"], check=True)
    return exe


def fq(n, seed, L=(20, 60), eol="\n", last_eol=True, names=None):
    rnd = random.Random(seed); out = []
    for i in range(n):
        l = rnd.randint(*L)
        name = names[i % len(names)] if names else "r%d_%d" % (seed, i)
        out.append("@%s%s%s%s+%s%s%s" % (name, eol, "".join(rnd.choice("ACGTNacgt") for _ in range(l)), eol, eol, "".join(chr(rnd.randint(33, 73)) for _ in range(l)), eol))
    s = "".join(out)
    if not last_eol and s.endswith(eol):
        s = s[:-len(eol)]
    return s.encode()


def run(exe, mode, threads, batch, pe, files, chunk=None):
    env = dict(os.environ)
    if chunk:
        env["KART_B200_IO_CHUNK"] = str(chunk)
    return subprocess.run([exe, mode, str(threads), str(batch), str(pe)] + files, capture_output=True, env=env, check=True).stdout


CASES = {
    "plain": (fq(37, 1), fq(37, 2)),
    "no_final_newline": (fq(9, 3, last_eol=False), fq(9, 4, last_eol=False)),
    "crlf": (fq(8, 5, eol="\r\n"), fq(8, 6, eol="\r\n")),
    "odd_names": (fq(12, 7, names=["a/1", "b c", "@@x", "t\tz", "", ">q", "n/"]), fq(12, 8, names=["a/2", "b d"])),
    "unequal_counts": (fq(11, 9), fq(7, 10)),
    "truncated_entry": (fq(5, 11) + b"@tail\nACGT\n+\n", fq(5, 12) + b"@tail\nACGT"),
    "empty_seq_in_mate1": (fq(4, 13) + b"@e\n\n+\n\n" + fq(4, 14), fq(9, 15)),
    "empty_seq_in_mate2": (fq(9, 16), fq(3, 17) + b"@e\n\n+\n\n" + fq(5, 18)),
    "short_and_long_qual": (b"@a\nACGTACGT\n+\nIII\n@b\nACGT\n+\nIIIIIIII\n@c\nAC\n+\n\n" + fq(3, 19), fq(6, 20)),
    "nul_in_qual": (b"@a\nACGTACGT\n+\nII\x00IIIII\n" + fq(3, 21), fq(4, 22)),
    "empty_file": (b"@x\nACGT\n+\nIIII\n", b""),
    "one_line": (b"@only", b"@only\nAC"),
    "blank_tail": (fq(3, 23) + b"\n\n", fq(3, 24) + b"\n"),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_block_reader_equals_entry_reader(io_check, tmp_path, case):
    a, b = CASES[case]
    f1, f2 = str(tmp_path / "a.fq"), str(tmp_path / "b.fq")
    for f, d in ((f1, a), (f2, b)):
        with open(f, "wb") as fh:
            fh.write(d)
    for files, pe in (([f1, f2], 1), ([f1], 1), ([f1], 0), ([f2, f1], 1)):
        if not open(files[0], "rb").read(1):
            continue
        for batch, chunk in itertools.product((2, 6, 1000), (None, 16, 101)):
            want = run(io_check, "serial", 1, batch, pe, files)
            for threads in (1, 3):
                got = run(io_check, "blocks", threads, batch, pe, files, chunk)
                assert got == want, (case, files, pe, batch, chunk, threads)


def test_block_reader_gz_and_many_refills(io_check, tmp_path):
    a, b = fq(3000, 31, L=(30, 200)), fq(3000, 32, L=(30, 200))
    f1, f2 = str(tmp_path / "a.fq.gz"), str(tmp_path / "b.fq.gz")
    for f, d in ((f1, a), (f2, b)):
        with gzip.open(f, "wb") as fh:
            fh.write(d)
    want = run(io_check, "serial", 1, 512, 1, [f1, f2])
    assert want.count(b"\n") == 6000 + 12
    for threads, chunk in ((4, 4096), (8, 333), (2, None)):
        assert run(io_check, "blocks", threads, 512, 1, [f1, f2], chunk) == want


def bgzf_bytes(data: bytes, block: int = 60000, level: int = 6) -> bytes:
    """The container bgzip / htslib write: self-contained gzip members of at most 64 KB with a 'BC' extra field, then the empty end marker."""
    import struct
    import zlib
    out = []
    for at in list(range(0, len(data), block)) + [None]:
        piece = b"" if at is None else data[at:at + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = c.compress(piece) + c.flush()
        bsize = 18 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + body + struct.pack("<II", zlib.crc32(piece) & 0xFFFFFFFF, len(piece)))
    return b"".join(out)


def test_bgzf_input_is_inflated_in_parallel_and_equals_plain(io_check, tmp_path):
    """A chain of BGZF blocks goes through GzInput's parallel inflate; the batches must equal the ones of the same text read plainly, whatever
    the request size (requests smaller than a block use the carry buffer) and the worker count; a damaged block ends the input like gzread."""
    a, b = fq(4000, 41, L=(30, 200)), fq(4000, 42, L=(30, 200))
    p1, p2 = str(tmp_path / "a.fq"), str(tmp_path / "b.fq")
    z1, z2 = str(tmp_path / "a.fq.gz"), str(tmp_path / "b.fq.gz")
    for f, d in ((p1, a), (p2, b)):
        open(f, "wb").write(d)
    for f, d, blk in ((z1, a, 60000), (z2, b, 777)):
        open(f, "wb").write(bgzf_bytes(d, blk))
    assert gzip.open(z1, "rb").read() == a and gzip.open(z2, "rb").read() == b   # valid gzip for everybody else
    want = run(io_check, "serial", 1, 512, 1, [p1, p2])
    assert want.count(b"\n") == 8000 + 16
    for threads, chunk in ((4, 4096), (8, 333), (2, None), (1, 100000)):
        assert run(io_check, "blocks", threads, 512, 1, [z1, z2], chunk) == want
    assert run(io_check, "serial", 3, 512, 1, [z1, z2]) == want
    env = dict(os.environ, KART_B200_NO_BGZF="1")   # the same files through zlib's gzread
    assert subprocess.run([io_check, "blocks", "4", "512", "1", z1, z2], capture_output=True, env=env, check=True).stdout == want
    # damage one block's payload: the reader must stop there (CRC / inflate error), not deliver garbage
    raw = bytearray(open(z1, "rb").read()); raw[len(raw) // 2] ^= 0x55
    bad = str(tmp_path / "bad.fq.gz"); open(bad, "wb").write(bytes(raw))
    good = run(io_check, "blocks", 4, 512, 0, [z1]).split(b"\n")
    got = run(io_check, "blocks", 4, 512, 0, [bad]).split(b"\n")
    reads = lambda lines: [ln for ln in lines if ln.startswith(b"[")]
    assert 100 < len(reads(got)) < len(reads(good)) and reads(got)[:-1] == reads(good)[:len(reads(got)) - 1]   # everything in front of the damage, nothing made up behind it
