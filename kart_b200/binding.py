"""ctypes binding of the C ABI in include/kart_b200.h (kart_b200/libkartb200.so).

This is the host-side mirror used by tests and bench.py; the production host is the C++ CLI (kart_b200/host).
There is no CPU path: if the CUDA library or a CUDA device is missing, construction fails loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .index import KartIndex

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libkartb200.so")


class KbIndexHost(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt", C.c_void_p), ("bwt_words", C.c_uint64),
                ("sa", C.c_void_p), ("n_sa", C.c_uint64), ("sa_intv", C.c_int32), ("pac", C.c_void_p), ("l_pac", C.c_int64),
                ("n_chr", C.c_int32), ("chr_len", C.c_void_p)]


class KbParams(C.Structure):
    _fields_ = [("min_seed_len", C.c_int32), ("max_gaps", C.c_int32), ("max_insert", C.c_int32), ("pacbio", C.c_int32),
                ("paired", C.c_int32), ("multihit", C.c_int32)]


class KbReads(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("seq", C.c_void_p), ("seq_off", C.c_void_p)]


class KbReadsPacked(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("code", C.c_void_p), ("n_words", C.c_uint64), ("seq_off", C.c_void_p), ("exc", C.c_void_p), ("n_exc", C.c_uint64)]


class KbResults(C.Structure):
    _fields_ = [("aln", C.c_void_p), ("pairs", C.c_void_p), ("cigar", C.c_void_p), ("cap_cigar", C.c_uint32), ("n_cigar", C.c_uint32)]


ALN_DTYPE = np.dtype([("pos", "<i8"), ("mate_pos", "<i8"), ("kind", "<i4"), ("flag", "<i4"), ("chr", "<i4"), ("mapq", "<i4"),
                      ("score", "<i4"), ("sub_score", "<i4"), ("tlen", "<i4"), ("fwd", "<i4"), ("cig_off", "<u4"), ("cig_len", "<i4")])
EXTRA_DTYPE = np.dtype([("read", "<u4"), ("rank", "<u4"), ("aln", ALN_DTYPE)])   # kb_extra_t
PAIR_DTYPE = np.dtype([("counted", "<i4"), ("absdist", "<i4"), ("est_lo", "<i4"), ("est_hi", "<i4")])
SEG_DTYPE = np.dtype([("gpos", "<i8"), ("rpos", "<i4"), ("rlen", "<i4"), ("glen", "<i4"), ("simple", "<i4")])
CAND_DTYPE = np.dtype([("diff", "<i8"), ("score", "<i4"), ("mate", "<i4"), ("seg_start", "<u4"), ("nseg", "<i4")])
REPORT_DTYPE = np.dtype([("pos", "<i8"), ("aln", "<i4"), ("flag", "<i4"), ("mate", "<i4"), ("chr", "<i4"), ("cig_off", "<u4"),
                         ("cig_len", "<i4"), ("fwd", "<i4"), ("pad", "<i4")])
HIT_DTYPE = np.dtype([("x0", "<u8"), ("rpos", "<u4"), ("len_freq", "<u4")])   # KbHit: len << 8 | freq
DBG_FRAG_DTYPE = np.dtype([("read", "<u4"), ("rpos", "<i4"), ("rlen", "<i4"), ("glen", "<i4"), ("mode", "<i4"), ("pad", "<i4"), ("gpos", "<i8")])   # kb_dbg_frag_t
DBG_OUT_DTYPE = np.dtype([("info", "<i4"), ("aux", "<i4"), ("score", "<i4"), ("n_ops", "<i4"), ("ops_off", "<u4"), ("nruns", "<i4"), ("ident", "<i4"),
                          ("aligned", "<i4"), ("g_first", "<i8"), ("g_end", "<i8")])   # kb_dbg_frag_out_t
RES_DTYPE = np.dtype([("score", "<i4"), ("sub", "<i4"), ("mapq", "<i4"), ("ncan", "<i4"), ("best", "<i4"), ("rep_off", "<u4")])
CIGAR_OPS = "MIDNSHP=X"

EXPORTS = ["kb_device_count", "kb_init", "kb_destroy", "kb_strerror", "kb_last_error", "kb_upload_index", "kb_clone_index", "kb_set_params", "kb_get_min_seed_len",
           "kb_map_chunk", "kb_map_chunk_packed", "kb_map_chunk_begin", "kb_map_chunk_begin_packed", "kb_map_chunk_end", "kb_stage_reads", "kb_stage_reads_packed", "kb_packed_words", "kb_pack_reads", "kb_run", "kb_fetch_results", "kb_fetch_extra", "kb_stage_ms", "kb_work", "kb_cuda_stream", "kb_debug_fetch", "kb_debug_align", "kb_host_alloc", "kb_host_free", "kb_host_register", "kb_host_unregister",
           "kb_index_build", "kb_index_free", "kb_index_build_error"]


class KartB200Error(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    path = path or os.environ.get("KART_B200_LIB") or DEFAULT_LIB   # the override is for A/B runs of an older build (scripts/gpu_ab.py)
    if not os.path.exists(path):
        raise KartB200Error("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.kb_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.kb_destroy.argtypes = [C.c_void_p]
    lib.kb_destroy.restype = None
    lib.kb_strerror.restype = C.c_char_p
    lib.kb_strerror.argtypes = [C.c_int]
    lib.kb_last_error.restype = C.c_char_p
    lib.kb_last_error.argtypes = [C.c_void_p]
    lib.kb_upload_index.argtypes = [C.c_void_p, C.POINTER(KbIndexHost), C.c_int]
    lib.kb_set_params.argtypes = [C.c_void_p, C.POINTER(KbParams)]
    lib.kb_clone_index.argtypes = [C.c_void_p, C.c_void_p]
    lib.kb_get_min_seed_len.argtypes = [C.c_void_p]
    lib.kb_map_chunk.argtypes = [C.c_void_p, C.POINTER(KbReads), C.c_void_p, C.POINTER(KbResults)]
    lib.kb_stage_reads.argtypes = [C.c_void_p, C.POINTER(KbReads), C.c_void_p]
    lib.kb_map_chunk_packed.argtypes = [C.c_void_p, C.POINTER(KbReadsPacked), C.c_void_p, C.POINTER(KbResults)]
    lib.kb_stage_reads_packed.argtypes = [C.c_void_p, C.POINTER(KbReadsPacked), C.c_void_p]
    lib.kb_map_chunk_begin.argtypes = [C.c_void_p, C.POINTER(KbReads), C.c_void_p, C.POINTER(KbResults), C.POINTER(C.c_int)]
    lib.kb_map_chunk_begin_packed.argtypes = [C.c_void_p, C.POINTER(KbReadsPacked), C.c_void_p, C.POINTER(KbResults), C.POINTER(C.c_int)]
    lib.kb_map_chunk_end.argtypes = [C.c_void_p, C.c_int]
    lib.kb_packed_words.restype = C.c_uint64
    lib.kb_packed_words.argtypes = [C.POINTER(KbReads)]
    lib.kb_pack_reads.argtypes = [C.POINTER(KbReads), C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(KbReadsPacked)]
    lib.kb_run.argtypes = [C.c_void_p]
    lib.kb_fetch_results.argtypes = [C.c_void_p, C.POINTER(KbResults)]
    lib.kb_fetch_extra.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.kb_stage_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.kb_work.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.kb_cuda_stream.restype = C.c_void_p
    lib.kb_cuda_stream.argtypes = [C.c_void_p]
    lib.kb_debug_fetch.restype = C.c_int64
    lib.kb_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
    lib.kb_debug_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32]
    return lib


class Mapper:
    """One device context: upload an index, then map chunks of reads (kb_map_chunk)."""

    def __init__(self, device: int = 0, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        h = C.c_void_p()
        rc = self.lib.kb_init(device, C.byref(h))
        if rc != 0:
            raise KartB200Error("kb_init: " + self.lib.kb_strerror(rc).decode())
        self.h = h
        self.index = None
        self.params = KbParams(0, 5, 1500, 0, 0, 0)
        self.n_reads = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.kb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise KartB200Error("%s: %s (%s)" % (what, self.lib.kb_strerror(rc).decode(), self.lib.kb_last_error(self.h).decode()))

    def upload_index(self, idx: KartIndex, expand_sa: bool = False):
        hi = KbIndexHost()
        hi.primary = idx.primary
        for i in range(5):
            hi.L2[i] = idx.L2[i]
        hi.seq_len = idx.seq_len
        hi.bwt = idx.bwt.ctypes.data
        hi.bwt_words = len(idx.bwt)
        hi.sa = idx.sa.ctypes.data
        hi.n_sa = len(idx.sa)
        hi.sa_intv = idx.sa_intv
        hi.pac = idx.pac.ctypes.data
        hi.l_pac = idx.l_pac
        hi.n_chr = len(idx.chr_len)
        hi.chr_len = idx.chr_len_arr.ctypes.data
        self._check(self.lib.kb_upload_index(self.h, C.byref(hi), 1 if expand_sa else 0), "kb_upload_index")
        self.index = idx
        self.set_params()

    def clone_index_from(self, other: "Mapper"):
        """kb_clone_index: this device's replica of the index `other` holds, copied device to device."""
        self._check(self.lib.kb_clone_index(self.h, other.h), "kb_clone_index")
        self.index = other.index
        self.set_params()

    def set_params(self, pacbio: bool = False, paired: bool = False, max_gaps: int = 5, multihit: bool = False, min_seed_len: int = 0):
        self.params = KbParams(min_seed_len, max_gaps, 1500, int(pacbio), int(paired), int(multihit))
        self._check(self.lib.kb_set_params(self.h, C.byref(self.params)), "kb_set_params")

    @property
    def min_seed_len(self) -> int:
        return self.lib.kb_get_min_seed_len(self.h)

    @staticmethod
    def pack_reads(reads):
        """reads: 2-D uint8 array [n, L] or list of bytes -> (flat uint8, offsets uint64)."""
        if isinstance(reads, np.ndarray) and reads.ndim == 2:
            n, L = reads.shape
            return np.ascontiguousarray(reads).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
        lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        flat = np.frombuffer(b"".join(bytes(r) for r in reads), dtype=np.uint8).copy() if len(reads) else np.zeros(0, np.uint8)
        return flat, off

    def _reads_struct(self, flat, off):
        r = KbReads()
        r.n_reads = len(off) - 1
        r.seq = flat.ctypes.data
        r.seq_off = off.ctypes.data
        return r

    def _results(self, n, cap_cigar, out=None):
        """Result buffers: freshly allocated, or views of caller-owned (e.g. pinned) buffers `out = (aln, pairs, cigar)`."""
        if out is not None:
            aln, pairs, cig = out
            assert len(aln) >= n and len(pairs) >= n // 2
            aln, pairs = aln[:n], pairs[:max(n // 2, 1)]
        else:
            aln = np.empty(n, dtype=ALN_DTYPE)
            pairs = np.empty(max(n // 2, 1), dtype=PAIR_DTYPE)
            cig = np.empty(max(cap_cigar, 1), dtype=np.uint32)
        res = KbResults(aln.ctypes.data, pairs.ctypes.data, cig.ctypes.data, len(cig), 0)
        return aln, pairs, cig, res

    def _est(self, n, est):
        if not self.params.paired:
            return None
        if isinstance(est, np.ndarray) and est.dtype == np.int32 and len(est) == n // 2 and est.flags.c_contiguous:
            return est
        return np.ascontiguousarray(np.broadcast_to(np.asarray(1500 if est is None else est, dtype=np.int32), (n // 2,)))

    def pack(self, flat, off, code=None, threads=8):
        """kb_pack_reads: (KbReadsPacked, the arrays it borrows). code: optional caller-owned (e.g. pinned) uint64 buffer."""
        reads = self._reads_struct(flat, off)
        nw = int(self.lib.kb_packed_words(C.byref(reads)))
        if code is None:
            code = np.zeros(nw, dtype=np.uint64)
        assert len(code) >= nw
        exc = np.zeros(max(1024, len(flat) // 64), dtype=np.uint64)
        pk = KbReadsPacked()
        rc = self.lib.kb_pack_reads(C.byref(reads), code.ctypes.data, exc.ctypes.data, len(exc), threads, C.byref(pk))
        if rc == -6:
            exc = np.zeros(int(pk.n_exc), dtype=np.uint64)
            rc = self.lib.kb_pack_reads(C.byref(reads), code.ctypes.data, exc.ctypes.data, len(exc), threads, C.byref(pk))
        self._check(rc, "kb_pack_reads")
        return pk, (code, exc, off, flat)

    def map_chunk(self, flat, off, est=None, out=None, packed=None):
        """Host buffers in, host buffers out (the drop-in call). Returns (aln, pairs, cigar). packed: a (KbReadsPacked, keep-alive)
        pair from pack() sends the 2-bit form over PCIe instead of the text (kb_map_chunk_packed); True packs here first."""
        n = len(off) - 1
        self.n_reads = n
        est_arr = self._est(n, est)
        if packed is None and os.environ.get("KART_TEST_PACKED"):
            packed = True
        if packed is True:
            packed = self.pack(flat, off)
        reads = self._reads_struct(flat, off)
        aln, pairs, cig, res = self._results(n, 4 * n + 1024, out)
        if packed:
            rc = self.lib.kb_map_chunk_packed(self.h, C.byref(packed[0]), est_arr.ctypes.data if est_arr is not None else None, C.byref(res))
        else:
            rc = self.lib.kb_map_chunk(self.h, C.byref(reads), est_arr.ctypes.data if est_arr is not None else None, C.byref(res))
        if rc == -6 and out is None:   # KB_ECAPACITY: results are still on the device, fetch again with a big enough cigar buffer
            aln, pairs, cig, res = self._results(n, int(res.n_cigar))
            rc = self.lib.kb_fetch_results(self.h, C.byref(res))
        self._check(rc, "kb_map_chunk")
        return aln, pairs, cig[:res.n_cigar]

    def map_chunk_begin(self, flat, off, est=None, out=None, packed=None):
        """kb_map_chunk_begin[_packed]: returns a handle for map_chunk_end(). The buffers are kept alive by the handle."""
        n = len(off) - 1
        est_arr = self._est(n, est)
        if packed is True:
            packed = self.pack(flat, off)
        reads = self._reads_struct(flat, off)
        aln, pairs, cig, res = self._results(n, 4 * n + 1024, out)
        ticket = C.c_int(-1)
        if packed:
            rc = self.lib.kb_map_chunk_begin_packed(self.h, C.byref(packed[0]), est_arr.ctypes.data if est_arr is not None else None, C.byref(res), C.byref(ticket))
        else:
            rc = self.lib.kb_map_chunk_begin(self.h, C.byref(reads), est_arr.ctypes.data if est_arr is not None else None, C.byref(res), C.byref(ticket))
        self._check(rc, "kb_map_chunk_begin")
        return (ticket.value, aln, pairs, cig, res, (flat, off, est_arr, packed, reads))

    def map_chunk_end(self, handle):
        ticket, aln, pairs, cig, res, _keep = handle
        self._check(self.lib.kb_map_chunk_end(self.h, ticket), "kb_map_chunk_end")
        self.n_reads = len(aln)
        return aln, pairs, cig[:res.n_cigar]

    def stage(self, flat, off, est=None, packed=False):
        n = len(off) - 1
        self.n_reads = n
        self._keep = (flat, off)
        est_arr = self._est(n, est)
        if packed or os.environ.get("KART_TEST_PACKED"):
            pk, keep = self.pack(flat, off)
            self._check(self.lib.kb_stage_reads_packed(self.h, C.byref(pk), est_arr.ctypes.data if est_arr is not None else None), "kb_stage_reads_packed")
            return
        reads = self._reads_struct(flat, off)
        self._check(self.lib.kb_stage_reads(self.h, C.byref(reads), est_arr.ctypes.data if est_arr is not None else None), "kb_stage_reads")

    def run(self):
        self._check(self.lib.kb_run(self.h), "kb_run")

    def fetch(self):
        n = self.n_reads
        aln, pairs, cig, res = self._results(n, 4 * n + 1024)
        rc = self.lib.kb_fetch_results(self.h, C.byref(res))
        if rc == -6:
            aln, pairs, cig, res = self._results(n, int(res.n_cigar))
            rc = self.lib.kb_fetch_results(self.h, C.byref(res))
        self._check(rc, "kb_fetch_results")
        return aln, pairs, cig[:res.n_cigar]

    def fetch_extra(self):
        """-m: the further lines of the last chunk as a structured array (read, rank, aln), sorted by (read, rank)."""
        n = C.c_uint32(0)
        rc = self.lib.kb_fetch_extra(self.h, None, 0, C.byref(n))
        ext = np.zeros(n.value, dtype=EXTRA_DTYPE)
        if rc == -6:
            rc = self.lib.kb_fetch_extra(self.h, ext.ctypes.data, n.value, C.byref(n))
        self._check(rc, "kb_fetch_extra")
        return ext

    def stage_ms(self):
        a = np.zeros(10, dtype=np.float32)
        self.lib.kb_stage_ms(self.h, a.ctypes.data, 10)
        return dict(zip(["fm_seed", "sa_locate", "cand_pair", "rescue", "segments", "align", "assemble", "finalize", "total", "nw"], [float(x) for x in a]))

    def work(self):
        a = np.zeros(8, dtype=np.uint64)
        self.lib.kb_work(self.h, a.ctypes.data, 8)
        return dict(zip(["ext_steps", "occ_blocks", "lf_steps", "nw_cells", "seeds", "nw_calls", "rescues", "launches"], [int(x) for x in a]))

    def debug(self, what: int, dtype, max_items: int):
        buf = np.zeros(max_items, dtype=dtype)
        got = self.lib.kb_debug_fetch(self.h, what, buf.ctypes.data, buf.nbytes)
        if got < 0:
            raise KartB200Error("kb_debug_fetch(%d): %s" % (what, self.lib.kb_strerror(int(got)).decode()))
        return buf[:got // buf.itemsize]

    def debug_align(self, frags):
        """Stage-level test hook (kb_debug_align): frags = [(read, rpos, rlen, gpos, glen, mode)] on the staged reads. Returns
        (kb_dbg_frag_out_t array, per-fragment cigar strings)."""
        spec = np.zeros(len(frags), dtype=DBG_FRAG_DTYPE)
        for i, (r, rpos, rlen, gpos, glen, mode) in enumerate(frags):
            spec[i] = (r, rpos, rlen, glen, mode, 0, gpos)
        out = np.zeros(len(frags), dtype=DBG_OUT_DTYPE)
        cap = int((spec["rlen"] + spec["glen"] + 4).sum())
        ops = np.zeros(cap, dtype=np.uint32)
        self._check(self.lib.kb_debug_align(self.h, spec.ctypes.data, len(frags), out.ctypes.data, ops.ctypes.data, cap), "kb_debug_align")
        return out, [cigar_string(ops, int(o["ops_off"]), int(o["n_ops"])) if o["n_ops"] >= 0 else None for o in out]

    def hits(self):
        """The searches of the last run that yield seeds, per read: [(rpos, len, freq, x0)] (BWT_Search results, bwt_search.cpp:171-181)."""
        n = self.n_reads
        mh = int(self.debug(13, np.int32, 1)[0])
        nh = self.debug(12, np.int32, n)
        h = self.debug(11, HIT_DTYPE, n * mh).reshape(n, mh)
        return [[(int(x["rpos"]), int(x["len_freq"]) >> 8, int(x["len_freq"]) & 255, int(x["x0"])) for x in h[r, :nh[r]]] for r in range(n)]

    # ---- dumps in the oracle's text format (oracle/kart_oracle.h) ----
    def dump_state(self):
        n = self.n_reads
        cnt = self.debug(9, np.uint32, 16)
        st = {"n_seeds": self.debug(0, np.int32, n), "seed_off": self.debug(1, np.uint32, n), "segs": self.debug(2, SEG_DTYPE, int(cnt[0])),
              "n_cands": self.debug(3, np.int32, n), "cand_off": self.debug(4, np.uint32, n), "cands": self.debug(5, CAND_DTYPE, int(cnt[1])),
              "reports": self.debug(6, REPORT_DTYPE, int(cnt[1])), "res": self.debug(7, RES_DTYPE, n), "cigar": self.debug(8, np.uint32, int(cnt[2]))}
        return st


def cigar_string(cig: np.ndarray, off: int, n: int) -> str:
    return "".join("%d%s" % (int(c) >> 4, CIGAR_OPS[int(c) & 15]) for c in cig[off:off + n])


def dump_read(st, r: int) -> str:
    """'R'/'A' lines of read r, identical to oracle dump_read()."""
    rd = st["res"][r]
    out = ["R %d %d %d %d %d" % (rd["score"], rd["sub"], rd["mapq"], rd["ncan"], rd["best"])]
    for i in range(int(rd["ncan"])):
        a = st["reports"][int(rd["rep_off"]) + i]
        ln = "A %d %d %d" % (i, a["aln"], a["mate"])
        if (rd["score"] == 0 and i == 0) or (rd["score"] > 0 and i == rd["best"]):
            ln += " F%d" % a["flag"]
        if a["aln"] > 0:
            ln += " %d %d %d %s" % (a["fwd"], a["chr"], a["pos"], cigar_string(st["cigar"], int(a["cig_off"]), int(a["cig_len"])))
        out.append(ln)
    return "\n".join(out) + "\n"


def dump_cands(st, r: int) -> str:
    out = []
    for i in range(int(st["n_cands"][r])):
        c = st["cands"][int(st["cand_off"][r]) + i]
        out.append("C %d %d %d %d" % (c["score"], c["diff"], c["mate"], c["nseg"]))
        for k in range(int(c["nseg"])):
            s = st["segs"][int(c["seg_start"]) + k]
            out.append("S %d %d %d %d %d" % (s["rpos"], s["rlen"], s["glen"], s["gpos"], s["simple"]))
    return "\n".join(out) + ("\n" if out else "")
