"""Host-side loader for the BWA-format index files Kart reads (reference src/bwt_index.cpp:16-36, 38-90, 103-122, 230-259).

Returns plain numpy arrays; nothing here computes on the hot path."""
from __future__ import annotations

import numpy as np


class KartIndex:
    """.bwt/.sa/.pac/.ann of one index prefix, as the reference holds them in bwt_t / bntseq_t."""

    def __init__(self, prefix: str):
        self.prefix = prefix
        raw = np.fromfile(prefix + ".bwt", dtype=np.uint8)
        hdr = raw[:40].view(np.uint64)
        self.primary = int(hdr[0])
        self.L2 = [0] + [int(x) for x in hdr[1:5]]
        self.seq_len = self.L2[4]
        self.bwt = np.ascontiguousarray(raw[40:].view(np.uint32))
        sa_raw = np.fromfile(prefix + ".sa", dtype=np.uint64)
        self.sa_intv = int(sa_raw[5])
        n_sa = (self.seq_len + self.sa_intv) // self.sa_intv
        self.sa = np.empty(n_sa, dtype=np.uint64)
        self.sa[0] = np.uint64(0xFFFFFFFFFFFFFFFF)          # bwt_restore_sa: sa[0] = -1
        body = sa_raw[7:7 + n_sa - 1]
        self.sa[1:1 + len(body)] = body
        with open(prefix + ".ann") as fh:
            tok = fh.readline().split()
            self.l_pac, n_seqs = int(tok[0]), int(tok[1])
            self.chr_names, self.chr_len = [], []
            for _ in range(n_seqs):
                self.chr_names.append(fh.readline().split()[1])
                self.chr_len.append(int(fh.readline().split()[1]))
        pac = np.fromfile(prefix + ".pac", dtype=np.uint8)
        need = self.l_pac // 4 + 1
        self.pac = np.zeros(need, dtype=np.uint8)
        self.pac[:min(need, len(pac))] = pac[:need]
        self.chr_len_arr = np.asarray(self.chr_len, dtype=np.int64)

    @property
    def min_seed_len(self) -> int:                           # src/Mapping.cpp:645
        m = 13
        while m < 16 and not (2 * self.l_pac < 4 ** m):
            m += 1
        return m
