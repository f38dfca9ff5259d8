"""TEST / BENCH INFRASTRUCTURE — the literal "reference CPU `lofreq call`" of BASELINE.json's metric.

The reference cannot be built from source here (htslib is not vendored, no network), but its own release tarball
dist/lofreq_star-2.1.4_linux-x86-64.tgz holds a statically linked `lofreq` that runs in this image.
oracle/Makefile extracts that binary to oracle/_ref/lofreq-2.1.4 (git-ignored; travels to the GPU box like the other
compiled checkers).  This module writes a synthetic, coordinate-sorted BAM of the same column model as the benchmark
(uniform depth, per-base errors with probability 10^(-q/10), 1 % variant positions with log-uniform AF) with a small
pure-Python BGZF/BAM writer, runs `lofreq call [-B] --no-default-filter` on it and reports pileup columns per second.
2.1.4 and HEAD share pruned_calc_prob_dist / poissbin / snpcaller unchanged."""
import os
import re
import struct
import subprocess
import tempfile
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(HERE, "_ref", "lofreq-2.1.4")


def have_cli():
    return os.path.exists(BINARY) and os.access(BINARY, os.X_OK)


# ---- BGZF / BAM ---------------------------------------------------------------------------------------------------
def _bgzf_block(data):
    comp = zlib.compressobj(1, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = 12 + 6 + len(body) + 8 - 1            # header (12) + extra (6) + body + crc/isize (8), minus one
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def write_bam(path, ref_name, ref_len, starts, seqs, quals, reverse, mapq=60):
    """starts int[n] (sorted), seqs uint8[n][L] as 0..3 = A,C,G,T, quals uint8[n][L], reverse bool[n]"""
    n, L = seqs.shape
    text = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:%s\tLN:%d\n" % (ref_name, ref_len)).encode()
    name = ref_name.encode() + b"\0"
    head = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", len(name)) + name + struct.pack("<i", ref_len)
    code = np.array([1, 2, 4, 8], np.uint8)[seqs]                 # 4-bit codes of =ACMGRSVTWYHKDBN
    if L % 2:
        code = np.concatenate([code, np.zeros((n, 1), np.uint8)], axis=1)
    packed = (code[:, 0::2] << 4) | code[:, 1::2]
    cigar = struct.pack("<I", (L << 4) | 0)                       # L x M
    out = bytearray(head)
    blocks = []

    def flush(force=False):
        nonlocal out
        while len(out) >= 60000 or (force and len(out)):
            blocks.append(_bgzf_block(bytes(out[:60000])))
            out = out[60000:]
    for i in range(n):
        rname = b"r%d\0" % i
        pos = int(starts[i])
        rec = (struct.pack("<iiBBHHHiiii", 0, pos, len(rname), mapq, _reg2bin(pos, pos + L), 1, 16 if reverse[i] else 0, L, -1, -1, 0) +
               rname + cigar + packed[i].tobytes() + quals[i].tobytes())
        out += struct.pack("<i", len(rec)) + rec
        if len(out) >= 60000:
            flush()
    flush(force=True)
    with open(path, "wb") as f:
        for b in blocks:
            f.write(b)
        f.write(BGZF_EOF)


# ---- synthetic data of the benchmark's column model ---------------------------------------------------------------
def make_dataset(tmpdir, n_cols=20000, depth=500, read_len=100, qual=30, seed=20261017):
    """reference of n_cols + read_len bases; reads start at every position, depth/read_len of them, so that every
    position past the first read_len has exactly `depth` reads; returns (fasta, bam, number of full-depth columns)"""
    rng = np.random.default_rng(seed)
    ref_len = n_cols + read_len
    ref = rng.integers(0, 4, ref_len).astype(np.uint8)
    per_pos = depth // read_len
    starts = np.repeat(np.arange(0, ref_len - read_len + 1), per_pos)
    n = len(starts)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    seqs = ref[idx]
    # sequencing errors
    err = rng.random(seqs.shape) < 10.0 ** (-qual / 10.0)
    seqs = np.where(err, (seqs + rng.integers(1, 4, seqs.shape)) & 3, seqs).astype(np.uint8)
    # variant positions: 1 %, AF log-uniform in [0.5 %, 50 %]
    var_pos = np.flatnonzero(rng.random(ref_len) < 0.01)
    af = 0.005 * 100.0 ** rng.random(len(var_pos))
    is_var = np.zeros(ref_len, bool)
    is_var[var_pos] = True
    af_at = np.zeros(ref_len)
    af_at[var_pos] = af
    carry = is_var[idx] & (rng.random(seqs.shape) < af_at[idx])
    seqs = np.where(carry, (ref[idx] + 1) & 3, seqs).astype(np.uint8)
    quals = np.full(seqs.shape, qual, np.uint8)
    reverse = rng.random(n) < 0.5
    fasta = os.path.join(tmpdir, "ref.fa")
    with open(fasta, "w") as f:
        f.write(">chrS\n")
        s = "".join("ACGT"[b] for b in ref)
        for i in range(0, len(s), 60):
            f.write(s[i:i + 60] + "\n")
    bam = os.path.join(tmpdir, "reads.bam")
    write_bam(bam, "chrS", ref_len, starts, seqs, quals, reverse)
    return fasta, bam, ref_len


def run_cli(n_cols=20000, depth=500, baq=False, qual=30, keep=None):
    """-> dict(value columns/s, seconds, columns, tests, variants, cmd)"""
    if not have_cli():
        raise FileNotFoundError(BINARY)
    with tempfile.TemporaryDirectory() as td:
        fasta, bam, ref_len = make_dataset(td, n_cols, depth, qual=qual)
        vcf = os.path.join(td, "out.vcf")
        cmd = [BINARY, "call", "--verbose", "--no-default-filter", "-f", fasta, "-o", vcf]
        if not baq:
            cmd.insert(2, "-B")
        cmd.append(bam)
        # main_call() ends with system("lofreq filter ...") (lofreq_call.c:1506-1552): the binary must be on PATH as `lofreq`
        bindir = os.path.join(td, "bin")
        os.makedirs(bindir)
        os.symlink(BINARY, os.path.join(bindir, "lofreq"))
        env = dict(os.environ, PATH=bindir + os.pathsep + os.environ.get("PATH", ""))
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("lofreq call failed: " + r.stderr[-2000:])
        m = re.search(r"Number of substitution tests performed: (\d+)", r.stderr)
        with open(vcf) as f:
            nvar = sum(1 for ln in f if not ln.startswith("#"))
        if keep:
            import shutil
            shutil.copy(vcf, keep)
    return dict(value=ref_len / dt, unit="columns/s", seconds=dt, columns=ref_len, depth=depth, cores=1,
                tests=int(m.group(1)) if m else None, variants=nvar,
                kind="cli: prebuilt lofreq 2.1.4 (`lofreq call%s --no-default-filter`) on a synthetic BAM written in-process; BAM decode, "
                     "pileup%s and the per-column test, single thread" % (" -B" if not baq else "", " + BAQ" if baq else ""))


if __name__ == "__main__":
    import json
    import sys
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    for baq in (False, True):
        print(json.dumps(run_cli(n, baq=baq)))
