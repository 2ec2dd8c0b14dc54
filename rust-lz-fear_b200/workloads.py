"""Synthetic workloads of BASELINE.json's configs, generated with torch so that the
same code makes small CPU cases for the tests and multi-GiB device-resident cases for bench.py.

  seq50(...)   config 2: LZ4 blocks synthesised DIRECTLY as sequences — L in U{4..12} random literal
               bytes, match length M in U{4..12}, offset U{1..min(out_pos, 65535)} — each block
               decoding to exactly `block_size` bytes and ending in a literal-only sequence of
               >= 12 bytes, so the blocks are spec-conformant (~50 % literal / 50 % match bytes).
  text(...)    config 3: Zipf(s=1) words from a 4096-word vocabulary, space separated (ratio ~2).
  lowent(...)  config 4/5: 4-symbol alphabet, run lengths U{1..64}.
  random(...)  config 4: uniform bytes (stored-block fallback).
"""
import torch

_U8 = torch.uint8


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) & 0x7FFFFFFFFFFFFFFF)
    return g


def random_bytes(n, seed, device="cpu"):
    return torch.randint(0, 256, (int(n),), dtype=_U8, device=device, generator=_gen(seed, device))


def lowent(n, seed, device="cpu"):
    """bytes from a 4-symbol alphabet with run lengths U{1..64}."""
    n = int(n)
    g = _gen(seed, device)
    nruns = n // 24 + 64                      # mean run 32.5
    out = torch.empty(0, dtype=_U8, device=device)
    while out.numel() < n:
        runs = torch.randint(1, 65, (nruns,), device=device, generator=g)
        syms = torch.randint(0, 4, (nruns,), device=device, generator=g).to(_U8) + 65
        out = torch.cat([out, torch.repeat_interleave(syms, runs)])
    return out[:n].contiguous()


class TextSource:
    """Zipf(s=1) sampling of a fixed 4096-word vocabulary (word lengths U{2..11}, letters a-z)."""

    def __init__(self, seed=0x4C5A0003, device="cpu", vocab=4096):
        self.device = device
        g = _gen(seed, "cpu")
        lens = torch.randint(2, 12, (vocab,), generator=g)
        letters = torch.randint(97, 123, (vocab, 12), generator=g).to(_U8)
        letters[torch.arange(12)[None, :] >= lens[:, None]] = 32        # pad with the separating space
        self.lens = (lens + 1).to(device)                                # word + one space
        self.letters = letters.to(device)
        w = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64)
        self.weights = (w / w.sum()).to(torch.float32).to(device)
        self.g = _gen(seed ^ 0x5DEECE66D, device)

    def make(self, n):
        n = int(n)
        nwords = n // 5 + 64                  # mean word+space is ~7.5 bytes under Zipf; generous
        parts, total = [], 0
        while total < n:
            ids = torch.multinomial(self.weights, nwords, replacement=True, generator=self.g)
            lens = self.lens[ids]
            starts = torch.cumsum(lens, 0) - lens
            word_of_byte = torch.repeat_interleave(torch.arange(nwords, device=self.device), lens)
            pos_in_word = torch.arange(word_of_byte.numel(), device=self.device) - starts[word_of_byte]
            chunk = self.letters[ids[word_of_byte], pos_in_word]
            parts.append(chunk)
            total += chunk.numel()
        return torch.cat(parts)[:n].contiguous()


def text(n, seed=0x4C5A0003, device="cpu"):
    return TextSource(seed, device).make(n)


def seq50_blocks(nblocks, seed=0x4C5A0002, device="cpu", block_size=65536, stride=None):
    """-> (comp u8[nblocks*stride], in_off i64[nblocks], in_len i32[nblocks]).  Block b occupies
    comp[b*stride : b*stride + in_len[b]] and decodes to exactly `block_size` bytes."""
    nblocks = int(nblocks)
    B = int(block_size)
    assert B >= 64
    stride = int(stride or B)
    g = _gen(seed, device)
    S = (B - 12) // 8 + 1                      # sequences never exceed this (>= 8 output bytes each)
    comp = torch.randint(0, 256, (nblocks, stride), dtype=_U8, device=device, generator=g)   # literals = noise
    L = torch.randint(4, 13, (nblocks, S), device=device, generator=g)
    M = torch.randint(4, 13, (nblocks, S), device=device, generator=g)
    R = torch.randint(0, 1 << 30, (nblocks, S), device=device, generator=g)
    out_after = torch.cumsum(L + M, 1)
    out_before = out_after - (L + M)
    # keep a sequence while it still leaves >= 12 bytes for the closing literal run (24 keeps the tail a
    # little longer so that some tails need an LSIC byte)
    keep = out_after <= (B - 24)
    nseq = keep.sum(1)                                                   # per block
    csz = torch.where(keep, 1 + L + 2, torch.zeros_like(L))
    c_after = torch.cumsum(csz, 1)
    c_before = c_after - csz
    hist = torch.clamp(out_before + L, max=65535)                        # bytes addressable behind the match
    D = 1 + R % hist
    token = ((L << 4) | (M - 4)).to(_U8)
    rows = torch.arange(nblocks, device=device)[:, None].expand(-1, S)
    comp[rows[keep], c_before[keep]] = token[keep]
    off_pos = (c_before + 1 + L)
    comp[rows[keep], off_pos[keep]] = (D & 0xFF).to(_U8)[keep]
    comp[rows[keep], (off_pos + 1)[keep]] = (D >> 8).to(_U8)[keep]
    # closing literal-only sequence
    last = (nseq - 1).clamp(min=0)
    r = torch.arange(nblocks, device=device)
    out_end = torch.where(nseq > 0, out_after[r, last], torch.zeros_like(nseq))
    c_end = torch.where(nseq > 0, c_after[r, last], torch.zeros_like(nseq))
    tail = B - out_end                                                   # 24 .. 47 literal bytes
    assert int(tail.min()) >= 12 and int(tail.max()) < 15 + 255
    comp[r, c_end] = (torch.clamp(tail, max=15) << 4).to(_U8)
    has_lsic = tail >= 15
    comp[r[has_lsic], (c_end + 1)[has_lsic]] = (tail - 15).to(_U8)[has_lsic]
    in_len = (c_end + 1 + has_lsic.to(c_end.dtype) + tail)
    assert int(in_len.max()) <= stride
    in_off = torch.arange(nblocks, device=device, dtype=torch.int64) * stride
    return comp.reshape(-1), in_off, in_len.to(torch.int32)


def mixed_blocks(nblocks, block_size, seed=0x4C5A0004, device="cpu"):
    """config 4: block b's class = b mod 3 -> random / text / lowent.  -> u8[nblocks*block_size]"""
    src = TextSource(seed, device)
    out = torch.empty(nblocks * block_size, dtype=_U8, device=device)
    for b in range(nblocks):
        cls = b % 3
        if cls == 0:
            blk = random_bytes(block_size, seed + b, device)
        elif cls == 1:
            blk = src.make(block_size)
        else:
            blk = lowent(block_size, seed + b, device)
        out[b * block_size:(b + 1) * block_size] = blk
    return out
