#!/usr/bin/env python
"""Sector floor of execute_cache_moves for a given move list.

The DRAM traffic no implementation of the reference's cache layout can go below: every distinct 32-byte
sector that holds a byte the moves READ, every sector they WRITE, plus a fill read for every written sector
that is only partly overwritten (L2 has to fetch it before the partial write can go back to DRAM).

Layout (vllm/attention/ops/paged_attn.py:256-284; one block = 16 tokens of one kv head, 2-byte elements):
  k_cache [NB, 16 j, 16 t, 8]  token t of block b: sixteen 16-byte chunks at  b*4096 + j*256 + t*16
                               -> 16 sectors per token, a sector is shared by the token pair (2i, 2i+1)
  v_cache [NB, 128 d, 16 t]    token t of block b: 128 two-byte elements at   b*4096 + d*32  + t*2
                               -> ALL 128 sectors of the block for any one token
  kv_metrics / kv_position [NB, 16] 4-byte: sector = 8 consecutive slots
  cache_moves_idx [N, 2] int32: 8 bytes per move, read once

    sector_floor(moves) -> dict(read_bytes, fill_bytes, write_bytes, floor_bytes, per_move, ...)
`moves` = int array [n, 2] of (dst slot, src slot), the concatenation of every head's [0, count) entries.
Pure numpy; used by bench.py (per CR: floor_bytes, frac_of_floor) and runnable on its own for a synthetic
post-prefill schedule:  python tools/moves_floor.py [--ctx 6000 --keep 1488 --heads 64]
"""
import argparse
import json

import numpy as np

B = 16
SECTOR = 32
ALGORITHMIC_BYTES_PER_MOVE = 1048


def sector_floor(moves: np.ndarray, fetch_sectors: int = 1) -> dict:
    """fetch_sectors: DRAM read granularity in 32 B sectors.  1 = the layout's floor; 2 = what the device reads
    with the default cudaLimitMaxL2FetchGranularity of 64 B (an L2 miss brings in the aligned sector PAIR), which
    doubles the cost of the 16-byte K chunks and of the 4-byte metric / position reads.  Writes stay per sector."""
    moves = np.asarray(moves).reshape(-1, 2).astype(np.int64)
    n = moves.shape[0]
    if n == 0:
        return dict(moves=0, read_bytes=0, fill_bytes=0, write_bytes=0, floor_bytes=0, per_move=0.0)
    dst, src = moves[:, 0], moves[:, 1]
    db, do = dst // B, dst % B
    sb, so = src // B, src % B
    # ---- V: whole blocks ----
    v_src_blocks = np.unique(sb)
    v_dst_blocks, inv = np.unique(db, return_inverse=True)
    dmask = np.zeros(v_dst_blocks.shape[0], np.int64)
    np.bitwise_or.at(dmask, inv, 1 << do)
    v_full = dmask == 0xffff
    # a destination block that is not fully overwritten must be fetched; if it is also a source block the fetch
    # is already counted
    v_fill_blocks = np.setdiff1d(v_dst_blocks[~v_full], v_src_blocks, assume_unique=True)
    v_read = v_src_blocks.shape[0] * 128
    v_fill = v_fill_blocks.shape[0] * 128
    v_write = v_dst_blocks.shape[0] * 128
    # ---- K: sector (block, j, pair) for j in 0..15; count per (block, pair) and multiply by 16 ----
    fs = int(fetch_sectors)
    k_src_pairs = np.unique(sb * 8 + so // 2)
    dpair, pinv = np.unique(db * 8 + do // 2, return_inverse=True)
    pmask = np.zeros(dpair.shape[0], np.int64)
    np.bitwise_or.at(pmask, pinv, 1 << (do % 2))
    # reads happen in granules of fs sectors (a K row [j] of a block = 8 sectors = 16 tokens)
    k_src_gran = np.unique(k_src_pairs // fs)
    k_fill_gran = np.setdiff1d(np.unique(dpair[pmask != 3] // fs), k_src_gran, assume_unique=True)
    k_read = k_src_gran.shape[0] * 16 * fs
    k_fill = k_fill_gran.shape[0] * 16 * fs
    k_write = dpair.shape[0] * 16
    # ---- metrics + positions: sector = 8 slots, two arrays ----
    m_src = np.unique(src // 8)
    m_dst, minv = np.unique(dst // 8, return_inverse=True)
    mm = np.zeros(m_dst.shape[0], np.int64)
    np.bitwise_or.at(mm, minv, 1 << (dst % 8))
    m_src_gran = np.unique(m_src // fs)
    m_fill = np.setdiff1d(np.unique(m_dst[mm != 0xff] // fs), m_src_gran, assume_unique=True)
    mp_read = 2 * m_src_gran.shape[0] * fs
    mp_fill = 2 * m_fill.shape[0] * fs
    mp_write = 2 * m_dst.shape[0]
    list_bytes = 8 * n
    read = (v_read + k_read + mp_read) * SECTOR + list_bytes
    fill = (v_fill + k_fill + mp_fill) * SECTOR
    write = (v_write + k_write + mp_write) * SECTOR
    floor = read + fill + write
    return dict(moves=int(n), read_bytes=int(read), fill_bytes=int(fill), write_bytes=int(write),
                floor_bytes=int(floor), per_move=float(floor) / n,
                algorithmic_bytes=int(n) * ALGORITHMIC_BYTES_PER_MOVE,
                amplification=float(floor) / (n * ALGORITHMIC_BYTES_PER_MOVE),
                src_blocks=int(v_src_blocks.shape[0]), dst_blocks=int(v_dst_blocks.shape[0]),
                full_dst_blocks=int(v_full.sum()))


def gather_moves(cache_moves: np.ndarray, counts: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """[0, count) of every head segment of the K4 output (cache_moves_idx [N,2], counts / offsets [S,L,H])."""
    cache_moves = np.asarray(cache_moves).reshape(-1, 2)
    c, o = np.asarray(counts).reshape(-1).astype(np.int64), np.asarray(offsets).reshape(-1).astype(np.int64)
    idx = np.repeat(o - np.concatenate([[0], np.cumsum(c)[:-1]]), c) + np.arange(int(c.sum()))
    return cache_moves[idx]


def synthetic_schedule(ctx: int, keep: int, heads: int, protected: int = 32, seed: int = 0) -> np.ndarray:
    """Post-prefill compaction of `heads` heads of `ctx` tokens down to `keep` each (uniformly random evictions
    outside the protected window, scattered physical blocks), in K4's order (kvcompress_eviction_kernels.cu:252-288:
    k-th smallest evicted index below the new length <- k-th largest survivor above it)."""
    rng = np.random.RandomState(seed)
    nblk = (ctx + B - 1) // B
    perm = rng.permutation(heads * nblk).reshape(heads, nblk)
    out = []
    for h in range(heads):
        ev = np.sort(rng.choice(ctx - protected, ctx - keep, replace=False))
        new_n = keep
        dst = ev[ev < new_n]
        surv = np.setdiff1d(np.arange(new_n, ctx), ev, assume_unique=True)[::-1]
        m = dst.shape[0]
        assert surv.shape[0] == m
        phys = lambda i: perm[h, i // B].astype(np.int64) * B + i % B
        out.append(np.stack([phys(dst), phys(surv[:m])], 1))
    return np.concatenate(out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ctx", type=int, default=6000)
    ap.add_argument("--heads", type=int, default=64)
    ap.add_argument("--keep", type=int, nargs="*", default=[128, 368, 736, 1488])
    a = ap.parse_args()
    for keep in a.keep:
        mv = synthetic_schedule(a.ctx, keep, a.heads)
        r = sector_floor(mv)
        r["per_move_fetch64"] = sector_floor(mv, 2)["per_move"]
        print(json.dumps(dict(keep=keep, **r)))
