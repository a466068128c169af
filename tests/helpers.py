"""Shared test helpers: seeded inputs and canonical forms."""
import numpy as np
import torch


def dense_cube(n, batch=0):
    g = np.arange(n)
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    return np.stack([np.full(n ** 3, batch), xx.ravel(), yy.ravel(), zz.ravel()], 1).astype(np.int32)


def random_sparse_coords(rng, n, extent=24, batches=2, negative=True, duplicates=0):
    lo = -extent // 2 if negative else 0
    c = np.unique(np.concatenate([rng.integers(0, batches, (n, 1)), rng.integers(lo, lo + extent, (n, 3))], 1), axis=0)
    rng.shuffle(c)
    if duplicates:
        c = np.concatenate([c, c[rng.integers(0, c.shape[0], duplicates)]], 0)
        rng.shuffle(c)
    return c.astype(np.int32)


def pair_set(pairs, cin, cout):
    """canonical kernel map: sorted set of (k, coord_in, coord_out) triples"""
    out = set()
    for k, (ii, oo) in enumerate(pairs):
        ii = np.asarray(ii.cpu() if isinstance(ii, torch.Tensor) else ii)
        oo = np.asarray(oo.cpu() if isinstance(oo, torch.Tensor) else oo)
        for a, b in zip(cin[ii], cout[oo]):
            out.add((k, tuple(a), tuple(b)))
    return out


def pairs_array(pairs):
    """per-offset (in,out) pairs sorted by out row — comparable row-for-row when row orders agree"""
    res = []
    for ii, oo in pairs:
        ii = np.asarray(ii.cpu() if isinstance(ii, torch.Tensor) else ii).astype(np.int64)
        oo = np.asarray(oo.cpu() if isinstance(oo, torch.Tensor) else oo).astype(np.int64)
        o = np.argsort(oo, kind="stable")
        res.append(np.stack([ii[o], oo[o]], 1))
    return res


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()
