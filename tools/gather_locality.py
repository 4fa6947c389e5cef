#!/usr/bin/env python
"""CPU analysis (oracle tables, no GPU): how many distinct 32-byte sectors / 128-byte lines one warp-level gather
instruction of the d_k kernels touches when lane = consecutive rows, per index column and for the whole warp.
Usage: python tools/gather_locality.py [n]"""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import ddf_oracle as o
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
s, c, w = o.large_hypercube_tables(3, (n, n, n))
m = o.build_manifold("k", s, c, w)
for R in (1, 2, 3):
    B = o.boundary(o.Pr, R, m).tocsc(); B.sort_indices()
    W = R + 1
    idx = B.indices.reshape(-1, W).astype(np.int64)
    nr = idx.shape[0] // 32 * 32
    g = idx[:nr].reshape(-1, 32, W)
    for name, div in (("32B sectors", 4), ("128B lines", 16)):
        per = [np.mean([len(np.unique(g[k, :, p] // div)) for k in range(0, g.shape[0], 7)]) for p in range(W)]
        allp = np.mean([len(np.unique(g[k] // div)) for k in range(0, g.shape[0], 7)])
        print(f"d{R-1}: distinct {name} per warp gather, by index column: {[round(x,1) for x in per]}; whole warp (all columns): {allp:.1f}")
