#!/usr/bin/env python
"""Hand-derived known-answer case: the unit square cut into triangles {1,2,3} and {2,3,4} (1-based vertex ids).
Everything below follows the reference's RULES (src/Manifolds.jl:717-768), written out without the oracle:
faces of a simplex = its vertex set minus the n-th smallest vertex, sign (-1)^(n-1); face id = lexicographic
rank among the unique faces; d_0 = adjoint of boundary_1 (src/ManifoldOps.jl:46)."""
import itertools
import json
import os

tris = [(1, 2, 3), (2, 3, 4)]
edges = sorted({f for t in tris for f in itertools.combinations(t, 2)})
edge_id = {e: i + 1 for i, e in enumerate(edges)}
boundary2 = []          # per triangle column: rows (edge ids ascending) and signs
for t in tris:
    col = []
    for n in range(1, 4):                      # omit the n-th smallest vertex
        face = tuple(v for k, v in enumerate(t, 1) if k != n)
        col.append((edge_id[face], (-1) ** (n - 1)))
    col.sort()
    boundary2.append({"rows": [r for r, _ in col], "signs": [s for _, s in col]})
boundary1 = [{"rows": [a, b], "signs": [-1, 1]} for a, b in edges]   # omit 1st (-> vertex b, +) / 2nd (-> a, -), rows ascending
x = [1.0, 4.0, 9.0, 16.0]
d0x = [x[b - 1] - x[a - 1] for a, b in edges]
out = {"vertices": 4, "coords": [[0, 0], [1, 0], [0, 1], [1, 1]], "triangles": [list(t) for t in tris],
       "edges": [list(e) for e in edges], "boundary1": boundary1, "boundary2": boundary2, "x": x, "d0x": d0x,
       "area_per_triangle": 0.5}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "two_triangles.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
