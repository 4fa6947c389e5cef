#!/usr/bin/env python
"""Closed-form Hodge stars and Laplacian of ONE regular simplex with unit edges (simplex_manifold(Val(D)),
src/ManifoldConstructors.jl:30-74; zero weights), worked out by hand from the definitions the reference uses
(circumcentric duals, src/Manifolds.jl:1160-1223: dualvol_R = 1/(D-R) * sum over cofaces of dualvol_{R+1} * distance
between circumcentres) -- NOT from the oracle:

D = 2 (equilateral triangle, area A = sqrt(3)/4, centroid = circumcentre):
  dualvol_2 = 1; dualvol_1 = centroid -> edge midpoint = sqrt(3)/6;
  dualvol_0 = 1/2 * 2 edges * dualvol_1 * (half edge = 1/2) = sqrt(3)/12 (= A/3).
D = 3 (regular tetrahedron, V = sqrt(2)/12, inradius r = sqrt(6)/12):
  dualvol_3 = 1; dualvol_2 = circumcentre -> face centroid = r = sqrt(6)/12;
  dualvol_1 = 1/2 * 2 faces * dualvol_2 * (face centroid -> edge midpoint = sqrt(3)/6) = sqrt(2)/24 (= V/2);
  dualvol_0 = 1/3 * 3 edges * dualvol_1 * 1/2 = sqrt(2)/48 (= V/4).
hodge(Pr,R) = dualvol_R / vol_R (src/ManifoldOps.jl:58-66); laplace(Pr,0) = -hodge0^-1 d0' hodge1 d0
(src/ManifoldOps.jl:110-144): off-diagonal hodge1/hodge0 = 2, diagonal -2 D.
Writes regular_simplex.json next to this file."""
import json
import math
import os

s2, s3, s6 = math.sqrt(2), math.sqrt(3), math.sqrt(6)
out = {
    "2": {"vol": [1.0, 1.0, s3 / 4], "dualvol": [s3 / 12, s3 / 6, 1.0],
          "hodge": [s3 / 12, s3 / 6, 4 / s3], "laplace_offdiag": 2.0, "laplace_diag": -4.0},
    "3": {"vol": [1.0, 1.0, s3 / 4, s2 / 12], "dualvol": [s2 / 48, s2 / 24, s6 / 12, 1.0],
          "hodge": [s2 / 48, s2 / 24, s2 / 3, 6 * s2], "laplace_offdiag": 2.0, "laplace_diag": -6.0},
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "regular_simplex.json"), "w") as f:
    json.dump(out, f, indent=1)
