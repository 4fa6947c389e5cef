#!/usr/bin/env python
"""Known answers the REFERENCE ITSELF holds for the weighted circumcentre (src/Algorithms.jl:56-96, row A6 of SURVEY 8a):
the Mathematica notebook src/hypercube-weights.nb (outputs saved in the file, Mathematica 12.1) derives the level weights
of large_hypercube_manifold (src/ManifoldConstructors.jl:253-271).  For the Kuhn path simplex
x_0 = (0,...,0), x_1 = (0,...,0,1), ..., x_D = (1,...,1) with weights (0, w_2, ..., w_D, 0) it solves
|x_i - c|^2 - w_i = cr2 for the centre c and cr2 (Out[12], Out[19], Out[26], Out[33]: c as a function of the weights), then
the weights that put c at the barycentre (Out[13], Out[20], Out[27], Out[34]) and the resulting centre / cr2 (Out[14],
Out[21], Out[28], Out[35]).  This script transcribes those OUTPUT cells (rationals) into hypercube_weights.json; nothing
here is computed by the oracle.

NOTE on the notebook's equation: it is written `Total[(xs[[i]] - cc)^2 - ws[[i]]] == cr2`, and Mathematica subtracts the
scalar ws[[i]] from EVERY component before summing, so what it solves is |x_i - c|^2 - D w_i = cr2: the notebook's w
enters with a factor D.  The reference's Julia code solves |x_i - c|^2 - w_i = r^2 (src/Algorithms.jl:56-96,
b_i = x_i.x_i - w_i) and the generator feeds it the notebook's numbers unchanged (ManifoldConstructors.jl:253-260), so in
the reference the weighted circumcentre of a Kuhn simplex is NOT its barycentre (D = 2: (5/12, 7/12) instead of
(1/3, 2/3)) although the comment at :252 says so.  The product follows the Julia code; the pin applies the factor D when
it compares the oracle's solver with the notebook's outputs."""
import json
import os
from fractions import Fraction as F

out = {
    # D: weights of the interior levels, centre, cr2   (Out[13]/[14], [20]/[21], [27]/[28], [34]/[35])
    "solved": {
        "2": {"w": [F(-1, 6)], "centre": [F(1, 3), F(2, 3)], "cr2": F(5, 9)},
        "3": {"w": [F(-1, 6), F(-1, 6)], "centre": [F(1, 4), F(1, 2), F(3, 4)], "cr2": F(7, 8)},
        "4": {"w": [F(-3, 20), F(-1, 5), F(-3, 20)], "centre": [F(1, 5), F(2, 5), F(3, 5), F(4, 5)], "cr2": F(6, 5)},
        "5": {"w": [F(-2, 15), F(-1, 5), F(-1, 5), F(-2, 15)],
              "centre": [F(1, 6), F(1, 3), F(1, 2), F(2, 3), F(5, 6)], "cr2": F(55, 36)},
    },
    # the general solutions (Out[12], [19], [26], [33]) as coefficient rows: centre_k = 1/2 (1 + D * sum_j M[k][j] w_{j+2}),
    # cr2 = D/4 * (1 + 2 D * (sum w_j^2 - sum w_j w_{j+1}))
    "general": {
        "2": {"M": [[1], [-1]]},
        "3": {"M": [[0, 1], [1, -1], [-1, 0]]},
        "4": {"M": [[0, 0, 1], [0, 1, -1], [1, -1, 0], [-1, 0, 0]]},
        "5": {"M": [[0, 0, 0, 1], [0, 0, 1, -1], [0, 1, -1, 0], [1, -1, 0, 0], [-1, 0, 0, 0]]},
    },
    # D = 1 (Out[5]): centre 1/2, cr2 1/4, no free weight
    "d1": {"centre": [F(1, 2)], "cr2": F(1, 4)},
}


def enc(o):
    if isinstance(o, F):
        return [o.numerator, o.denominator]
    if isinstance(o, dict):
        return {k: enc(v) for k, v in o.items()}
    if isinstance(o, list):
        return [enc(v) for v in o]
    return o


with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hypercube_weights.json"), "w") as f:
    json.dump(enc(out), f, indent=1)
