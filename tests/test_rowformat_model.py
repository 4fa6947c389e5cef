"""Host-side model of the row storages of d_k (csrc/ops.cu: k_pack_fixed, k_bits_scan / k_bits_pack): the encodings
are lossless on the oracle's tables, and the storage the model predicts for the fixtures of the GPU parity tests is
the one those tests assert on the device.  No GPU needed."""
import numpy as np
import pytest
import scipy.spatial

import ddf_oracle as o


def table(om, k):
    """Rows of d_k = columns of the boundary matrix of level k+1: face ids ascending (Manifolds.jl:735-757)."""
    B = o.boundary(o.Pr, k + 1, om).tocsc()
    B.sort_indices()
    return B.indices.reshape(-1, k + 2).astype(np.int64)


def bit_fields(idx):
    """Differences the bit-packed rows store and the widths k_bits_scan would choose for them."""
    nr, W = idx.shape
    pad = (-nr) % 32
    i0 = np.concatenate([idx[:, 0], np.full(pad, np.iinfo(np.int64).max)]).reshape(-1, 32)
    base = i0.min(axis=1)
    d = np.empty_like(idx)
    d[:, 0] = idx[:, 0] - np.repeat(base, 32)[:nr]
    d[:, 1:] = np.diff(idx, axis=1)
    bits = [max(1, int(d[:, p].max()).bit_length()) for p in range(W)]
    return base, d, bits


def encode_bits(idx):
    base, d, bits = bit_fields(idx)
    shifts = np.concatenate([[0], np.cumsum(bits)[:-1]])
    words = np.zeros(idx.shape[0], dtype=object)
    for p in range(idx.shape[1]):
        words = words + (d[:, p].astype(object) << int(shifts[p]))
    return base, words, bits, shifts


def decode_bits(base, words, bits, shifts, W):
    out = np.empty((len(words), W), dtype=np.int64)
    grp = np.repeat(base, 32)[:len(words)]
    for p in range(W):
        f = np.array([(int(w) >> int(shifts[p])) & ((1 << bits[p]) - 1) for w in words], dtype=np.int64)
        out[:, p] = (grp if p == 0 else out[:, p - 1]) + f
    return out


def prefix_packs(idx):
    """k_pack_fixed: idx0 against the idx0 of the FIRST row of its 32-row group in 8 bits, idx1 - idx0 in 24."""
    nr = idx.shape[0]
    b = np.repeat(idx[::32, 0], 32)[:nr]
    d = idx[:, 0] - b
    off = idx[:, 1] - idx[:, 0]
    return bool((d >= 0).all() and d.max() <= 255 and (off > 0).all() and off.max() < (1 << 24))


def predicted_format(idx):
    W = idx.shape[1]
    _, _, bits = bit_fields(idx)
    total = sum(bits)
    if W >= 3 and total <= 32:
        return "bits32"
    if W == 4 and total <= 64:
        return "bits64"
    if prefix_packs(idx):
        return "packed"
    if W == 3 and total <= 64:
        return "bits64"
    return "explicit"


def zoo():
    rng = np.random.default_rng(78)
    simp, coords, weights = o.large_hypercube_tables(3, (8, 8, 8))
    perm = rng.permutation(coords.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    pts = rng.random((30000, 2))
    tri = np.sort(scipy.spatial.Delaunay(pts).simplices.astype(np.int64), axis=1)
    return {
        "kuhn3n6": (o.large_hypercube_tables(3, (6, 6, 6)), {0: "packed", 1: "bits32", 2: "bits32"}),
        "kuhn3n24": (o.large_hypercube_tables(3, (24, 24, 24)), {0: "packed", 1: "bits32", 2: "bits64"}),
        "shuffled": ((inv[simp], coords[perm], weights[perm]), {0: "packed", 1: "bits32", 2: "bits64"}),
        "delaunay2": ((tri, pts, np.zeros(len(pts))), {1: "bits64"}),
        "kuhn2n74": (o.large_hypercube_tables(2, (74, 74)), {0: "packed", 1: "bits32"}),
    }


@pytest.mark.parametrize("name", ["kuhn3n6", "kuhn3n24", "shuffled", "delaunay2", "kuhn2n74"])
def test_row_storage_model(name):
    (s_, c_, w_), expect = zoo()[name]
    om = o.build_manifold(name, s_, c_, w_)
    for k, fmt in expect.items():
        idx = table(om, k)
        assert (np.diff(idx, axis=1) > 0).all()                       # ascending face ids: every difference positive
        assert predicted_format(idx) == fmt, (name, k)
        if idx.shape[1] >= 3:
            sub = idx[: 32 * 40]                                      # whole groups: the round trip is per group
            base, words, bits, shifts = encode_bits(sub)
            assert sum(bits) <= 64
            assert np.array_equal(decode_bits(base, words, bits, shifts, idx.shape[1]), sub), (name, k)


def test_prefix_rule_holds_below_the_top_level():
    """The first face of a row is the row's prefix, so idx0 is non-decreasing down the rows of every level below the
    top (lexicographic numbering) -- on ordered, shuffled and Delaunay meshes alike."""
    for name, ((s_, c_, w_), _) in zoo().items():
        om = o.build_manifold(name, s_, c_, w_)
        for k in range(om.D - 1):
            idx = table(om, k)
            assert (np.diff(idx[:, 0]) >= 0).all(), (name, k)
