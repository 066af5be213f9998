"""Seeded synthetic assembly graphs of the shape BASELINE.json names.

Nodes are strand pairs (2k, 2k+1) laid out in genome order (graph_parser.py:174-181 of the
reference); out-degrees of the even nodes follow a truncated discrete power law; an edge (u, v)
goes ``band``-geometrically forward in genome order (or, with probability ``p_long``, to a
uniformly random node -- a repeat-induced overlap), and is mirrored as the reverse-complement
edge (v^1, u^1) with the next edge id (graph_parser.py:300-326).  Multi-edges can occur; there
are no self loops unless a long-range edge lands on its own source.
"""
import numpy as np


def power_law_degrees(n, total, alpha=2.2, dmax=4096, rng=None):
    """``n`` integer degrees >= 1 from p(d) ~ d^-alpha on [1, dmax], rescaled to sum to ``total``."""
    rng = rng or np.random.default_rng(0)
    if total < n:
        raise ValueError('need at least one out-edge per even node')
    u = rng.random(n)
    a1 = 1.0 - alpha
    d = np.floor(((dmax + 1.0) ** a1 * u + (1.0 - u)) ** (1.0 / a1))  # inverse-CDF of the continuous law
    d = np.clip(d, 1, dmax)
    d = np.maximum(1, np.floor(d * (total / d.sum()) + rng.random(n))).astype(np.int64)
    diff = int(total - d.sum())
    while diff != 0:  # spread the rounding remainder over random nodes
        k = min(abs(diff), n)
        idx = rng.choice(n, size=k, replace=False)
        if diff > 0:
            d[idx] += 1
            diff -= k
        else:
            ok = idx[d[idx] > 1]
            d[ok] -= 1
            diff += ok.size
    return d


def make_assembly_graph(num_nodes, num_edges, seed=0, band=64, alpha=2.2, p_long=0.01):
    """Return ``(src, dst)`` int32 arrays of length ``num_edges`` (even) over ``num_nodes`` (even)."""
    if num_nodes % 2 or num_edges % 2:
        raise ValueError('num_nodes and num_edges must be even (strand pairs)')
    rng = np.random.default_rng(seed)
    half_n, half_e = num_nodes // 2, num_edges // 2
    deg = power_law_degrees(half_n, half_e, alpha=alpha, rng=rng)
    u = np.repeat(np.arange(half_n, dtype=np.int64), deg)          # read index of the even source
    delta = rng.geometric(1.0 / band, size=half_e)                  # >= 1, mean = band
    v = (u + delta) % half_n
    if p_long > 0:
        far = rng.random(half_e) < p_long
        v[far] = rng.integers(0, half_n, size=int(far.sum()))
    src = np.empty(num_edges, dtype=np.int32)
    dst = np.empty(num_edges, dtype=np.int32)
    src[0::2], dst[0::2] = 2 * u, 2 * v                            # edge k:   (u, v)
    src[1::2], dst[1::2] = 2 * v + 1, 2 * u + 1                    # edge k+1: (v^1, u^1)
    return src, dst


def make_features(src, dst, num_nodes, seed=0):
    """Inputs with the reference's semantics: x = [z(in_deg), z(out_deg)] with the unbiased std
    (inference.py:416-420); e[:,0] ~ N(0,1) (z-scored overlap length, utils/data_utils.py:36),
    e[:,1] ~ U(0.9, 1.0) (overlap similarity, graph_parser.py:110-113)."""
    rng = np.random.default_rng(seed + 1)
    indeg = np.bincount(dst, minlength=num_nodes).astype(np.float64)
    outdeg = np.bincount(src, minlength=num_nodes).astype(np.float64)

    def z(a):
        s = a.std(ddof=1)
        return (a - a.mean()) / (s if s > 0 else 1.0)

    x = np.stack((z(indeg), z(outdeg)), axis=1).astype(np.float32)
    e = np.empty((src.shape[0], 2), dtype=np.float32)
    e[:, 0] = rng.standard_normal(src.shape[0], dtype=np.float32)
    e[:, 1] = rng.uniform(0.9, 1.0, size=src.shape[0]).astype(np.float32)
    return x, e


def tiny_adversarial_graph():
    """8-node hand graph: node 7 has no in-edges, node 6 no out-edges, a self loop (3->3),
    a multi-edge (0->1 twice), a 2-cycle (4<->5)."""
    src = np.array([0, 0, 1, 2, 3, 3, 4, 5, 7, 2, 1, 5], dtype=np.int32)
    dst = np.array([1, 1, 2, 3, 3, 4, 5, 4, 0, 6, 6, 6], dtype=np.int32)
    return src, dst, 8
