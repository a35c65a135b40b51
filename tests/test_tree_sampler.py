"""The tcgen05 encoder draws each head's bin by walking a binary tree of partial sums (csrc/encode_tc.cu: tree32 /
sample_regs) instead of the sequential inverse-CDF scan of the oracle (oracle/ref_model.py:sample_bins_cdf).  Both
compute  bin = #{k : cumsum(e)_k <= u * sum(e)}  and differ only in how the float32 partial sums are associated, so
they may disagree when u * sum(e) falls within a rounding error of a CDF edge.  This float32 restatement of the tree
walk measures that: the draws agree except for ~1e-6 of the cases, and then differ by one bin."""
import numpy as np

F = np.float32


def _tree_draw(e, u):
    """e: [P, 32] float32 weights, u: [P] uniforms -> bins, float32 arithmetic associated as in sample_regs / tree32."""
    s2 = (e[:, 0::2] + e[:, 1::2]).astype(F)
    s4 = (s2[:, 0::2] + s2[:, 1::2]).astype(F)
    s8 = (s4[:, 0::2] + s4[:, 1::2]).astype(F)
    s16 = (s8[:, 0::2] + s8[:, 1::2]).astype(F)
    tot = (s16[:, 0] + s16[:, 1]).astype(F)
    t = (u * tot).astype(F)
    rows = np.arange(e.shape[0])
    node = np.zeros(e.shape[0], np.int64)            # index of the current node at the current level
    for level in (s16, s8, s4, s2, e):               # left child weight decides the next bit
        left = level[rows, 2 * node]
        go_right = t >= left
        t = np.where(go_right, (t - left).astype(F), t)
        node = 2 * node + go_right
    return node, tot


def test_tree_walk_equals_sequential_scan_except_at_cdf_edges():
    rng = np.random.default_rng(0)
    p = 400_000
    logits = (rng.standard_normal((p, 32)) * rng.uniform(0.2, 6.0, (p, 1))).astype(F)      # flat to sharp heads
    e = np.exp2(logits - logits.max(1, keepdims=True)).astype(F)
    u = rng.random(p, dtype=F)
    tree, tot = _tree_draw(e, u)
    cum = np.cumsum(e, axis=1, dtype=F)                                                       # sequential float32 scan
    seq = (cum <= (u * cum[:, -1]).astype(F)[:, None]).sum(1)
    seq = np.minimum(seq, 31)
    tree = np.minimum(tree, 31)
    diff = np.abs(tree - seq)
    assert diff.max() <= 1
    assert (diff != 0).mean() < 2e-5, f"{int((diff != 0).sum())} of {p} draws differ"
    # and the tree is an exact sampler of ITS OWN partial sums: the chosen bin brackets t in the tree's arithmetic
    hist = np.bincount(tree, minlength=32) / p
    want = (e / e.sum(1, keepdims=True)).mean(0)
    np.testing.assert_allclose(hist, want, atol=4e-3)
