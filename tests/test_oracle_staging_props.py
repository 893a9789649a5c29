"""CPU: size-independent properties of the staging oracle (oracle/staging_ref.py) -- the checker the GPU staging tests
trust.  Runs anywhere (no GPU, no /root/reference)."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import staging_ref

tokens = st.lists(st.integers(0, 5), min_size=0, max_size=24)


@settings(max_examples=150, deadline=None)
@given(tokens, tokens, tokens)
def test_levenshtein_is_a_metric(a, b, c):
    d = staging_ref.levenshtein
    assert d(a, a) == 0
    assert d(a, b) == d(b, a)
    assert (d(a, b) == 0) == (a == b)
    assert abs(len(a) - len(b)) <= d(a, b) <= max(len(a), len(b))
    assert d(a, c) <= d(a, b) + d(b, c)


@settings(max_examples=100, deadline=None)
@given(tokens, st.integers(0, 23), st.integers(0, 5))
def test_levenshtein_single_edits_cost_at_most_one(a, pos, tok):
    d = staging_ref.levenshtein
    pos = min(pos, len(a))
    assert d(a, a[:pos] + [tok] + a[pos:]) == 1            # insertion
    if a:
        p = min(pos, len(a) - 1)
        assert d(a, a[:p] + a[p + 1:]) == 1                 # deletion
        assert d(a, a[:p] + [tok] + a[p + 1:]) == (a[p] != tok)  # substitution


def test_error_rates_on_known_cases():
    m = staging_ref.compute_ed_metrics([["a", "b", "c"], ["x"]], [["a", "c"], ["x"]])
    assert m == {"sym-er": 100.0 * 1 / 4, "seq-er": 50.0}


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 40), st.integers(1, 40), st.integers(2, 9)), min_size=1, max_size=5), st.integers(0, 10 ** 6))
def test_collate_round_trip(shapes, seed):
    """cropping the padded batch gives the samples back, everything else is the pad value, frame counts follow
    ceil(h/16)*ceil(w/8), and y_in / y_out are the transcript shifted by one"""
    g = torch.Generator().manual_seed(seed)
    batch = []
    for h, w, n in shapes:
        x = torch.rand(1, h, w, generator=g)
        batch.append((x, staging_ref.number_of_frames(x), torch.randint(1, 50, (n,), generator=g)))
    for pad in (0.0, 1.0):
        xp, xl, y_in, y_out = staging_ref.ar_batch_preparation_unimodal(batch, pad)
        assert xp.shape == (len(batch), 1, max(s[0] for s in shapes), max(s[1] for s in shapes))
        total_inside = 0
        for b, ((h, w, n), (x, nf, y)) in enumerate(zip(shapes, batch)):
            assert torch.equal(xp[b, :, :h, :w], x)
            total_inside += h * w
            assert int(xl[b]) == -(-h // 16) * -(-w // 8)
            assert y_in[b, : n - 1].tolist() == y[:-1].tolist() and y_out[b, : n - 1].tolist() == y[1:].tolist()
            assert not y_in[b, n - 1:].any() and not y_out[b, n - 1:].any()
        outside = xp.numel() - total_inside
        mask = torch.ones_like(xp, dtype=torch.bool)
        for b, (h, w, n) in enumerate(shapes):
            mask[b, :, :h, :w] = False
        assert int(mask.sum()) == outside and bool((xp[mask] == pad).all())
