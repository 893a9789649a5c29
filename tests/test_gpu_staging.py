"""-m gpu: the steps either side of the model against oracle/staging_ref.py (itself pinned
to the reference's own collate / metrics / weighted_prediction code in tests/test_oracle_staging_pin.py).
Integer and byte work: bit-exact.  Late-fusion decode: identical token sequences in fp32."""
import random

import pytest
import torch

from oracle import restate, staging_ref, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ragged(seed, shapes, multimodal=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for h, w, n in shapes:
        xi = torch.rand(1, h, w, generator=g)
        y = torch.randint(1, 90, (n,), generator=g)
        if multimodal:
            xa = torch.rand(1, 195, max(1, w // 2 + 3), generator=g)
            out.append((xi, staging_ref.number_of_frames(xi), xa, staging_ref.number_of_frames(xa), y))
        else:
            out.append((xi, staging_ref.number_of_frames(xi), y))
    return out


SHAPES = [(40, 70, 9), (33, 128, 4), (64, 17, 12), (1, 1, 2), (128, 1024, 40)]


@pytest.mark.parametrize("shapes", [SHAPES, [(5, 7, 3)], [(16, 8, 2), (16, 8, 2)], [(3, 1021, 5), (7, 1023, 2)]])
def test_collate_unimodal_bit_exact(shapes):
    from omr_a2s_multimodal_transformer_b200 import staging

    b = _ragged(0, shapes)
    for pad, fn in [(1.0, staging.ar_batch_preparation_image), (0.0, staging.ar_batch_preparation_audio)]:
        ref = staging_ref.ar_batch_preparation_unimodal(b, pad)
        got = fn(b, device=DEV)
        for r, o in zip(ref, got):
            assert o.is_cuda and r.dtype == o.dtype and r.shape == o.shape and torch.equal(r, o.cpu())
    nf = staging.number_of_frames([s[0] for s in b], device=DEV)
    assert nf.cpu().tolist() == [s[1] for s in b]
    # samples that already live on the device take the same path
    bd = [(x.to(DEV), xl, y.to(DEV)) for x, xl, y in b]
    got = staging.ar_batch_preparation_image(bd, device=DEV)
    for r, o in zip(staging_ref.ar_batch_preparation_unimodal(b, 1.0), got):
        assert torch.equal(r, o.cpu())


def test_collate_multimodal_bit_exact():
    from omr_a2s_multimodal_transformer_b200 import staging

    b = _ragged(1, SHAPES, multimodal=True)
    for r, o in zip(staging_ref.ar_batch_preparation_multimodal(b), staging.ar_batch_preparation_multimodal(b, device=DEV)):
        assert r.dtype == o.dtype and r.shape == o.shape and torch.equal(r, o.cpu())


def test_collate_rejects_cpu_device_and_bad_shapes():
    from omr_a2s_multimodal_transformer_b200 import staging

    with pytest.raises(RuntimeError):
        staging.ar_batch_preparation_image(_ragged(0, [(5, 7, 3)]), device="cpu")
    with pytest.raises(ValueError):
        staging.pad_batch_inputs([torch.rand(2, 5, 7)], device=DEV)


def test_levenshtein_and_error_rates_match_reference_algorithm():
    from omr_a2s_multimodal_transformer_b200 import staging

    rnd = random.Random(5)

    def seq(n, k=7):
        return [f"t{rnd.randrange(k)}" for _ in range(n)]

    y_true = [seq(rnd.randrange(1, 60)) for _ in range(40)]
    y_pred = [seq(rnd.randrange(0, 60)) for _ in range(40)]
    # edge cases: identical, empty hypothesis, single tokens, the grandstaff maximum length (1268), near-identical long pair
    long_a = seq(1268, 50)
    long_b = list(long_a)
    for i in range(0, 1268, 97):
        long_b[i] = "zz"
    del long_b[500:520]
    y_true += [["a", "b", "c"], ["a", "b"], ["a"], ["a"], long_a, seq(1268, 3), seq(700)]
    y_pred += [["a", "b", "c"], [], ["a"], ["b"], long_b, seq(1100, 3), seq(1268)]
    ed, sums = staging.edit_distances(y_true, y_pred, device=DEV)
    ref = [staging_ref.levenshtein(t, h) for t, h in zip(y_true, y_pred)]
    assert ed.cpu().tolist() == ref
    assert sums.cpu().tolist() == [sum(ref), sum(len(t) for t in y_true), sum(e > 0 for e in ref)]
    assert staging.compute_ed_metrics(y_true, y_pred, device=DEV) == staging_ref.compute_ed_metrics(y_true, y_pred)
    # symmetry: a size-independent property of the distance
    ed_t, _ = staging.edit_distances(y_pred, y_true, device=DEV)
    assert ed_t.cpu().tolist() == ref


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("alpha", [0.5, 0.2, 1.0])
def test_mix_argmax_step_matches_torch(dtype, alpha):
    from omr_a2s_multimodal_transformer_b200 import ops

    g = torch.Generator().manual_seed(3)
    b, v = 5, 6997
    la = (torch.randn(b, v, generator=g) * 3).to(dtype).to(DEV)
    lb = (torch.randn(b, v, generator=g) * 3).to(dtype).to(DEV)
    la[1, 10] = la[1, 4000] = 30.0  # a tie: the first maximum wins, as torch.argmax
    lb[1, 10] = lb[1, 4000] = 30.0
    tok = torch.zeros(b, dtype=torch.int64, device=DEV)
    val = torch.zeros(b, dtype=torch.float32, device=DEV)
    fin = torch.zeros(b, dtype=torch.int32, device=DEV)
    fin[4] = 1
    out_t = torch.full((b, 3), -1, dtype=torch.int64, device=DEV)
    out_v = torch.zeros((b, 3), dtype=torch.float32, device=DEV)
    p = alpha * la.float().softmax(-1) + (1 - alpha) * lb.float().softmax(-1)
    eos = int(p[2].argmax())
    ops.mix_argmax_step(la, lb, alpha, tok, val, fin, eos, 0, out_t, out_v, 1)
    torch.cuda.synchronize()
    exp = p.argmax(-1)
    assert tok[:4].tolist() == exp[:4].tolist() and tok[1].item() == 10
    assert tok[4].item() == 0 and val[4].item() == 0.0  # finished rows emit PAD
    assert fin.tolist() == [0, 0, 1, 0, 1]  # the row that drew EOS becomes finished
    assert torch.allclose(val[:4], p.max(-1).values[:4], rtol=1e-4, atol=1e-7)
    assert out_t[:, 1].tolist() == tok.tolist() and out_t[:, 0].tolist() == [-1] * b


@pytest.mark.parametrize("alpha", [0.5, 0.2])
def test_weighted_prediction_tokens_identical_fp32(alpha):
    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.tiny_vocab(61)
    models, sds = [], []
    for seed, hw in ((21, (64, 128)), (22, (48, 96))):
        m = pkg.Transformer(hw[0], hw[1], 14, w2i, i2w)
        sd = synth.synth_state_dict(m.state_dict(), seed=seed)
        m.load_state_dict(sd)
        m = m.to(DEV).eval()
        m.set_compute_dtype(torch.float32)
        models.append(m)
        sds.append(sd)
    g = torch.Generator().manual_seed(2)
    xi, xa = torch.rand(3, 1, 64, 128, generator=g), torch.rand(3, 1, 48, 96, generator=g)
    sos, eos = w2i["<sos>"], w2i["<eos>"]
    refs = []
    for b in range(3):
        mi = restate.encode_to_memory(sds[0], "encoder.", "pos_2d.pe", xi[b:b + 1])
        ma = restate.encode_to_memory(sds[1], "encoder.", "pos_2d.pe", xa[b:b + 1])
        refs.append(staging_ref.weighted_greedy_decode(sds[0], sds[1], mi, ma, sos, eos, 14, alpha))
    for use_graph in (False, True):
        toks, vals, lens = pkg.weighted_prediction_batch(xi.to(DEV), xa.to(DEV), models[0], models[1], alpha, use_graph=use_graph)
        seqs, probs = pkg.BatchedGreedyDecoder.to_lists(toks, vals, lens)
        for b in range(3):
            assert seqs[b] == refs[b][0], (use_graph, b, seqs[b], refs[b][0])
            assert max(abs(p - q) for p, q in zip(probs[b], refs[b][1])) < 1e-4
    words = pkg.weighted_prediction(xi[:1].to(DEV), xa[:1].to(DEV), models[0], models[1], alpha)
    assert [w2i[w] for w in words] == refs[0][0]
    with pytest.raises(AssertionError):
        pkg.weighted_prediction(xi.to(DEV), xa.to(DEV), models[0], models[1], alpha)
