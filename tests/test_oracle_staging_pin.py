"""CPU, build container only: pins oracle/staging_ref.py (collate, Levenshtein / Sym-ER / Seq-ER, weighted late-fusion
decode) against the reference's own code (skipped where /root/reference is absent)."""
import ast
import os
import random

import pytest
import torch

from oracle import restate, shim, staging_ref, synth

pytestmark = pytest.mark.skipif(not shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    import sys

    r = shim.load_reference()
    r.pre = sys.modules["src.data.preprocessing"]
    r.metrics = sys.modules["src.utils.metrics"]
    return r


def _ragged_batch(seed=0, multimodal=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for h, w, n in [(40, 70, 9), (33, 128, 4), (64, 17, 12), (1, 1, 2)]:
        xi = torch.rand(1, h, w, generator=g)
        y = torch.randint(1, 90, (n,), generator=g)
        if multimodal:
            xa = torch.rand(1, 24, w + 5, generator=g)
            out.append((xi, staging_ref.number_of_frames(xi), xa, staging_ref.number_of_frames(xa), y))
        else:
            out.append((xi, staging_ref.number_of_frames(xi), y))
    return out


def test_collate_matches_reference(ref):
    b = _ragged_batch()
    for fn_ref, fn in [(ref.pre.ar_batch_preparation_image, lambda x: staging_ref.ar_batch_preparation_unimodal(x, 1.0)),
                       (ref.pre.ar_batch_preparation_audio, lambda x: staging_ref.ar_batch_preparation_unimodal(x, 0.0))]:
        for r, o in zip(fn_ref(b), fn(b)):
            assert r.dtype == o.dtype and torch.equal(r, o)
    bm = _ragged_batch(1, multimodal=True)
    for r, o in zip(ref.pre.ar_batch_preparation_multimodal(bm), staging_ref.ar_batch_preparation_multimodal(bm)):
        assert r.dtype == o.dtype and torch.equal(r, o)


def test_ed_metrics_match_reference(ref):
    rnd = random.Random(3)
    y_true = [[rnd.randrange(6) for _ in range(rnd.randrange(1, 30))] for _ in range(25)] + [[1, 2, 3]]
    y_pred = [[rnd.randrange(6) for _ in range(rnd.randrange(0, 30))] for _ in range(25)] + [[1, 2, 3]]
    assert ref.metrics.compute_ed_metrics(y_true, y_pred) == staging_ref.compute_ed_metrics(y_true, y_pred)


def _reference_weighted_prediction():
    """the reference's own ``weighted_prediction`` function, compiled from its source file without running the script's
    import-time side effects (fire / wandb / seeding)"""
    path = os.path.join(shim.REFERENCE_ROOT, "src", "multimodal", "weighted_multimodal", "test.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "weighted_prediction")
    from typing import List

    ns = {"torch": torch, "List": List, "EOS_TOKEN": "<eos>", "SOS_TOKEN": "<sos>"}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["weighted_prediction"]


def test_weighted_prediction_matches_reference(ref):
    w2i, i2w = synth.tiny_vocab(61)
    models, sds = [], []
    for seed in (21, 22):
        m = ref.Transformer(64, 128, 14, w2i, i2w)
        sd = synth.synth_state_dict(m.state_dict(), seed=seed)
        m.load_state_dict(sd)
        models.append(m.eval())
        sds.append(sd)
    g = torch.Generator().manual_seed(2)
    xi, xa = torch.rand(1, 1, 64, 128, generator=g), torch.rand(1, 1, 48, 96, generator=g)
    fn = _reference_weighted_prediction()
    for alpha in (0.5, 0.2):
        with torch.no_grad():
            words = fn(xi, xa, models[0], models[1], alpha)
        mi = restate.encode_to_memory(sds[0], "encoder.", "pos_2d.pe", xi)
        ma = restate.encode_to_memory(sds[1], "encoder.", "pos_2d.pe", xa)
        toks, _ = staging_ref.weighted_greedy_decode(sds[0], sds[1], mi, ma, w2i["<sos>"], w2i["<eos>"], 14, alpha)
        assert [w2i[w] for w in words] == toks
