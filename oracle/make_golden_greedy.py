"""TEST INFRASTRUCTURE ONLY.  Long greedy-decode fixtures: tests/golden/greedy_long_{c1,c3}.pt.

Runs the REAL reference's own batch-1 ``validation_step`` loop (src/transformer/model.py:170-199 unimodal,
:592-617 multimodal) in fp32 on the synthetic weights / inputs of oracle/synth.py:

* ``greedy_long_c1``: BASELINE config 1 -- every one of the 4 image-only samples, to <eos> or max_seq_len = 1268;
* ``greedy_long_c3``: BASELINE config 3's shapes (S = 2337) -- both samples of the c3 fixture batch, 640 steps.

Next to each token stream the fixture stores the reference's top-2 logit margin at every step (one teacher-forced
decoder pass over the emitted sequence; the causal mask makes position t of that pass the step-t logits), so that a
parity test can tell a real divergence from a step where the reference itself sits within fp32 rounding of a tie.

Run in the build container:  python -m oracle.make_golden_greedy [c1|c3]
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shim, synth  # noqa: E402
from oracle.make_golden import OUT, builder  # noqa: E402


def _margins(model, memory, tokens, sos):
    """top-2 logit margin of the reference at every emitted step (fp32, teacher-forced on its own output)"""
    y_in = torch.tensor([[sos] + tokens[:-1]], dtype=torch.int64)
    with torch.no_grad():
        logits = model.decoder(tgt=y_in, memory=memory, memory_len=None)[0]  # [V, T]
    top2 = logits.topk(2, dim=0).values
    assert top2.shape[1] == len(tokens)
    tf = logits.argmax(0).tolist()
    if tf != tokens:  # only possible at a near-tie: the margin recorded for that step says so
        bad = [t for t in range(len(tokens)) if tf[t] != tokens[t]]
        print(f"  teacher-forced pass picks another token at steps {bad[:8]} (margins {[float(top2[0, t] - top2[1, t]) for t in bad[:8]]})")
    return (top2[0] - top2[1]).tolist(), float(logits.abs().max())


def run(name, build, samples, steps):
    """samples(w2i) -> list of (validation batch, memory_fn(model))"""
    ref = shim.load_reference()  # noqa: F841
    torch.manual_seed(0)
    model, w2i = build(ref)
    model.eval()
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=build.seed))
    saved = model.max_seq_len
    model.max_seq_len = steps
    seqs, margins, scale = [], [], []
    for vb, mem_fn in samples(w2i):
        t0 = time.time()
        model.Y, model.YHat = [], []
        with torch.no_grad():
            model.validation_step(vb, 0)
        toks = [w2i[t] for t in model.YHat[0]]
        with torch.no_grad():
            mg, mx = _margins(model, mem_fn(model), toks, w2i["<sos>"])
        seqs.append(toks)
        margins.append(mg)
        scale.append(mx)
        print(f"{name}: sample {len(seqs) - 1}: {len(toks)} tokens, eos={toks[-1] == w2i['<eos>']}, "
              f"min top-2 margin {min(mg):.3e} (|logit| max {mx:.2f}), {time.time() - t0:.0f}s", flush=True)
    model.max_seq_len = saved
    torch.save({"name": name, "steps": steps, "greedy": seqs, "margins": [torch.tensor(m, dtype=torch.float32) for m in margins],
                "logit_scale": scale}, os.path.join(OUT, f"{name}.pt"))


def main():
    which = sys.argv[1:] or ["c1", "c3"]
    dummy_y = lambda w2i: torch.tensor([[w2i["<sos>"], 3, 4, w2i["<eos>"]]])  # noqa: E731
    if "c1" in which:
        def c1(w2i):
            x = synth.synth_unimodal_batch(4, 128, 1024, [257, 200, 128, 64], w2i, frame_lens=[1024, 1024, 896, 768])[0]
            out = []
            for i in range(4):
                xi = x[i:i + 1]
                out.append(((xi, dummy_y(w2i)),
                            lambda m, xi=xi: m.pos_2d(m.encoder(xi)).flatten(2).permute(0, 2, 1).contiguous()))
            return out

        run("greedy_long_c1", builder("uni", 0, hw=(128, 1024), max_len=1268, vocab="real"), c1, 1268)
    if "c3" in which:
        def c3(w2i):
            b = synth.synth_multimodal_batch(2, (128, 1024), (195, 808), [300, 129], w2i, img_frame_lens=[1024, 700],
                                             aud_frame_lens=[800, 1313])
            out = []
            for i in range(2):
                xi, xa = b[0][i:i + 1], b[2][i:i + 1]
                out.append(((xi, xa, dummy_y(w2i)), lambda m, xi=xi, xa=xa: m.encoder_forward(xi, xa, None, None, False)[0]))
            return out

        run("greedy_long_c3", builder("mm", 6, img=(128, 1024), aud=(195, 808), max_len=1268, mixer="concat", vocab="real"),
            c3, 640)


if __name__ == "__main__":
    main()
