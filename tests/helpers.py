"""Shared helpers for the parity tests (oracle on CPU vs CUDA modules on the same weights/inputs)."""
import torch

from oracle import restate, synth


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build_multimodal(mixer="concat", window=-1, vocab=97, img=(64, 128), aud=(48, 96), max_len=40, seed=3, dtype=torch.float32):
    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.tiny_vocab(vocab)
    m = pkg.MultimodalTransformer(img[0], img[1], aud[0], aud[1], max_len, w2i, i2w, mixer_type=mixer, attn_window=window)
    sd = synth.synth_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    m = m.to("cuda:0").eval()
    m.set_compute_dtype(dtype)
    return m, sd, w2i


def build_unimodal(window=-1, vocab=97, hw=(64, 128), max_len=40, seed=4, dtype=torch.float32):
    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.tiny_vocab(vocab)
    m = pkg.Transformer(hw[0], hw[1], max_len, w2i, i2w, attn_window=window)
    sd = synth.synth_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    m = m.to("cuda:0").eval()
    m.set_compute_dtype(dtype)
    return m, sd, w2i


def oracle_grads(loss_fn, sd):
    """loss_fn(sd_requiring_grad) -> scalar; returns (loss, {key: grad})"""
    sdg = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and not k.endswith(".pe") else v) for k, v in sd.items()}
    loss = loss_fn(sdg)
    loss.backward()
    return float(loss.detach()), {k: v.grad for k, v in sdg.items() if torch.is_floating_point(v) and v.grad is not None}


def grads_rel(ga, gb):
    """global L2-relative error of gradient dict ga against gb"""
    num = den = 0.0
    for k, b in gb.items():
        if k not in ga or ga[k] is None:
            continue
        a, b = ga[k].double().reshape(-1), b.double().reshape(-1)
        num += float((a - b).pow(2).sum())
        den += float(b.pow(2).sum())
    return (num / max(den, 1e-300)) ** 0.5


def oracle_truth_and_floors(forward, y_out, sd):
    """forward(sd, dtype) -> logits [B,V,T] of the oracle restatement.

    Returns a dict with the fp64 oracle ("truth": logits, loss, grads) and the reference's OWN precision
    noise floors against it: fp32 oracle vs fp64 and autocast-bf16 oracle vs fp64,
    for logits (L2-rel) and gradients (global L2-rel).  The CUDA path is held to the north-star tolerances
    (1e-4 fp32 / 1e-2 bf16) or a small multiple of these floors, whichever is larger: random-init gradients
    through 9 InstanceNorm'd blocks are cancellation-dominated and the reference itself misses 1e-4.
    """
    from oracle import restate

    out = {}
    with torch.no_grad():
        l64 = forward(sd, torch.float64)
        l32 = forward(sd, torch.float32)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            lbf = forward(sd, torch.float32)
    out["logits"] = l64
    out["floor_logits_fp32"] = rel_err(l32, l64)
    out["floor_logits_bf16"] = rel_err(lbf.float(), l64)
    loss64, g64 = oracle_grads(lambda s: restate.ce_loss(forward(s, torch.float64), y_out), sd)
    _, g32 = oracle_grads(lambda s: restate.ce_loss(forward(s, torch.float32), y_out), sd)

    def _bf(s):
        with torch.autocast("cpu", dtype=torch.bfloat16):
            return restate.ce_loss(forward(s, torch.float32).float(), y_out)

    _, gbf = oracle_grads(_bf, sd)
    out["loss"], out["grads"] = loss64, g64
    out["floor_grads_fp32"] = grads_rel(g32, g64)
    out["floor_grads_bf16"] = grads_rel(gbf, g64)
    return out


def grad_report(model, ref_grads):
    """global L2-rel error, cosine and the worst per-tensor relative error of model.grad vs the oracle"""
    num = den = dot = na = 0.0
    worst = ("", 0.0)
    missing = []
    for k, p in model.named_parameters():
        if k not in ref_grads:
            continue
        if p.grad is None:
            missing.append(k)
            continue
        a, b = p.grad.detach().double().cpu().reshape(-1), ref_grads[k].double().reshape(-1)
        num += float((a - b).pow(2).sum())
        den += float(b.pow(2).sum())
        dot += float((a * b).sum())
        na += float(a.pow(2).sum())
        r = float((a - b).norm() / (b.norm() + 1e-30))
        if r > worst[1] and float(b.norm()) > 1e-12:
            worst = (k, r)
    return dict(global_rel=(num / max(den, 1e-300)) ** 0.5, cos=dot / max((na * den) ** 0.5, 1e-300), worst=worst, missing=missing)


# ---- ReLU decisions: gradients are discontinuous at ReLU boundaries ------------------------------------------------
# With ~1e6 pre-activations of O(1) scale, about one per batch lands within fp32 rounding of zero, and the side of zero a
# correctly-rounded fp32 kernel puts it on decides a mask entry that can carry percents of a layer's gradient.  The fp32
# gradient tests therefore (1) read the product's own ReLU decisions back (ops.RELU_PROBE), (2) replay them in the fp64
# oracle (restate.RELU_HOOK), (3) bound how many decisions differ from the oracle's own and how far from zero those
# pre-activations are, and (4) hold the gradients to the tight tolerance given identical decisions.
def capture_relu_masks(fn):
    """run fn() with the product's ReLU probe armed -> (fn's result, [bool mask per ReLU, in call order] on the CPU)"""
    from omr_a2s_multimodal_transformer_b200 import ops

    ops.RELU_PROBE = []
    try:
        out = fn()
        torch.cuda.synchronize()
        return out, [(t > 0).cpu() for t in ops.RELU_PROBE]
    finally:
        ops.RELU_PROBE = None


class replay_relu_masks:
    """context manager: every ReLU of oracle/restate.py takes the decision recorded in ``masks`` (product layouts: NHWC for
    feature maps, [rows, width] for the feed-forward blocks).  ``flips`` counts the decisions that differ from the oracle's
    own sign test and ``max_abs_flipped`` is the largest oracle pre-activation among them."""

    def __init__(self, masks):
        self.masks, self.i, self.flips, self.max_abs_flipped, self.total = masks, 0, 0, 0.0, 0

    def _hook(self, x):
        m = self.masks[self.i % len(self.masks)]
        self.i += 1
        m = m.permute(0, 3, 1, 2) if x.dim() == 4 else m.reshape(x.shape)
        assert m.shape == x.shape, (tuple(m.shape), tuple(x.shape), "ReLU call order of product and oracle differ")
        flipped = (x.detach() > 0) != m
        n = int(flipped.sum())
        if n:
            self.flips += n
            self.max_abs_flipped = max(self.max_abs_flipped, float(x.detach()[flipped].abs().max()))
        self.total += x.numel()
        return x * m.to(x.dtype)

    def __enter__(self):
        restate.RELU_HOOK = self._hook
        return self

    def __exit__(self, *exc):
        restate.RELU_HOOK = None
        return False
