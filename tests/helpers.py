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
    return float(loss), {k: v.grad for k, v in sdg.items() if torch.is_floating_point(v) and v.grad is not None}


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
