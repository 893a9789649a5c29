"""BASELINE config 5 (-m gpu): the scaled-up decoder (d_model 512, 8 heads x 64, ff 512, last DSCBlock 128 -> 512,
PositionalEncoding2D(512)) composed from the same classes the way SURVEY.md section 8d describes
(``enc.dscblocks[3] = DSCBlock(128, 512, stride=(1, 1))``), against the oracle restatement on the same weights.
Sizes are reduced (2 decoder layers, short sequences) so the CPU oracle finishes in seconds; widths are the real ones."""
import pytest
import torch

from oracle import restate, synth
from tests.helpers import grad_report, oracle_truth_and_floors, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D, NHEAD, LAYERS = 512, 8, 2


def _build(dtype, max_len=48, hw=(64, 128), vocab=131, window=-1):
    import omr_a2s_multimodal_transformer_b200 as pkg
    from omr_a2s_multimodal_transformer_b200.decoder import Decoder
    from omr_a2s_multimodal_transformer_b200.encoder import Encoder
    from omr_a2s_multimodal_transformer_b200.model import PositionalEncoding2D

    w2i, i2w = synth.tiny_vocab(vocab)
    m = pkg.Transformer(hw[0], hw[1], max_len, w2i, i2w, attn_window=window)
    m.encoder = Encoder(1, out_channels=D)
    m.pos_2d = PositionalEncoding2D(D, -(-hw[0] // 16), -(-hw[1] // 8))
    m.decoder = Decoder(len(w2i), max_len, len(w2i), embedding_dim=D, ff_dim=D, nhead=NHEAD, num_transformer_layers=LAYERS,
                        padding_idx=0, attn_window=window)
    sd = synth.synth_state_dict(m.state_dict(), seed=11)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    m.set_compute_dtype(dtype)
    return m, sd, w2i


def _oracle_forward(x, xl, y_in, window=-1):
    def f(s, dt):
        mem = restate.encode_to_memory(s, "encoder.", "pos_2d.pe", x.to(dt))
        return restate.decoder_forward(s, "decoder.", y_in, mem, xl, window, nhead=NHEAD, num_layers=LAYERS)

    return f


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_c5_logits_loss_grads(dtype):
    m, sd, w2i = _build(dtype)
    x, xl, y_in, y_out = synth.synth_unimodal_batch(3, 64, 128, [40, 23, 9], w2i)
    truth = oracle_truth_and_floors(_oracle_forward(x, xl, y_in), y_out, sd)
    if dtype == torch.float32:
        tol_logit, tol_grad = max(1e-4, 4 * truth["floor_logits_fp32"]), max(1e-4, 4 * truth["floor_grads_fp32"])
    else:
        tol_logit, tol_grad = max(1e-2, 1.5 * truth["floor_logits_bf16"]), max(1e-2, 1.5 * truth["floor_grads_bf16"])
    with torch.no_grad():
        logits = m(x.to(DEV), xl.to(DEV), y_in.to(DEV))
    assert logits.shape == truth["logits"].shape
    assert rel_err(logits.float(), truth["logits"]) < tol_logit
    m.zero_grad(set_to_none=True)
    loss = m.decoder.loss(tgt=y_in.to(DEV), memory=m.encode(x.to(DEV)), memory_len=xl.to(DEV), targets=y_out.to(DEV))
    loss.backward()
    assert abs(float(loss) - truth["loss"]) < tol_logit * max(1.0, abs(truth["loss"]))
    rep = grad_report(m, truth["grads"])
    assert not rep["missing"], rep
    assert rep["global_rel"] < tol_grad and rep["cos"] > 1 - tol_grad, (rep, tol_grad)


def test_c5_greedy_tokens_identical_fp32():
    m, sd, w2i = _build(torch.float32, max_len=20)
    x = torch.rand(2, 1, 64, 128, generator=torch.Generator().manual_seed(5))
    sos, eos = w2i["<sos>"], w2i["<eos>"]
    refs = []
    for b in range(2):
        mem = restate.encode_to_memory(sd, "encoder.", "pos_2d.pe", x[b:b + 1])
        refs.append(restate.greedy_decode(sd, mem, sos, eos, 20, nhead=NHEAD, num_layers=LAYERS))
    toks, vals, lens = m._decoder_runner().decode(m.encode(x.to(DEV)), sos, eos, 0)
    seqs, probs = m._decoder_runner().to_lists(toks, vals, lens)
    for b in range(2):
        assert seqs[b] == refs[b][0], (b, seqs[b], refs[b][0])
        assert max(abs(p - q) for p, q in zip(probs[b], refs[b][1])) < 1e-3


def test_c5_bf16_long_sequence_runs_at_full_widths():
    """max_len x 2 (T = 2535 of the 2536 positions) through the tensor-core kernels: finite loss and gradients"""
    m, sd, w2i = _build(torch.bfloat16, max_len=2536)
    x, xl, y_in, y_out = synth.synth_unimodal_batch(2, 64, 128, [2536, 700], w2i)
    m.zero_grad(set_to_none=True)
    loss = m.decoder.loss(tgt=y_in.to(DEV), memory=m.encode(x.to(DEV)), memory_len=xl.to(DEV), targets=y_out.to(DEV))
    loss.backward()
    assert torch.isfinite(loss)
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
