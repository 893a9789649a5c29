"""-m gpu: the fused decoder links of csrc/ln_fused.cu (OMR_FUSE_DECODER_LINKS, on by default since round 2: first run
green on a B200 there, 24.19 -> 24.04 ms/step) and the double-buffered input prefetch of GraphedTrainStep.
Each fused kernel against the pair of kernels it replaces (same seeds -> the same dropout mask), then the whole decoder
in train mode, fused against unfused."""
import pytest
import torch

from oracle import synth
from tests.helpers import build_unimodal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("d", [256, 512])
def test_dropout_add_layernorm_fwd_matches_the_two_kernels(dtype, tol, d):
    from omr_a2s_multimodal_transformer_b200 import ops

    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 5, d, generator=g).to(DEV).to(dtype)
    res = torch.randn(37, 5, d, generator=g).to(DEV).to(dtype)
    gamma, beta = torch.rand(d, generator=g).to(DEV) + 0.5, torch.randn(d, generator=g).to(DEV)
    p, seed = 0.1, 12345
    y1, s1, st1 = ops.dropout_add_layernorm_fwd(x, res, gamma, beta, 1e-5, True, p, seed)
    xd = ops.dropout(x, p, seed)
    y0, s0, st0 = ops.add_layernorm_fwd(xd, res, gamma, beta, 1e-5, True)
    # identical mask: the dropped positions of s - res coincide
    assert torch.equal((s1.float() - res.float()) == 0, (s0.float() - res.float()) == 0) or dtype == torch.bfloat16
    assert float((y1.float() - y0.float()).norm() / y0.float().norm()) < tol
    assert float((s1.float() - s0.float()).norm() / s0.float().norm()) < tol
    assert torch.allclose(st1, st0, rtol=1e-3, atol=1e-3)
    keep = float((ops.dropout(torch.ones_like(x), p, seed) > 0).float().mean())
    assert abs(keep - 0.9) < 0.01


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
def test_layernorm_bwd_dropout_matches_the_two_kernels(dtype, tol):
    from omr_a2s_multimodal_transformer_b200 import ops

    g = torch.Generator().manual_seed(1)
    d = 256
    x = torch.randn(64, 9, d, generator=g).to(DEV).to(dtype)
    res = torch.randn(64, 9, d, generator=g).to(DEV).to(dtype)
    gamma, beta = torch.rand(d, generator=g).to(DEV) + 0.5, torch.randn(d, generator=g).to(DEV)
    dy = torch.randn(64, 9, d, generator=g).to(DEV).to(dtype)
    p, seed = 0.1, 777
    _, s, st = ops.add_layernorm_fwd(x, res, gamma, beta, 1e-5, True)
    dg0, db0, dg1, db1 = (torch.zeros(d, device=DEV) for _ in range(4))
    ds0 = ops.layernorm_bwd(dy, s, st, gamma, dg0, db0)
    da0 = ops.dropout(ds0, p, seed)
    dbias = torch.zeros(d, device=DEV)
    ds1, da1 = ops.layernorm_bwd_dropout(dy, s, st, gamma, dg1, db1, p, seed, dbias=dbias)
    if dtype == torch.float32:
        assert torch.equal(ds1, ds0) and torch.equal(da1 == 0, da0 == 0)
    else:  # the 16-byte-wide bf16 kernel sums a row in another order: the last bit of a few values may differ
        assert float((ds1.float() - ds0.float()).norm() / ds0.float().norm()) < 2e-3
        assert float(((da1 == 0) != (da0 == 0)).float().mean()) < 1e-4
    assert float((da1.float() - da0.float()).norm() / da0.float().norm()) < tol
    # the bias gradient of the linear layer in front of the dropout = column sums of da, accumulated on the way
    ref_b = da1.float().reshape(-1, d).sum(0)
    assert torch.allclose(dbias, ref_b, rtol=2e-3, atol=2e-3)
    assert torch.allclose(dg1, dg0, rtol=1e-4, atol=1e-4) and torch.allclose(db1, db0, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [4096, 4097])
def test_mask_scale_is_relu_and_dropout_backward(dtype, n):
    from omr_a2s_multimodal_transformer_b200 import ops

    g = torch.Generator().manual_seed(2)
    h = torch.randn(n, generator=g).to(DEV).to(dtype)
    hmid = torch.relu(h)
    p, seed = 0.1, 99
    hdrop = ops.dropout(hmid.view(1, -1), p, seed).view(-1)
    dh = torch.randn(n, generator=g).to(DEV).to(dtype)
    ref = ops.relu_bwd(hmid, ops.dropout(dh.clone().view(1, -1), p, seed).view(-1))
    got = ops.mask_scale(dh.clone(), hdrop, 1.0 / (1.0 - p))
    assert torch.equal(got, ref)


def test_decoder_train_step_fused_links_match_unfused(monkeypatch):
    """Train-mode decoder step (same seeds => same dropout masks in all four runs): fused and unfused links agree to
    rounding in fp32; in bf16 the two paths round at different places (the fused kernels keep the dropped sublayer output
    and the LayerNorm gradient in fp32 registers), so each is judged against the fp32 result: the fused path must not be
    further from it than the unfused one (+25 %), and both stay inside the bf16 gradient tolerance of the model tests."""
    results = {}
    for dtype in (torch.float32, torch.bfloat16):
        for fuse in ("0", "1"):
            monkeypatch.setenv("OMR_FUSE_DECODER_LINKS", fuse)
            m, sd, w2i = build_unimodal(dtype=dtype)
            x, xl, y_in, y_out = synth.synth_unimodal_batch(3, 64, 128, [20, 12, 7], w2i)
            with torch.no_grad():
                mem = m.encode(x.to(DEV))
            m.decoder.train()
            m.zero_grad(set_to_none=True)
            loss = m.decoder.loss(tgt=y_in.to(DEV), memory=mem, memory_len=xl.to(DEV), targets=y_out.to(DEV))
            loss.backward()
            torch.cuda.synchronize()
            results[(dtype, fuse)] = (float(loss.detach()), {k: p.grad.detach().double().cpu()
                                                             for k, p in m.decoder.named_parameters() if p.grad is not None})

    def dist(a, b):
        (la, ga), (lb, gb) = results[a], results[b]
        assert set(ga) == set(gb)
        num = sum(float((ga[k] - gb[k]).pow(2).sum()) for k in gb)
        den = sum(float(gb[k].pow(2).sum()) for k in gb)
        return abs(la - lb) / max(1.0, abs(lb)), (num / den) ** 0.5

    f32, bf = torch.float32, torch.bfloat16
    dl, dg = dist((f32, "1"), (f32, "0"))
    assert dl < 1e-5 and dg < 1e-5, (dl, dg)
    (l_un, g_un), (l_fu, g_fu) = dist((bf, "0"), (f32, "0")), dist((bf, "1"), (f32, "0"))
    assert l_fu < 2e-2 and g_fu < 5e-2 and g_fu < 1.25 * g_un + 2e-3, ((l_un, g_un), (l_fu, g_fu))


def test_graphed_step_prefetch_matches_plain_loading():
    """GraphedTrainStep(double_buffer=True): a step replayed on a batch prefetched through the copy stream equals the same
    step with the batch loaded on the compute stream (eval mode: deterministic up to atomic ordering)"""
    import omr_a2s_multimodal_transformer_b200 as pkg
    from tests.helpers import build_multimodal

    finals = []
    for double in (False, True):
        m, sd, w2i = build_multimodal(dtype=torch.bfloat16)
        dp = pkg.DataParallel(m, broadcast=False)
        opt = m.configure_optimizers()
        batches = [[t.pin_memory() for t in synth.synth_multimodal_batch(3, (64, 128), (48, 96), [20, 12, 7], w2i, seed=s)]
                   for s in (1, 2, 3)]
        dev0 = [t.to(DEV) for t in batches[0]]

        def step(bt):
            xi, xli, xa, xla, y_in, y_out = bt
            dp.zero_grad()
            mem, xl = m._memory(xi, xa, xli, xla, "both")
            loss = m.decoder.loss(tgt=y_in, memory=mem, memory_len=xl, targets=y_out)
            loss.backward()
            dp.sync_gradients()
            opt.step()
            return loss

        stepper = pkg.GraphedTrainStep(step, dev0, opt, variants=2, warmup=1, double_buffer=double)
        losses = []
        if double:
            stepper.prefetch(batches[0])
            for k in range(6):
                loss = stepper()
                stepper.prefetch(batches[(k + 1) % 3])
                losses.append(float(loss))
        else:
            for k in range(6):
                losses.append(float(stepper(batches[k % 3])))
        finals.append(losses)
    assert max(abs(a - b) for a, b in zip(*finals)) < 2e-2 * max(finals[0]), finals
