"""Tensor-core (tcgen05 / TMA / TMEM) kernels against fp32 torch references on bf16-rounded inputs (-m gpu).
Every case also asserts that the call really was served by the tensor-core kernel (omr_tc_call_count), so a
silent fall-back to the CUDA-core path cannot pass."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + 7 * sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.fixture(scope="module")
def ops():
    from omr_a2s_multimodal_transformer_b200 import _lib, ops as o

    _lib.load().omr_set_tensor_core_path(1)
    return o


def tc_calls():
    from omr_a2s_multimodal_transformer_b200 import _lib

    return _lib.tc_call_count()


# M, N, K  (tails in every dimension, the decoder / DSC / classifier shapes, a long-K case)
GEMM_SHAPES = [(128, 256, 256), (300, 256, 256), (1000, 768, 256), (257, 512, 256), (4096, 96, 128), (130, 7040, 256),
               (512, 128, 128), (333, 64, 72), (2048, 256, 6997 + 3), (96, 32, 512)]


@pytest.mark.parametrize("shape", GEMM_SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_tn_forward(ops, shape, out_dtype):
    """y = act(x W^T + b): both operands K-major (nn.Linear forward)"""
    m, n, k = shape
    x = rnd(m, k, seed=1).bfloat16()
    w = rnd(n, k, seed=2, scale=1 / math.sqrt(k)).bfloat16()
    b = rnd(n, seed=3, scale=0.1)
    y = torch.empty(m, n, dtype=out_dtype, device=DEV)
    n0 = tc_calls()
    ops.gemm(x, w, y, m, n, k, trans_b=True, lda=k, ldb=k, ldc=n, bias=b, bias_mode=1, relu=True)
    assert tc_calls() == n0 + 1, "tensor-core kernel did not run"
    ref = torch.relu(x.float() @ w.float().t() + b)
    assert rel_err(y.float(), ref) < (2e-5 if out_dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("shape", [(1000, 256, 768), (300, 256, 256), (4096, 256, 512), (200, 128, 7040)])
def test_gemm_nn_dgrad(ops, shape):
    """dx = dy W: A K-major, B MN-major (rows of W are the reduction index)"""
    m, n, k = shape  # dx [m, n], dy [m, k], W [k, n]
    dy = rnd(m, k, seed=4).bfloat16()
    w = rnd(k, n, seed=5, scale=1 / math.sqrt(k)).bfloat16()
    dx = torch.empty(m, n, dtype=torch.bfloat16, device=DEV)
    n0 = tc_calls()
    ops.gemm(dy, w, dx, m, n, k, lda=k, ldb=n, ldc=n)
    assert tc_calls() == n0 + 1
    assert rel_err(dx.float(), dy.float() @ w.float()) < 6e-3
    # accumulate into an existing bf16 tensor (dx = ds + dh W1 in the decoder backward)
    base = rnd(m, n, seed=6).bfloat16()
    acc = base.clone()
    ops.gemm(dy, w, acc, m, n, k, lda=k, ldb=n, ldc=n, accumulate=True)
    assert rel_err(acc.float(), base.float() + dy.float() @ w.float()) < 8e-3


@pytest.mark.parametrize("shape", [(256, 256, 16384), (768, 256, 5000), (512, 256, 74784 // 8), (7000, 256, 1024), (256, 128, 640)])
def test_gemm_tn_wgrad_split_k(ops, shape):
    """dW += dy^T x: both operands MN-major, fp32 output, long reduction split across CTAs with atomics"""
    n_out, k_in, rows = shape
    dy = rnd(rows, n_out, seed=7).bfloat16()
    x = rnd(rows, k_in, seed=8).bfloat16()
    dw = torch.zeros(n_out, k_in, dtype=torch.float32, device=DEV)
    n0 = tc_calls()
    ops.gemm(dy, x, dw, n_out, k_in, rows, trans_a=True, lda=n_out, ldb=k_in, ldc=k_in, accumulate=True)
    assert tc_calls() == n0 + 1
    ref = dy.float().t() @ x.float()
    assert rel_err(dw, ref) < 2e-5
    ops.gemm(dy, x, dw, n_out, k_in, rows, trans_a=True, lda=n_out, ldb=k_in, ldc=k_in, accumulate=True)
    assert rel_err(dw, 2 * ref) < 2e-5


def test_gemm_strided_views(ops):
    """row-sliced weights / column-sliced outputs as used for the packed in-proj (ld != width)"""
    m, d = 700, 256
    x = rnd(m, d, seed=9).bfloat16()
    w = rnd(3 * d, d, seed=10, scale=1 / 16).bfloat16()
    b = rnd(3 * d, seed=11, scale=0.1)
    out = torch.zeros(m, 2 * d, dtype=torch.bfloat16, device=DEV)
    n0 = tc_calls()
    ops.linear_fwd(x, w[d:], b[d:], out=out)
    assert tc_calls() == n0 + 1
    assert rel_err(out.float(), x.float() @ w[d:].float().t() + b[d:]) < 6e-3


# N, H, W, Ci, Co, stride  -- GrandStaff-like widths (multiples of 128 and the odd audio sizes), all channel pairs
CONV_TC_CASES = [
    (2, 8, 256, 16, 16, (1, 1)),
    (1, 12, 300, 16, 32, (1, 1)),
    (2, 9, 256, 32, 32, (2, 2)),
    (1, 13, 101, 32, 64, (1, 1)),
    (2, 10, 202, 64, 64, (2, 2)),
    (1, 7, 128, 64, 128, (1, 1)),
    (2, 16, 128, 128, 128, (2, 1)),
    (1, 25, 101, 128, 128, (1, 1)),
    (1, 5, 808, 16, 16, (1, 1)),
    (3, 6, 20, 32, 32, (1, 1)),
]


@pytest.mark.parametrize("case", CONV_TC_CASES)
def test_conv3x3_tc_fwd_dgrad(ops, case):
    import torch.nn.functional as F

    n, h, w, ci, co, st = case
    x = rnd(n, ci, h, w, seed=21).bfloat16()
    wt = rnd(co, ci, 3, 3, seed=22, scale=1.0 / math.sqrt(9 * ci))
    b = rnd(co, seed=23, scale=0.1)
    wq = wt.bfloat16().float()
    xr = x.float().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wq, b, stride=st, padding=1))
    xn = x.permute(0, 2, 3, 1).contiguous()
    wp = ops.pack_conv_weight(wt.contiguous(), torch.bfloat16, False)
    wpt = ops.pack_conv_weight(wt.contiguous(), torch.bfloat16, True)
    n0 = tc_calls()
    y = ops.conv3x3_fwd(xn, wp, b, st, relu=True)
    assert tc_calls() == n0 + 1, "tensor-core conv did not run"
    assert rel_err(y.permute(0, 3, 1, 2).float(), yr) < 6e-3
    gy = rnd(*yr.shape, seed=24).bfloat16()
    (gx_ref,) = torch.autograd.grad(F.conv2d(xr, wq, None, stride=st, padding=1), xr, gy.float())
    dy = gy.permute(0, 2, 3, 1).contiguous()
    n0 = tc_calls()
    dx = ops.conv3x3_dgrad(dy, wpt, (h, w), st)
    assert tc_calls() == n0 + 1, "tensor-core dgrad did not run"
    assert rel_err(dx.permute(0, 3, 1, 2).float(), gx_ref) < 6e-3


# a larger multi-sample case per channel width so that a CTA's contiguous tile range crosses sample boundaries (flushes)
FUSED_CASES = CONV_TC_CASES + [(5, 40, 384, 16, 16, (1, 1)), (6, 33, 260, 32, 32, (1, 1)), (5, 24, 200, 64, 64, (1, 1)),
                               (4, 30, 150, 32, 32, (2, 2)), (3, 21, 131, 64, 64, (2, 2))]


@pytest.mark.parametrize("case", FUSED_CASES)
def test_conv3x3_fused_epilogue_sums(ops, case):
    """The per-channel reductions accumulated in the convolution epilogues equal the same sums taken over the STORED
    result: InstanceNorm statistics with the forward (in_sums), bias-gradient column sums (colsum, with the fused ReLU mask)
    and InstanceNorm backward sums (in_bsums) with the data gradient.  C > 64 takes the separate-pass fallback: same contract."""
    n, h, w, ci, co, st = case
    x = rnd(n, h, w, ci, seed=31).bfloat16().relu()
    wt = rnd(co, ci, 3, 3, seed=32, scale=1.0 / math.sqrt(9 * ci))
    b = rnd(co, seed=33, scale=0.1)
    wp = ops.pack_conv_weight(wt.contiguous(), torch.bfloat16, False)
    wpt = ops.pack_conv_weight(wt.contiguous(), torch.bfloat16, True)
    # forward + statistics
    y_plain = ops.conv3x3_fwd(x, wp, b, st, relu=True)
    sums = ops.in_sums_buffer(n, co, DEV)
    sums.fill_(float("nan"))  # the entry point must initialise it
    y = ops.conv3x3_fwd(x, wp, b, st, relu=True, in_sums=sums)
    assert torch.equal(y, y_plain)
    yd = y.double()
    want = torch.stack([yd.sum(dim=(1, 2)), (yd * yd).sum(dim=(1, 2))], dim=-1)
    assert rel_err(sums, want) < 2e-6
    # data gradient + bias-gradient column sums (mask = the conv's forward input, a ReLU output)
    ho, wo = y.shape[1], y.shape[2]
    dy = rnd(n, ho, wo, co, seed=34).bfloat16()
    dx_plain = ops.conv3x3_dgrad(dy, wpt, (h, w), st, mask=x, mask_scale=2.0)
    col = torch.full((ci,), 3.0, device=DEV)
    dx = ops.conv3x3_dgrad(dy, wpt, (h, w), st, mask=x, mask_scale=2.0, colsum=col)
    assert torch.equal(dx, dx_plain)
    assert rel_err(col - 3.0, dx.double().sum(dim=(0, 1, 2))) < 1e-4  # fp32 atomics into a buffer that starts at 3.0
    # data gradient + InstanceNorm backward sums
    xin = rnd(n, h, w, ci, seed=35).bfloat16()
    bs = ops.in_sums_buffer(n, ci, DEV)
    bs.fill_(float("nan"))
    dx2_plain = ops.conv3x3_dgrad(dy, wpt, (h, w), st)
    dx2 = ops.conv3x3_dgrad(dy, wpt, (h, w), st, in_x=xin, in_bsums=bs)
    assert torch.equal(dx2, dx2_plain)
    d2 = dx2.double()
    want_b = torch.stack([d2.sum(dim=(1, 2)), (d2 * xin.double()).sum(dim=(1, 2))], dim=-1)
    assert float((bs - want_b).abs().max()) < 2e-5 * float(want_b.abs().max()) + 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("c,hw", [(16, (40, 384)), (64, (24, 200)), (128, (9, 33))])
def test_instnorm_with_sums_from_the_convolutions(ops, dtype, c, hw):
    """instnorm_fwd / instnorm_bwd fed with the sums the convolution epilogues produce (sums_ready) == their own reduction passes"""
    n = 3
    x = rnd(n, hw[0], hw[1], c, seed=41).to(dtype).relu()
    dy = rnd(n, hw[0], hw[1], c, seed=42).to(dtype)
    y0, st0 = ops.instnorm_fwd(x, 1e-3)
    xd = x.double()
    sums = torch.stack([xd.sum(dim=(1, 2)), (xd * xd).sum(dim=(1, 2))], dim=-1).contiguous()
    y1, st1 = ops.instnorm_fwd(x, 1e-3, sums=sums)
    assert rel_err(st1, st0) < 1e-6 and rel_err(y1.float(), y0.float()) < (1e-6 if dtype == torch.float32 else 4e-3)
    dx0 = ops.instnorm_bwd(dy, x, st0, relu_mask=True, mask_scale=2.0)
    dd = dy.double()
    bsums = torch.stack([dd.sum(dim=(1, 2)), (dd * xd).sum(dim=(1, 2))], dim=-1).contiguous()
    col = torch.zeros(c, device=DEV)
    dx1 = ops.instnorm_bwd(dy, x, st0, relu_mask=True, mask_scale=2.0, sums=bsums, colsum=col)
    assert rel_err(dx1.float(), dx0.float()) < (2e-5 if dtype == torch.float32 else 4e-3)
    assert float((col.double() - dx1.double().sum(dim=(0, 1, 2))).abs().max()) < 1e-4 * float(dx1.double().abs().sum(dim=(0, 1, 2)).max())


@pytest.mark.parametrize("case", CONV_TC_CASES + [(4, 64, 512, 16, 16, (1, 1)), (2, 32, 256, 64, 128, (2, 2)), (2, 20, 70, 32, 16, (1, 1)),
                                                  (3, 17, 99, 32, 32, (1, 1))])
def test_conv3x3_tc_wgrad(ops, case):
    import torch.nn.functional as F

    n, h, w, ci, co, st = case
    x = rnd(n, ci, h, w, seed=31).bfloat16()
    wt = rnd(co, ci, 3, 3, seed=32, scale=1.0 / math.sqrt(9 * ci)).requires_grad_(True)
    yr = F.conv2d(x.float(), wt, None, stride=st, padding=1)
    gy = rnd(*yr.shape, seed=33).bfloat16()
    (gw_ref,) = torch.autograd.grad(yr, wt, gy.float())
    xn = x.permute(0, 2, 3, 1).contiguous()
    dy = gy.permute(0, 2, 3, 1).contiguous()
    dw = torch.zeros(co, ci, 3, 3, dtype=torch.float32, device=DEV)
    db = torch.zeros(co, dtype=torch.float32, device=DEV)
    n0 = tc_calls()
    ops.conv3x3_wgrad(xn, dy, dw, db, st, accumulate=True)
    assert tc_calls() == n0 + 1, "tensor-core wgrad did not run"
    assert rel_err(dw, gw_ref) < 1e-4
    assert rel_err(db, gy.float().sum(dim=(0, 2, 3))) < 1e-4
    ops.conv3x3_wgrad(xn, dy, dw, db, st, accumulate=True)
    assert rel_err(dw, 2 * gw_ref) < 1e-4
    ops.conv3x3_wgrad(xn, dy, dw, db, st, accumulate=False)
    assert rel_err(dw, gw_ref) < 1e-4


def _attn_ref(q, k, v, scale, causal, window, key_bias):
    """q [B,Tq,H,64], k/v [B,Tk,H,64] fp32 -> (o [B,Tq,H*64], lse [B,H,Tq]) with the reference mask algebra"""
    b, tq, h, d = q.shape
    tk = k.shape[1]
    s = torch.einsum("bthd,bshd->bhts", q, k) * scale
    if key_bias is not None:
        s = s + key_bias[:, None, None, :]
    if causal:
        i = torch.arange(tq, device=q.device)[:, None] + (tk - tq)
        j = torch.arange(tk, device=q.device)[None, :]
        ok = j <= i
        if window > 0:
            ok = ok & (j >= i - window)
        s = s.masked_fill(~ok, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhts,bshd->bthd", p, v).reshape(b, tq, h * d)
    return o, lse


ATTN_CASES = [
    # B, H, Tq, Tk, causal, window, bias kind
    (2, 4, 300, 300, True, 0, "plus1"),
    (1, 4, 512, 512, True, 100, None),
    (2, 4, 200, 700, False, 0, "neginf"),
    (1, 2, 129, 2337, False, 0, "plus1"),
    (3, 4, 64, 64, True, 5, None),
    (1, 4, 1, 37, False, 0, "neginf"),
    # more work items than SMs: the persistent backward walks several (key tile, head, batch) items per CTA -- K/V double
    # buffering, the dK / dV drain between items, dead (fully masked) key tiles, items of different length (causal)
    (20, 4, 200, 700, False, 0, "neginf"),
    (40, 4, 300, 300, True, 0, None),
]


@pytest.mark.parametrize("case", ATTN_CASES)
def test_attention_fwd_tc(ops, case):
    from omr_a2s_multimodal_transformer_b200.ops import AttnSpec

    b, h, tq, tk, causal, window, bias_kind = case
    d = h * 64
    self_attn = tq == tk and causal
    if self_attn:
        qkv = rnd(b, tq, 3 * d, seed=41).bfloat16()
        qb, kb_, vb, qo, ko, vo = qkv, qkv, qkv, 0, d, 2 * d
    else:
        qb = rnd(b, tq, d, seed=42).bfloat16()
        kv = rnd(b, tk, 2 * d, seed=43).bfloat16()
        kb_, vb, qo, ko, vo = kv, kv, 0, 0, d
    key_bias = None
    if bias_kind is not None:
        lens = torch.randint(tk // 2, tk + 1, (b,), generator=torch.Generator().manual_seed(5)).to(DEV)
        pad = torch.arange(tk, device=DEV)[None, :] >= lens[:, None]
        key_bias = torch.zeros(b, tk, device=DEV).masked_fill(pad, 1.0 if bias_kind == "plus1" else float("-inf"))
    spec = AttnSpec(h, 64, causal=causal, window=window, key_bias=key_bias)
    n0 = tc_calls()
    o, lse = ops.attn_fwd(qb, qo, kb_, ko, vb, vo, spec)
    assert tc_calls() == n0 + 1, "tensor-core attention did not run"
    q4 = qb[:, :, qo:qo + d].float().reshape(b, tq, h, 64)
    k4 = kb_[:, :, ko:ko + d].float().reshape(b, tk, h, 64)
    v4 = vb[:, :, vo:vo + d].float().reshape(b, tk, h, 64)
    o_ref, lse_ref = _attn_ref(q4, k4, v4, 0.125, causal, window, key_bias)
    assert rel_err(o.float(), o_ref) < 8e-3
    assert float((lse - lse_ref).abs().max()) < 2e-3


@pytest.mark.parametrize("case", ATTN_CASES)
def test_attention_bwd_tc(ops, case):
    from omr_a2s_multimodal_transformer_b200.ops import AttnSpec

    b, h, tq, tk, causal, window, bias_kind = case
    d = h * 64
    self_attn = tq == tk and causal
    if self_attn:
        qkv = rnd(b, tq, 3 * d, seed=41).bfloat16()
        qb, kb_, vb, qo, ko, vo = qkv, qkv, qkv, 0, d, 2 * d
    else:
        qb = rnd(b, tq, d, seed=42).bfloat16()
        kv = rnd(b, tk, 2 * d, seed=43).bfloat16()
        kb_, vb, qo, ko, vo = kv, kv, 0, 0, d
    key_bias = None
    if bias_kind is not None:
        lens = torch.randint(tk // 2, tk + 1, (b,), generator=torch.Generator().manual_seed(5)).to(DEV)
        pad = torch.arange(tk, device=DEV)[None, :] >= lens[:, None]
        key_bias = torch.zeros(b, tk, device=DEV).masked_fill(pad, 1.0 if bias_kind == "plus1" else float("-inf"))
    spec = AttnSpec(h, 64, causal=causal, window=window, key_bias=key_bias)
    o, lse = ops.attn_fwd(qb, qo, kb_, ko, vb, vo, spec)
    do = rnd(b, tq, d, seed=44).bfloat16()
    q4 = qb[:, :, qo:qo + d].float().reshape(b, tq, h, 64).requires_grad_(True)
    k4 = kb_[:, :, ko:ko + d].float().reshape(b, tk, h, 64).requires_grad_(True)
    v4 = vb[:, :, vo:vo + d].float().reshape(b, tk, h, 64).requires_grad_(True)
    o_ref, _ = _attn_ref(q4, k4, v4, 0.125, causal, window, key_bias)
    gq, gk, gv = torch.autograd.grad(o_ref, (q4, k4, v4), do.float())
    if self_attn:
        dqkv = torch.zeros_like(qkv)
        dqb, dkb, dvb = dqkv, dqkv, dqkv
    else:
        dqb = torch.zeros_like(qb)
        dkv = torch.zeros_like(kv)
        dkb, dvb = dkv, dkv
    n0 = tc_calls()
    ops.attn_bwd(qb, qo, kb_, ko, vb, vo, o, do, lse, dqb, qo, dkb, ko, dvb, vo, spec)
    assert tc_calls() == n0 + 1, "tensor-core attention backward did not run"
    assert rel_err(dqb[:, :, qo:qo + d].float(), gq.reshape(b, tq, d)) < 1.5e-2
    assert rel_err(dkb[:, :, ko:ko + d].float(), gk.reshape(b, tk, d)) < 1.5e-2
    assert rel_err(dvb[:, :, vo:vo + d].float(), gv.reshape(b, tk, d)) < 1.5e-2


@pytest.mark.parametrize("shape", [(3, 4, 330, 260), (2, 4, 1313, 1024), (5, 4, 128, 257)])
def test_attention_mixer_block_mask_tc(ops, shape):
    """The attention mixers' block mask (reference model.py:340-352: pairs with query >= q_len AND key >= kv_len are excluded,
    repeated head-major so that (b, h) takes the lengths of sample (b H + h) mod B) on the tcgen05 flash kernels, forward
    and backward, against fp32 torch with the same mask."""
    from omr_a2s_multimodal_transformer_b200.ops import AttnSpec

    b, h, tq, tk = shape
    d = h * 64
    qb = rnd(b, tq, d, seed=52).bfloat16()
    kv = rnd(b, tk, 2 * d, seed=53).bfloat16()
    g = torch.Generator().manual_seed(8)
    q_len = torch.randint(tq // 3, tq + 1, (b,), generator=g).to(torch.int32).to(DEV)
    kv_len = torch.randint(tk // 3, tk + 1, (b,), generator=g).to(torch.int32).to(DEV)
    spec = AttnSpec(h, 64, q_len=q_len, kv_len=kv_len, quirk_mod=b)
    n0 = tc_calls()
    o, lse = ops.attn_fwd(qb, 0, kv, 0, kv, d, spec)
    assert tc_calls() == n0 + 1, "tensor-core attention did not take the mixer block mask"
    q4 = qb.float().reshape(b, tq, h, 64).requires_grad_(True)
    k4 = kv[:, :, :d].float().reshape(b, tk, h, 64).requires_grad_(True)
    v4 = kv[:, :, d:].float().reshape(b, tk, h, 64).requires_grad_(True)
    s = torch.einsum("bthd,bshd->bhts", q4, k4) * 0.125
    idx = (torch.arange(b, device=DEV)[:, None] * h + torch.arange(h, device=DEV)[None, :]) % b  # [B,H] -> sample whose lengths apply
    lq, lk = q_len[idx].long(), kv_len[idx].long()
    masked = (torch.arange(tq, device=DEV)[None, None, :, None] >= lq[:, :, None, None]) & \
             (torch.arange(tk, device=DEV)[None, None, None, :] >= lk[:, :, None, None])
    s = s.masked_fill(masked, float("-inf"))
    o_ref = torch.einsum("bhts,bshd->bthd", torch.softmax(s, dim=-1), v4).reshape(b, tq, d)
    assert rel_err(o.float(), o_ref) < 8e-3
    assert float((lse - torch.logsumexp(s, dim=-1)).abs().max()) < 2e-3
    do = rnd(b, tq, d, seed=54).bfloat16()
    gq, gk, gv = torch.autograd.grad(o_ref, (q4, k4, v4), do.float())
    dq, dkv = torch.zeros_like(qb), torch.zeros_like(kv)
    n0 = tc_calls()
    ops.attn_bwd(qb, 0, kv, 0, kv, d, o, do, lse, dq, 0, dkv, 0, dkv, d, spec)
    assert tc_calls() == n0 + 1, "tensor-core attention backward did not take the mixer block mask"
    assert rel_err(dq.float(), gq.reshape(b, tq, d)) < 1.5e-2
    assert rel_err(dkv[:, :, :d].float(), gk.reshape(b, tk, d)) < 1.5e-2
    assert rel_err(dkv[:, :, d:].float(), gv.reshape(b, tk, d)) < 1.5e-2


def test_gemm_batched_logits_layout_tc(ops):
    """the public forward()'s class-major logits [B,V,T] = W [V,D] @ hidden[b]^T + bias[:, None] (reference decoder.py:145-146)
    are served by the tensor-core GEMM, one launch per batch element"""
    b, v, t, d = 3, 6997, 128, 256
    w = rnd(v, d, seed=61, scale=1 / 16).bfloat16()
    hidden = rnd(b, t, d, seed=62).bfloat16()
    bias = rnd(v, seed=63, scale=0.1)
    out = torch.empty(b, v, t, dtype=torch.bfloat16, device=DEV)
    n0 = tc_calls()
    ops.gemm(w, hidden, out, v, t, d, trans_b=True, lda=d, ldb=d, ldc=t, batch=b, stride_b=t * d, stride_c=v * t, bias=bias, bias_mode=2)
    assert tc_calls() == n0 + 1, "batched GEMM fell back to the CUDA-core kernel"
    ref = torch.einsum("vd,btd->bvt", w.float(), hidden.float()) + bias[None, :, None]
    assert rel_err(out.float(), ref) < 6e-3


@pytest.mark.parametrize("shape", [(300, 97), (1000, 6997), (129, 64)])
def test_proj_ce_fused_tc(ops, shape):
    """Classifier fused with the softmax cross-entropy (csrc/projce_tc.cu; reference decoder.py:145-146 +
    model.py:109,444): loss, per-row lse and all three gradients against torch on the same bf16 inputs -- ragged row /
    class tails, ignored rows, a target in every 32-column slot of a tile."""
    rows, v = shape
    d = 256
    x = rnd(rows, d, seed=51, scale=1.0).bfloat16()
    w = rnd(v, d, seed=52, scale=0.08).bfloat16()
    bias = rnd(v, seed=53, scale=0.5).float()
    gen = torch.Generator().manual_seed(54)
    tg = torch.randint(0, v, (rows,), generator=gen)
    tg[::7] = 0  # ignore_index rows
    tg[1] = v - 1
    tg = tg.to(DEV)
    assert ops.proj_ce_supported(torch.bfloat16, d)
    n0 = tc_calls()
    loss_out, row_lse = ops.proj_ce_fwd(x, w, bias, tg, 0)
    assert tc_calls() == n0 + 1, "fused projection + cross-entropy did not run on the tensor-core kernel"
    xr, wr, br = x.float().requires_grad_(True), w.float().requires_grad_(True), bias.clone().requires_grad_(True)
    logits = xr @ wr.t() + br
    ref = torch.nn.functional.cross_entropy(logits, tg, ignore_index=0)
    assert abs(float(loss_out[0]) - float(ref)) < 2e-3 * max(1.0, abs(float(ref)))
    assert int(loss_out[1]) == int((tg != 0).sum())
    assert float((row_lse - torch.logsumexp(logits.detach(), dim=1)).abs().max()) < 5e-3
    gscale = torch.tensor([1.7], device=DEV)
    (1.7 * ref).backward()
    dx = ops.proj_ce_bwd_dx(x, w, bias, tg, row_lse, loss_out, gscale, 0)
    dw = torch.zeros(v, d, device=DEV)
    db = torch.zeros(v, device=DEV)
    ops.proj_ce_bwd_dw(x, w, bias, tg, row_lse, loss_out, gscale, 0, dw, db)
    assert rel_err(dx.float(), xr.grad) < 1.5e-2
    assert rel_err(dw, wr.grad) < 1.5e-2
    assert rel_err(db, br.grad) < 1.5e-2
    # accumulation into dw / db
    ops.proj_ce_bwd_dw(x, w, bias, tg, row_lse, loss_out, gscale, 0, dw, db)
    assert rel_err(dw, 2 * wr.grad) < 1.5e-2 and rel_err(db, 2 * br.grad) < 1.5e-2
