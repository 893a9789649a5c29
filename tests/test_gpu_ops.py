"""Kernel-level parity (-m gpu): every C-ABI entry point against the plain torch op it replaces, on
the same seeded inputs, in fp32 (tight tolerance) and bf16 (tolerance of the storage type)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]


def tol(dtype, fp32=2e-5, bf16=2e-2):
    return fp32 if dtype == torch.float32 else bf16


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev())


@pytest.fixture(scope="module")
def ops():
    from omr_a2s_multimodal_transformer_b200 import ops as o

    return o


def to_nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def to_nchw(x):
    return x.permute(0, 3, 1, 2).float()


CONV_CASES = [
    # N, H, W, Ci, Co, stride
    (2, 9, 13, 1, 16, (1, 1)),
    (2, 9, 70, 1, 16, (1, 1)),   # first layer, strip-walking weight gradient (ragged last strip)
    (1, 3, 97, 1, 16, (1, 1)),
    (2, 12, 20, 16, 16, (1, 1)),
    (1, 11, 17, 16, 32, (2, 2)),
    (2, 8, 10, 32, 64, (2, 2)),
    (1, 7, 9, 64, 128, (2, 1)),
    (1, 5, 6, 128, 128, (1, 1)),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3x3_fwd_dgrad_wgrad(ops, dtype, case):
    n, h, w, ci, co, st = case
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * ci))
    b = rnd(co, seed=3, scale=0.1)
    xq, wq = x.to(dtype).float(), wt.to(dtype).float()
    xr = xq.clone().requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wr, br, stride=st, padding=1))
    xn = to_nhwc(x, dtype)
    wp = ops.pack_conv_weight(wt.contiguous(), dtype, False)
    wpt = ops.pack_conv_weight(wt.contiguous(), dtype, True)
    y = ops.conv3x3_fwd(xn, wp, b, st, relu=True)
    assert rel_err(to_nchw(y), yr) < tol(dtype)
    gy = rnd(*yr.shape, seed=4)
    gyq = gy.to(dtype).float()
    yr.backward(gyq)
    dy = to_nhwc(gy, dtype)
    dz = ops.relu_bwd(y, dy)
    dx = ops.conv3x3_dgrad(dz, wpt, (h, w), st)
    assert rel_err(to_nchw(dx), xr.grad) < tol(dtype, 5e-5, 3e-2)
    dw = torch.zeros_like(wt)
    db = torch.zeros_like(b)
    ops.conv3x3_wgrad(xn, dz, dw, db, st, accumulate=True)
    assert rel_err(dw, wr.grad) < tol(dtype, 5e-5, 3e-2)
    assert rel_err(db, br.grad) < tol(dtype, 5e-5, 3e-2)
    ops.conv3x3_wgrad(xn, dz, dw, db, st, accumulate=True)  # accumulation doubles
    assert rel_err(dw, 2 * wr.grad) < tol(dtype, 5e-5, 3e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 8, 16, 128), (1, 13, 11, 256), (3, 5, 7, 32), (8, 13, 101, 256), (16, 8, 128, 128)])
def test_dwconv3x3(ops, dtype, shape):
    n, h, w, c = shape
    x = rnd(n, c, h, w, seed=5)
    wt = rnd(c, 1, 3, 3, seed=6, scale=1 / 3)
    b = rnd(c, seed=7, scale=0.1)
    xq, wq = x.to(dtype).float(), wt.to(dtype).float()
    xr, wr, br = xq.clone().requires_grad_(True), wq.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, padding=1, groups=c)
    xn = to_nhwc(x, dtype)
    wp = ops.pack_dw_weight(wt.contiguous(), dtype)
    y = ops.dwconv3x3_fwd(xn, wp, b)
    assert rel_err(to_nchw(y), yr) < tol(dtype)
    gy = rnd(*yr.shape, seed=8)
    yr.backward(gy.to(dtype).float())
    dy = to_nhwc(gy, dtype)
    dx = ops.dwconv3x3_dgrad(dy, wp)
    assert rel_err(to_nchw(dx), xr.grad) < tol(dtype, 5e-5, 3e-2)
    dw, db = torch.zeros_like(wt), torch.zeros_like(b)
    ops.dwconv3x3_wgrad(xn, dy, dw, db)
    assert rel_err(dw, wr.grad) < tol(dtype, 5e-5, 3e-2)
    assert rel_err(db, br.grad) < tol(dtype, 5e-5, 3e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 24, 16), (3, 9, 7, 32), (2, 8, 16, 128), (1, 13, 11, 256), (2, 4, 6, 512)])
def test_instnorm(ops, dtype, shape):
    n, h, w, c = shape
    x = rnd(n, c, h, w, seed=9) * 1.5 + 0.7
    xq = x.to(dtype).float()
    xr = xq.clone().requires_grad_(True)
    yr = F.instance_norm(xr, eps=1e-3)
    xn = to_nhwc(x, dtype)
    y, stats = ops.instnorm_fwd(xn, 1e-3)
    assert rel_err(to_nchw(y), yr) < tol(dtype, 2e-5, 1e-2)
    gy = rnd(*yr.shape, seed=10)
    yr.backward(gy.to(dtype).float())
    dx = ops.instnorm_bwd(to_nhwc(gy, dtype), xn, stats)
    assert rel_err(to_nchw(dx), xr.grad) < tol(dtype, 1e-4, 3e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("d", [64, 256, 512])
def test_add_layernorm(ops, dtype, d):
    rows = 37
    x, r = rnd(rows, d, seed=11), rnd(rows, d, seed=12)
    g, b = 1 + 0.1 * rnd(d, seed=13), 0.1 * rnd(d, seed=14)
    xq, rq = x.to(dtype), r.to(dtype)
    s_ref = (xq.float() + rq.float()).to(dtype).float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.layer_norm(s_ref, (d,), gr, br, 1e-5)
    y, s, st = ops.add_layernorm_fwd(xq.contiguous(), rq.contiguous(), g, b, 1e-5, True)
    assert rel_err(y, yr) < tol(dtype, 1e-5, 1e-2)
    assert rel_err(s, s_ref) < 1e-6
    gy = rnd(rows, d, seed=15)
    yr.backward(gy.to(dtype).float())
    dg, db = torch.zeros_like(g), torch.zeros_like(b)
    ds = ops.layernorm_bwd(gy.to(dtype).contiguous(), s, st, g, dg, db)
    assert rel_err(ds, s_ref.grad) < tol(dtype, 5e-5, 2e-2)
    assert rel_err(dg, gr.grad) < tol(dtype, 5e-5, 2e-2)
    assert rel_err(db, br.grad) < tol(dtype, 5e-5, 2e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mnk", [(70, 50, 33), (128, 256, 256), (5, 6997, 256), (300, 768, 256), (64, 64, 16)])
def test_linear_fwd_dgrad_wgrad(ops, dtype, mnk):
    m, n, k = mnk
    x, w, b = rnd(m, k, seed=16), rnd(n, k, seed=17, scale=1 / math.sqrt(k)), rnd(n, seed=18, scale=0.1)
    xq, wq = x.to(dtype), w.to(dtype)
    xr, wr, br = xq.float().requires_grad_(True), wq.float().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.relu(F.linear(xr, wr, br))
    y = ops.linear_fwd(xq.contiguous(), wq.contiguous(), b, relu=True)
    assert rel_err(y, yr) < tol(dtype)
    gy = rnd(m, n, seed=19)
    yr.backward(gy.to(dtype).float())
    dz = ops.relu_bwd(y, gy.to(dtype).contiguous())
    dx = ops.linear_dgrad(dz, wq.contiguous())
    assert rel_err(dx, xr.grad) < tol(dtype, 5e-5, 3e-2)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.linear_wgrad(xq.contiguous(), dz, dw, db)
    assert rel_err(dw, wr.grad) < tol(dtype, 5e-5, 3e-2)
    assert rel_err(db, br.grad) < tol(dtype, 5e-5, 3e-2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_batched_bias_rows(ops, dtype):
    """the [B,V,T] classifier layout: C_b[V,T] = W[V,D] @ H_b[T,D]^T + bias[V]"""
    b, v, t, d = 3, 97, 21, 64
    w, h, bias = rnd(v, d, seed=20, scale=0.2), rnd(b, t, d, seed=21), rnd(v, seed=22)
    wq, hq = w.to(dtype).contiguous(), h.to(dtype).contiguous()
    out = torch.empty(b, v, t, dtype=dtype, device=dev())
    ops.gemm(wq, hq, out, v, t, d, trans_b=True, lda=d, ldb=d, ldc=t, batch=b, stride_b=t * d, stride_c=v * t, bias=bias,
             bias_mode=2)
    ref = torch.einsum("vd,btd->bvt", wq.float(), hq.float()) + bias[None, :, None]
    assert rel_err(out, ref) < tol(dtype)


def ref_attention(q, k, v, scale, bias=None, causal=False, window=0, q_len=None, kv_len=None, quirk_mod=0):
    """q [B,H,Tq,hd] etc. fp32 -> o [B,H,Tq,hd]"""
    B, H, Tq, _ = q.shape
    Tk = k.shape[2]
    s = (q @ k.transpose(-1, -2)) * scale
    if bias is not None:
        s = s + bias[:, None, None, :]
    off = Tk - Tq
    i = torch.arange(Tq, device=q.device)[:, None]
    j = torch.arange(Tk, device=q.device)[None, :]
    if causal:
        ok = j <= i + off
        if window > 0:
            ok = ok & (j >= i + off - window)
        s = s.masked_fill(~ok, float("-inf"))
    if q_len is not None:
        for b in range(B):
            for h in range(H):
                sidx = (b * H + h) % quirk_mod if quirk_mod > 0 else b
                m = (i >= int(q_len[sidx])) & (j >= int(kv_len[sidx]))
                s[b, h] = s[b, h].masked_fill(m, float("-inf"))
    return torch.softmax(s, dim=-1) @ v


ATTN_CASES = [
    dict(B=2, H=4, Tq=37, Tk=37, causal=True),
    dict(B=2, H=2, Tq=130, Tk=130, causal=True, window=20),
    dict(B=3, H=4, Tq=19, Tk=150, bias="inf"),
    dict(B=3, H=4, Tq=140, Tk=700, bias="inf"),  # whole 128-key tiles masked out: skipped by the tensor-core kernels
    dict(B=2, H=4, Tq=70, Tk=70, causal=True, bias="one"),
    dict(B=3, H=4, Tq=45, Tk=67, quirk=True),
    dict(B=1, H=4, Tq=1, Tk=33),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", ATTN_CASES)
def test_attention_fwd_bwd(ops, dtype, case):
    B, H, Tq, Tk, hd = case["B"], case["H"], case["Tq"], case["Tk"], 64
    D = H * hd
    causal, window = case.get("causal", False), case.get("window", 0)
    self_attn = Tq == Tk and causal
    g = torch.Generator().manual_seed(Tq * 7 + Tk)
    if self_attn:
        qkv = torch.randn(B, Tq, 3 * D, generator=g).to(dev()).to(dtype).contiguous()
        qb, kb, vb, qo, ko, vo = qkv, qkv, qkv, 0, D, 2 * D
    else:
        qb = torch.randn(B, Tq, D, generator=g).to(dev()).to(dtype).contiguous()
        kvb = torch.randn(B, Tk, 2 * D, generator=g).to(dev()).to(dtype).contiguous()
        kb, vb, qo, ko, vo = kvb, kvb, 0, 0, D
    bias = None
    if case.get("bias") == "inf":
        lens = torch.tensor([Tk, Tk - 40, 5][:B])
        bias = torch.zeros(B, Tk)
        for b in range(B):
            bias[b, lens[b]:] = float("-inf")
        bias = bias.to(dev())
    elif case.get("bias") == "one":
        bias = (torch.rand(B, Tk, generator=g) < 0.3).float().to(dev())
    q_len = kv_len = None
    quirk_mod = 0
    if case.get("quirk"):
        q_len = torch.tensor([Tq, 20, 33][:B], dtype=torch.int32, device=dev())
        kv_len = torch.tensor([Tk, 50, 9][:B], dtype=torch.int32, device=dev())
        quirk_mod = B
    spec = ops.AttnSpec(H, hd, causal=causal, window=window, key_bias=bias, q_len=q_len, kv_len=kv_len, quirk_mod=quirk_mod)
    o, lse = ops.attn_fwd(qb, qo, kb, ko, vb, vo, spec)

    def heads(buf, off):
        return buf[:, :, off:off + D].float().view(B, -1, H, hd).transpose(1, 2).contiguous().requires_grad_(True)

    qr, kr, vr = heads(qb, qo), heads(kb, ko), heads(vb, vo)
    oref = ref_attention(qr, kr, vr, 1 / math.sqrt(hd), bias, causal, window, q_len, kv_len, quirk_mod)
    o_cmp = o.float().view(B, Tq, H, hd).transpose(1, 2)
    assert rel_err(o_cmp, oref) < tol(dtype, 2e-5, 1.5e-2)
    go = torch.randn(B, Tq, D, generator=g).to(dev()).to(dtype).contiguous()
    oref.backward(go.float().view(B, Tq, H, hd).transpose(1, 2))
    if self_attn:
        dqkv = torch.zeros_like(qkv)
        ops.attn_bwd(qb, qo, kb, ko, vb, vo, o, go, lse, dqkv, 0, dqkv, D, dqkv, 2 * D, spec)
        dq, dk, dv = dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]
    else:
        dqb, dkvb = torch.zeros_like(qb), torch.zeros_like(kvb)
        ops.attn_bwd(qb, qo, kb, ko, vb, vo, o, go, lse, dqb, 0, dkvb, 0, dkvb, D, spec)
        dq, dk, dv = dqb, dkvb[:, :, :D], dkvb[:, :, D:]

    def unheads(t):
        return t.transpose(1, 2).reshape(B, -1, D)

    assert rel_err(dq, unheads(qr.grad)) < tol(dtype, 1e-4, 3e-2)
    assert rel_err(dk, unheads(kr.grad)) < tol(dtype, 1e-4, 3e-2)
    assert rel_err(dv, unheads(vr.grad)) < tol(dtype, 1e-4, 3e-2)


def _extract_attn_dropout_mask(ops, dtype, B, H, Tq, Tk, causal, p, seed):
    """keep mask [B,H,Tq,Tk] of the attention-probability dropout, read back through the kernel itself: with Q = K = 0 the
    probabilities are uniform over the visible keys, and one-hot V blocks turn O into rows of P o M / (1-p)."""
    hd, D = 64, H * 64
    q = torch.zeros(B, Tq, D, device=dev(), dtype=dtype)
    mask = torch.zeros(B, H, Tq, Tk, dtype=torch.bool)
    for blk in range(0, Tk, hd):
        kv = torch.zeros(B, Tk, 2 * D, device=dev(), dtype=dtype)
        for h in range(H):
            for dcol in range(min(hd, Tk - blk)):
                kv[:, blk + dcol, D + h * hd + dcol] = 1.0
        spec = ops.AttnSpec(H, hd, causal=causal, dropout_p=p, seed=seed)
        o, _ = ops.attn_fwd(q, 0, kv, 0, kv, D, spec)
        o = o.float().view(B, Tq, H, hd).transpose(1, 2).cpu()  # [B,H,Tq,hd]
        n = min(hd, Tk - blk)
        mask[:, :, :, blk:blk + n] = o[:, :, :, :n] > 0
    return mask


@pytest.mark.parametrize("case", [dict(Tq=130, Tk=200, causal=False), dict(Tq=160, Tk=160, causal=True)])
def test_attention_dropout_fwd_bwd(ops, case):
    """attention-probability dropout: the fp32 CUDA-core kernels and the bf16 tcgen05 kernels draw the SAME keep mask (a pure
    function of seed, head, query, key), the forward and both backward kernels agree with a torch reference that applies
    that mask to softmax(S) (nn.MultiheadAttention's dropout), and the drop rate is the requested one."""
    B, H, Tq, Tk, hd, p, seed = 2, 4, case["Tq"], case["Tk"], 64, 0.25, 1234
    causal = case["causal"]
    D = H * hd
    m32 = _extract_attn_dropout_mask(ops, torch.float32, B, H, Tq, Tk, causal, p, seed)
    m16 = _extract_attn_dropout_mask(ops, torch.bfloat16, B, H, Tq, Tk, causal, p, seed)
    vis = torch.ones(Tq, Tk, dtype=torch.bool).tril(Tk - Tq) if causal else torch.ones(Tq, Tk, dtype=torch.bool)
    assert torch.equal(m32 & vis, m16 & vis)
    rate = 1.0 - float(m32[:, :, vis].float().mean())
    assert abs(rate - p) < 0.01, rate
    assert not torch.equal(m32[0, 0], m32[0, 1]) and not torch.equal(m32[0, 0], m32[1, 0])  # per (batch, head) streams
    other = _extract_attn_dropout_mask(ops, torch.float32, B, H, Tq, Tk, causal, p, seed + 1)
    assert not torch.equal(other & vis, m32 & vis)
    keep = m32.to(dev()).float() / (1.0 - round(p * 32768) / 32768.0)  # 15-bit uniforms (csrc/attn_drop.cuh)
    for dtype in DTYPES:
        g = torch.Generator().manual_seed(5)
        if causal:
            qkv = torch.randn(B, Tq, 3 * D, generator=g).to(dev()).to(dtype).contiguous()
            qb, kb, vb, qo, ko, vo = qkv, qkv, qkv, 0, D, 2 * D
        else:
            qb = torch.randn(B, Tq, D, generator=g).to(dev()).to(dtype).contiguous()
            kvb = torch.randn(B, Tk, 2 * D, generator=g).to(dev()).to(dtype).contiguous()
            kb, vb, qo, ko, vo = kvb, kvb, 0, 0, D
        spec = ops.AttnSpec(H, hd, causal=causal, dropout_p=p, seed=seed)
        o, lse = ops.attn_fwd(qb, qo, kb, ko, vb, vo, spec)

        def heads(buf, off):
            return buf[:, :, off:off + D].float().view(B, -1, H, hd).transpose(1, 2).contiguous().requires_grad_(True)

        qr, kr, vr = heads(qb, qo), heads(kb, ko), heads(vb, vo)
        sc = qr @ kr.transpose(-1, -2) / math.sqrt(hd)
        if causal:
            sc = sc.masked_fill(~vis.to(dev()), float("-inf"))
        oref = (torch.softmax(sc, dim=-1) * keep) @ vr
        assert rel_err(o.float().view(B, Tq, H, hd).transpose(1, 2), oref) < tol(dtype, 2e-5, 2e-2)
        go = torch.randn(B, Tq, D, generator=g).to(dev()).to(dtype).contiguous()
        oref.backward(go.float().view(B, Tq, H, hd).transpose(1, 2))
        if causal:
            dqkv = torch.zeros_like(qkv)
            ops.attn_bwd(qb, qo, kb, ko, vb, vo, o, go, lse, dqkv, 0, dqkv, D, dqkv, 2 * D, spec)
            dq, dk, dv = dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:]
        else:
            dqb, dkvb = torch.zeros_like(qb), torch.zeros_like(kvb)
            ops.attn_bwd(qb, qo, kb, ko, vb, vo, o, go, lse, dqb, 0, dkvb, 0, dkvb, D, spec)
            dq, dk, dv = dqb, dkvb[:, :, :D], dkvb[:, :, D:]

        def unheads(t):
            return t.transpose(1, 2).reshape(B, -1, D)

        assert rel_err(dq, unheads(qr.grad)) < tol(dtype, 1e-4, 4e-2)
        assert rel_err(dk, unheads(kr.grad)) < tol(dtype, 1e-4, 4e-2)
        assert rel_err(dv, unheads(vr.grad)) < tol(dtype, 1e-4, 4e-2)
    # the state is one-shot: a following call without dropout is unaffected
    spec0 = ops.AttnSpec(H, hd, causal=causal)
    o0, _ = ops.attn_fwd(qb, qo, kb, ko, vb, vo, spec0)
    o1, _ = ops.attn_fwd(qb, qo, kb, ko, vb, vo, spec0)
    assert torch.equal(o0, o1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cfg", [(4, 4, 300, 0), (32, 4, 2337, 0), (2, 4, 77, 20), (1, 4, 1, 0)])
def test_attn_decode(ops, dtype, cfg):
    B, H, Tk, window = cfg
    hd, D = 64, H * 64
    tmax = Tk + 9
    q = rnd(B, 3 * D, seed=30).to(dtype).contiguous()
    cache = rnd(B, tmax, 2 * D, seed=31).to(dtype).contiguous()
    esz = q.element_size()
    o = torch.empty(B, D, dtype=dtype, device=dev())
    ws = torch.empty(ops.attn_decode_ws_floats(B, H), dtype=torch.float32, device=dev())
    pos = torch.tensor([Tk - 1], dtype=torch.int32, device=dev())
    for pos_dev, tk_arg in ((None, Tk), (pos, tmax)):
        o.zero_()
        ops.attn_decode(q.data_ptr(), 3 * D, cache.data_ptr(), tmax * 2 * D, 2 * D, cache.data_ptr() + D * esz, tmax * 2 * D,
                        2 * D, o, None, ws, B, H, tk_arg, hd, window, dtype, pos_dev=pos_dev)
        qr = q[:, :D].float().view(B, 1, H, hd).transpose(1, 2)
        kr = cache[:, :Tk, :D].float().view(B, Tk, H, hd).transpose(1, 2)
        vr = cache[:, :Tk, D:].float().view(B, Tk, H, hd).transpose(1, 2)
        oref = ref_attention(qr, kr, vr, 1 / math.sqrt(hd), None, True, window)
        assert rel_err(o.float().view(B, H, hd), oref[:, :, 0]) < tol(dtype, 2e-5, 1.5e-2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_cross_entropy(ops, dtype):
    rows, v = 50, 997
    logits = (rnd(rows, v, seed=40) * 3).to(dtype).contiguous()
    tg = torch.randint(1, v, (rows,), generator=torch.Generator().manual_seed(1)).to(dev())
    tg[::7] = 0
    lr = logits.float().requires_grad_(True)
    ref = F.cross_entropy(lr, tg, ignore_index=0)
    loss_out, lse = ops.ce_fwd(logits, tg, 0)
    assert abs(float(loss_out[0]) - float(ref)) < tol(dtype, 1e-5, 1e-5) * max(1.0, abs(float(ref)))
    assert int(loss_out[1]) == int((tg != 0).sum())
    (ref * 1.7).backward()
    gs = torch.tensor([1.7], device=dev())
    dl = ops.ce_bwd(logits, tg, lse, loss_out, gs, 0, inplace=False)
    assert rel_err(dl, lr.grad) < tol(dtype, 1e-5, 1e-2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_embed_pe_and_bwd(ops, dtype):
    b, t, d, v = 3, 17, 64, 50
    table = rnd(v, d, seed=41)
    table[0] = 0
    pe = rnd(40, d, seed=42)
    tok = torch.randint(0, v, (b, t), generator=torch.Generator().manual_seed(2)).to(dev())
    out = ops.embed_pe_fwd(tok, table.to(dtype).contiguous(), pe, 3)
    ref = table.to(dtype).float()[tok] + pe[3:3 + t][None]
    assert rel_err(out, ref) < tol(dtype, 1e-6, 5e-3)
    g = rnd(b, t, d, seed=43).to(dtype).contiguous()
    dt = torch.zeros(v, d, device=dev())
    ops.embed_bwd(tok, g, dt, 0)
    ref_dt = torch.zeros(v, d, device=dev())
    ref_dt.index_add_(0, tok.reshape(-1), g.float().reshape(-1, d))
    ref_dt[0] = 0
    assert rel_err(dt, ref_dt) < 1e-5


def test_masks_pe2d_copy_dropout(ops):
    lens = torch.tensor([5, 0, 9], dtype=torch.int32, device=dev())
    bias = torch.full((3, 12), 7.0, device=dev())
    ops.key_bias_from_lengths(bias, lens, 2, 9, float("-inf"))
    ref = torch.full((3, 12), 7.0)
    for b, l in enumerate([5, 0, 9]):
        ref[b, 2:11] = 0
        ref[b, 2 + l:11] = float("-inf")
    assert torch.equal(bias.cpu(), ref)
    tok = torch.tensor([[3, 0, 5], [0, 0, 1]], device=dev())
    assert torch.equal(ops.key_bias_from_tokens(tok, 0, 1.0).cpu(), torch.tensor([[0., 1, 0], [1, 1, 0]]))
    for dtype in DTYPES:
        x = rnd(2, 3, 5, 8, seed=50).to(dtype).contiguous()
        pe = rnd(4, 7, 8, seed=51)
        out = torch.zeros(2, 20, 8, dtype=dtype, device=dev())
        ops.pe2d_add(x, pe, out, 4)
        ref = (x.float() + pe[:3, :5][None]).reshape(2, 15, 8)
        assert rel_err(out[:, 4:19], ref) < tol(dtype, 1e-6, 5e-3)
        assert float(out[:, :4].abs().sum()) == 0 and float(out[:, 19:].abs().sum()) == 0
        assert torch.equal(ops.copy_rows(out, 4, 15), out[:, 4:19])
        big = torch.ones(4, 32, 32, 16, dtype=dtype, device=dev())
        y = ops.dropout(big, 0.5, 1234)
        keep = (y != 0).float().mean().item()
        assert abs(keep - 0.5) < 0.01 and abs(float(y.float().max()) - 2.0) < 1e-3
        assert torch.equal(ops.dropout(big, 0.5, 1234), y) and not torch.equal(ops.dropout(big, 0.5, 1235), y)
        y2 = ops.dropout(big, 0.25, 77, channelwise=True)
        per = y2.float().amax(dim=(1, 2))  # [N,C]: whole planes kept or dropped
        assert torch.equal(per, y2.float().amin(dim=(1, 2)))
        assert 0.55 < (per != 0).float().mean().item() < 0.95


def test_argmax_step_and_kv_append(ops):
    for dtype in DTYPES:
        b, v = 5, 6997
        logits = rnd(b, v, seed=60).to(dtype).contiguous()
        logits[1, 100] = logits[1, 4000] = 50.0  # tie -> first index
        logits[2, 6835] = 60.0  # eos
        tok = torch.zeros(b, dtype=torch.int64, device=dev())
        val = torch.zeros(b, device=dev())
        fin = torch.tensor([0, 0, 0, 1, 0], dtype=torch.int32, device=dev())
        out_t = torch.zeros(b, 4, dtype=torch.int64, device=dev())
        out_v = torch.zeros(b, 4, device=dev())
        ops.argmax_step(logits, tok, val, fin, 6835, 0, out_t, out_v, 2)
        ref = logits.float().argmax(dim=1)
        ref[3] = 0
        assert torch.equal(tok, ref) and int(tok[1]) == 100
        assert fin.tolist() == [0, 0, 1, 1, 0]
        assert torch.equal(out_t[:, 2], tok) and float(out_v[2, 2]) == 60.0
        cache = torch.zeros(3, 6, 8, dtype=dtype, device=dev())
        src = rnd(3, 20, seed=61).to(dtype).contiguous()
        ops.kv_append(src.data_ptr() + 4 * src.element_size(), 20, cache, 5, dtype)
        assert torch.equal(cache[:, 5], src[:, 4:12]) and float(cache[:, :5].abs().sum()) == 0


def test_fused_adam_matches_torch():
    from omr_a2s_multimodal_transformer_b200 import FusedAdam

    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, device=dev())) for s in [(33, 7), (128,), (16, 4, 3, 3)]]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a, b = FusedAdam(ps, lr=1e-3), torch.optim.Adam(qs, lr=1e-3)
    for it in range(5):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p) * (it + 1)
            p.grad, q.grad = g.clone(), g.clone()
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        assert rel_err(p, q) < 1e-6


def test_fused_adam_checkpoint_resume_matches_torch():
    """step, save, load into a FRESH optimizer, step: the bias-correction step and both moments survive the round trip
    (torch.optim.Adam checkpoint layout in both directions), and the kernel binds the loaded moment tensors"""
    from omr_a2s_multimodal_transformer_b200 import FusedAdam

    torch.manual_seed(1)
    shapes = [(33, 7), (128,), (16, 4, 3, 3)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev())) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a, b = FusedAdam(ps, lr=1e-3), torch.optim.Adam(qs, lr=1e-3)
    grads = [[torch.randn(s, device=dev()) * (it + 1) for s in shapes] for it in range(6)]

    def run(opt, params, its):
        for it in its:
            for p, g in zip(params, grads[it]):
                p.grad = g.clone()
            opt.step()

    run(a, ps, range(3))
    run(b, qs, range(3))
    sd = a.state_dict()
    assert all(float(st["step"]) == 3.0 for st in sd["state"].values())
    # resume in a fresh FusedAdam: without the step the first updates would be ~0.3x too small (bias correction at t=1)
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a2 = FusedAdam(ps2, lr=1e-3)
    a2.load_state_dict(sd)
    run(a2, ps2, range(3, 6))
    run(b, qs, range(3, 6))
    for p, q in zip(ps2, qs):
        assert rel_err(p, q) < 1e-6
    # loading into an optimizer whose pointer table already exists must re-bind the moments, and a torch.optim.Adam
    # checkpoint loads as well
    qs3 = [torch.nn.Parameter(q.detach().clone()) for q in qs]
    ps3 = [torch.nn.Parameter(q.detach().clone()) for q in qs]
    a3, b3 = FusedAdam(ps3, lr=1e-3), torch.optim.Adam(qs3, lr=1e-3)
    run(a3, ps3, range(1))  # builds the table with its own (wrong) moments
    for p, q in zip(ps3, qs):
        p.data.copy_(q.data)
    import copy

    # (Optimizer.load_state_dict keeps tensors that already match the parameter's dtype/device: without the deep copies
    # a3 and b3 would share -- and both update -- b's moment buffers)
    a3.load_state_dict(copy.deepcopy(b.state_dict()))
    b3.load_state_dict(copy.deepcopy(b.state_dict()))
    run(a3, ps3, range(2))
    run(b3, qs3, range(2))
    for p, q in zip(ps3, qs3):
        assert rel_err(p, q) < 1e-6
