#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_fused_links.py tests/test_gpu_tc.py -m gpu -q -x -k "fused or proj_ce or layernorm" 2>&1 | tail -4
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from omr_a2s_multimodal_transformer_b200 import ops
dev='cuda:0'
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
M, d, V = 16384, 256, 6997
x = torch.randn(M, d, device=dev).bfloat16(); res = torch.randn(M, d, device=dev).bfloat16()
gamma = torch.rand(d, device=dev) + 0.5; beta = torch.randn(d, device=dev)
y, s, st = ops.dropout_add_layernorm_fwd(x, res, gamma, beta, 1e-5, True, 0.1, 7)
dg, db, dbias = (torch.zeros(d, device=dev) for _ in range(3))
dy = torch.randn(M, d, device=dev).bfloat16()
print("ln fwd us", timed(lambda: ops.dropout_add_layernorm_fwd(x, res, gamma, beta, 1e-5, True, 0.1, 7)))
print("ln bwd us", timed(lambda: ops.layernorm_bwd_dropout(dy, s, st, gamma, dg, db, 0.1, 7, dbias=dbias)))
w = (torch.randn(V, d, device=dev) * 0.05).bfloat16(); bias = torch.randn(V, device=dev) * 0.1
tg = torch.randint(1, V, (M,), device=dev)
lo, lse = ops.proj_ce_fwd(x, w, bias, tg, 0)
g = torch.ones(1, device=dev); dw = torch.zeros(V, d, device=dev); dbv = torch.zeros(V, device=dev)
print("proj_ce fwd us", timed(lambda: ops.proj_ce_fwd(x, w, bias, tg, 0), 20))
print("proj_ce dx us", timed(lambda: ops.proj_ce_bwd_dx(x, w, bias, tg, lse, lo, g, 0), 20))
print("proj_ce dw us", timed(lambda: ops.proj_ce_bwd_dw(x, w, bias, tg, lse, lo, g, 0, dw, dbv), 20))
PY
