import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '.')
import torch
from omr_a2s_multimodal_transformer_b200 import ops
dev = torch.device("cuda", 0)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (n, h, w, c) in ((32, 195, 808, 32), (32, 128, 1024, 16), (32, 98, 404, 64), (32, 49, 202, 128)):
    x = torch.randn(n, h, w, c, device=dev).bfloat16(); dy = torch.randn(n, h, w, c, device=dev).bfloat16()
    sums = torch.zeros(n, c, 2, dtype=torch.float64, device=dev)
    sums[:, :, 0] = x.double().sum((1, 2)); sums[:, :, 1] = (x.double() ** 2).sum((1, 2))
    y, stats = ops.instnorm_fwd(x, 1e-5, sums=sums)
    bs = torch.zeros(n, c, 2, dtype=torch.float64, device=dev)
    bs[:, :, 0] = dy.double().sum((1, 2)); bs[:, :, 1] = (dy.double() * x.double()).sum((1, 2))
    cs = torch.zeros(c, device=dev)
    t0 = timed(lambda: ops.instnorm_bwd(dy, x, stats, relu_mask=True, mask_scale=2.0, sums=bs))
    t1 = timed(lambda: ops.instnorm_bwd(dy, x, stats, relu_mask=True, mask_scale=2.0, sums=bs, colsum=cs))
    print(f"{n}x{h}x{w}x{c}: IN bwd {t0:6.1f} us, with colsum {t1:6.1f} us")
