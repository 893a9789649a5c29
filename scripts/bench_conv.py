#!/usr/bin/env python
"""Time the encoder's convolution entry points (forward, data gradient, weight gradient) and its element-wise tail
(InstanceNorm fwd/bwd, dropout) on the layer shapes of BASELINE config 3 (bf16, batch 32).  CUDA events, 3 warm-up + 5
timed calls per shape; every activation is larger than or comparable to the L2 and each call streams fresh data.

    python scripts/bench_conv.py [fwd] [dgrad] [wgrad] [tail]        (default: all)
    OMR_CONV_DEBUG=1|2|4 python scripts/bench_conv.py fwd            (diagnostic switches of the halo kernel)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from omr_a2s_multimodal_transformer_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
B = 32
# (H, W, Ci, Co, sh, sw) of the input: image encoder then audio encoder (C_in = 1 first layers excluded)
IMG = [(128, 1024, 16, 16, 1, 1), (128, 1024, 16, 32, 1, 1), (128, 1024, 32, 32, 1, 1), (128, 1024, 32, 32, 2, 2),
       (64, 512, 32, 64, 1, 1), (64, 512, 64, 64, 1, 1), (64, 512, 64, 64, 2, 2), (32, 256, 64, 128, 1, 1),
       (32, 256, 128, 128, 1, 1), (32, 256, 128, 128, 2, 2), (16, 128, 128, 128, 1, 1), (16, 128, 128, 128, 2, 1)]
AUD = [(195, 808, 16, 16, 1, 1), (195, 808, 16, 32, 1, 1), (195, 808, 32, 32, 1, 1), (195, 808, 32, 32, 2, 2),
       (98, 404, 32, 64, 1, 1), (98, 404, 64, 64, 1, 1), (98, 404, 64, 64, 2, 2), (49, 202, 64, 128, 1, 1),
       (49, 202, 128, 128, 1, 1), (49, 202, 128, 128, 2, 2), (25, 101, 128, 128, 1, 1), (25, 101, 128, 128, 2, 1)]
# calls per training step of each shape (conv2+conv3 share 16->16 in block 0, conv1 of block 4 + conv2 share 128->128 ...)
MULT = {(16, 16, 1): 2, (128, 128, 1): 2}
PEAK_HBM = 6455.9


def timeit(fn, n=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    which = set(sys.argv[1:]) or {"fwd", "dgrad", "wgrad", "tail"}
    shapes = IMG + AUD
    if os.environ.get("BENCH_CONV_FEW"):
        shapes = [IMG[0], IMG[2], IMG[3], IMG[5], AUD[0], AUD[2]]
    tot = {k: 0.0 for k in ("fwd", "dgrad", "wgrad", "in_fwd", "in_bwd", "drop")}
    for (h, w, ci, co, sh, sw) in shapes:
        ho, wo = -(-h // sh), -(-w // sw)
        x = torch.rand(B, h, w, ci, device=dev).to(torch.bfloat16)
        dy = torch.randn(B, ho, wo, co, device=dev).to(torch.bfloat16)
        wt = (torch.randn(co, ci, 3, 3, device=dev) * 0.05)
        wp = ops.pack_conv_weight(wt, torch.bfloat16, False)
        wpt = ops.pack_conv_weight(wt, torch.bfloat16, True)
        bias = torch.zeros(co, device=dev)
        dw, db = torch.zeros(co, ci, 3, 3, device=dev), torch.zeros(co, device=dev)
        mult = MULT.get((ci, co, sh), 1)
        fl = 2.0 * B * ho * wo * 9 * ci * co
        by = 2.0 * B * (h * w * ci + ho * wo * co)
        line = f"{h:3d}x{w:<4d} {ci:3d}->{co:<3d} s{sh}{sw} x{mult} floor {by / PEAK_HBM / 1e3:6.1f} us |"
        if "fwd" in which:
            ms = timeit(lambda: ops.conv3x3_fwd(x, wp, bias, (sh, sw), True))
            tot["fwd"] += ms * mult
            line += f" fwd {ms * 1e3:7.1f} us {fl / ms / 1e9:6.1f} TF/s {by / ms / 1e6:5.0f} GB/s |"
        if "dgrad" in which:
            ms = timeit(lambda: ops.conv3x3_dgrad(dy, wpt, (h, w), (sh, sw), mask=x, mask_scale=2.0))
            tot["dgrad"] += ms * mult
            line += f" dgrad {ms * 1e3:7.1f} us {fl / ms / 1e9:6.1f} TF/s |"
        if "wgrad" in which:
            ms = timeit(lambda: ops.conv3x3_wgrad(x, dy, dw, db, (sh, sw), True))
            tot["wgrad"] += ms * mult
            line += f" wgrad {ms * 1e3:7.1f} us {fl / ms / 1e9:6.1f} TF/s |"
        print(line, flush=True)
        if "tail" in which and sh == 1 and ci == co:  # the conv2 output of a block: IN fwd/bwd and (sometimes) dropout
            ms1 = timeit(lambda: ops.instnorm_fwd(x, 1e-3))
            y, st = ops.instnorm_fwd(x, 1e-3)
            dyy = torch.randn_like(x)
            ms2 = timeit(lambda: ops.instnorm_bwd(dyy, x, st, relu_mask=True, mask_scale=1.0))
            ms3 = timeit(lambda: ops.dropout(x, 0.5, 1234))
            n = x.numel() * 2
            tot["in_fwd"] += ms1
            tot["in_bwd"] += ms2
            tot["drop"] += ms3
            print(f"      tail on [{B},{h},{w},{ci}] ({n / 1e6:.0f} MB): IN fwd {ms1 * 1e3:6.1f} us ({3 * n / ms1 / 1e6:5.0f} GB/s) "
                  f"IN bwd {ms2 * 1e3:6.1f} us ({5 * n / ms2 / 1e6:5.0f} GB/s) dropout {ms3 * 1e3:6.1f} us ({2 * n / ms3 / 1e6:5.0f} GB/s)", flush=True)
        del x, dy
    print("per-step totals (ms): " + "  ".join(f"{k} {v:.3f}" for k, v in tot.items() if v))


if __name__ == "__main__":
    main()
