#!/usr/bin/env python
"""Run the small fp32 model; at every InstanceNorm call compare the in-tree library with a second build."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from tests.helpers import build_multimodal
from omr_a2s_multimodal_transformer_b200 import ops
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
old = ctypes.CDLL(os.path.join(here, "omr_a2s_multimodal_transformer_b200", "libomr_old.so.keep"))
P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
old.omr_instnorm_fwd.argtypes = [I, P, P, P, P, I, I, I, F, P]
old.omr_instnorm_bwd.argtypes = [I, P, P, P, P, P, I, I, I, I, F, P]
DEV = "cuda:0"
KEEP = []
MODE = sys.argv[1] if len(sys.argv) > 1 else "cmp"   # cmp: run both, use new | useold: run both, use old | onlyold
f0, b0 = ops.instnorm_fwd, ops.instnorm_bwd
def fwd(x, eps):
    if MODE != "onlyold":
        y, stats = f0(x, eps)
    n, h, w, c = x.shape
    y2 = torch.empty_like(x); st2 = torch.empty((n, c, 2), dtype=torch.float32, device=x.device); ws = torch.empty((n, c, 2), dtype=torch.float64, device=x.device)
    old.omr_instnorm_fwd(0, x.data_ptr(), y2.data_ptr(), st2.data_ptr(), ws.data_ptr(), n, h * w, c, eps, None)
    torch.cuda.synchronize()
    KEEP.append((tuple(x.shape), y, y2, y.clone(), x, x.clone()))
    if MODE == "newclone":
        return y.clone(), st2
    if MODE == "oldvals_in_new":
        y.copy_(y2)
        return y, st2
    if MODE == "newvals_in_old":
        y2.copy_(y)
        return y2, st2
    if MODE.startswith("mix"):
        return (y if "y=new" in MODE else y2), (stats if "st=new" in MODE else st2)
    if MODE != "cmp":
        return y2, st2
    return y, stats
def bwd(dy, x, stats, relu_mask=False, mask_scale=1.0):
    if MODE != "onlyold":
        dx = b0(dy, x, stats, relu_mask, mask_scale)
    n, h, w, c = x.shape
    dx2 = torch.empty_like(x); ws = torch.empty((n, c, 2), dtype=torch.float64, device=x.device)
    old.omr_instnorm_bwd(0, dy.data_ptr(), x.data_ptr(), stats.data_ptr(), dx2.data_ptr(), ws.data_ptr(), n, h * w, c, int(relu_mask), mask_scale, None)
    torch.cuda.synchronize()
    if MODE.startswith("mix") or MODE in ("newclone", "oldvals_in_new", "newvals_in_old"):
        return dx if "dx=new" in MODE else dx2
    if MODE != "cmp":
        return dx2
    sn, so = dx.double().sum(dim=(1, 2)), dx2.double().sum(dim=(1, 2))
    d = (dx.double() - dx2.double())
    print("bwd", tuple(x.shape), relu_mask, "max|dx-dx2| %.2e  max|sum_new - sum_old| %.2e  max|sum_old| %.2e  mean diff %.2e  frac nonzero diff %.3f" % (
        float(d.abs().max()), float((sn - so).abs().max()), float(so.abs().max()), float(d.mean()), float((d != 0).double().mean())))
    return dx
ops.instnorm_fwd, ops.instnorm_bwd = fwd, bwd
import omr_a2s_multimodal_transformer_b200.encoder as enc
m, sd, w2i = build_multimodal(mixer="concat", dtype=torch.float32)
xi, xli, xa, xla, y_in, y_out = synth.synth_multimodal_batch(3, (64, 128), (48, 96), [20, 12, 7], w2i)
if MODE == "twice":
    DG0 = ops.conv3x3_dgrad
    logs = {}
    for MODE in ("useold", "newvals_in_old"):
        rec = []
        def bwd_rec(dy, x, stats, relu_mask=False, mask_scale=1.0, _b=bwd):
            rec.append(("bwd_dy", tuple(x.shape), dy.clone()))
            return _b(dy, x, stats, relu_mask, mask_scale)
        def fwd_rec(x, eps, _f=fwd):
            rec.append(("fwd_x", tuple(x.shape), x.clone()))
            return _f(x, eps)
        ops.instnorm_fwd, ops.instnorm_bwd = fwd_rec, bwd_rec
        def dgrad_rec(dy, wpt, in_hw, stride=(1, 1), mask=None, mask_scale=1.0, _d=DG0):
            out = _d(dy, wpt, in_hw, stride, mask, mask_scale)
            rec.append(("dgrad_in", tuple(dy.shape) + tuple(stride), dy.clone()))
            if mask is not None:
                rec.append(("dgrad_mask", tuple(mask.shape), (mask > 0).float()))
                rec.append(("dgrad_maskval", tuple(mask.shape), mask.clone()))
            rec.append(("dgrad_out", tuple(out.shape), out.clone()))
            return out
        ops.conv3x3_dgrad = dgrad_rec
        m.zero_grad(set_to_none=True)
        mem, xl = m._memory(xi.to(DEV), xa.to(DEV), xli.to(DEV), xla.to(DEV), "both")
        rec.append(("memory", tuple(mem.shape), mem.detach().clone()))
        loss = m.decoder.loss(tgt=y_in.to(DEV), memory=mem, memory_len=xl, targets=y_out.to(DEV))
        rec.append(("loss", (), loss.detach().clone()))
        loss.backward()
        logs[MODE] = rec
    for (ka, sa, ta), (kb, sb, tb) in zip(logs["useold"], logs["newvals_in_old"]):
        extra = ""
        if ka == "dgrad_mask":
            extra = "  flipped mask entries %d of %d" % (int((ta != tb).sum()), ta.numel())
        if ka == "dgrad_maskval":
            extra = "  entries with 0 < |x| < 1e-5: %d / %d ; exact zeros %d" % (int(((ta.abs() < 1e-5) & (ta != 0)).sum()), int(((tb.abs() < 1e-5) & (tb != 0)).sum()), int((ta == 0).sum()))
        print(ka, sa, "rel diff %.2e" % float((ta.double() - tb.double()).norm() / (ta.double().norm() + 1e-300)), extra)
    sys.exit(0)
m.zero_grad(set_to_none=True)
mem, xl = m._memory(xi.to(DEV), xa.to(DEV), xli.to(DEV), xla.to(DEV), "both")
loss = m.decoder.loss(tgt=y_in.to(DEV), memory=mem, memory_len=xl, targets=y_out.to(DEV))
loss.backward()
from tests.helpers import oracle_truth_and_floors, grad_report
from oracle import restate
truth = oracle_truth_and_floors(lambda s_, dt: restate.multimodal_forward(s_, xi, xli, xa, xla, y_in, mixer_type="concat", dtype=dt), y_out, sd)
print(MODE, grad_report(m, truth["grads"]))
torch.cuda.synchronize()
for shp, y, y2, yc, x, xc in KEEP:
    print(shp, "y vs y2 now %.2e   y vs its own copy %.2e   x vs its copy %.2e   storage_off %d  ptr%%256 %d" % (
        float((y - y2).abs().max()), float((y - yc).abs().max()), float((x - xc).abs().max()), y.storage_offset(), y.data_ptr() % 256))
