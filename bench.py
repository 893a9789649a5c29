#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C2|C5|C5long]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Default workload = BASELINE.json config 3 ("C3" in SURVEY.md section 8d): multimodal concat model, full training step
(forward + backward + fused Adam, train mode with dropout) in bf16, batch 32 PER GPU, image 1x128x1024,
audio 1x195x808, target length 512, grandstaff vocabulary (V=6997), data-parallel over N GPUs (weak scaling).
`--config C2` (audio-only A2S, batch 16) and `--config C5` / `C5long` (scaled-up decoder: d_model 512, 8 heads, 8 layers,
max_len 2536; T = 1024 / 2535) run the other training configurations of BASELINE.json through the same protocol.

One JSON line is printed by rank 0; `value` is whole-job train samples/s with inputs resident in HBM, `e2e` the same
through the public call with pinned host buffers (H2D of every step's batch -- on a copy stream, overlapped with the
previous step -- and D2H of the loss inside the timed region).  The line also carries
  * `modality_drop`: the same step timed with the reference's 0.2 teacher-forcing modality drop (model.py:561-575; every
    rank draws independently, the dropped encoder's gradient bucket is all-reduced as zeros),
  * `decode`: the greedy-decode leg (C4: batch 32 per GPU, S=2337, forced full length),
  * `roofline`: the dominant kernel of the step against the measured peaks, `breakdown_ms`,
  * `library_baseline` (stock-PyTorch twin on the same GPU) and `cpu_baseline` (the reference's CPU path, bounded
    sample, rank 0 / N=1 only).

`--impl reference` times the reference's own CPU implementation on the same workload shape with a bounded batch: the
UNMODIFIED reference modules from oracle/_ref (a git-ignored verbatim copy made by oracle/build_ref.py in the build
container; `kind: "reference"`), else the oracle port (oracle/restate.py; `kind: "port"`).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

IMG_HW, AUD_HW, T_LEN, BATCH = (128, 1024), (195, 808), 512, 32
DEC_BATCH, MAX_LEN = 32, 1268
METRIC, UNIT = "train_samples_per_s", "samples/s"

# BASELINE.json training configurations; "kind": mm = image+audio concat model, audio = unimodal A2S
CONFIGS = {
    "C3": dict(kind="mm", t=512, batch=32, d=256, heads=4, layers=8, max_len=1268,
               desc="C3 multimodal concat train step (fwd+bwd+fused Adam, train mode), image 1x128x1024 + audio 1x195x808, T=512, V=6997"),
    "C2": dict(kind="audio", t=512, batch=16, d=256, heads=4, layers=8, max_len=1268,
               desc="C2 audio-only A2S train step (fwd+bwd+fused Adam, train mode), spectrogram 1x195x808, T=512, V=6997"),
    "C5": dict(kind="mm", t=1024, batch=32, d=512, heads=8, layers=8, max_len=2536,
               desc="C5 scaled-up multimodal train step (d_model 512, 8 heads, ff 512, 8 layers, max_len 2536), image 1x128x1024 + audio 1x195x808, T=1024, V=6997"),
    "C5long": dict(kind="mm", t=2535, batch=32, d=512, heads=8, layers=8, max_len=2536,
                   desc="C5 scaled-up multimodal train step (d_model 512, 8 heads, ff 512, 8 layers, max_len 2536), image 1x128x1024 + audio 1x195x808, T=2535, V=6997"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------------------
# synthetic batches: U[0,1) pixels, random lengths, uniform non-special token ids
# ------------------------------------------------------------------------------------------------------------
def make_batch(b, w2i, seed, t_len=T_LEN, img=IMG_HW, aud=AUD_HW):
    g = torch.Generator().manual_seed(seed)
    sos, eos = w2i["<sos>"], w2i["<eos>"]
    xi = torch.rand(b, 1, *img, generator=g)
    xa = torch.rand(b, 1, *aud, generator=g)
    li_max = -(-img[0] // 16) * -(-img[1] // 8)
    la_max = -(-aud[0] // 16) * -(-aud[1] // 8)
    xli = torch.randint(li_max // 2, li_max + 1, (b,), generator=g, dtype=torch.int32)
    xla = torch.randint(la_max // 2, la_max + 1, (b,), generator=g, dtype=torch.int32)
    lens = torch.randint(t_len // 4, t_len + 1, (b,), generator=g)
    lens[0] = t_len
    y = torch.zeros(b, t_len + 1, dtype=torch.int64)
    for i in range(b):
        n = int(lens[i])
        body = torch.randint(1, min(sos, eos), (n - 1,), generator=g)
        y[i, 0] = sos
        y[i, 1:n] = body
        y[i, n] = eos
    return xi, xli, xa, xla, y[:, :-1].contiguous(), y[:, 1:].contiguous()


def config_batch(cfg, b, w2i, seed):
    xi, xli, xa, xla, y_in, y_out = make_batch(b, w2i, seed, t_len=cfg["t"])
    return [xa, xla, y_in, y_out] if cfg["kind"] == "audio" else [xi, xli, xa, xla, y_in, y_out]


def train_flops_per_sample(cfg=None):
    """3 x forward: encoders 16.110 (image) + 19.551 (audio) GFLOP, decoder closed form"""
    cfg = cfg or CONFIGS["C3"]
    t, d, L, v = cfg["t"], cfg["d"], cfg["layers"], 6997
    ff = d
    s = 1313 if cfg["kind"] == "audio" else 1024 + 1313
    enc = 19.551e9 if cfg["kind"] == "audio" else 16.110e9 + 19.551e9
    dec = L * (8 * t * d * d + 4 * t * t * d + 4 * t * d * d + 4 * s * d * d + 4 * t * s * d + 4 * t * d * ff) + 2 * t * d * v
    return 3.0 * (enc + dec)


def build_model(cfg, w2i, i2w):
    import omr_a2s_multimodal_transformer_b200 as pkg

    if cfg["kind"] == "audio":
        return pkg.Transformer(AUD_HW[0], AUD_HW[1], cfg["max_len"], w2i, i2w, teacher_forcing_prob=0.2)
    m = pkg.MultimodalTransformer(IMG_HW[0], IMG_HW[1], AUD_HW[0], AUD_HW[1], cfg["max_len"], w2i, i2w, teacher_forcing_prob=0.2,
                                  teacher_forcing_modality_prob=0.2)
    if cfg["d"] != 256:  # C5: the composition SURVEY.md section 8d describes (last DSC block -> d_model, wider PE and decoder)
        from omr_a2s_multimodal_transformer_b200.decoder import Decoder
        from omr_a2s_multimodal_transformer_b200.encoder import Encoder
        from omr_a2s_multimodal_transformer_b200.model import PositionalEncoding2D

        d = cfg["d"]
        m.image_encoder = Encoder(1, out_channels=d)
        m.audio_encoder = Encoder(1, out_channels=d)
        m.image_pos_2d = PositionalEncoding2D(d, -(-IMG_HW[0] // 16), -(-IMG_HW[1] // 8))
        m.audio_pos_2d = PositionalEncoding2D(d, -(-AUD_HW[0] // 16), -(-AUD_HW[1] // 8))
        m.decoder = Decoder(len(w2i), cfg["max_len"], len(w2i), embedding_dim=d, ff_dim=d, nhead=cfg["heads"],
                            num_transformer_layers=cfg["layers"], padding_idx=m.padding_idx)
    return m


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], 0.0, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own modules (oracle/_ref) or their oracle port, on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(w2i, batch, cfg=None):
    """-> (step() -> loss, kind) for one full training step (forward, CE, backward, Adam) of the reference on the CPU.

    kind "reference": the unmodified reference classes (src/transformer/model.py) through their own ``training_step`` and
    ``configure_optimizers`` (train mode, teacher-forcing noise 0.2, modality drop off like the GPU arm's headline);
    kind "port": oracle/restate.py, when no copy of the reference is available on this machine."""
    from oracle import shim, synth

    cfg = cfg or CONFIGS["C3"]
    i2w = {v: k for k, v in w2i.items()}
    bt = config_batch(cfg, batch, w2i, seed=1)
    if shim.reference_available():
        import contextlib

        ref = shim.load_reference()
        torch.manual_seed(0)
        with contextlib.redirect_stdout(sys.stderr):  # the reference's constructors print their summaries: keep stdout for the JSON line
            if cfg["kind"] == "audio":
                model = ref.Transformer(AUD_HW[0], AUD_HW[1], cfg["max_len"], w2i, i2w, teacher_forcing_prob=0.2)
            else:
                model = ref.MultimodalTransformer(IMG_HW[0], IMG_HW[1], AUD_HW[0], AUD_HW[1], cfg["max_len"], w2i, i2w,
                                                  teacher_forcing_prob=0.2, teacher_forcing_modality_prob=0.0)
            if cfg["d"] != 256:  # the C5 composition of SURVEY.md section 8d, from the reference's own classes
                d = cfg["d"]
                for enc in (model.image_encoder, model.audio_encoder):
                    enc.dscblocks[3] = ref.DSCBlock(128, d, stride=(1, 1))
                model.image_pos_2d = ref.PositionalEncoding2D(d, -(-IMG_HW[0] // 16), -(-IMG_HW[1] // 8))
                model.audio_pos_2d = ref.PositionalEncoding2D(d, -(-AUD_HW[0] // 16), -(-AUD_HW[1] // 8))
                model.decoder = ref.Decoder(output_size=len(w2i), max_seq_len=cfg["max_len"], num_embeddings=len(w2i),
                                            embedding_dim=d, ff_dim=d, nhead=cfg["heads"], num_transformer_layers=cfg["layers"],
                                            padding_idx=model.padding_idx)
        model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
        model.train()
        opt = model.configure_optimizers()
        def step():
            opt.zero_grad()
            loss = model.training_step(tuple(bt), 0)
            loss.backward()
            opt.step()
            return float(loss.detach())

        return step, "reference"
    from oracle import restate

    tmpl = build_model(cfg, w2i, i2w)
    sd = synth.synth_state_dict(tmpl.state_dict(), seed=0)
    del tmpl
    sdg = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and not k.endswith(".pe") else v) for k, v in sd.items()}
    opt = torch.optim.Adam([v for v in sdg.values() if v.requires_grad], lr=1e-4)
    if cfg["d"] != 256:
        raise SystemExit("the oracle port covers the d_model 256 configurations; C5 on the CPU needs oracle/_ref (python -m oracle.build_ref)")

    def step():
        opt.zero_grad()
        if cfg["kind"] == "audio":
            logits = restate.unimodal_forward(sdg, bt[0], bt[1], bt[2])
        else:
            logits = restate.multimodal_forward(sdg, bt[0], bt[1], bt[2], bt[3], bt[4])
        loss = restate.ce_loss(logits, bt[-1])
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, "port"


def _all_host_threads():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm takes all the host threads it may use
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:
        torch.set_num_threads(avail)
    return torch.get_num_threads()


def _ref_what(kind):
    return ("the UNMODIFIED reference modules (oracle/_ref = verbatim copy of the reference's src/, torch CPU fp32, train mode, "
            "its own training_step + Adam)" if kind == "reference" else "oracle port of the reference modules (oracle/restate.py), torch CPU fp32")


def run_reference(args, w2i, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _all_host_threads()
    b = args.cpu_batch
    step, kind = cpu_reference_step_fn(w2i, b, cfg)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = b * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "batch": b, "note": "reference CPU path = " + _ref_what(kind)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} full training steps at batch {b} (same per-sample shapes as the GPU arm)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def load_traffic_table():
    """DRAM bytes per launch of the step's kernels, from the committed `ncu --set full` capture of this round
    (profiles/r02_dram_traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep): {entry point: {shape key: bytes}}"""
    p = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    if os.path.isfile(p):
        try:
            with open(p) as f:
                return json.load(f)
        except Exception:
            return {}
    return {}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS), help="BASELINE.json training configuration")
    ap.add_argument("--batch", type=int, default=None, help="training batch per GPU (default: the configuration's)")
    ap.add_argument("--cpu-batch", type=int, default=4, help="batch of the CPU reference sample (SURVEY.md section 8d: B=4)")
    ap.add_argument("--modality-drop", type=float, default=0.2,
                    help="probability of the teacher-forcing modality drop in the second timed leg (0 = skip the leg)")
    ap.add_argument("--variants", type=int, default=8, help="captured graphs with independent MixDropout draws")
    ap.add_argument("--no-decode", action="store_true", help="skip the greedy-decode leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--decode-steps", type=int, default=MAX_LEN)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-library", action="store_true", help="skip the stock-PyTorch (cuDNN/cuBLAS/SDPA) GPU baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="drive the step eagerly from Python instead of replaying CUDA graphs")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="e2e leg: copy the host batch on the compute stream in front of the replay instead of prefetching the "
                         "NEXT step's batch on a copy stream while the current step runs")
    ap.add_argument("--prefetch", action="store_true", help="(default since round 2; kept for old command lines)")
    args = ap.parse_args()
    from oracle import synth  # vocabulary loader + synthetic weights only (test infrastructure, not on the timed path)

    cfg = CONFIGS[args.config]
    w2i, i2w = synth.load_vocab()
    if args.impl == "reference":
        run_reference(args, w2i, cfg)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import omr_a2s_multimodal_transformer_b200 as pkg
    from omr_a2s_multimodal_transformer_b200 import _lib

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    mm = cfg["kind"] == "mm"
    model = build_model(cfg, w2i, i2w).to(dev)
    model.set_compute_dtype(dtype)
    model.train()
    dp = pkg.DataParallel(model, broadcast=world > 1)
    opt = model.configure_optimizers()
    opt.grad_scale = dp.grad_scale
    b = args.batch or cfg["batch"]
    prefetch = not args.no_prefetch and not args.no_graph
    host = [t.pin_memory() for t in config_batch(cfg, b, w2i, seed=100 + rank)]
    resident = [t.to(dev) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)
    stream = torch.cuda.current_stream(dev)

    def step(batch, modality="both"):
        dp.zero_grad()
        if mm:
            xi, xli, xa, xla, y_in, y_out = batch
            y_in = model.apply_teacher_forcing(y_in)
            mem, xl = model._memory(xi, xa, xli, xla, modality)
        else:
            x, xl, y_in, y_out = batch
            y_in = model.apply_teacher_forcing(y_in)
            mem = model.encode(x)
        loss = model.decoder.loss(tgt=y_in, memory=mem, memory_len=xl, targets=y_out)
        loss.backward()
        dp.sync_gradients()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    hostt = {}

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        hostt["issue_ms"] = (time.perf_counter() - t0) * 1e3 / k  # host time to ENQUEUE one step (no sync inside)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step(resident)
    stepper = None
    launches_per_step = None
    drop_p = args.modality_drop if mm else 0.0
    modes = ["both", "image", "audio"] if drop_p > 0 else ["both"]
    if not args.no_graph:
        # the public fast path for static shapes: the whole step (fwd + bwd + all-reduce + Adam) as replayed CUDA graphs;
        # one set of graphs per teacher-forcing modality (the reference's host-side draw), `variants` MixDropout draws each
        n0 = _lib.launch_count()
        stepper = pkg.GraphedTrainStep(step, resident, opt, variants=args.variants, warmup=1, double_buffer=prefetch, modes=modes,
                                       mode_variants={"image": 2, "audio": 2})
        # launches of ONE "both" step: count an eager one (same code path as the captured step)
        n0 = _lib.launch_count()
        step(resident)
        launches_per_step = _lib.launch_count() - n0
        for _ in range(args.warmup):
            stepper(mode="both")
    run_resident = (lambda: stepper(mode="both")) if stepper is not None else (lambda: step(resident))
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = _lib.launch_count()
    ms = timed(run_resident, args.steps)
    launches = (launches_per_step * args.steps) if stepper is not None else (_lib.launch_count() - n0)
    host_issue_ms = hostt["issue_ms"]
    clocks = sampler.stop() if sampler else None
    value = world * b * args.steps / (ms / 1e3)

    def e2e_step():
        if stepper is not None and prefetch:
            loss = stepper(mode="both")  # replays on the batch prefetched during the previous step ...
            stepper.prefetch(host)       # ... and copies the next one (pinned host -> the other input set) beside it
        elif stepper is not None:
            loss = stepper(host, mode="both")  # pinned host batch -> static device inputs (async H2D), then one graph replay
        else:
            loss = step([t.to(dev, non_blocking=True) for t in host])
        return float(loss.item())  # device -> host read of the step's result

    if stepper is not None and prefetch:
        stepper.prefetch(host)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_val = world * b * args.steps / (ms_e2e / 1e3)

    # second timed leg: the reference's teacher-forcing modality drop (every rank draws on the host, per step)
    moddrop = None
    if drop_p > 0:
        rng = random.Random(1234 + rank)
        draws = {"both": 0, "image": 0, "audio": 0}

        def drop_step():
            m = ("image" if rng.random() < 0.5 else "audio") if rng.random() < drop_p else "both"
            draws[m] += 1
            if stepper is not None:
                stepper(mode=m)
            else:
                step(resident, m)

        if stepper is not None:  # every input set holds the resident batch for this leg
            for s_in in stepper.inputs:
                for s, t in zip(s_in, resident):
                    s.copy_(t)
        k = max(args.steps, 20)
        for m_ in ("image", "audio"):  # warm both rare graphs once
            (stepper(mode=m_) if stepper is not None else step(resident, m_))
        ms_drop = timed(drop_step, k)
        moddrop = {"p": drop_p, "value": world * b * k / (ms_drop / 1e3), "unit": UNIT, "ms_per_step": ms_drop / k, "steps": k,
                   "draws_rank0": dict(draws),
                   "what": "same step with the reference's per-step modality draw (model.py:561-575) made independently on every rank; "
                           "a dropped encoder's gradient bucket is all-reduced as zeros"}

    # per-entry-point device times of ONE more step: CUDA events around every C-ABI call on the launching stream, with the
    # side streams switched off so that every kernel is timed alone (no overlap inflating its duration)
    barrier()
    saved_env = {k: os.environ.get(k) for k in ("OMR_OVERLAP_ENCODERS", "OMR_OVERLAP_DECODER", "OMR_OVERLAP_ENCODER_WGRAD")}
    for k in saved_env:
        os.environ[k] = "0"
    step(resident)
    _lib.prof_start()
    step(resident)
    shaped = _lib.prof_stop(by_shape=True)
    for k, v in saved_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    prof = {}
    for (name, _shape), d in shaped.items():
        a = prof.setdefault(name, {"calls": 0, "launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for k in a:
            a[k] += d[k]
    tot_ms = sum(d["ms"] for d in prof.values()) or 1e-9
    pk = peaks()
    ridge = pk["tc"] * 1e12 / (pk["hbm"] * 1e9)
    traffic_table = load_traffic_table()

    def roof_of(name, shape, d):
        ai = d["flops"] / max(d["bytes"], 1.0)
        if d["flops"] > 0 and ai > ridge * 0.25:
            r = {"bound": "tensor", "achieved": d["flops"] / (d["ms"] * 1e-3) / 1e12, "peak": pk["tc"], "unit": "TFLOP/s"}
        else:
            r = {"bound": "hbm", "achieved": d["bytes"] / (d["ms"] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
        r["frac"] = r["achieved"] / r["peak"]
        key = ",".join(str(int(v)) for v in [int(v) for v in shape if abs(int(v)) < (1 << 20)][-12:])  # = launch_shape below
        tr = (traffic_table.get(name) or {}).get(key)
        r.update({"kernel": name, "launch_shape": [int(v) for v in shape if abs(int(v)) < (1 << 20)][-12:],
                  "calls_per_step": d["calls"], "avg_ms": d["ms"] / max(d["calls"], 1),
                  "algorithmic_per_launch": {"flops": d["flops"] / max(d["calls"], 1), "bytes": d["bytes"] / max(d["calls"], 1)},
                  "share_of_step": d["ms"] / tot_ms,
                  "peak_source": pk["src"] + (" (sustained bf16)" if r["bound"] == "tensor" else ""),
                  "traffic": tr, "traffic_source": "profiles/r02_dram_traffic.json (ncu --set full of this kernel at this shape, this round)" if tr else None})
        return r

    # the dominant kernel = the (entry point, launch shape) group with the largest share of the step
    ranked = sorted(shaped.items(), key=lambda kv: -kv[1]["ms"])
    if os.environ.get("OMR_BENCH_TABLE"):  # every (entry point, launch shape) group of the profiled step, for the work list
        with open(os.environ["OMR_BENCH_TABLE"], "w") as f:
            for (n, sh), d in ranked:
                f.write(f"{n:28s} calls {d['calls']:3d} ms {d['ms']:7.3f} avg_us {1e3 * d['ms'] / max(d['calls'], 1):8.1f} "
                        f"TF/s {d['flops'] / (d['ms'] * 1e-3) / 1e12 if d['ms'] else 0:7.1f} GB/s {d['bytes'] / (d['ms'] * 1e-3) / 1e9 if d['ms'] else 0:7.0f} "
                        f"shape {[int(v) for v in sh if abs(int(v)) < (1 << 20)][-12:]}\n")
    (top_name, top_shape), top = ranked[0]
    roof = roof_of(top_name, top_shape, top)
    roof["timing"] = "CUDA events around each C-ABI call of one eager step, single stream (kernels timed alone)"
    if top_name in ("omr_attn_fwd", "omr_attn_bwd") and mm:
        # the algorithmic figure counts the full Tq x Tk rectangle (as the reference executes it); key tiles whose every key is
        # masked (padded tail of the fused memory) are skipped by the kernel: executed-work utilisation beside it
        # fused memory = [1024 image | 1313 audio] positions in 128-key tiles (8 + 11); live tiles of a sample = those that
        # hold at least one unmasked frame: ceil(xli / 128) + ceil(xla / 128)
        xli, xla = host[1].tolist(), host[3].tolist()
        live = sum(-(-li // 128) + -(-la // 128) for li, la in zip(xli, xla))
        frac_exec = min(1.0, live / (19 * len(xli)))
        roof["executed_fraction_of_rectangle"] = frac_exec if 2337 in top_shape else 1.0
        roof["achieved_on_executed_work"] = roof["achieved"] * roof["executed_fraction_of_rectangle"]
    top5 = [{k: v for k, v in roof_of(n, s, d).items() if k in ("kernel", "launch_shape", "bound", "achieved", "unit", "frac", "avg_ms", "calls_per_step", "share_of_step")}
            for (n, s), d in ranked[:6]]
    breakdown = {k: {"ms": round(d["ms"], 3), "calls": d["calls"],
                     "tflops": round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 2) if d["flops"] else None,
                     "gbs": round(d["bytes"] / (d["ms"] * 1e-3) / 1e9, 1) if d["bytes"] else None}
                 for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]}

    decode = None
    if not args.no_decode and cfg["d"] == 256:
        model.eval()
        torch.cuda.empty_cache()
        db = DEC_BATCH
        xi, _, xa, _, _, _ = make_batch(db, w2i, seed=500 + rank)
        with torch.no_grad():
            if mm:
                mem, _ = model._memory(xi.to(dev), xa.to(dev), None, None, "both")
            else:
                mem = model.encode(xa.to(dev))
            runner = model._decoder_runner()
            nsteps = min(args.decode_steps, cfg["max_len"])
            # warm-up = one whole untimed decode: the batch above is synthesised on the CPU for about a second, the idle GPU
            # drops its clocks, and a 16-step warm-up (a few ms) did not bring them back -- the timed launch then ran up
            # to 25 % slower on some runs (380 vs 310 ms, same code, same box)
            runner.decode(mem, w2i["<sos>"], w2i["<eos>"], 0, max_steps=nsteps, stop_at_eos=False)
            holder = {}

            def dec():
                holder["out"] = runner.decode(mem, w2i["<sos>"], w2i["<eos>"], 0, max_steps=nsteps, stop_at_eos=False)

            ms_dec = timed(dec, 1)
            toks = holder["out"][0]
        s_mem = int(mem.shape[1])
        bytes_tok = 8 * 2 * s_mem * 256 * 2 + 8 * 2 * (nsteps / 2) * 256 * 2 + 11.97e6 / db  # cross-K/V + mean self-KV + weights / batch
        decode = {"metric": "greedy_decode_tokens_per_s", "value": world * db * toks.shape[1] / (ms_dec / 1e3), "unit": "tokens/s",
                  "batch_per_gpu": db, "steps": int(toks.shape[1]), "memory_len": s_mem, "ms": ms_dec,
                  "includes": "cross-K/V projection of the memory + ONE launch of the persistent decode kernel (all steps)",
                  "algorithmic_bytes_per_token": bytes_tok,
                  "hbm_roofline_tokens_per_s_per_gpu": pk["hbm"] * 1e9 / bytes_tok}
        decode["frac_of_hbm_roofline"] = decode["value"] / world / decode["hbm_roofline_tokens_per_s_per_gpu"]
        decode["warmup"] = "one whole untimed decode of the same length"
        decode["bound"] = ("what one SM takes in from L2 (the attention phases take the same time at batch 8, 16 and 32; "
                           "DESIGN.md section 4, profiles/r02_decode_ncu_full.txt), not HBM")
        model.train()

    library = None
    graphed = stepper is not None
    ngraphs = stepper.num_graphs() if stepper is not None else 0
    if stepper is not None:
        stepper.release()  # the captured graphs hold NCCL kernels and the step's memory pool
        stepper = None
    torch.cuda.synchronize(dev)
    if rank == 0 and world == 1 and not args.no_library and args.config == "C3":
        # "library kernels to beat": the same architecture from stock torch.nn modules (oracle/torch_twin.py), bf16
        # autocast, same batch, forward + backward + fused torch Adam, on this GPU
        try:
            from oracle.torch_twin import TwinMultimodal

            torch.cuda.empty_cache()
            twin = TwinMultimodal(len(w2i), MAX_LEN, IMG_HW, AUD_HW).to(dev).train()
            topt = torch.optim.Adam(twin.parameters(), lr=1e-4, fused=True)
            xi, xli, xa, xla, y_in, y_out = resident

            def tstep():
                topt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    logits = twin(xi, xli, xa, xla, y_in)
                loss = torch.nn.functional.cross_entropy(logits.float(), y_out, ignore_index=0)
                loss.backward()
                topt.step()

            for _ in range(3):
                tstep()
            ms_lib = timed(tstep, args.steps)
            library = {"value": b * args.steps / (ms_lib / 1e3), "unit": UNIT, "ms_per_step": ms_lib / args.steps,
                       "what": "stock torch.nn twin of the reference (cuDNN conv / InstanceNorm, nn.TransformerDecoder SDPA, cuBLAS, "
                               "CrossEntropyLoss, fused Adam), bf16 autocast, same batch on this GPU, eager",
                       "torch": torch.__version__}
            del twin, topt
            torch.cuda.empty_cache()
        except Exception as e:  # a reported baseline must never take the benchmark down
            library = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cores = _all_host_threads()
            cstep, kind = cpu_reference_step_fn(w2i, args.cpu_batch, cfg)
            cstep()
            t0 = time.perf_counter()
            n = 0
            while n < 2 or (time.perf_counter() - t0 < 12.0 and n < 8):
                cstep()
                n += 1
            dtc = time.perf_counter() - t0
            cpu = {"value": args.cpu_batch * n / dtc, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{n} full training steps at batch {args.cpu_batch} of the same per-sample shapes: {_ref_what(kind)}"}
        except Exception as e:
            cpu = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        flops_step = train_flops_per_sample(cfg) * b
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": cfg["desc"], "name": args.config,
                       "batch_per_gpu": b, "global_batch": b * world, "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations, several GB) is far larger than the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "input_copy": ("every step's pinned host batch is copied on a copy stream while the previous step runs (double-buffered "
                                   "graph inputs)") if prefetch else "on the compute stream in front of the step"},
            "gpu_launches": int(launches),
            "step_driver": (f"cuda-graph replay ({ngraphs} captured graphs: {args.variants} MixDropout draws x input sets per modality mode, "
                            "picked at random per step; device-side dropout seeds)") if graphed else "eager python",
            "host_issue_ms_per_step": host_issue_ms,
            "clocks": clocks,
            "model_tflops_per_gpu": flops_step / (ms / args.steps * 1e-3) / 1e12,
            "model_tc_frac_of_sustained_peak": flops_step / (ms / args.steps * 1e-3) / 1e12 / pk["tc"],
            "roofline": roof,
            "top_kernels": top5,
            "breakdown_ms": breakdown,
            "breakdown_note": "single-stream eager step: per-entry-point times are un-overlapped and add up to more than the replayed step",
            "modality_drop": moddrop,
            "decode": decode,
            "library_baseline": library,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear NCCL down normally: the captured graphs (which hold the communicator's kernels) were released above.  A
        # watchdog still guarantees that the process ends if ncclCommDestroy were to wait (seen in round 1 WITH live graphs).
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        wd = threading.Timer(60.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.destroy_process_group()
        wd.cancel()


if __name__ == "__main__":
    main()
