"""-m gpu, needs >= 2 CUDA devices (skipped otherwise; run with ``gpurun --gpus 2``): data-parallel equivalence of the
CUDA path over NCCL (ddp.DataParallel + BucketReducer + GraphedTrainStep + FusedAdam), SURVEY.md section 8e:

* after graph-REPLAYED training steps the parameters of the two ranks are bit-identical;
* the all-reduced gradient equals (1e-5, fp32) the gradient ONE process computes on the concatenated batch, with the
  reference's local-mean loss turned into the global-batch mean by ``ddp.token_weight``;
* a step in which only rank 0 drops a modality (``apply_teacher_forcing_modality``, reference model.py:561-575: that
  rank's audio-encoder bucket never becomes ready and is reduced as zeros) neither deadlocks nor desynchronises the ranks;
* the process group is destroyed normally after the captured graphs are released (no hard exit).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model_and_batch(dev):
    import omr_a2s_multimodal_transformer_b200 as pkg
    from oracle import synth

    w2i, i2w = synth.tiny_vocab(97)
    m = pkg.MultimodalTransformer(64, 128, 48, 96, 40, w2i, i2w)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=3))
    m = m.to(dev).eval()  # eval: no dropout, so the 1-process run on the concatenated batch is comparable
    m.set_compute_dtype(torch.float32)
    # 4 samples with different token counts (20+12 on rank 0, 7+15 on rank 1: token_weight matters)
    batch = synth.synth_multimodal_batch(4, (64, 128), (48, 96), [20, 12, 7, 15], w2i, seed=5)
    return m, [t.to(dev) for t in batch]


def _flat_params(m):
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


def _step_fn(m, dp, opt, weighted=True, modality="both"):
    def step(batch):
        xi, xli, xa, xla, y_in, y_out = batch
        dp.zero_grad()
        mem, xl = m._memory(xi, xa, xli, xla, modality)
        loss = m.decoder.loss(tgt=y_in, memory=mem, memory_len=xl, targets=y_out)
        (loss * dp.token_weight(y_out, 0) if weighted else loss).backward()
        dp.sync_gradients()
        opt.step()
        return loss

    return step


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    out = {"rank": rank}
    try:
        import omr_a2s_multimodal_transformer_b200 as pkg

        m, full = _model_and_batch(dev)
        shard = [t[2 * rank: 2 * rank + 2].contiguous() for t in full]
        dp = pkg.DataParallel(m, broadcast=True)
        opt = m.configure_optimizers()
        opt.grad_scale = dp.grad_scale
        step = _step_fn(m, dp, opt)
        step(shard)  # eager step 1 (builds caches / Adam state); its all-reduced gradient is the one compared below
        out["grad_step1"] = (dp.arena.flat.detach() * dp.grad_scale).cpu()
        stepper = pkg.GraphedTrainStep(step, shard, opt, variants=2, warmup=1)  # one more eager step inside
        for _ in range(3):
            stepper(shard)
        torch.cuda.synchronize(dev)
        flat = _flat_params(m)
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        out["identical_after_replay"] = bool(all(torch.equal(gathered[0], g) for g in gathered[1:]))
        out["params"] = flat.cpu()
        # zero-bucket step: rank 0 keeps only the image modality, rank 1 both (eager; the decision is per rank)
        stepper = None  # release the graphs (they hold NCCL kernel nodes) before anything else uses the communicator
        torch.cuda.synchronize(dev)
        step_drop = _step_fn(m, dp, opt, modality="image" if rank == 0 else "both")
        step_drop(shard)
        torch.cuda.synchronize(dev)
        flat = _flat_params(m)
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        out["identical_after_modality_drop"] = bool(all(torch.equal(gathered[0], g) for g in gathered[1:]))
        out["finite"] = bool(torch.isfinite(flat).all())
        torch.cuda.synchronize(dev)
        dist.barrier()
    except Exception as e:  # report instead of hanging the parent
        import traceback

        out["error"] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        q.put(out)
        dist.destroy_process_group()  # must return: a hang here fails the test through the join timeout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices (gpurun --gpus 2)")
def test_two_gpu_nccl_training_matches_one_process_on_the_concatenated_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        r = q.get(timeout=600)
        results[r["rank"]] = r
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, f"rank process did not exit cleanly (exitcode {p.exitcode}): NCCL teardown hung?"
    for r in results.values():
        assert "error" not in r, r["error"]
        assert r["identical_after_replay"] and r["identical_after_modality_drop"] and r["finite"], {k: v for k, v in r.items() if k not in ("params", "grad_step1")}
    assert torch.equal(results[0]["params"], results[1]["params"])
    assert torch.equal(results[0]["grad_step1"], results[1]["grad_step1"])

    # one process, the concatenated batch, plain loss (= mean over ALL non-pad tokens)
    import omr_a2s_multimodal_transformer_b200 as pkg

    dev = torch.device("cuda", 0)
    m, full = _model_and_batch(dev)
    dp = pkg.DataParallel(m, broadcast=False)
    opt = m.configure_optimizers()
    step = _step_fn(m, dp, opt, weighted=False)
    step(full)
    g1 = dp.arena.flat.detach().cpu()
    g2 = results[0]["grad_step1"]
    rel = float((g1.double() - g2.double()).norm() / g1.double().norm())
    assert rel < 1e-5, f"all-reduced gradient differs from the 1-process gradient on the concatenated batch: {rel}"
    for _ in range(4):  # the workers ran 1 eager + 1 warm-up + 3 replayed steps = 5 optimizer steps
        step(full)
    torch.cuda.synchronize(dev)
    p1, p2 = _flat_params(m).cpu().double(), results[0]["params"].double()
    # Adam turns a gradient into lr * g / (|g| + eps): entries whose gradient is mathematically zero (biases in front of an
    # InstanceNorm) move by +-lr per step on rounding noise alone, in either run; hence 1e-3 here, 1e-5 on the gradient
    assert float((p1 - p2).norm() / p1.norm()) < 1e-3
