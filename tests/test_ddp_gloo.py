"""CPU (gloo, world_size 2): the ordered, overlapped bucket all-reduce behind the data-parallel trainer
(omr_a2s_multimodal_transformer_b200/ddp.py).  The ranks mark their buckets ready in DIFFERENT orders and
one rank never marks one of them (an encoder whose modality was dropped receives no gradient, reference
model.py:561-575): the collectives must still be issued in the same order everywhere, the sums must be exact
and nothing may deadlock."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from omr_a2s_multimodal_transformer_b200.ddp import BucketReducer

        torch.manual_seed(rank)
        sizes = [1000, 37, 512]
        ok = True
        for step in range(3):
            buckets = [torch.full((n,), float(rank + 1 + 10 * i + step)) for i, n in enumerate(sizes)]
            if rank == 1 and step == 1:
                buckets[1].zero_()  # this rank's "audio encoder" got no gradient this step
            red = BucketReducer(buckets)
            order = [0, 1, 2] if rank == 0 else [2, 0, 1]
            for i in order:
                if rank == 1 and step == 1 and i == 1:
                    continue  # never becomes ready: reduced as zeros at finish()
                red.mark_ready(i)
            red.finish()
            for i, b in enumerate(buckets):
                want = sum((r + 1 + 10 * i + step) for r in range(world))
                if step == 1 and i == 1:
                    want -= (1 + 1 + 10 * i + step)
                ok = ok and bool(torch.all(b == want))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_bucket_reducer_orders_collectives_and_sums():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, "worker crashed or timed out"
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(0, True), (1, True)]


def test_grad_arena_layout_and_buckets():
    """all gradients are views of one flat buffer, buckets are contiguous slices in backward-completion order"""
    import omr_a2s_multimodal_transformer_b200 as pkg
    from oracle import synth

    w2i, i2w = synth.tiny_vocab(31)
    m = pkg.MultimodalTransformer(32, 64, 32, 64, 12, w2i, i2w)
    dp = pkg.DataParallel(m, broadcast=False)
    assert dp.world == 1 and dp.reducer is None and dp.grad_scale == 1.0
    n_dec = sum(p.numel() for p in m.decoder.parameters())
    assert dp.buckets[0].numel() >= n_dec and len(dp.buckets) == 3
    assert sum(b.numel() for b in dp.buckets) == dp.arena.flat.numel()
    for p in m.parameters():
        assert p.grad is not None and p.grad.untyped_storage().data_ptr() == dp.arena.flat.untyped_storage().data_ptr()
    dp.arena.flat.fill_(1.0)
    dp.zero_grad()
    assert float(dp.arena.flat.abs().sum()) == 0.0


def _tw_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from omr_a2s_multimodal_transformer_b200.ddp import token_weight

        # rank 0 holds 5 real targets, rank 1 holds 2 (0 = PAD); per-token losses are known numbers
        y = torch.tensor([[3, 4, 5, 0], [6, 7, 0, 0]]) if rank == 0 else torch.tensor([[9, 0, 0, 0], [8, 0, 0, 0]])
        tok_loss = torch.arange(8, dtype=torch.float32).reshape(2, 4) + 10 * rank
        w = torch.ones(1, requires_grad=True)
        local = (tok_loss * w)[y != 0].mean()  # what CrossEntropyLoss(ignore_index=0) returns on this rank
        (local * token_weight(y, 0)).backward()
        g = w.grad.clone()
        dist.all_reduce(g)
        g /= world  # gradient mean over ranks (DataParallel folds 1/world into Adam)
        q.put((rank, float(g), float(token_weight(y, 0))))
    finally:
        dist.destroy_process_group()


def test_token_weight_reproduces_the_global_batch_mean():
    """loss_r * n_r * world / sum(n) averaged over ranks == the mean over ALL non-pad targets of the concatenated batch"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tw_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, "worker crashed or timed out"
    res = sorted(q.get(timeout=5) for _ in range(world))
    vals0 = [0.0, 1.0, 2.0, 4.0, 5.0]  # rank 0: positions with y != 0
    vals1 = [10.0, 14.0]               # rank 1
    want = sum(vals0 + vals1) / 7
    for rank, g, s in res:
        assert abs(g - want) < 1e-5, (rank, g, want)
    assert abs(res[0][2] - 5 * 2 / 7) < 1e-6 and abs(res[1][2] - 2 * 2 / 7) < 1e-6


def test_token_weight_is_one_without_a_process_group():
    from omr_a2s_multimodal_transformer_b200.ddp import token_weight

    assert float(token_weight(torch.tensor([[1, 2, 0]]), 0)) == 1.0


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import omr_a2s_multimodal_transformer_b200 as pkg
        from oracle import synth

        w2i, i2w = synth.tiny_vocab(31)
        torch.manual_seed(100 + rank)  # ranks start from DIFFERENT weights: the wrapper must broadcast rank 0's
        m = pkg.MultimodalTransformer(32, 64, 32, 64, 12, w2i, i2w)
        dp = pkg.DataParallel(m, broadcast=True)
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same_weights = all(torch.equal(gathered[0], g) for g in gathered)
        ok = True
        for step in range(2):
            dp.zero_grad()
            # what the backward kernels do: write gradients straight into the arena views
            for i, p in enumerate(m.parameters()):
                p.grad.fill_(float(rank + 1) * (1 + (i % 3)) + step)
            skip_audio = rank == 1 and step == 1  # modality dropped on this rank: that encoder never reports
            if skip_audio:
                for p in m.audio_encoder.parameters():
                    p.grad.zero_()
            # completion order of the real backward: decoder first, then the encoders (either order)
            m.decoder._bwd_done_cb()
            order = [m.image_encoder, m.audio_encoder] if rank == 0 else [m.audio_encoder, m.image_encoder]
            for enc in order:
                if skip_audio and enc is m.audio_encoder:
                    continue
                enc._bwd_done_cb()
            dp.sync_gradients()
            audio = {id(p) for p in m.audio_encoder.parameters()}
            for i, p in enumerate(m.parameters()):
                want = sum(float(r + 1) * (1 + (i % 3)) + step for r in range(world))
                if step == 1 and id(p) in audio:
                    want -= float(1 + 1) * (1 + (i % 3)) + step
                ok = ok and bool(torch.all(p.grad == want))
        q.put((rank, same_weights, ok, dp.grad_scale))
    finally:
        dist.destroy_process_group()


def test_data_parallel_wrapper_broadcasts_and_reduces_arena_buckets():
    """DataParallel end to end on the host side: parameter broadcast, gradient arena, bucket callbacks in rank-dependent
    order, a rank whose audio encoder got no gradient, SUM all-reduce with the mean left to the optimizer (grad_scale)"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, "worker crashed or timed out"
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(0, True, True, 0.5), (1, True, True, 0.5)]
