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
