"""world_size-2 gloo test of the N>1 path's only communication: scatter of the request batch from
rank 0 before the loop, gather of the decoded images after it (no collective inside the loop)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from powerpaint_b200.parallel import gather_images, scatter_requests, shard_ranges

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = 3
    full_img = torch.arange(world * b * 3 * 4 * 4, dtype=torch.float32).reshape(world * b, 3, 4, 4)
    full_emb = torch.arange(world * b * 5 * 2, dtype=torch.float32).reshape(world * b, 5, 2)
    tensors = [full_img, full_emb] if rank == 0 else None
    img, emb = scatter_requests(tensors, [(b, 3, 4, 4), (b, 5, 2)], [torch.float32, torch.float32], "cpu")
    s, e = shard_ranges(world * b, world)[rank]
    ok = torch.equal(img, full_img[s:e]) and torch.equal(emb, full_emb[s:e])
    out = gather_images((img * 2).to(torch.uint8))
    if rank == 0:
        ok = ok and torch.equal(out, (full_img * 2).to(torch.uint8))
    else:
        ok = ok and out is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
