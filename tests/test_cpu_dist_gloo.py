"""CPU, world_size 2 over gloo: the multi-GPU host logic (sharding, the single conditioning broadcast, per-global-sample
seeding) gives every rank exactly the rows a single process would have."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, q):
    sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
    from kandinsky2 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    B = 5
    cond = {"image_emb": torch.zeros(2 * B, 16), "pooled": torch.zeros(2 * B, 8, dtype=torch.float16)}
    if rank == 0:
        g = torch.Generator().manual_seed(7)
        cond["image_emb"].copy_(torch.randn(2 * B, 16, generator=g))
        cond["pooled"].copy_(torch.randn(2 * B, 8, generator=g).half())
    parallel.broadcast_conditioning(cond, src=0)
    lo, hi = parallel.shard_range(B, rank, world_size)
    noise = parallel.sample_noise(range(lo, hi), (4, 2, 2), base_seed=99)
    # numpy arrays travel by value (torch tensors travel as file descriptors served by THIS process, which may have
    # exited by the time the parent unpickles them)
    q.put((rank, lo, hi, cond["image_emb"].numpy().copy(), cond["pooled"].numpy().copy(), noise.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_broadcast_world2():
    sys.path.insert(0, os.path.join(ROOT, "kandinsky-2_b200"))
    from kandinsky2 import parallel
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(7)
    emb = torch.randn(10, 16, generator=g)
    pooled = torch.randn(10, 8, generator=g).half()
    ref_noise = parallel.sample_noise(range(5), (4, 2, 2), base_seed=99)
    covered = []
    for rank, lo, hi, e, p, noise in got:
        e, p, noise = torch.from_numpy(e), torch.from_numpy(p), torch.from_numpy(noise)
        assert torch.equal(e, emb) and torch.equal(p, pooled)          # one broadcast delivered everything
        assert torch.equal(noise, ref_noise[lo:hi])                    # per-global-sample seeds: world-size independent
        covered += list(range(lo, hi))
    assert covered == list(range(5))
    assert parallel.shard_range(16, 3, 8) == (6, 8) and parallel.shard_range(5, 0, 2) == (0, 3)
