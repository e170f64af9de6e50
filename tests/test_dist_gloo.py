"""world_size-2 gloo test of the multi-process plumbing used by bench.py (N > 1 path)."""
import os

import torch
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from w2v2_speaker_b200 import dist_utils as du
    from oracle.params import make_inputs
    assert du.init("gloo")
    assert du.env() == (rank, rank, world)
    wav, labels = make_inputs(2, 400, seed=du.shard_seed(1234))
    # ranks must see different utterances (independent shards) ...
    sig = float(wav.abs().sum())
    gathered = [None] * world
    torch.distributed.all_gather_object(gathered, sig)
    # ... and the reported time is the max over ranks
    t = du.max_over_ranks(10.0 + rank)
    du.barrier()
    # the trainer's gradient exchange: per-layer spans sent deepest layer first, then the complement --
    # together exactly one sum over the whole flat buffer
    n = 1000
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    sent = []
    for lo, hi in [(700, 1000), (400, 700), (100, 400)]:
        torch.distributed.all_reduce(flat[lo:hi])
        sent.append((lo, hi))
    rest = du.remaining_spans(sent, n)
    assert rest == [(0, 100)]
    for lo, hi in rest:
        torch.distributed.all_reduce(flat[lo:hi])
    assert torch.equal(flat, torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world)))
    assert du.remaining_spans([], 5) == [(0, 5)] and du.remaining_spans([(0, 5)], 5) == []
    assert du.remaining_spans([(3, 4), (1, 2)], 6) == [(0, 1), (2, 3), (4, 6)]
    q.put((rank, gathered, t, du.global_batch(64)))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_max_time():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, t, gb in out:
        assert gathered[0] != gathered[1]
        assert t == 11.0
        assert gb == 128
