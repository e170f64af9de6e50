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


def _trainer_worker(rank, world, port, q):
    """The real FlatAdamTrainer.step on two gloo ranks with the kernels stubbed out (tests/dryrun.py): the flat gradient
    is pre-filled with ones and nothing writes to it (the stubs compute nothing, the fused Adam that would clear it is a
    stub too), so after the step every element must be exactly `world` -- reduced once, by the per-layer spans sent from
    inside the backward or by the complement sent after it, never twice and never skipped."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from dryrun import dry_library
    from w2v2_speaker_b200 import dist_utils as du
    from w2v2_speaker_b200.optim.loss import CrossEntropyLoss
    from w2v2_speaker_b200.speaker_module import Wav2vec2FCModule, Wav2vec2FCModuleConfig
    from w2v2_speaker_b200.trainer import FlatAdamTrainer
    assert du.init("gloo")
    results = {}
    with dry_library() as lib:
        for pooling in ("mean", "first+cls"):
            cfg = Wav2vec2FCModuleConfig(stat_pooling_type=pooling, test_stat_pooling_type=pooling, mask_time_prob=0.0,
                                         layerdrop=0.0 if pooling == "mean" else 0.3)
            m = Wav2vec2FCModule(cfg, 50, CrossEntropyLoss).train()
            m.on_train_start()
            tr = FlatAdamTrainer(m, lr=1e-4)
            assert tr.world == world
            for step in range(2):
                tr.flat_g.fill_(1.0)
                lib.calls.clear()
                tr.step(torch.randn(2, 1, 8000), torch.tensor([1, 2]))
                n0 = tr.n0
                enc = tr.flat_g[:n0]
                # the encoder segment is written by stubs only; the head segment also receives autograd's (garbage)
                # gradients before the exchange, so only its all-reduce COUNT can be checked, through the encoder part
                results[(pooling, step)] = (float(enc.min()), float(enc.max()), lib.calls.count("w2v2_adam_step_ex"))
            tr.synchronize()
    q.put((rank, results))
    torch.distributed.destroy_process_group()


def test_two_rank_trainer_reduces_every_gradient_element_once():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_trainer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, results in out:
        assert len(results) == 4
        for key, (lo, hi, adams) in results.items():
            assert lo == float(world) and hi == float(world), (rank, key, lo, hi)
            assert adams == 2, (rank, key)
