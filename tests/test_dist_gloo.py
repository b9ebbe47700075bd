"""world_size-2 gloo run (CPU) of the N>1 host logic: global-noise slicing + the single all-gather
give exactly the 1-rank result, including a ragged split (B=5 over 2 ranks)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio2photoreal_b200.dist import global_noise, global_noise_tape, sample_sharded


def _fake_loop(shape, noise, y):
    # stands in for sampler.ddim_sample_loop: a row-wise deterministic function of (noise row, conditioning row)
    return noise * 2.0 + y["audio_embed"].view(shape[0], 1, 1, 1)


def _fake_stochastic_loop(shape, noise, y, lo, hi, tape):
    # stands in for p_sample_loop(noise_tape=tape): every step adds ITS rows of the global per-step noise
    assert tape.shape[1] == hi - lo == shape[0]
    return noise + tape.sum(0) + float(lo) * 0 + y["audio_embed"].view(shape[0], 1, 1, 1)


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y = {"audio_embed": torch.arange(B, dtype=torch.float32), "tag": "shared"}
    out = sample_sharded(_fake_loop, (B, 3, 1, 7), y, seed=10, device="cpu")
    out2 = sample_sharded(_fake_stochastic_loop, (B, 3, 1, 7), y, seed=10, device="cpu", noise_tape_steps=4)
    if rank == 0:
        q.put((out, out2))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_sharding_equals_single_rank():
    for B in (4, 5):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
        for p in procs:
            p.start()
        got, got2 = q.get(timeout=120)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        y = {"audio_embed": torch.arange(B, dtype=torch.float32)}
        ref = _fake_loop((B, 3, 1, 7), global_noise((B, 3, 1, 7), 10, "cpu"), y)
        assert torch.equal(got, ref)
        # stochastic samplers: the per-step noise is the GLOBAL tape sliced by rows (not one tape per rank)
        ref2 = _fake_stochastic_loop((B, 3, 1, 7), global_noise((B, 3, 1, 7), 10, "cpu"), y, 0, B,
                                     global_noise_tape(4, (B, 3, 1, 7), 10, "cpu"))
        assert torch.equal(got2, ref2)
