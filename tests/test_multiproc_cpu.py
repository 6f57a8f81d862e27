"""N>1 host logic on CPU (gloo, world_size 2): streams are independent replicas, the only collective is the
broadcast of the shared prompt embedding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from live2diff_b200.schedule import RingSchedule, stream_constants
    from live2diff_b200.stream_pipeline import broadcast_prompt

    prompt = torch.arange(77 * 96, dtype=torch.float32).reshape(1, 77, 96) / 100 if rank == 0 else None
    got = broadcast_prompt(prompt, (1, 77, 96), torch.device("cpu"), src=0)
    # every replica advances its own ring schedule independently (different frame counts per stream)
    rs = RingSchedule(2, 16, 8)
    for _ in range(5 + 7 * rank):
        rs.advance()
    # max-over-ranks reduction used by bench.py for the timed region
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = (got.float().sum().item(), rs.update_idx, float(t.item()), stream_constants([30, 40]).timesteps)
    dist.barrier()
    dist.destroy_process_group()


def test_prompt_broadcast_and_independent_streams():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ref = (torch.arange(77 * 96, dtype=torch.float32) / 100).half().float().sum().item()
    assert abs(out[0][0] - ref) < 1e-3 and out[0][0] == out[1][0]          # both ranks hold rank 0's embedding
    assert out[0][1] != out[1][1]                                           # streams are at different frames
    assert out[0][2] == out[1][2] == 11.0                                   # max over ranks
    assert out[0][3] == out[1][3] == [399, 199]
