"""World-size-2 Gloo (CPU) test of the N>1 host logic: weight broadcast from rank 0, batch sharding, MAX time reduction,
detection all-gather — the same code bench.py runs over NCCL on the GPU box."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, init_file, out_file):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from yolo_master_b200 import parallel as P
    from yolo_master_b200.nn.modules import Conv
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
    m = Conv(16, 32, 3, 1)
    with torch.no_grad():
        m.bn.running_mean.uniform_(-1, 1)
    before = m.conv.weight.clone()
    n = P.broadcast_module_state(m, src=0)
    gathered = [torch.empty_like(m.conv.weight) for _ in range(world)]
    dist.all_gather(gathered, m.conv.weight.data)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    changed = not torch.equal(before, m.conv.weight) if rank != 0 else torch.equal(before, m.conv.weight)
    lo, hi = P.shard_range(7, rank, world)
    tmax = P.max_over_ranks(10.0 + rank, "cpu")
    dets = P.gather_detections(torch.full((2, 3, 6), float(rank)))
    ok = same and changed and n == len(list(m.parameters())) + len(list(m.buffers())) and tmax == 10.0 + world - 1 \
        and (lo, hi) == ((0, 4) if rank == 0 else (4, 7)) and dets.shape == (4, 3, 6) and dets[2:].eq(1).all() and dets[:2].eq(0).all()
    if rank == 0:
        t = torch.tensor([1.0 if ok else 0.0])
    else:
        t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out_file, "w").write(str(float(t)))
    dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    with tempfile.TemporaryDirectory() as d:
        init, out = os.path.join(d, "init"), os.path.join(d, "out")
        mp.spawn(_worker, args=(2, init, out), nprocs=2, join=True)
        assert open(out).read() == "1.0"


def test_shard_range_covers_batch():
    from yolo_master_b200.parallel import shard_range
    for gb in (1, 7, 32, 128):
        for w in (1, 2, 4, 8):
            spans = [shard_range(gb, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
