"""Replica mode host logic on CPU: world_size-2 gloo processes (no GPU needed)."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qwen_megakernel.replicas import assign, combine, throughput


def test_assign_is_a_partition():
    utts = list(range(23))
    for world in (1, 2, 4, 8):
        parts = [assign(utts, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == utts
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_combine_without_process_group_is_identity():
    assert combine(200, 350.0) == (200.0, 350.0)
    assert abs(throughput(200, 400.0) - 500.0) < 1e-9


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        utts = list(range(11))
        mine = assign(utts, world, rank)
        units, ms = combine(len(mine) * 10, 100.0 + 50.0 * rank)      # rank 1 is the slow one
        dist.barrier()
        import bench                                                   # the bench's own helpers on the same process group
        per_rank = bench.gather_floats(100.0 + 50.0 * rank, world, torch.device("cpu"))
        cores = bench.pin_rank_to_cores(rank, world)
        out.put((rank, mine, units, ms, throughput(len(mine) * 10, 100.0 + 50.0 * rank), per_rank, cores))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_aggregate():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, u0, t0, f0, pr0, c0), (r1, m1, u1, t1, f1, pr1, c1) = res
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1)
    assert u0 == u1 == 110.0 and t0 == t1 == 150.0          # sum of units, max of time
    assert abs(f0 - 110.0 / 0.150) < 1e-6 and f0 == f1
    assert pr0 == pr1 == [100.0, 150.0]                      # every rank's own time travels with the line (bench.per_rank_ms)
    assert c0 and c1 and (len(c0) == 1 or not set(c0) & set(c1))   # ranks are pinned to disjoint host cores


def test_bench_median_and_reference_config_match():
    """bench.py helpers that need no GPU: the median over repetitions, and the two arms print IDENTICAL config dicts."""
    import bench
    assert bench._median([3.0, 1.0, 2.0]) == 2.0 and bench._median([4.0, 1.0, 2.0, 3.0]) == 2.5
    assert bench.talker_bytes(0) == 887_228_928 + 2 * 114_688 and bench.cp_frame_bytes() == 2_583_052_288
    src = open(bench.__file__).read()
    assert src.count('"frames_per_utterance": K') == 1 and "config[\"" not in src.split("def main()")[1], \
        "config must be built once and never edited per arm (the driver compares the two arms' dicts)"
