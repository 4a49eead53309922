"""N>1 host logic on CPU: world_size-2 gloo run of the sharding + max-over-ranks aggregation that
bench.py uses under torchrun, and the --impl reference rule (rank 0 works, other ranks exit 0)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
    from imagekit_cuda.sharding import aggregate_throughput, shard_indices
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(1024, world, rank)
    local_units = float(len(mine)) * 0.12       # 0.12 output MP per thumbnail (cfg3)
    local_ms = 10.0 + 5.0 * rank                # rank 1 is slower: the job time is the max
    total, ms, rate = aggregate_throughput(local_units, local_ms, dist)
    dist.barrier()
    q.put((rank, mine[:3], len(mine), total, ms, rate))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, first0, n0, total0, ms0, rate0), (r1, first1, n1, total1, ms1, rate1) = res
    assert first0 == [0, 2, 4] and first1 == [1, 3, 5] and n0 + n1 == 1024
    assert total0 == total1 == pytest.approx(1024 * 0.12)
    assert ms0 == ms1 == 15.0                    # max over ranks, not rank 0's own 10 ms
    assert rate0 == pytest.approx(1024 * 0.12 / 0.015)


def test_shards_partition_the_batch():
    sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
    from imagekit_cuda.sharding import shard_indices
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_indices(1000, world, r))
        assert seen == list(range(1000))
        sizes = [len(shard_indices(1000, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_reference_arm_only_rank0_works():
    """bench.py --impl reference under torchrun: ranks other than 0 exit 0 without work or output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_prints_contract_line():
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
