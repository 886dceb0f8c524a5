"""N>1 host logic on CPU (gloo, world_size 2): the plan broadcast makes every rank hold
rank 0's derived integers, and the stream shards partition the batch."""
import ctypes as C
import os

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import minimodem_b200 as mm
    from minimodem_b200 import dist as mdist
    mode = "1200" if rank == 0 else "300"          # the non-root's own derivation must be discarded
    p = mm.rx_params(mm.rx_config_for_mode(mode, 48000))
    got = mdist.broadcast_params(p, src=0)
    lo, hi = mdist.shard_range(65537, rank, world)
    out.put((rank, bytes(got), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_plan_broadcast_and_shards_world2():
    import minimodem_b200 as mm
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes(mm.rx_params(mm.rx_config_for_mode("1200", 48000)))
    assert all(r[1] == want for r in res)
    p = mm.RxParams.from_buffer_copy(res[1][1])
    assert (p.fftsize, p.b_mark, p.b_space, p.expect_nsamples, p.bit_nsamples) == (240, 6, 11, 440, 40)
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 65537
    assert abs((res[0][3] - res[0][2]) - (res[1][3] - res[1][2])) <= 1


def test_shard_range_partitions():
    from minimodem_b200.dist import shard_range
    for n in (0, 1, 7, 8, 65536, 1048576 + 3):
        for w in (1, 2, 4, 8):
            cuts = [shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
