"""Host-side logic of the multi-GPU paths, on CPU: the sharding plan (ltxv_parallel_plan) and the world_size-2 `gloo`
bootstrap that exchanges the communicator's 64-byte IPC handles (candle_video_b200.exchange_handles).  No GPU work:
the data path itself (peer stores + flag barrier) is covered by tests/test_gpu_multi.py on a multi-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import candle_video_b200 as cv


@pytest.mark.parametrize("S", [4992, 13680, 256])
@pytest.mark.parametrize("n", [1, 2, 4, 8])
@pytest.mark.parametrize("cfg", [False, True])
def test_plan_tiles_the_sequence(S, n, cfg):
    plans = [cv.parallel_plan(n, r, S, cfg) for r in range(n)]
    groups = plans[0]["cfg_groups"]
    assert groups == (2 if (cfg and n % 2 == 0) else 1)
    sp = plans[0]["sp_size"]
    assert groups * sp == n
    for b in range(groups):
        cover = []
        for r, p in enumerate(plans):
            assert p["cfg_groups"] == groups and p["sp_size"] == sp
            if p["branch"] != b:
                continue
            assert p["sp_rank"] == r - b * sp
            assert p["local_tokens"] == S // sp
            cover.append((p["token0"], p["token0"] + p["local_tokens"]))
        cover.sort()
        assert cover[0][0] == 0 and cover[-1][1] == S
        assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))
    # partner of a rank in the CFG exchange = same token shard in the other branch group
    if groups == 2:
        for r, p in enumerate(plans):
            q = plans[(r + sp) % n]
            assert q["token0"] == p["token0"] and q["branch"] != p["branch"]


def test_plan_rejects_indivisible_sequence():
    with pytest.raises(cv.LtxvError, match="not divisible"):
        cv.parallel_plan(8, 0, 4991, False)
    with pytest.raises(cv.LtxvError, match="invalid rank"):
        cv.parallel_plan(2, 2, 64, False)


def test_comm_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(cv.LtxvError):
        cv.PeerComm(1, 0, 0, heap_bytes=1 << 20)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([(rank * 37 + i) & 0xFF for i in range(64)])
    allh = cv.exchange_handles(mine)
    plan = cv.parallel_plan(world, rank, 4992, True)
    q.put((rank, allh, plan))
    dist.barrier()
    dist.destroy_process_group()


def test_handle_exchange_over_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = b"".join(bytes([(r * 37 + i) & 0xFF for i in range(64)]) for r in range(world))
    for rank, allh, plan in res:
        assert allh == want                      # every rank sees every handle, indexed by rank
        assert plan["branch"] == rank and plan["sp_size"] == 1 and plan["local_tokens"] == 4992
