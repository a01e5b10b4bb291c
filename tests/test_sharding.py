"""The multi-GPU path is counter-range sharding plus ONE broadcast of key||iv.
Covered here on CPU with world_size = 2 over gloo: rank 0 broadcasts the key material, each rank
runs its shard_plan() range (the checker stands in for the kernel: same first_block argument the
GPU path passes to uaes_ctr_crypt_range), and the concatenation must equal the unsharded stream."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from util import ROOT, Oracle, rnd

sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    orc = Oracle()
    kiv = torch.tensor(list(rnd("sh-key", 16) + rnd("sh-iv", 12)) if rank == 0 else [0] * 28, dtype=torch.uint8)
    dist.broadcast(kiv, src=0)
    key, iv = bytes(kiv[:16].tolist()), bytes(kiv[16:].tolist())
    total_blocks = 4096 + 3
    first, n = bench.shard_plan(total_blocks, world)[rank]
    data = orc.splitmix(7, first * 2, n * 2)
    out = orc.ctr(key, iv, data, first_block=first)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, n, out))
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def test_two_rank_counter_range_sharding():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = tmp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    orc = Oracle()
    total_blocks = 4096 + 3
    whole = orc.ctr(rnd("sh-key", 16), rnd("sh-iv", 12), orc.splitmix(7, 0, total_blocks * 2))
    assert [g[0] for g in gathered] == [0, total_blocks // 2]
    assert sum(g[1] for g in gathered) == total_blocks
    assert b"".join(g[2] for g in gathered) == whole


def test_shard_plan_covers_and_crosses_2_32():
    import bench
    # BASELINE config 5: 128 GiB = 2^33 blocks over 8 ranks; rank 3 crosses the 2^32 counter carry
    plan = bench.shard_plan(1 << 33, 8)
    assert plan[0] == (0, 1 << 30) and plan[7] == (7 << 30, 1 << 30)
    assert sum(n for _, n in plan) == 1 << 33
    assert plan[3][0] < (1 << 32) - 1 <= plan[3][0] + plan[3][1]
    plan = bench.shard_plan(1003, 4)
    assert [f for f, _ in plan] == [0, 250, 500, 750] and plan[-1][1] == 253


def _gcm_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    orc = Oracle()
    key, nonce, aad = rnd("gs-key", 16), rnd("gs-nonce", 12), rnd("gs-aad", 20)
    total = 16 * 1000 + 5                                    # ragged: the last shard owns the tail
    shards = bench.gcm_shards(total, world)
    off, n = shards[rank]
    data = rnd("gs-data", total)[off:off + n]
    # what uaes_gcm_shard does on a GPU, restated with the checker: CTR from J0 + 1 + first_block,
    # then the shard's GHASH contribution from a zero state
    ct = orc.ctr(key, nonce, data, first_block=1 + off // 16)
    H = orc.encrypt_block(key, bytes(16))
    part = torch.tensor(list(orc.ghash_absorb(H, ct, len(ct))), dtype=torch.uint8)
    gathered = [torch.zeros(16, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, part)                          # the path's one exchange: 16 B per rank
    cts = [None] * world
    dist.all_gather_object(cts, ct)
    if rank == 0:
        # what uaes_gcm_combine does: sum Z_r * H^(blocks after shard r), AAD in front, lengths, E(J0)
        after = bench.gcm_blocks_after(shards, total)
        state = orc.gf128_mul(orc.gf128_pow(H, (total + 15) // 16), orc.ghash_absorb(H, aad, len(aad)))
        for z, a in zip(gathered, after):
            term = orc.gf128_mul(orc.gf128_pow(H, a), bytes(z.tolist()))
            state = bytes(x ^ y for x, y in zip(state, term))
        lens = (len(aad) * 8).to_bytes(8, "big") + (total * 8).to_bytes(8, "big")
        state = orc.ghash_absorb(H, lens, 16, state)
        ej0 = orc.encrypt_block(key, nonce + b"\0\0\0\1")
        q.put((b"".join(cts), bytes(x ^ y for x, y in zip(state, ej0))))
    dist.destroy_process_group()


def test_two_rank_gcm_shards_and_16_byte_gather():
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = tmp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_gcm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ct, tag = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = Oracle().gcm_encrypt(rnd("gs-key", 16), rnd("gs-nonce", 12), rnd("gs-aad", 20), rnd("gs-data", 16 * 1000 + 5))
    assert ct == want[:-16] and tag == want[-16:]
