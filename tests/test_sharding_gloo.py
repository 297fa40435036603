"""N>1 host path on CPU: world_size-2 gloo.  Each rank verifies its contiguous shard (with the CPU
oracle standing in for the GPU), the packed verdict masks are all-gathered, and the quorum stream is
sharded by instance.  Checks the sharding / packing / gather logic that bench.py and libsbv.so use."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from consensus_b200 import sharding
    from oracle import corpus
    from oracle import ecdsa_ref as ref

    n = 1003  # ragged on purpose: not a multiple of 32 * world
    b = corpus.make_batch(oracle.P256, n=n, K=8, seed=77, corrupt_rate=4)
    lo, hi = sharding.shard_range(n, rank, world)
    local = oracle.verify_batch(oracle.P256, b["r"][lo:hi], b["s"][lo:hi], b["qx"][lo:hi], b["qy"][lo:hi], b["digest"][lo:hi], nthreads=2)
    full = sharding.gather_verdicts(local, n, rank, world)
    want = oracle.verify_batch(oracle.P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"], nthreads=2)
    ok_verdicts = bool(np.array_equal(full, want))

    # quorum stream sharded by instance: n=7 nodes, votes grouped by instance
    rng = np.random.default_rng(5)
    I, N = 101, 7
    q_, f_ = ref.compute_quorum(N)
    inst = np.repeat(np.arange(I), N - 1).astype(np.uint32)
    sender = np.tile(np.arange(1, N), I).astype(np.uint16)
    signer = sender.copy()
    signer[rng.random(sender.size) < 0.1] = 0
    dm = (rng.random(sender.size) < 0.9).astype(np.uint8)
    sig_ok = (rng.random(sender.size) < 0.85).astype(np.uint8)
    mask, ilo, ihi = sharding.shard_instances(inst, I, rank, world)
    counts_local = np.zeros(ihi - ilo, np.int64)
    for i in range(ilo, ihi):
        sel = inst == i
        votes = list(zip(sender[sel].tolist(), signer[sel].tolist(), dm[sel].tolist(), sig_ok[sel].tolist()))
        counts_local[i - ilo] = ref.count_commit_votes(votes, self_id=0)
    reached_local = (counts_local >= q_ - 1).astype(np.uint8)
    reached = sharding.gather_verdicts(reached_local, I, rank, world)
    want_r = []
    for i in range(I):
        sel = inst == i
        votes = list(zip(sender[sel].tolist(), signer[sel].tolist(), dm[sel].tolist(), sig_ok[sel].tolist()))
        want_r.append(int(ref.count_commit_votes(votes, self_id=0) >= q_ - 1))
    ok_quorum = reached.tolist() == want_r and bool(mask.sum() == (ihi - ilo) * (N - 1))
    # the vote stream as sbv_verify_quorum takes it on a rank: shard-local instance ids, counts local to the shard
    vlo, vhi, ilo2, ihi2, local_inst = sharding.shard_votes(inst, I, rank, world)
    cnt2, _ = ref.count_commit_votes_batch(local_inst, sender[vlo:vhi], signer[vlo:vhi], dm[vlo:vhi], sig_ok[vlo:vhi], ihi2 - ilo2, q_ - 1,
                                           np.zeros(ihi2 - ilo2, np.uint16))
    ok_quorum = ok_quorum and (ilo2, ihi2) == (ilo, ihi) and cnt2.tolist() == counts_local.tolist() and vhi - vlo == int(mask.sum())
    if rank == 0:
        q.put((ok_verdicts, ok_quorum, int(want.sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_multi_rank_shard_and_gather(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() * 7 + world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0], "gathered verdict mask differs from the single-process verdicts"
    assert res[1], "instance-sharded quorum bits differ"
    assert 0 < res[2] < 1003


def test_pack_unpack_roundtrip():
    from consensus_b200 import sharding
    rng = np.random.default_rng(1)
    for n in [0, 1, 31, 32, 33, 1000]:
        ok = (rng.random(n) < 0.5).astype(np.uint8)
        w = sharding.pack_bits(ok, (n + 31) // 32 + 1)
        assert np.array_equal(sharding.unpack_bits(w, n), ok)
    assert [sharding.shard_range(10, g, 4) for g in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
