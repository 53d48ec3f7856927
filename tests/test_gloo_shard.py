"""Multi-GPU host logic on CPU: world_size 2 over gloo.  Batch sharding must partition the batch,
each rank's results must equal the unsharded computation (ciphertexts are independent), and the
timing reduction must be a MAX over ranks."""
import importlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = importlib.util.spec_from_file_location("shard", os.path.join(ROOT, "toy-heaan-ckks_b200", "shard.py"))
        shard = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(shard)
        import oracle as orc

        lo, hi = shard.shard_range(batch, rank, world)
        n, l = 64, 2
        moduli = orc.generate_primes(40, l, n)
        ob = orc.Basis(n, moduli)
        rng = np.random.default_rng(77)  # every rank derives the same global batch, then keeps its slice
        qq = np.array(moduli, dtype=np.uint64)
        data = (rng.integers(0, 1 << 63, size=(6, batch, l, n), dtype=np.uint64) % qq[:, None]).astype(np.uint64)
        ka, kb = data[4, :l], data[5, :l]
        mine = []
        for i in range(lo, hi):
            m0, m1 = ob.mul_ciphertexts_gadget(data[0, i], data[1, i], data[2, i], data[3, i], ka, kb)
            r0, _, _ = ob.rescale_ciphertext(m0, m1)
            mine.append(r0)
        local = torch.from_numpy(np.stack(mine).astype(np.int64)) if mine else torch.zeros((0, l - 1, n), dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        gathered = [torch.zeros((int(s.item()), l - 1, n), dtype=torch.int64) for s in sizes]
        dist.all_gather(gathered, local) if len(set(int(s.item()) for s in sizes)) == 1 else None
        ms = shard.max_over_ranks(10.0 + rank)
        if rank == 0:
            ok = True
            if len(set(int(s.item()) for s in sizes)) == 1:
                full = torch.cat(gathered).numpy().astype(np.uint64)
                for i in range(batch):
                    m0, m1 = ob.mul_ciphertexts_gadget(data[0, i], data[1, i], data[2, i], data[3, i], ka, kb)
                    r0, _, _ = ob.rescale_ciphertext(m0, m1)
                    ok &= bool(np.array_equal(full[i], r0))
            q.put((ok, ms, [int(s.item()) for s in sizes], shard.aggregate_rate(hi - lo, world, ms)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    spec = importlib.util.spec_from_file_location("shard", os.path.join(ROOT, "toy-heaan-ckks_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    for batch in (0, 1, 7, 8, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            got = [shard.shard_range(batch, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == batch
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo():
    world, batch = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok, ms, sizes, rate = q.get(timeout=10)
    assert ok and sizes == [2, 2]
    assert ms == 11.0  # max over ranks, not rank 0's own 10.0
    assert abs(rate - 2 * 2 / 11e-3) < 1e-6
