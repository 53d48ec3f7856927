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


def _limb_worker(rank, world, port, q):
    """The limb-sharded protocol (toy-heaan-ckks_b200/csrc/limb_shard.inl) replayed with the oracle's
    per-limb arithmetic and gloo collectives: limb j on rank j mod world, all-gather of the digits of d2,
    key-switch on the own target limbs with the own key slices, broadcast of the dropped limb, rescale."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ck = importlib.import_module("toy-heaan-ckks_b200")
        import oracle as orc

        n, l = 64, 5
        moduli = orc.generate_primes(40, l, n)
        own = ck.owned_limbs(l, rank, world)
        assert own == list(range(rank, l, world))
        own_q = [moduli[j] for j in own]
        ob = orc.Basis(n, own_q)
        rng = np.random.default_rng(2024)  # the same global data on every rank
        qq = np.array(moduli, dtype=np.uint64)
        a0, a1, b0, b1 = ((rng.integers(0, 1 << 63, size=(l, n), dtype=np.uint64) % qq[:, None]) for _ in range(4))
        ka, kb = ((rng.integers(0, 1 << 63, size=(l, l, n), dtype=np.uint64) % qq[None, :, None]) for _ in range(2))
        mine = lambda x: np.ascontiguousarray(x[rank::world])
        # phase A (limb-local): tensor product in the coefficient domain
        d0 = ob.mul(mine(a0), mine(b0))
        d1 = ob.add(ob.mul(mine(a0), mine(b1)), ob.mul(mine(a1), mine(b0)))
        d2 = ob.mul(mine(a1), mine(b1))
        # all-gather of the digits: slot i <- limb i of d2 from rank i mod world
        gather = torch.zeros((l, n), dtype=torch.int64)
        for i in range(l):
            t = torch.from_numpy(d2[i // world].astype(np.int64)) if i % world == rank else torch.zeros(n, dtype=torch.int64)
            dist.broadcast(t, src=i % world)
            gather[i] = t
        digits = gather.numpy().astype(np.uint64)
        # phase B: key-switch on the own target limbs with the own key slices [digit][own limb]
        oq = np.array(own_q, dtype=np.uint64)[:, None]
        c0, c1 = d0, d1
        for i in range(l):
            alpha = (digits[i][None, :] % oq).astype(np.uint64)  # engine.rs:507-516
            c0 = ob.add(c0, ob.mul(alpha, np.ascontiguousarray(kb[i, rank::world])))
            c1 = ob.add(c1, ob.mul(alpha, np.ascontiguousarray(ka[i, rank::world])))
        # broadcast of the limb rescale drops, then poly.rs:214-225 on the own limbs
        owner = (l - 1) % world
        last = torch.zeros((2, n), dtype=torch.int64)
        if rank == owner:
            last = torch.from_numpy(np.stack([c0[-1], c1[-1]]).astype(np.int64))
        dist.broadcast(last, src=owner)
        last = last.numpy().astype(np.uint64)
        keep = [jl for jl, j in enumerate(own) if j != l - 1]
        ql = moduli[-1]
        res = []
        for comp, c in enumerate((c0, c1)):
            rows = []
            for jl in keep:
                qj = own_q[jl]
                inv = pow(ql % qj, -1, qj)
                rows.append(np.array([((int(x) - int(y) % qj) * inv) % qj for x, y in zip(c[jl], last[comp])], dtype=np.uint64))
            res.append(np.stack(rows))
        full = orc.Basis(n, moduli)
        m0, m1 = full.mul_ciphertexts_gadget(a0, a1, b0, b1, ka, kb)
        r0, r1, _ = full.rescale_ciphertext(m0, m1)
        kept = [j for j in own if j != l - 1]
        ok = bool(np.array_equal(res[0], r0[kept]) and np.array_equal(res[1], r1[kept]))
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        if rank == 0:
            q.put(oks)
    finally:
        dist.destroy_process_group()


def test_limb_sharded_protocol_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_limb_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) == [True, True]


def test_owned_limbs_partition_and_stay_balanced():
    ck = importlib.import_module("toy-heaan-ckks_b200")
    for l in range(1, 33):
        for world in (1, 2, 3, 4, 8):
            parts = [ck.owned_limbs(l, r, world) for r in range(world)]
            assert sorted(j for p in parts for j in p) == list(range(l))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1  # also after every drop_last, since it is the same rule at l-1
