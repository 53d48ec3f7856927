"""Limb-sharded mode (optional): the limbs of one batch spread over `world` ranks, digits
all-gathered and the dropped limb broadcast by stores into the peers' buffers, flag barrier in between.
Every output limb must equal the oracle's mul_ciphertexts_gadget / rescale_ciphertext words
(engine.rs:473-539, 263-282) and the unsharded device path.

The ranks of these tests share ONE GPU: in one process (direct pointers, ranks driven in lockstep with an
event barrier) or as two processes (CUDA IPC handles, the flag barrier in peer memory), which exercises the
code a multi-GPU box runs; tests/run_limb_shard.py is the torchrun driver for real peers."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import uniform_limbs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ct(gpu, sh, c0, c1, logp=30, logq=90):
    return gpu.Ciphertext(sh.scatter(c0), sh.scatter(c1), logp, logq)


def _assemble(parts, world, l):
    """Per-rank [batch, L_own, N] arrays -> [batch, L, N] in basis order."""
    batch, _, n = parts[0].shape
    out = np.zeros((batch, l, n), dtype=np.uint64)
    for r in range(world):
        out[:, r::world] = parts[r]
    return out


def _group(gpu, n, moduli, world, chunk):
    shards = [gpu.LimbShard(n, moduli, r, world, chunk=chunk) for r in range(world)]
    gpu.LimbShard.connect_local(shards)
    for s in shards:
        s.set_timeout_ms(30000)
    return shards


@pytest.mark.parametrize("n,bits,l,world,batch,chunk", [
    (1024, 40, 4, 2, 3, 2),     # two chunks, the second one ragged
    (4096, 61, 5, 3, 2, 8),     # lazy8 butterflies; ownership 2/2/1
    (16384, 30, 6, 4, 2, 1),    # 32-bit word path; ownership 2/2/1/1
    (256, 62, 2, 2, 2, 8),      # one limb per rank, Harvey butterflies
    (2048, 63, 3, 1, 5, 2),     # a group of one (no peers): strict arithmetic; three chunks through the pipeline
])
def test_limb_sharded_mul_relin_rescale_matches_oracle(gpu, orc, n, bits, l, world, batch, chunk):
    moduli = orc.generate_primes(bits, l, n)
    ob = orc.Basis(n, moduli)
    rng = np.random.default_rng(900 + n + world)
    a0, a1, b0, b1 = (uniform_limbs(rng, moduli, n, batch) for _ in range(4))
    ka, kb = uniform_limbs(rng, moduli, n, l), uniform_limbs(rng, moduli, n, l)
    shards = _group(gpu, n, moduli, world, chunk)
    assert [s.owned() for s in shards] == [list(range(r, l, world)) for r in range(world)]
    keys = [s.upload_key(ka, kb) for s in shards]
    cta = [_ct(gpu, s, a0, a1) for s in shards]
    ctb = [_ct(gpu, s, b0, b1) for s in shards]
    # without rescale
    prods = gpu.LimbShard.group_mul_relin_rescale(shards, cta, ctb, keys)
    for s in shards:
        s.check()
    g0 = _assemble([p.c0.channels() for p in prods], world, l)
    g1 = _assemble([p.c1.channels() for p in prods], world, l)
    exp = [ob.mul_ciphertexts_gadget(a0[i], a1[i], b0[i], b1[i], ka, kb) for i in range(batch)]
    for i in range(batch):
        assert np.array_equal(g0[i], exp[i][0]) and np.array_equal(g1[i], exp[i][1]), "mul_ciphertexts_gadget limbs differ"
    assert prods[0].logp == 60 and prods[0].logq == 90
    if l - 1 < world:
        with pytest.raises(gpu.RnsNttError):
            shards[0].drop_last()
        return
    # fused with rescale into the child level
    kids = [s.drop_last() for s in shards]
    res = gpu.LimbShard.group_mul_relin_rescale(shards, cta, ctb, keys, kids)
    for s in shards:
        s.check()
    r0 = _assemble([p.c0.channels() for p in res], world, l - 1)
    r1 = _assemble([p.c1.channels() for p in res], world, l - 1)
    for i in range(batch):
        o0, o1, dropped = ob.rescale_ciphertext(exp[i][0], exp[i][1])
        assert np.array_equal(r0[i], o0) and np.array_equal(r1[i], o1), "rescale_ciphertext limbs differ"
        assert res[0].logp == 60 - dropped and res[0].logq == 90 - dropped
    if world == 1:  # the one-call entry point (chunk pipeline over two streams; no peers to wait for)
        one = shards[0].mul_relin_rescale(cta[0], ctb[0], keys[0], kids[0])
        shards[0].check()
        assert np.array_equal(one.c0.channels(), r0) and np.array_equal(one.c1.channels(), r1)
    # and the unsharded device path agrees word for word
    gb = gpu.RnsBasis(n, moduli)
    full = gpu.CkksEngine.mul_relin_rescale(
        gpu.Ciphertext(gpu.RnsPoly.from_channels(a0, gb), gpu.RnsPoly.from_channels(a1, gb), 30, 90),
        gpu.Ciphertext(gpu.RnsPoly.from_channels(b0, gb), gpu.RnsPoly.from_channels(b1, gb), 30, 90),
        gpu.GadgetKey.upload(gb, ka, kb))
    assert np.array_equal(full.c0.channels(), r0) and np.array_equal(full.c1.channels(), r1)


def test_limb_sharded_chain_two_levels(gpu, orc):
    """horner_chain-style descent (examples/horner_chain.rs:211-265): two mul+rescale levels with a fresh key
    per level; ownership of the dropped limb moves from rank to rank."""
    n, l, world, batch = 1024, 6, 2, 2
    moduli = orc.generate_primes(40, l, n)
    rng = np.random.default_rng(4242)
    x0, x1, y0, y1 = (uniform_limbs(rng, moduli, n, batch) for _ in range(4))
    level = _group(gpu, n, moduli, world, 4)
    cx = [_ct(gpu, s, x0, x1) for s in level]
    cy = [_ct(gpu, s, y0, y1) for s in level]
    hx0, hx1, hy0, hy1 = x0, x1, y0, y1
    for depth in range(2):
        lv = l - depth
        mods = moduli[:lv]
        ob = orc.Basis(n, mods)
        ka, kb = uniform_limbs(rng, mods, n, lv), uniform_limbs(rng, mods, n, lv)
        keys = [s.upload_key(ka, kb) for s in level]
        kids = [s.drop_last() for s in level]
        out = gpu.LimbShard.group_mul_relin_rescale(level, cx, cy, keys, kids)
        for s in level:
            s.check()
        got0 = _assemble([o.c0.channels() for o in out], world, lv - 1)
        got1 = _assemble([o.c1.channels() for o in out], world, lv - 1)
        e0, e1 = [], []
        for i in range(batch):
            m0, m1 = ob.mul_ciphertexts_gadget(hx0[i], hx1[i], hy0[i], hy1[i], ka, kb)
            o0, o1, _ = ob.rescale_ciphertext(m0, m1)
            e0.append(o0)
            e1.append(o1)
        assert np.array_equal(got0, np.stack(e0)) and np.array_equal(got1, np.stack(e1)), f"level {depth}"
        # next level: square the result
        hx0 = hy0 = np.stack(e0)
        hx1 = hy1 = np.stack(e1)
        cx = cy = out
        level = kids


@pytest.mark.parametrize("n,bits,l,world,rots", [(1024, 30, 4, 2, [1, -3]), (4096, 61, 5, 3, [5]), (16384, 30, 8, 4, [64])])
def test_limb_sharded_rotate_matches_oracle(gpu, orc, n, bits, l, world, rots):
    """rotate_ciphertext (engine.rs:412-463): automorphism limb-local, rotated c1 pushed to every rank, key-switch."""
    moduli = orc.generate_primes(bits, l, n)
    ob = orc.Basis(n, moduli)
    rng = np.random.default_rng(77 + n)
    batch = 3
    c0, c1 = uniform_limbs(rng, moduli, n, batch), uniform_limbs(rng, moduli, n, batch)
    shards = _group(gpu, n, moduli, world, 2)  # two chunks
    cts = [_ct(gpu, s, c0, c1) for s in shards]
    for k in rots:
        ka, kb = uniform_limbs(rng, moduli, n, l), uniform_limbs(rng, moduli, n, l)
        keys = [s.upload_key(ka, kb, rotation=k) for s in shards]
        outs = gpu.LimbShard.group_rotate(shards, cts, keys)
        for s in shards:
            s.check()
        g0 = _assemble([o.c0.channels() for o in outs], world, l)
        g1 = _assemble([o.c1.channels() for o in outs], world, l)
        for i in range(batch):
            e0, e1 = ob.rotate_ciphertext(c0[i], c1[i], ka, kb, k)
            assert np.array_equal(g0[i], e0) and np.array_equal(g1[i], e1), f"rotation {k}"


def test_limb_sharded_phases_with_caller_run_collectives(gpu, orc):
    """peer_stores = 0: the kernels fill only the rank's own slots; the caller moves the digits and the
    dropped limb itself (here: device-to-device copies standing in for NCCL all-gather / broadcast)."""
    import ctypes as C

    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    n, l, world, batch = 1024, 5, 2, 2
    moduli = orc.generate_primes(61, l, n)
    ob = orc.Basis(n, moduli)
    rng = np.random.default_rng(31337)
    a0, a1, b0, b1 = (uniform_limbs(rng, moduli, n, batch) for _ in range(4))
    ka, kb = uniform_limbs(rng, moduli, n, l), uniform_limbs(rng, moduli, n, l)
    shards = [gpu.LimbShard(n, moduli, r, world, chunk=4) for r in range(world)]  # never connected
    kids = [s.drop_last() for s in shards]
    keys = [s.upload_key(ka, kb) for s in shards]
    cta = [_ct(gpu, s, a0, a1) for s in shards]
    ctb = [_ct(gpu, s, b0, b1) for s in shards]
    outs = [gpu.Ciphertext(gpu.RnsPoly.zero(k.local_basis(), batch), gpu.RnsPoly.zero(k.local_basis(), batch), 0, 0) for k in kids]
    with pytest.raises(gpu.RnsNttError):  # peer stores need a connected group
        shards[0].mul_phase(0, 0, batch, cta[0], ctb[0], keys[0], kids[0], outs[0], True)
    chunk = shards[0].chunk()
    bufs = [s.buffers() for s in shards]
    for r, s in enumerate(shards):
        s.mul_phase(0, 0, batch, cta[r], ctb[r], keys[r], kids[r], outs[r], False)
        s.check()
    # "all-gather": slot i of every rank <- slot i of its owner
    slot_bytes = chunk * n * 8
    for i in range(l):
        src = bufs[i % world][0] + i * slot_bytes
        for r in range(world):
            if r != i % world:
                assert rt.cudaMemcpy(bufs[r][0] + i * slot_bytes, src, slot_bytes, 3) == 0
    assert rt.cudaDeviceSynchronize() == 0  # device-to-device cudaMemcpy does not block the host
    for r, s in enumerate(shards):
        s.mul_phase(1, 0, batch, cta[r], ctb[r], keys[r], kids[r], outs[r], False)
        s.check()
    owner = (l - 1) % world
    for r in range(world):  # "broadcast" of the dropped limb
        if r != owner:
            assert rt.cudaMemcpy(bufs[r][2], bufs[owner][2], bufs[owner][3] * 8, 3) == 0
    assert rt.cudaDeviceSynchronize() == 0
    for r, s in enumerate(shards):
        s.mul_phase(2, 0, batch, cta[r], ctb[r], keys[r], kids[r], outs[r], False)
        s.check()
    r0 = _assemble([o.c0.channels() for o in outs], world, l - 1)
    r1 = _assemble([o.c1.channels() for o in outs], world, l - 1)
    for i in range(batch):
        m0, m1 = ob.mul_ciphertexts_gadget(a0[i], a1[i], b0[i], b1[i], ka, kb)
        o0, o1, _ = ob.rescale_ciphertext(m0, m1)
        assert np.array_equal(r0[i], o0) and np.array_equal(r1[i], o1)


def test_limb_sharded_argument_checks(gpu, orc):
    n = 1024
    moduli = orc.generate_primes(40, 3, n)
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.LimbShard(n, moduli, 0, 4)  # fewer limbs than ranks
    assert e.value.kind == "Unsupported"
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.LimbShard(64, orc.generate_primes(40, 3, 64), 0, 2)  # below the four-step path
    assert e.value.kind == "Unsupported"
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.LimbShard(n, [], 0, 1)
    assert e.value.kind == "EmptyBasis"
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.LimbShard(n, [moduli[0], 19], 0, 2)  # the whole basis is validated, not only the own share
    assert e.value.kind == "NonNttFriendlyModulus"
    s = gpu.LimbShard(n, moduli, 1, 2, chunk=2)
    assert s.owned() == [1] and s.local_basis().moduli() == [moduli[1]] and s.channel_count() == 3
    rng = np.random.default_rng(1)
    wrong = gpu.RnsBasis(n, moduli)
    full = gpu.RnsPoly.from_channels(uniform_limbs(rng, moduli, n, 1), wrong)
    key = s.upload_key(uniform_limbs(rng, moduli, n, 3), uniform_limbs(rng, moduli, n, 3))
    ct = gpu.Ciphertext(full, full, 0, 0)
    peer = gpu.LimbShard(n, moduli, 0, 2, chunk=2)
    gpu.LimbShard.connect_local([peer, s])
    with pytest.raises(gpu.RnsNttError) as e:
        s.mul_relin_rescale(ct, ct, key)  # polynomials of the whole basis, not of the share
    assert e.value.kind == "BasisMismatch"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, q):
    import importlib

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ck = importlib.import_module("toy-heaan-ckks_b200")
        import oracle as orc

        n, l, batch = 2048, 4, 5
        moduli = orc.generate_primes(61, l, n)
        rng = np.random.default_rng(555)  # the same global batch on every rank
        a0, a1, b0, b1 = (uniform_limbs(rng, moduli, n, batch) for _ in range(4))
        ka, kb = uniform_limbs(rng, moduli, n, l), uniform_limbs(rng, moduli, n, l)
        sh = ck.LimbShard(n, moduli, rank, world, device=0, chunk=2)  # three chunks through the pipeline
        sh.set_timeout_ms(60000)
        sh.connect_process_group()
        kid = sh.drop_last()
        key = sh.upload_key(ka, kb)
        cta = ck.Ciphertext(sh.scatter(a0), sh.scatter(a1), 30, 90)
        ctb = ck.Ciphertext(sh.scatter(b0), sh.scatter(b1), 30, 90)
        ok = True
        for _ in range(3):  # repeated calls reuse the exchange buffers: epochs and hazards
            out = sh.mul_relin_rescale(cta, ctb, key, kid)
            sh.check()
            g0, g1 = out.c0.channels(), out.c1.channels()
            ob = orc.Basis(n, moduli)
            for i in range(batch):
                m0, m1 = ob.mul_ciphertexts_gadget(a0[i], a1[i], b0[i], b1[i], ka, kb)
                o0, o1, _ = ob.rescale_ciphertext(m0, m1)
                ok &= bool(np.array_equal(g0[i], o0[rank::world]) and np.array_equal(g1[i], o1[rank::world]))
        ra, rb = uniform_limbs(rng, moduli, n, l), uniform_limbs(rng, moduli, n, l)
        rot = sh.rotate_ciphertext(cta, sh.upload_key(ra, rb, rotation=3))
        sh.check()
        for i in range(batch):
            e0, e1 = ob.rotate_ciphertext(a0[i], a1[i], ra, rb, 3)
            ok &= bool(np.array_equal(rot.c0.channels()[i], e0[rank::world]) and np.array_equal(rot.c1.channels()[i], e1[rank::world]))
        dist.barrier()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_limb_sharded_two_processes_cuda_ipc(gpu):
    """One process per rank as under torchrun: buffers exchanged as CUDA IPC handles over gloo."""
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert got == [(0, True), (1, True)]


def test_lost_peer_is_an_error_not_a_hang(gpu, orc):
    """Only rank 0 of a group of two makes the call: its flag barriers give up after the time limit and the stream
    drains instead of leaving a kernel spinning on the GPU.  The failure is CLOSED: the outputs are poisoned with
    all-ones words (never canonical), the first blocking call on the local context (download) returns NcclError, so
    do check() and every later call on the group (sticky) -- nothing that looks like ciphertext limbs gets out."""
    n, l = 1024, 4
    moduli = orc.generate_primes(40, l, n)
    rng = np.random.default_rng(3)
    shards = [gpu.LimbShard(n, moduli, r, 2, chunk=2) for r in range(2)]
    gpu.LimbShard.connect_local(shards)
    shards[0].set_timeout_ms(200)
    a0, a1 = uniform_limbs(rng, moduli, n, 2), uniform_limbs(rng, moduli, n, 2)
    key = shards[0].upload_key(uniform_limbs(rng, moduli, n, l), uniform_limbs(rng, moduli, n, l))
    ct = _ct(gpu, shards[0], a0, a1)
    out = shards[0].mul_relin_rescale(ct, ct, key)  # enqueues; nobody answers
    with pytest.raises(gpu.RnsNttError) as e:
        out.c0.channels()  # the first sync point sees the failure
    assert e.value.kind == "NcclError"
    with pytest.raises(gpu.RnsNttError) as e:
        shards[0].check()
    assert e.value.kind == "NcclError" and "timed out" in str(e.value)
    with pytest.raises(gpu.RnsNttError) as e:  # sticky: the group is dead until it is rebuilt
        shards[0].mul_relin_rescale(ct, ct, key)
    assert e.value.kind == "NcclError" and "re-create" in str(e.value)
    with pytest.raises(gpu.RnsNttError):
        shards[0].check()
    # the words behind the failed barrier were overwritten: read them through the raw device pointer
    import ctypes as C

    ptr = C.POINTER(C.c_uint64)()
    gpu._check(gpu._lib.ckks_poly_device_ptr(out.c0._h, C.byref(ptr)))
    import torch

    class _Buf:
        def __init__(self, addr, words):
            self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (addr, False), "version": 3}

    words = 2 * out.c0.channel_count() * n
    raw = torch.as_tensor(_Buf(C.cast(ptr, C.c_void_p).value, words), device="cuda:0").cpu().numpy().view(np.uint64)
    assert np.all(raw == np.uint64(0xFFFFFFFFFFFFFFFF))
    # recovery = a new group
    del out, ct, key, shards
    fresh = [gpu.LimbShard(n, moduli, r, 2, chunk=2) for r in range(2)]
    gpu.LimbShard.connect_local(fresh)
    fresh[0].check()


def test_limb_sharded_parity_on_every_gpu_of_the_box(gpu):
    """The limb-sharded mode on REAL peers: one process per GPU under torchrun (tests/run_limb_shard.py --full):
    every limb against the oracle (N=4096 61-bit, N=16384 30-bit) and, at N=65536, L=24, against the unsharded device
    path, with the fused peer-store exchange and with NCCL between the phases.  Needs at least two GPUs; on a
    one-GPU box the same code paths run with the ranks sharing the device (the tests above)."""
    import socket
    import subprocess
    import sys

    ndev = min(gpu.device_count(), 8)
    if ndev < 2:
        pytest.skip("one GPU: covered by the shared-device tests above")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ndev}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "run_limb_shard.py"), "--full"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"LIMB_SHARD_PARITY ok world={ndev}" in r.stdout, r.stdout[-2000:]


def test_batch_sharded_group_on_every_gpu_of_the_box(gpu, orc):
    """ckks_comm_* with one slot per physical GPU (skipped on a one-GPU box, where tests/test_gpu_engine.py runs the
    same entry points with the slots sharing the device): N=2^14, L=4, ragged batch, every word against the oracle."""
    ndev = min(gpu.device_count(), 8)
    if ndev < 2:
        pytest.skip("one GPU: covered by test_batch_sharded_group_matches_single_gpu_and_oracle")
    n, l, batch = 16384, 4, 2 * ndev + 1
    moduli = orc.generate_primes(30, l, n)
    ob = orc.Basis(n, moduli)
    rng = np.random.default_rng(99)
    a0, a1, b0, b1 = (uniform_limbs(rng, moduli, n, batch) for _ in range(4))
    ka, kb = uniform_limbs(rng, moduli, n, l), uniform_limbs(rng, moduli, n, l)
    grp = gpu.BatchShard(n, moduli, list(range(ndev)))
    key = grp.upload_key(ka, kb)
    o0 = np.zeros((batch, l - 1, n), dtype=np.uint64)
    o1 = np.zeros_like(o0)
    grp.mul_relin_rescale_host(key, a0, a1, b0, b1, o0, o1)
    for i in range(batch):
        m0, m1 = ob.mul_ciphertexts_gadget(a0[i], a1[i], b0[i], b1[i], ka, kb)
        e0, e1, _ = ob.rescale_ciphertext(m0, m1)
        assert np.array_equal(o0[i], e0) and np.array_equal(o1[i], e1), i
