#!/usr/bin/env python
"""Parity driver of the limb-sharded mode on REAL peers (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_limb_shard.py [--full]

Every rank holds the limbs j = rank (mod world) of the same synthetic batch and checks its share of
mul_ciphertexts_gadget + rescale_ciphertext (engine.rs:473-539, 263-282)
  * against the oracle at N=4096, L=6 (61-bit) and N=16384, L=8 (30-bit, 32-bit word path), and
  * with --full at N=65536, L=24 against the unsharded device path of the same GPU,
over two consecutive levels and repeated calls (buffer reuse), with the fused peer-store exchange and with
NCCL collectives between the phases.  Prints one line `LIMB_SHARD_PARITY ok ...` on rank 0; any mismatch
exits non-zero.  (pytest cannot span several GPUs' processes; tests/test_gpu_limb_shard.py covers the same
code with ranks sharing one GPU.)"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def uniform_limbs(rng, moduli, n, *lead):
    q = np.array(moduli, dtype=np.uint64)
    raw = rng.integers(0, 1 << 63, size=(*lead, len(moduli), n), dtype=np.uint64)
    return (raw % q[:, None]).astype(np.uint64)


class _DevBuf:
    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def nccl_step(ck, dist, torch, sh, kid, cta, ctb, key, batch, dev):
    """The three phases with NCCL collectives on the exchange buffers (peer_stores = 0)."""
    l, n, world, rank = sh.channel_count(), sh.degree, sh.world, sh.rank
    chunk = sh.chunk()
    gp, gw, lp, lw = sh.buffers()
    gt = torch.as_tensor(_DevBuf(gp, gw), device=dev).view(-1, chunk * n)
    lt = torch.as_tensor(_DevBuf(lp, lw), device=dev)
    out = ck.Ciphertext(ck.RnsPoly.zero(kid.local_basis(), batch), ck.RnsPoly.zero(kid.local_basis(), batch), 0, 0)
    for s0 in range(0, batch, chunk):
        cs = min(chunk, batch - s0)
        sh.mul_phase(0, s0, cs, cta, ctb, key, kid, out, False)
        for k0 in range(0, l, world):
            if k0 + world <= l:
                dist.all_gather_into_tensor(gt[k0 : k0 + world].view(-1), gt[k0 + rank])
            else:
                for i in range(k0, l):
                    dist.broadcast(gt[i], src=i % world)
        sh.mul_phase(1, s0, cs, cta, ctb, key, kid, out, False)
        dist.broadcast(lt, src=(l - 1) % world)
        sh.mul_phase(2, s0, cs, cta, ctb, key, kid, out, False)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also N=65536, L=24 against the unsharded device path")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ck = importlib.import_module("toy-heaan-ckks_b200")
    import oracle as orc

    checked = []
    stream = torch.cuda.current_stream()
    cases = [(4096, 61, 6, 3, 2), (16384, 30, 8, 2, 0)]
    if args.full:
        cases.append((65536, 61, 24, 3, 2))
    for n, bits, l, batch, chunk in cases:
        if l - 2 < world:
            continue
        moduli = orc.generate_primes(bits, l, n)
        rng = np.random.default_rng(1000 + n)  # the same global data on every rank
        a0, a1, b0, b1 = (uniform_limbs(rng, moduli, n, batch) for _ in range(4))
        sh = ck.LimbShard(n, moduli, rank, world, device=local, chunk=chunk)
        sh.local_basis().set_stream(stream.cuda_stream)
        sh.set_timeout_ms(60000)
        sh.connect_process_group()
        use_oracle = n < 65536
        gb = None if use_oracle else ck.RnsBasis(n, moduli, device=local)
        level, ha = sh, (a0, a1, b0, b1)
        ca = ck.Ciphertext(sh.scatter(a0), sh.scatter(a1), bits, bits * l)
        cb = ck.Ciphertext(sh.scatter(b0), sh.scatter(b1), bits, bits * l)
        for depth in range(2):
            lv = l - depth
            mods = moduli[:lv]
            ka, kb = uniform_limbs(rng, mods, n, lv), uniform_limbs(rng, mods, n, lv)
            key = level.upload_key(ka, kb)
            kid = level.drop_last()
            if use_oracle:
                ob = orc.Basis(n, mods)
                e0, e1 = [], []
                for i in range(batch):
                    m0, m1 = ob.mul_ciphertexts_gadget(ha[0][i], ha[1][i], ha[2][i], ha[3][i], ka, kb)
                    o0, o1, _ = ob.rescale_ciphertext(m0, m1)
                    e0.append(o0)
                    e1.append(o1)
                e0, e1 = np.stack(e0), np.stack(e1)
            else:
                fb = gb if depth == 0 else gb.drop_last(depth)
                full = ck.CkksEngine.mul_relin_rescale(
                    ck.Ciphertext(ck.RnsPoly.from_channels(ha[0], fb), ck.RnsPoly.from_channels(ha[1], fb), 0, 0),
                    ck.Ciphertext(ck.RnsPoly.from_channels(ha[2], fb), ck.RnsPoly.from_channels(ha[3], fb), 0, 0),
                    ck.GadgetKey.upload(fb, ka, kb))
                e0, e1 = full.c0.channels(), full.c1.channels()
                del full
            out = None
            for rep in range(2):  # repeated calls reuse the exchange buffers
                out = level.mul_relin_rescale(ca, cb, key, kid)
                level.check()
                g0, g1 = out.c0.channels(), out.c1.channels()
                if not (np.array_equal(g0, e0[:, rank::world]) and np.array_equal(g1, e1[:, rank::world])):
                    print(f"rank {rank}: MISMATCH peer-store exchange N={n} L={lv} rep={rep}", flush=True)
                    sys.exit(1)
            level.set_exchange(1)  # digits by the copy engines
            outc = level.mul_relin_rescale(ca, cb, key, kid)
            level.check()
            level.set_exchange(0)
            if not (np.array_equal(outc.c0.channels(), e0[:, rank::world]) and np.array_equal(outc.c1.channels(), e1[:, rank::world])):
                print(f"rank {rank}: MISMATCH copy-engine exchange N={n} L={lv}", flush=True)
                sys.exit(1)
            outn = nccl_step(ck, dist, torch, level, kid, ca, cb, key, batch, dev)
            level.check()
            if not (np.array_equal(outn.c0.channels(), e0[:, rank::world]) and np.array_equal(outn.c1.channels(), e1[:, rank::world])):
                print(f"rank {rank}: MISMATCH NCCL exchange N={n} L={lv}", flush=True)
                sys.exit(1)
            checked.append((n, lv))
            ha = (e0, e1, e0, e1)  # next level: square
            ca = cb = out
            level = kid
        dist.barrier()
    ok = torch.tensor([1], device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        print(f"LIMB_SHARD_PARITY ok world={world} ranks_ok={int(ok.item())} cases={checked}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
