#!/usr/bin/env python
"""bench.py -- ct-mults/s (mul_ciphertexts_gadget + rescale_ciphertext) at N=2^16, L=24.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One "step" = one pass of the hot path over one batch of synthetic ciphertexts (limbs uniform in
[0, q_i); BASELINE.json configs[3]: N=65536, L=24 61-bit primes, batch 256 per GPU).  Multi-GPU is
batch sharding: every rank runs the same per-GPU batch with no collective on the data path
("scaling": "weak"); value = ciphertexts all ranks processed / max-over-ranks device time.

Keys of the JSON line (see DESIGN.md "Measurement"):
  value        device-resident throughput (inputs in HBM when the timed region starts)
  e2e          same metric through the host-buffer C-ABI call, H2D/D2H inside the timed region
  roofline     dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM peak
  cpu_baseline the oracle (CPU restatement of the reference schedule) on this box's host cores
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (log2 N, L, prime bits, batch per GPU, e2e batch per GPU)
    "cfg4": (16, 24, 61, 256, 128),   # BASELINE.json configs[3]: the metric's configuration (default)
    "cfg3": (14, 8, 30, 1024, 256),   # configs[2]: rotation (automorphism + key-switch), see --op
    "cfg2": (12, 3, 40, 4096, 1024),  # configs[1]
    "cfg1": (4, 4, 31, 1, 1),         # configs[0]: examples/encrypt_mul as shipped (N=16, generate_primes(31, 4, 16), one ciphertext)
    "tiny": (12, 3, 40, 8, 8),
}


def algorithmic_bytes_per_ctmult(n: int, l: int, batch: int) -> float:
    """read both input ciphertexts, write the rescaled one, key once per batch."""
    return 8.0 * n * (4 * l + 2 * (l - 1)) + 16.0 * l * l * n / batch


def modmuls_per_ctmult(n: int, logn: int, l: int) -> float:
    ntt = n / 2 * logn + n
    return (4 * l + l * (l - 1) + 3 * l) * ntt + 4 * l * n + 2 * l * l * n + 2 * (l - 1) * n


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    REASONS = {
        0x0000000000000008: "hw_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000004: "sw_power_cap",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        s = sorted(self.samples)
        return {
            "sm_mhz": (s[len(s) // 2] if s else None),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference(args):
    """The reference algorithm's CPU implementation (oracle port: the Rust crate cannot be built in
    this image) on this box's host cores.  Rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc

    logn, l, bits, batch, _ = CONFIGS[args.config]
    n = 1 << logn
    cores = os.cpu_count() or 1
    moduli = orc.generate_primes(bits, l, n)
    ob = orc.Basis(n, moduli)
    rng = np.random.default_rng(1234)
    q = np.array(moduli, dtype=np.uint64)
    count = max(1, cores // l)  # ciphertext pairs per step; each spreads its limbs over cores/count threads

    def uni(*lead):
        return (rng.integers(0, 1 << 63, size=(*lead, l, n), dtype=np.uint64) % q[:, None]).astype(np.uint64)

    a0, a1, b0, b1 = uni(count), uni(count), uni(count), uni(count)
    ka, kb = uni(l), uni(l)
    for _ in range(args.warmup):
        ob.bench_mul_rescale(cores, a0[:1], a1[:1], b0[:1], b1[:1], ka, kb)  # warm-up: one unit, all threads
    total = 0.0
    for _ in range(args.steps):
        sec, _, _ = ob.bench_mul_rescale(cores, a0, a1, b0, b1, ka, kb)
        total += sec
    value = count * args.steps / total
    sample = f"{count} ciphertext pair(s) per step at full size N={n}, L={l}; limbs of each spread over {max(1, cores // count)} threads"
    line = {
        "impl": "reference",
        "metric": "ct-mults/sec (mul+relin+rescale)",
        "value": value,
        "unit": "ct-mult/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: N=2^{logn}, L={l}, {bits}-bit primes, mul_ciphertexts_gadget+rescale_ciphertext", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "ct-mult/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ct-mult/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc

    return orc


def cpu_baseline(config, op, sample, single_thread=True):
    """The oracle (C++ restatement of the reference schedule) on this box's host cores, fed the SAME inputs as the
    first pairs of the timed GPU batch (`sample`: host copies), so that its outputs double as the parity check.
    Returns (record, outputs)."""
    orc = _oracle()
    logn, l, bits, _, _ = CONFIGS[config]
    n = 1 << logn
    cores = os.cpu_count() or 1
    ob = orc.Basis(n, orc.generate_primes(bits, l, n))
    ins, ka, kb = sample["inputs"], sample["ka"], sample["kb"]
    count = ins[0].shape[0]

    def run(threads, k):
        if op == "rotate":
            return ob.bench_rotate(threads, ins[0][:k], ins[1][:k], ka, kb, 1)
        return ob.bench_mul_rescale(threads, ins[0][:k], ins[1][:k], ins[2][:k], ins[3][:k], ka, kb)

    timed = max(1, count - 1)  # the last sample is the last pair of the batch: parity only, untimed
    sec, o0, o1 = run(cores, timed)
    outs = [(o0, o1)]
    if count > timed:
        if op == "rotate":
            _, p0, p1 = ob.bench_rotate(cores, ins[0][timed:], ins[1][timed:], ka, kb, 1)
        else:
            _, p0, p1 = ob.bench_mul_rescale(cores, ins[0][timed:], ins[1][timed:], ins[2][timed:], ins[3][timed:], ka, kb)
        outs.append((p0, p1))
    unit = "ct-mult/s" if op == "mul" else "rotation/s"
    rec = {
        "value": timed / sec,
        "unit": unit,
        "cores": cores,
        "kind": "port",
        "sample": f"{timed} ciphertext pair(s) of the timed GPU batch at full size N={n}, L={l} ({sec:.1f} s of wall time on {cores} threads: "
        "the limbs of each unit are spread over the threads); oracle = C++ restatement of the reference schedule (the Rust crate cannot "
        "be built here).  The reference itself is single-threaded: `single_thread` is the like-for-like figure, `value` flatters the CPU",
    }
    if single_thread:
        s1, _, _ = run(1, 1)
        rec["single_thread"] = {"value": 1.0 / s1, "unit": unit, "cores": 1, "sample": f"1 pair, {s1:.1f} s"}
    return rec, (np_cat([o[0] for o in outs]), np_cat([o[1] for o in outs]))


def np_cat(xs):
    import numpy as np

    return np.concatenate(xs, axis=0)


# Algorithmic modmuls per launch of the key-switch kernels (SURVEY 8d counts a limb transform as N/2 log2 N + N:
# at N = 2^16 that is 5 N in the first pass -- 256-point negacyclic transform + four-step twiddle -- and 4 N in the
# second), plus one modmul per key multiply-accumulate.
def _ks1_modmuls(n, logn, l, batch):
    a1 = (logn + 1) // 2
    return _cs(n, l, batch) * l * (l - 1) * n * (a1 / 2 + 1)


def _ks2_modmuls(n, logn, l, batch):
    a2 = logn - (logn + 1) // 2
    cs = _cs(n, l, batch)
    return cs * l * (l - 1) * n * (a2 / 2) + 2 * cs * l * l * n + 2 * cs * l * n * (a2 / 2 + 1)


def ntt_point(ck, torch, dev, stream, hbm_peak, bits, logn, l, steps, warmup):
    """One point of BASELINE.json configs[4] (the metric's second half): limb-batched forward / inverse transforms of
    ~1 GiB of limbs (> L2), CUDA events around each call, median; 16 N algorithmic bytes per limb transform."""
    n = 1 << logn
    moduli = ck.generate_primes(bits, l, n)
    basis = ck.RnsBasis(n, moduli, device=dev.index)
    basis.set_stream(stream.cuda_stream)
    batch = max(1, (1 << 30) // (l * n * 8))
    qt = torch.tensor(moduli, dtype=torch.int64, device=dev)[:, None]
    t = torch.randint(0, 1 << 62, (batch, l, n), dtype=torch.int64, device=dev) % qt
    h = ck._vp()
    ck._check(ck._lib.ckks_poly_from_device(basis._h, batch, ck.C.cast(t.data_ptr(), ck._u64p), 0, ck.C.byref(h)))
    p = ck.RnsPoly(h, basis)
    del t
    rec = {"N": n, "L": l, "bits": bits, "batch": batch, "word": "u32" if bits <= 31 else "u64"}
    for name in ("fwd", "inv"):
        fn = p.to_ntt_domain if name == "fwd" else p.to_coeff_domain
        other = p.to_coeff_domain if name == "fwd" else p.to_ntt_domain
        if name == "inv":
            p.to_ntt_domain()
        for _ in range(max(1, warmup)):
            fn()
            other()
        times = []
        for _ in range(max(steps, 5)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            other()
        ms = sorted(times)[len(times) // 2]
        tr = batch * l / (ms * 1e-3)
        rec[name] = {"ms": ms, "transforms_per_s": tr, "gbs": tr * 16.0 * n / 1e9, "hbm_frac": tr * 16.0 * n / 1e9 / hbm_peak}
        if name == "inv":
            p.to_coeff_domain()
    del p, basis
    torch.cuda.empty_cache()
    return rec


def run_chain(ck, torch, dev, stream, uni_poly_at, basis, bits, l, n, batch, passes=2):
    """BASELINE.json configs[3] as worded: the horner_chain workload (examples/horner_chain.rs:211-278), x <- x * alpha
    + beta from L limbs down to 2, on a RESIDENT batch: per level one mul_ciphertexts_gadget + rescale_ciphertext with
    that level's gadget key and one add_ciphertexts.  Ciphertexts and keys are synthetic uniform limbs (the reference
    regenerates keys per level on the host; key generation is not part of the timed hot path)."""
    levels = list(range(l, 2, -1))  # limb count before each multiplication: L .. 3 -> ends with 2 primes
    bases = {l: basis}
    for k in range(l - 1, 1, -1):
        bases[k] = bases[k + 1].drop_last(1)
    keys, alphas, betas = {}, {}, {}
    for k in levels:
        ka, kb = uni_poly_at(bases[k], k, k), uni_poly_at(bases[k], k, k)
        keys[k] = ck.GadgetKey.from_polys(ka, kb)
        del ka, kb
    alpha_top = ck.Ciphertext(uni_poly_at(basis, l, batch), uni_poly_at(basis, l, batch), bits, bits * l)
    beta_top = ck.Ciphertext(uni_poly_at(basis, l, batch), uni_poly_at(basis, l, batch), bits, bits * l)
    x0 = ck.Ciphertext(uni_poly_at(basis, l, batch), uni_poly_at(basis, l, batch), bits, bits * l)
    torch.cuda.empty_cache()

    def at_level(ct_top, k):  # the reference re-encrypts alpha / beta at every level (horner_chain.rs:211-250): untimed here
        if k == l:
            return ct_top
        return ck.Ciphertext(ct_top.c0.mod_drop_last(basis=bases[k]), ct_top.c1.mod_drop_last(basis=bases[k]), bits, bits * k)

    per_level = {k: [] for k in levels}
    totals = []
    alloc = []
    WARM = 2  # warm-up passes: the arena reaches its steady-state footprint in the second one (one more 2 GiB segment)
    for it in range(passes + WARM):
        ck.alloc_stats(reset=True)
        ct = x0
        total = 0.0
        for k in levels:
            alpha_k, beta_k = at_level(alpha_top, k), at_level(beta_top, k - 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            ct = ck.CkksEngine.mul_relin_rescale(ct, alpha_k, keys[k], bases[k - 1])
            ct.logp = bits
            ct = ck.CkksEngine.add_ciphertexts(ct, beta_k)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            total += ms
            if it >= WARM:
                per_level[k].append(ms)
            del alpha_k, beta_k
        assert ct.c0.channel_count() == 2
        alloc.append(ck.alloc_stats())
        if it >= WARM:
            totals.append(total)
        del ct
    ms_chain = sum(totals) / len(totals)
    ms_best = sum(min(v) for v in per_level.values())
    return {
        "allocator_per_pass": alloc,
        "ms_per_chain_best_levels": ms_best,
        "ct_mults_per_s_best_levels": batch * len(levels) / (ms_best * 1e-3),
        "workload": f"horner_chain x <- x*alpha + beta, N={n}, L={l} -> 2 ({len(levels)} levels), resident batch of {batch} ciphertexts, "
        "per level: mul_ciphertexts_gadget + rescale_ciphertext + add_ciphertexts (CUDA events around each level; the level's alpha / beta "
        "ciphertexts are produced between the timed levels, as the reference re-encrypts them per level)",
        "ms_per_chain": ms_chain,
        "chains_per_s": batch / (ms_chain * 1e-3),
        "ct_mults_per_s": batch * len(levels) / (ms_chain * 1e-3),
        "per_level": {str(k): {"ms": sum(v) / len(v), "ms_min": min(v), "ct_mults_per_s": batch / (sum(v) / len(v) * 1e-3)} for k, v in per_level.items()},
    }


def run_extra_ops(ck, torch, dev, stream, basis, moduli, bits, l, n, batch, steps, hbm_peak):
    """BASELINE.json configs[1] as worded: "encrypt + homomorphic add + mul/relin/rescale".  encrypt (engine.rs:84-112,
    host-sampled u / e0 / e1 / m already resident, as north_star keeps sampling on the host) and add_ciphertexts
    (engine.rs:131-151) on a resident batch; CUDA events, median of `steps` calls."""
    qt = torch.tensor(moduli, dtype=torch.int64, device=dev)[:, None]

    def small_poly(nb, lo, hi):  # coefficients in [lo, hi) as canonical residues per limb (what from_coeffs produces)
        c = torch.randint(lo, hi, (nb, 1, n), dtype=torch.int64, device=dev)
        t = torch.where(c < 0, c + qt, c.expand(nb, l, n)).contiguous()
        h = ck._vp()
        ck._check(ck._lib.ckks_poly_from_device(basis._h, nb, ck.C.cast(t.data_ptr(), ck._u64p), 0, ck.C.byref(h)))
        return ck.RnsPoly(h, basis)

    def uni(nb):
        t = torch.randint(0, 1 << 62, (nb, l, n), dtype=torch.int64, device=dev) % qt
        h = ck._vp()
        ck._check(ck._lib.ckks_poly_from_device(basis._h, nb, ck.C.cast(t.data_ptr(), ck._u64p), 0, ck.C.byref(h)))
        return ck.RnsPoly(h, basis)

    pk_b, pk_a = uni(1), uni(1)
    u, e0, e1 = small_poly(batch, -1, 2), small_poly(batch, -16, 17), small_poly(batch, -16, 17)
    m = small_poly(batch, -(1 << 28), 1 << 28)  # |m| below every modulus of every config
    x = ck.Ciphertext(uni(batch), uni(batch), bits, bits * l)
    y = ck.Ciphertext(uni(batch), uni(batch), bits, bits * l)

    def timed(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(max(steps, 5)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    ms_enc = timed(lambda: ck.CkksEngine.encrypt(pk_b, pk_a, u, e0, e1, m, bits, bits * l))
    ms_add = timed(lambda: ck.CkksEngine.add_ciphertexts(x, y))
    poly_bytes = 8.0 * n * l
    return {
        "encrypt": {"ms": ms_enc, "per_s": batch / (ms_enc * 1e-3), "unit": "encryption/s",
                    "algorithmic_bytes": 6 * poly_bytes, "hbm_frac": batch / (ms_enc * 1e-3) * 6 * poly_bytes / (hbm_peak * 1e9),
                    "what": "c0 = pk_b*u + e0 + m, c1 = pk_a*u + e1 with u, e0, e1, m resident (read 4 polynomials, write 2; pk is shared)"},
        "add_ciphertexts": {"ms": ms_add, "per_s": batch / (ms_add * 1e-3), "unit": "ct-add/s",
                            "algorithmic_bytes": 6 * poly_bytes, "hbm_frac": batch / (ms_add * 1e-3) * 6 * poly_bytes / (hbm_peak * 1e9),
                            "what": "limb-wise add of c0, c1 (read 4 polynomials, write 2)"},
        "batch": batch,
    }


def run_b200(args):
    import numpy as np
    import torch

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    try:  # run (and first-touch the pinned staging buffers) on the CPUs/NUMA node closest to this GPU
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g

    if not os.environ.get("CKKS_B200_LIB"):
        g.build_cuda()
    ck = importlib.import_module("toy-heaan-ckks_b200")
    if args.host_chunk_mib:
        ck._check(ck._lib.ckks_set_host_chunk_mib(args.host_chunk_mib))
    if args.ks_scratch_mib:
        global KS_SCRATCH_MIB
        KS_SCRATCH_MIB = args.ks_scratch_mib
        ck._check(ck._lib.ckks_set_ks_scratch_mib(args.ks_scratch_mib))
    if args.ks_aux is not None:
        ck.set_ks_aux(args.ks_aux)
    logn, l, bits, batch, e2e_batch = CONFIGS[args.config]
    if args.batch:
        batch = args.batch
    if args.e2e_batch:
        e2e_batch = args.e2e_batch
    n = 1 << logn
    moduli = ck.generate_primes(bits, l, n)
    basis = ck.RnsBasis(n, moduli, device=local)
    child = basis.drop_last(1)
    stream = torch.cuda.current_stream()
    basis.set_stream(stream.cuda_stream)
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    op = args.op or ("rotate" if args.config == "cfg3" else "mul")
    n_in = 4 if op == "mul" else 2
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    cores = os.cpu_count() or 1
    # pairs whose host copies feed the oracle (timing sample + parity): the first `count` and the last of the batch
    count = max(1, cores // l)
    keep_idx = sorted(set(list(range(min(count, batch))) + [batch - 1])) if want_cpu else []

    def uni_poly_at(b, limbs, nb, keep=None):
        qt = torch.tensor(moduli[:limbs], dtype=torch.int64, device=dev)[:, None]
        t = torch.randint(0, 1 << 62, (nb, limbs, n), dtype=torch.int64, device=dev, generator=gen) % qt
        h = ck._vp()
        ck._check(ck._lib.ckks_poly_from_device(b._h, nb, ck.C.cast(t.data_ptr(), ck._u64p), 0, ck.C.byref(h)))
        p = ck.RnsPoly(h, b)
        if keep is not None:
            keep.append(t[keep_idx].cpu().numpy().view(np.uint64) if keep_idx and nb == batch else t.cpu().numpy().view(np.uint64))
        del t
        return p

    def uni_poly(nb, keep=None):
        return uni_poly_at(basis, l, nb, keep)

    host_keys, host_in = [], []
    ka, kb = uni_poly(l, host_keys if want_cpu else None), uni_poly(l, host_keys if want_cpu else None)
    rlk = ck.GadgetKey.from_polys(ka, kb)
    del ka, kb
    polys = [uni_poly(batch, host_in if want_cpu else None) for _ in range(4)]
    cta = ck.Ciphertext(polys[0], polys[1], bits, bits * l)
    ctb = ck.Ciphertext(polys[2], polys[3], bits, bits * l)
    del polys
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shard = importlib.import_module("toy-heaan-ckks_b200.shard")  # the package's batch-sharding helpers

    def max_over_ranks(ms: float) -> float:
        return shard.max_over_ranks(ms, device=dev)

    rlk.rotation = 1

    def step():
        if op == "rotate":
            return ck.CkksEngine.rotate_ciphertext(cta, rlk)
        return ck.CkksEngine.mul_relin_rescale(cta, ctb, rlk, child)

    out = None
    for _ in range(args.warmup):
        out = step()
    barrier()
    # ---- timed region: K steps, profiler hooks OFF, CUDA events on the library's stream ------------
    sampler = ClockSampler(local)
    ck._lib.ckks_prof_enable(0)
    launches0 = ck.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        out = step()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ck.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    value = shard.aggregate_rate(batch, world, ms_per_step)

    # device words of the pairs the oracle will recompute (taken from the last timed step's output)
    dev_rows = None
    if want_cpu and out is not None:
        outL = l - 1 if op == "mul" else l
        rows = []
        for comp in (out.c0, out.c1):
            ptr = ck.C.POINTER(ck.C.c_uint64)()
            ck._check(ck._lib.ckks_poly_device_ptr(comp._h, ck.C.byref(ptr)))
            addr = ck.C.cast(ptr, ck.C.c_void_p).value
            tt = torch.as_tensor(_DevBuf(addr, batch * outL * n), device=dev).view(batch, outL, n)
            rows.append(tt[keep_idx].cpu().numpy().view(np.uint64))
        dev_rows = rows
    del out

    # ---- per-kernel breakdown: a SEPARATE, untimed pass with every launch bracketed by CUDA events ----
    prof = {}
    if args.prof:
        ck._lib.ckks_prof_enable(1)
        for _ in range(max(1, min(args.steps, 3))):
            step()
        torch.cuda.synchronize()
        ck._lib.ckks_prof_enable(0)
        buf = ck.C.create_string_buffer(1 << 20)  # collect() drains the records: one call
        ck._lib.ckks_prof_collect(buf, len(buf))
        for ln in buf.value.decode().splitlines():
            name, rest = ln.split("=")
            cnt, ms = rest.split(",")
            prof[name] = (int(cnt), float(ms))

    # ---- end to end through the host-buffer C-ABI call -------------------------------------------
    while True:  # page-locked staging buffers; halve the e2e batch if the host cannot pin that much
        wi, wo = (e2e_batch, l, n), (e2e_batch, l - 1 if op == "mul" else l, n)
        try:
            hin = [ck.PinnedBuffer(wi) for _ in range(n_in)]
            hout = [ck.PinnedBuffer(wo) for _ in range(2)]
            break
        except ck.RnsNttError:
            hin = hout = None
            if e2e_batch <= 1:
                raise
            e2e_batch //= 2
    e2e_polys = [uni_poly(e2e_batch) for _ in range(n_in)]
    for hb, p in zip(hin, e2e_polys):  # host copies of the e2e inputs (outside the timed region)
        ck._check(ck._lib.ckks_poly_download(p._h, ck._ptr(hb.array)))
    h2d = n_in * hin[0].array.nbytes
    d2h = 2 * hout[0].array.nbytes

    def e2e_step():
        if op == "rotate":
            ck.rotate_host(basis, rlk, hin[0].array, hin[1].array, hout[0].array, hout[1].array)
        else:
            ck.mul_relin_rescale_host(basis, child, rlk, hin[0].array, hin[1].array, hin[2].array, hin[3].array, hout[0].array, hout[1].array)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        e2e_step()  # blocks until the results are in host memory
    ev1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)) / args.steps
    e2e_value = world * e2e_batch / (e2e_ms * 1e-3)

    # the host-buffer path must agree with the device-resident path on the same inputs
    if op == "rotate":
        ref = ck.CkksEngine.rotate_ciphertext(ck.Ciphertext(e2e_polys[0], e2e_polys[1], bits, bits * l), rlk)
    else:
        ref = ck.CkksEngine.mul_relin_rescale(ck.Ciphertext(e2e_polys[0], e2e_polys[1], bits, bits * l), ck.Ciphertext(e2e_polys[2], e2e_polys[3], bits, bits * l), rlk, child)
    assert np.array_equal(ref.c0.channels(), hout[0].array) and np.array_equal(ref.c1.channels(), hout[1].array), "host-buffer path and device path disagree"
    del ref, e2e_polys

    # ---- what the host's copy path alone sustains: the same bytes as one e2e step, H2D and D2H concurrently on two
    # streams in the pipeline's chunk size, all ranks at once (max over ranks) -> the ceiling of `e2e` on this box
    barrier()
    sec = float(ck._lib.ckks_bench_host_copy(local, ck._ptr(hin[0].array), hin[0].array.nbytes, n_in, ck._ptr(hout[0].array), hout[0].array.nbytes, 2, 3))
    copy_ms = max_over_ranks(sec * 1e3)
    barrier()
    copy_probe = {
        "ms_per_step_bytes": copy_ms,
        "h2d_gbs_per_gpu": h2d / (copy_ms * 1e-3) / 1e9,
        "d2h_gbs_per_gpu": d2h / (copy_ms * 1e-3) / 1e9,
        "aggregate_gbs": world * (h2d + d2h) / (copy_ms * 1e-3) / 1e9,
        "e2e_ceiling": world * e2e_batch / (copy_ms * 1e-3),
        "what": f"plain cudaMemcpyAsync of one e2e step's bytes ({h2d} B in, {d2h} B out per GPU), both directions concurrently, "
        f"{world} rank(s) at once, no kernels: the rate the host side alone allows",
    }
    # ---- N > 1: the same e2e metric through the SINGLE-PROCESS batch-sharded group (ckks_comm_*): rank 0 alone drives
    # every GPU of the job from one host batch (one pipeline thread per device); the other ranks wait at the barrier.
    e2e_single = None
    if world > 1 and op == "mul" and not args.no_e2e_single:
        barrier()
        if rank == 0:
            try:
                per_dev = max(1, min(e2e_batch, 32))
                tot = per_dev * world
                grp = ck.BatchShard(n, moduli, list(range(world)))
                rngk = np.random.default_rng(5)
                qn = np.array(moduli, dtype=np.uint64)[None, :, None]
                gka = rngk.integers(0, 1 << 62, size=(l, l, n), dtype=np.uint64) % qn
                gkb = rngk.integers(0, 1 << 62, size=(l, l, n), dtype=np.uint64) % qn
                gkey = grp.upload_key(gka, gkb)
                del gka, gkb
                gin = [ck.PinnedBuffer((tot, l, n)) for _ in range(4)]
                gout = [ck.PinnedBuffer((tot, l - 1, n)) for _ in range(2)]
                for gb_ in gin:  # replicate the per-rank e2e inputs: canonical words
                    for r0 in range(0, tot, e2e_batch):
                        m = min(e2e_batch, tot - r0)
                        gb_.array[r0 : r0 + m] = hin[gin.index(gb_)].array[:m]
                grp.mul_relin_rescale_host(gkey, gin[0].array, gin[1].array, gin[2].array, gin[3].array, gout[0].array, gout[1].array)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    grp.mul_relin_rescale_host(gkey, gin[0].array, gin[1].array, gin[2].array, gin[3].array, gout[0].array, gout[1].array)
                dt = (time.perf_counter() - t0) / args.steps
                e2e_single = {"value": tot / dt, "unit": "ct-mult/s", "devices": world, "batch": tot, "ms_per_step": dt * 1e3,
                              "h2d_bytes_per_step": 4 * gin[0].array.nbytes, "d2h_bytes_per_step": 2 * gout[0].array.nbytes,
                              "what": "ckks_comm_ct_mul_relin_rescale_host: ONE process, one host batch, every GPU of the job (host wall clock; the call blocks until the results are in host memory)"}
                del gkey, grp, gin, gout
            except Exception as exc:  # never lose the main line to the optional leg
                e2e_single = {"error": repr(exc)[:300]}
        barrier()
    del hin, hout

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_ct = algorithmic_bytes_per_ctmult(n, l, batch)
    mm_ct = modmuls_per_ctmult(n, logn, l)
    imad_peak = ck.modmul_peak(local, 4096)  # measured in this run: dependent-free exact Shoup modmuls
    imad_peak = max_over_ranks(imad_peak) if world > 1 else imad_peak
    line = {
        "metric": "ct-mults/sec (mul+relin+rescale)" if op == "mul" else "rotations/sec (automorphism + gadget key-switch)",
        "value": value,
        "unit": "ct-mult/s" if op == "mul" else "rotation/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {
            "workload": f"{args.config}: N=2^{logn}, L={l}, {bits}-bit primes, batch {batch} ciphertext pairs per GPU, "
            + ("mul_ciphertexts_gadget+rescale_ciphertext" if op == "mul" else "rotate_ciphertext(k=1)") + ", coefficient-domain in and out",
            "batch_per_gpu": batch,
            "e2e_batch_per_gpu": e2e_batch,
            "parallelism": f"batch-sharded x{world}, no data-path collective",
            "l2": f"inputs are {n_in * batch * l * n * 8 / 2**30:.1f} GiB per step (> 126 MB L2); no flush needed",
            "timing": "CUDA events on the library's stream, per-kernel event hooks off in the timed region (the kernel breakdown is a separate pass)",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "ct-mult/s" if op == "mul" else "rotation/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "host_copy_probe": copy_probe, "single_process_group": e2e_single},
        "gpu_launches": launches,
        "imad": {
            "peak_modmul_per_s": imad_peak,
            "peak_kind": "measured in this run (ckks_bench_modmul_peak: dependent-free exact 64-bit Shoup multiplies on every SM)",
        },
    }
    if op == "mul":
        rate = value / world
        line["ctmult"] = {
            "algorithmic_bytes": bytes_ct,
            "hbm_frac": rate * bytes_ct / (hbm_peak * 1e9),
            "modmuls": mm_ct,
            "modmul_rate": rate * mm_ct,
            "imad_frac": (rate * mm_ct / imad_peak) if imad_peak else None,
            "binding_roof": "imad",
            "note": "modmuls = the reference algorithm's count (SURVEY 8d); with the auxiliary-basis gadget product the library executes fewer, "
                    "cheaper multiplies for the same result, so imad_frac (reference-algorithm modmul rate / measured 64-bit modmul peak) may exceed 1",
        }
        line["imad"]["step_frac"] = line["ctmult"]["imad_frac"]
    if args.prof and prof:
        tot = sum(ms for _, ms in prof.values())
        top = max(prof.items(), key=lambda kv: kv[1][1])
        name, (cnt, ms) = top
        # algorithmic bytes per launch of the dominant kernel (DESIGN.md "Kernels"): an NTT pass is
        # half of a transform (16 N bytes per limb transform) -> 8 N per limb per pass;
        # elementwise kernels: the words they must read and write once.
        aux = "aux_mac" in prof  # the gadget product ran through the auxiliary 30-bit primes (csrc/aux_ks.cuh)
        global AUX_K, KS_WORD, KS_MUL
        KS_MUL = op == "mul"
        KS_WORD = 4 if bits <= 31 and logn >= 8 else 8  # the 32-bit word path (all q < 2^31, four-step transforms)
        AUX_K = _aux_k(logn, l, bits)
        table = KERNEL_BYTES_AUX if aux else KERNEL_BYTES
        per_launch = table.get(name, lambda **kw: None)(n=n, l=l, batch=batch)
        dur_s = ms * 1e-3 / cnt
        ach = per_launch / dur_s / 1e9 if per_launch else None
        kmm = {}
        if aux:
            mac_peak = ck.mac32_peak(local, 4096)
            c, m = prof["aux_mac"]
            macs = 2.0 * _cs_aux(n, l, batch, AUX_K) * AUX_K * l * l * n
            kmm["aux_mac"] = {"mac32_per_launch": macs, "mac32_per_s": macs / (m * 1e-3 / c), "mac32_peak_per_s": mac_peak,
                              "imad_frac": macs / (m * 1e-3 / c) / mac_peak if mac_peak else None,
                              "peak_kind": "measured in this run (ckks_bench_mac32_peak: 32x32->64-bit multiply-accumulates, IMAD.WIDE.U32, every SM)"}
            line["config"]["gadget_product"] = f"auxiliary basis: {AUX_K} NTT primes below 2^30, exact integer convolution + Garner CRT (csrc/aux_ks.cuh)"
        elif op == "mul" and logn >= 8:
            for kname, fn in (("ks_pass1", _ks1_modmuls), ("ks_pass2_tma", _ks2_modmuls), ("ks_pass2", _ks2_modmuls)):
                if kname in prof and imad_peak:
                    c, m = prof[kname]
                    kmm[kname] = {"modmuls_per_launch": fn(n, logn, l, batch), "modmul_per_s": fn(n, logn, l, batch) / (m * 1e-3 / c),
                                  "imad_frac": fn(n, logn, l, batch) / (m * 1e-3 / c) / imad_peak}
        line["imad"]["kernels"] = kmm
        traffic, traffic_src = ncu_traffic(name, args.config, op, batch)
        top_imad = kmm.get(name, {}).get("imad_frac")
        hbm_frac = (ach / hbm_peak if ach else None)
        line["roofline"] = {
            "kernel": name,
            # the binding roof of the dominant kernel: the integer (IMAD) pipe for the key-switch kernels, HBM otherwise
            "bound": "imad" if (top_imad is not None and hbm_frac is not None and top_imad > hbm_frac) else "hbm",
            "achieved": ach,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": hbm_frac,
            "hbm_frac": hbm_frac,
            "imad_frac": top_imad,
            "imad_achieved_modmul_per_s": kmm.get(name, {}).get("modmul_per_s"),
            "imad_achieved_mac32_per_s": kmm.get(name, {}).get("mac32_per_s"),
            "imad_peak_mac32_per_s": kmm.get(name, {}).get("mac32_peak_per_s"),
            "imad_peak_modmul_per_s": imad_peak,
            "traffic": traffic,
            "traffic_source": traffic_src,
            "peak_kind": peak_kind,
            "share_of_step": ms / tot,
            "avg_launch_ms": ms / cnt,
            "launches": cnt,
            "algorithmic_bytes_per_launch": per_launch,
        }
        line["kernels"] = {k: {"launches": c, "ms": round(m, 3), "share": round(m / tot, 4)} for k, (c, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    # ---- the metric's second half: limb NTT achieved HBM GB/s vs peak, at the metric's shape ----------
    if rank == 0 and world == 1 and not args.no_ntt and logn >= 12:
        del cta, ctb
        torch.cuda.empty_cache()
        line["ntt"] = {
            "what": "standalone limb-batched to_ntt_domain / to_coeff_domain, ~1 GiB of limbs per call, 16 N algorithmic bytes per limb transform, "
            f"fractions of the {peak_kind} HBM peak {hbm_peak} GB/s",
            "points": [ntt_point(ck, torch, dev, stream, hbm_peak, b_, logn, l, args.steps, 2) for b_ in (61, 30)],
        }
        if not args.no_chain and op == "mul" and world == 1 and args.config == "cfg4":
            line["chain"] = run_chain(ck, torch, dev, stream, uni_poly_at, basis, bits, l, n, min(batch, 256))
    if rank == 0 and (args.ops or args.config == "cfg2"):
        line["ops"] = run_extra_ops(ck, torch, dev, stream, basis, moduli, bits, l, n, batch, args.steps, hbm_peak)
    if want_cpu:
        sample = {"inputs": host_in, "ka": host_keys[0], "kb": host_keys[1]}
        rec, (o0, o1) = cpu_baseline(args.config, op, sample, single_thread=not args.no_single_thread)
        line["cpu_baseline"] = rec
        equal = bool(np.array_equal(o0, dev_rows[0]) and np.array_equal(o1, dev_rows[1]))
        line["parity"] = {"checked_pairs": len(keep_idx), "pair_indices": keep_idx, "equal": equal,
                          "what": "every output word of these pairs of the TIMED batch (last timed step) against the oracle on the same inputs"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if want_cpu and not line["parity"]["equal"]:
        raise SystemExit("bench.py: PARITY FAILURE -- the GPU output differs from the oracle on the timed batch")


def ncu_traffic(kernel, config, op, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, read from the committed `ncu --set full`
    summary (profiles/*.json, newest round first) of the same configuration; (None, reason) if there is none."""
    if not (config == "cfg4" and op == "mul" and batch >= 64 and KS_SCRATCH_MIB == 8192):
        return None, "no capture for this configuration"
    import glob

    want = {"ks_pass2_tma": "ks_pass2_kernel", "ks_pass2": "ks_pass2_kernel", "ks_pass1": "ks_pass1_kernel", "aux_mac": "aux_mac_kernel",
            "aux_crt": "aux_crt_kernel"}.get(kernel)
    if not want:
        return None, "no capture of this kernel"
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_ks_kernels*final*.json")), reverse=True):
        try:
            with open(path) as f:
                rows = json.load(f)
            for r in rows:
                if want in r.get("Kernel Name", ""):
                    def gb(v):
                        x, u = v.split()
                        return float(x) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                    return gb(r["dram__bytes_read.sum"]) + gb(r["dram__bytes_write.sum"]), os.path.relpath(path, ROOT) + " (grid " + r.get("Grid Size", "?") + ")"
        except Exception:
            continue
    return None, "no capture found under profiles/"


class _DevBuf:
    """A library-owned device buffer seen by torch (for the NCCL variant of the limb-sharded mode)."""

    def __init__(self, ptr: int, words: int):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def run_limb_sharded(args):
    """Optional limb-sharded mode: ONE batch whose limbs are spread over the GPUs (limb j on
    GPU j mod N), strong scaling.  --comm peer: digits all-gathered / dropped limb broadcast by stores into
    peer HBM from the producing kernels + flag barrier (one library call per step);  --comm nccl: the same
    phases with torch.distributed all_gather_into_tensor / broadcast on the exchange buffers."""
    import numpy as np
    import torch

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g

    g.build_cuda()
    ck = importlib.import_module("toy-heaan-ckks_b200")
    logn, l, bits, batch, _ = CONFIGS[args.config]
    if args.batch:
        batch = args.batch
    n = 1 << logn
    moduli = ck.generate_primes(bits, l, n)
    sh = ck.LimbShard(n, moduli, rank, world, device=local, chunk=args.ls_chunk)
    kid = sh.drop_last()
    if args.comm == "ce":
        sh.set_exchange(1)
    basis = sh.local_basis()
    stream = torch.cuda.current_stream()
    basis.set_stream(stream.cuda_stream)
    if world > 1:
        sh.connect_process_group()
    dev = torch.device("cuda", local)
    own = sh.owned()
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)  # every rank holds limbs of the SAME ciphertexts; the values are synthetic anyway
    qt = torch.tensor([moduli[j] for j in own], dtype=torch.int64, device=dev)[:, None]

    def uni_poly(b):
        t = torch.randint(0, 1 << 62, (b, len(own), n), dtype=torch.int64, device=dev, generator=gen) % qt
        h = ck._vp()
        ck._check(ck._lib.ckks_poly_from_device(basis._h, b, ck.C.cast(t.data_ptr(), ck._u64p), 0, ck.C.byref(h)))
        p = ck.RnsPoly(h, basis)
        del t
        return p

    rng = np.random.default_rng(7 + rank)
    qn = np.array([moduli[j] for j in own], dtype=np.uint64)[None, :, None]
    ka = rng.integers(0, 1 << 62, size=(l, len(own), n), dtype=np.uint64) % qn
    kb = rng.integers(0, 1 << 62, size=(l, len(own), n), dtype=np.uint64) % qn
    key = sh.upload_key(ka, kb)
    del ka, kb
    cta = ck.Ciphertext(uni_poly(batch), uni_poly(batch), bits, bits * l)
    ctb = ck.Ciphertext(uni_poly(batch), uni_poly(batch), bits, bits * l)
    torch.cuda.empty_cache()
    chunk = sh.chunk()
    if args.comm == "nccl" and world > 1:
        gp, gw, lp, lw = sh.buffers()
        gt = torch.as_tensor(_DevBuf(gp, gw), device=dev).view(l, chunk * n)
        lt = torch.as_tensor(_DevBuf(lp, lw), device=dev)
        owner = (l - 1) % world

    def step():
        if args.comm != "nccl" or world == 1:
            return sh.mul_relin_rescale(cta, ctb, key, kid)
        out = ck.Ciphertext(ck.RnsPoly.zero(kid.local_basis(), batch), ck.RnsPoly.zero(kid.local_basis(), batch), 0, 0)
        for s0 in range(0, batch, chunk):
            cs = min(chunk, batch - s0)
            sh.mul_phase(0, s0, cs, cta, ctb, key, kid, out, False)
            for k0 in range(0, l, world):  # slots k0..k0+world-1 are owned by ranks 0..world-1
                if k0 + world <= l:
                    dist.all_gather_into_tensor(gt[k0 : k0 + world].view(-1), gt[k0 + rank])
                else:
                    for i in range(k0, l):
                        dist.broadcast(gt[i], src=i % world)
            sh.mul_phase(1, s0, cs, cta, ctb, key, kid, out, False)
            dist.broadcast(lt, src=owner)
            sh.mul_phase(2, s0, cs, cta, ctb, key, kid, out, False)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sh.check()
    sampler = ClockSampler(local)
    ck._lib.ckks_prof_enable(1 if args.prof else 0)
    launches0 = ck.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    sh.check()
    launches = ck.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ck._lib.ckks_prof_enable(0)
    prof = {}
    if args.prof:
        buf = ck.C.create_string_buffer(1 << 20)
        ck._lib.ckks_prof_collect(buf, len(buf))
        for ln in buf.value.decode().splitlines():
            name, rest = ln.split("=")
            cnt, kms = rest.split(",")
            prof[name] = (int(cnt), float(kms))
    ms_per_step = ms / args.steps
    value = batch / (ms_per_step * 1e-3)
    tot = sum(m for _, m in prof.values()) or 1.0
    line = {
        "metric": "ct-mults/sec (mul+relin+rescale)",
        "value": value,
        "unit": "ct-mult/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {
            "workload": f"{args.config}: N=2^{logn}, L={l}, {bits}-bit primes, ONE batch of {batch} ciphertext pairs, limbs spread over "
            f"{world} GPU(s), mul_ciphertexts_gadget+rescale_ciphertext",
            "parallelism": f"limb-sharded x{world} (limb j on GPU j mod {world}); exchange: "
            + {"peer": "stores into peer HBM from the producing kernels + flag barrier", "ce": "digits pushed by the copy engines, dropped limb by peer stores, flag barrier",
               "nccl": "NCCL all-gather / broadcast between the phases"}[args.comm],
            "chunk": chunk,
            "exchange_bytes_per_ct_per_gpu": (len(own) * (world - 1) * n * 8) + (2 * (world - 1) * n * 8 if rank == (l - 1) % world else 0),
            "key_bytes_per_gpu": 2 * l * len(own) * n * 8,
        },
        "clocks": clocks,
        "gpu_launches": launches,
        "kernels": {k: {"launches": c, "ms": round(m, 3), "share": round(m / tot, 4)} for k, (c, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])},
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    del cta, ctb, key
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ntt_sweep(args):
    """BASELINE.json configs[4]: standalone limb-batched NTT / INTT, N = 2^12..2^16, L in {1, 8, 24, 32},
    30-bit and 61-bit chains; batch sized to ~1 GiB of limbs.  transforms/s and achieved GB/s
    (algorithmic 16 N bytes per limb transform) against the measured HBM peak."""
    import torch

    import __graft_entry__ as g

    g.build_cuda()
    ck = importlib.import_module("toy-heaan-ckks_b200")
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rows = []
    for bits in (30, 61):
        for logn in (12, 13, 14, 15, 16):
            for l in (1, 8, 24, 32):
                n = 1 << logn
                try:
                    moduli = ck.generate_primes(bits, l, n)
                except ck.RnsNttError:
                    continue
                basis = ck.RnsBasis(n, moduli)
                basis.set_stream(stream.cuda_stream)
                batch = max(1, (1 << 30) // (l * n * 8))
                qt = torch.tensor(moduli, dtype=torch.int64, device=dev)[:, None]
                t = torch.randint(0, 1 << 62, (batch, l, n), dtype=torch.int64, device=dev) % qt
                h = ck._vp()
                ck._check(ck._lib.ckks_poly_from_device(basis._h, batch, ck.C.cast(t.data_ptr(), ck._u64p), 0, ck.C.byref(h)))
                p = ck.RnsPoly(h, basis)
                del t
                res = {}
                for name in ("fwd", "inv"):
                    fn = p.to_ntt_domain if name == "fwd" else p.to_coeff_domain
                    other = p.to_coeff_domain if name == "fwd" else p.to_ntt_domain
                    if name == "inv":
                        p.to_ntt_domain()
                    for _ in range(args.warmup):
                        fn()
                        other()
                    times = []
                    for _ in range(max(args.steps, 3)):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        fn()
                        e1.record(stream)
                        torch.cuda.synchronize()
                        times.append(e0.elapsed_time(e1))
                        other()
                    ms = sorted(times)[len(times) // 2]
                    tr = batch * l / (ms * 1e-3)
                    res[name] = {"ms": ms, "transforms_per_s": tr, "gbs": tr * 16.0 * n / 1e9, "frac": tr * 16.0 * n / 1e9 / hbm_peak}
                    if name == "inv":
                        p.to_coeff_domain()
                rows.append({"bits": bits, "logn": logn, "L": l, "batch": batch, **{k + "_" + kk: vv for k, v in res.items() for kk, vv in v.items()}})
                del p, basis
                torch.cuda.empty_cache()
    best = max(rows, key=lambda r: r["fwd_gbs"])
    print(json.dumps({"metric": "limb NTT achieved HBM GB/s (16 N bytes per transform)", "value": best["fwd_gbs"], "unit": "GB/s", "n_gpus": 1,
                      "peak": hbm_peak, "peak_kind": peak_kind, "frac": best["fwd_frac"], "best": best, "sweep": rows,
                      "config": {"workload": "NTT/INTT sweep N=2^12..2^16, L in {1,8,24,32}, 30- and 61-bit chains, ~1 GiB of limbs per point (> L2)"}}), flush=True)


# Algorithmic bytes per launch of each kernel as launched by the batched ct-mult.
# The fused pipeline works on chunks of cs = min(batch, 4 GiB / (L^2 N 8)) ciphertexts (ks_chunk in
# csrc/ckks_b200.cu); w = 8-byte words (61-bit chain; 4 for the internal scratch of the 32-bit path).
KS_SCRATCH_MIB = 8192  # the library's default (ckks_set_ks_scratch_mib); --ks-scratch-mib changes both


KS_MUL = True  # ct-mult (NTT-domain digit limb and d0 / d1 are read by ks_pass2) or rotation (they are not)
KS_WORD = 8  # bytes per word of the key-switch scratch and of the resident key: 4 on the 32-bit word path (set in run_b200)


def _cs(n, l, batch):
    c = max(1, min(batch, (KS_SCRATCH_MIB << 20) // (l * l * n * 8)))
    return min(c, 512) if KS_WORD == 4 else c  # the library caps 32-bit-word contexts at 512 ciphertexts per launch (ks_chunk)


def _ks2(n, l, batch, w=None):
    w = KS_WORD if w is None else w
    cs = _cs(n, l, batch)
    return cs * l * (l - 1 if KS_MUL else l) * n * w + 2 * l * l * n * w + (cs * l * n * 8 * 3 if KS_MUL else 0) + 2 * cs * l * n * w


def _aux_k(logn, l, bits):
    """Auxiliary 30-bit primes of the exact multi-modular gadget product (csrc/aux_ks.inl aux_get): enough for
    P > 2 * L * N * q_max^2."""
    need = max(0, (l - 1).bit_length()) + logn + 2 * bits + 1.5
    k = 1
    while k * 29.49 < need:  # the auxiliary primes lie just below 2^29.5 (csrc/aux_crt.cuh AUX_P_BOUND)
        k += 1
    return k


def _cs_aux(n, l, batch, k):
    return max(1, min(batch, (KS_SCRATCH_MIB << 20) // (4 * k * l * n * 4)))  # x, rb, ra and the transposed intermediate


AUX_K = 5  # set from the configuration in run_b200


def _aux_bytes(what):
    def f(n, l, batch):
        k = AUX_K
        cs = _cs_aux(n, l, batch, k)
        if what == "mac":  # x once, both key halves once per launch, both sums out
            return 4.0 * n * (cs * k * l + 2 * k * l * l + 2 * cs * l * k)
        if what == "crt":  # two launches per chunk (last limb, then the others): the average of the two
            return (4.0 * n * 2 * cs * l * k + 8.0 * n * 2 * cs * l + 8.0 * n * 2 * cs * (l - 1) + 8.0 * n * 4 * cs) / 2
        if what == "pass":  # one 32-bit pass over cs * l * k rows: half of a transform's read-once + write-once (SURVEY 8d convention)
            return 4.0 * n * cs * l * k
        if what == "ks1":  # digits in (u64, once: the K-1 re-reads are L2 hits), first-pass rows out (u32)
            return 8.0 * n * cs * l + 4.0 * n * cs * l * k
        if what == "u64pass":
            return 8.0 * n * l * cs
        if what == "tensor":
            return 8.0 * n * l * cs * 7
        raise KeyError(what)
    return f


KERNEL_BYTES_AUX = {
    "aux_mac": _aux_bytes("mac"), "aux_crt": _aux_bytes("crt"), "aux_inv_pass1": _aux_bytes("pass"), "aux_inv_pass2": _aux_bytes("pass"),
    "aux_fwd_pass2": _aux_bytes("pass"), "ks_pass1": _aux_bytes("ks1"), "ntt_fwd_pass1": _aux_bytes("u64pass"),
    "ntt_fwd_pass2": _aux_bytes("u64pass"), "ntt_inv_pass1": _aux_bytes("u64pass"), "ntt_inv_pass2": _aux_bytes("u64pass"),
    "tensor": _aux_bytes("tensor"),
}

KERNEL_BYTES = {
    # an NTT pass is half of a limb transform (16 N bytes per transform) -> 8 N per limb per pass
    "ntt_fwd_pass1": lambda n, l, batch: 8.0 * n * l * _cs(n, l, batch),
    "ntt_fwd_pass2": lambda n, l, batch: 8.0 * n * l * _cs(n, l, batch),
    "ntt_inv_pass1": lambda n, l, batch: 8.0 * n * l * _cs(n, l, batch),
    "ntt_inv_pass2": lambda n, l, batch: 8.0 * n * l * _cs(n, l, batch),
    "ntt_inv_pass1_rescale": lambda n, l, batch: 8.0 * n * _cs(n, l, batch) * (3 * (l - 1)),
    # ks_pass1: read each digit limb once (its L-1 re-reads are L2 hits), write the (i, j) scratch slabs
    "ks_pass1": lambda n, l, batch: 1.0 * n * _cs(n, l, batch) * (8 * l + KS_WORD * l * (l - 1)),
    # ks_pass2: scratch slabs + key once per launch + NTT-domain digit limb + d0/d1 in + two outputs
    "ks_pass2": lambda n, l, batch: float(_ks2(n, l, batch)),
    "ks_pass2_tma": lambda n, l, batch: float(_ks2(n, l, batch)),
    "ks_mac": lambda n, l, batch: 8.0 * n * l * batch * 5 + 16.0 * n * l,
    "digit_broadcast": lambda n, l, batch: 8.0 * n * batch * (l + 1),
    "tensor": lambda n, l, batch: 8.0 * n * l * _cs(n, l, batch) * 7,
    "rescale": lambda n, l, batch: 8.0 * n * batch * (2 * l - 1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="ciphertext pairs per GPU (default: the config's)")
    ap.add_argument("--e2e-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prof", dest="prof", action="store_false")
    ap.add_argument("--ops", action="store_true", help="also time encrypt and add_ciphertexts on the resident batch (default for cfg2)")
    ap.add_argument("--no-e2e-single", action="store_true", help="N > 1: skip the single-process ckks_comm_* e2e leg on rank 0")
    ap.add_argument("--no-ntt", action="store_true", help="skip the limb-NTT record (the metric's second half)")
    ap.add_argument("--no-chain", action="store_true", help="skip the horner_chain sub-record (cfg4, 1 GPU)")
    ap.add_argument("--no-single-thread", action="store_true", help="skip the single-threaded oracle timing (~40 s at cfg4)")
    ap.add_argument("--imad", action="store_true", help="also run the integer-pipe microbenchmark")
    ap.add_argument("--op", default="", choices=["", "mul", "rotate"], help="hot-path operation (default: mul; rotate for cfg3)")
    ap.add_argument("--host-chunk-mib", type=int, default=0, help="pipeline chunk of the host-buffer entry point")
    ap.add_argument("--ks-aux", type=int, default=None, choices=[0, 1, 2],
                    help="gadget product through auxiliary 30-bit primes: 0 never, 1 automatic (library default), 2 whenever possible")
    ap.add_argument("--ks-scratch-mib", type=int, default=0, help="key-switch scratch per chunk of ciphertexts (default 8192)")
    ap.add_argument("--ntt-sweep", action="store_true", help="BASELINE.json configs[4]: limb-batched NTT/INTT sweep instead of the ct-mult bench")
    ap.add_argument("--limb-sharded", action="store_true", help="optional limb-sharded mode (one batch, limbs spread over the GPUs)")
    ap.add_argument("--comm", default="peer", choices=["peer", "ce", "nccl"],
                    help="limb-sharded exchange: stores into peer HBM from the producing kernels, copy engines for the digits, or NCCL collectives")
    ap.add_argument("--ls-chunk", type=int, default=0, help="limb-sharded: ciphertexts per pass (0 = automatic)")
    args = ap.parse_args()
    if args.limb_sharded:
        run_limb_sharded(args)
    elif args.ntt_sweep:
        run_ntt_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
