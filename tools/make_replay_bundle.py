"""Writes a replay bundle (limb dumps of inputs, keys and outputs of mul_ciphertexts_gadget +
rescale_ciphertext and rotate_ciphertext) so that a machine with the Rust toolchain can run the same
inputs through the real crate and compare word for word.   python tools/make_replay_bundle.py OUTDIR [gpu|oracle]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
spec = importlib.util.spec_from_file_location("limbdump", os.path.join(ROOT, "toy-heaan-ckks_b200", "limbdump.py"))
limbdump = importlib.util.module_from_spec(spec)
spec.loader.exec_module(limbdump)


def main():
    out = sys.argv[1]
    backend = sys.argv[2] if len(sys.argv) > 2 else "oracle"
    os.makedirs(out, exist_ok=True)
    import oracle as orc

    for name, n, bits, l in (("cfg1_n16", 16, 31, 4), ("n1024_40bit", 1024, 40, 3)):
        moduli = orc.generate_primes(bits, l, n)
        rng = np.random.default_rng(2024)
        q = np.array(moduli, dtype=np.uint64)
        uni = lambda *lead: (rng.integers(0, 1 << 63, size=(*lead, l, n), dtype=np.uint64) % q[:, None]).astype(np.uint64)
        a0, a1, b0, b1, ka, kb = uni(), uni(), uni(), uni(), uni(l), uni(l)
        if backend == "gpu":
            ck = importlib.import_module("toy-heaan-ckks_b200")
            gb = ck.RnsBasis(n, moduli)
            key = ck.GadgetKey.upload(gb, ka, kb, rotation=1)
            P = lambda x: ck.RnsPoly.from_channels(x, gb)
            cta, ctb = ck.Ciphertext(P(a0), P(a1), 30, 90), ck.Ciphertext(P(b0), P(b1), 30, 90)
            r = ck.CkksEngine.mul_relin_rescale(cta, ctb, key)
            r0, r1 = r.c0.channels()[0], r.c1.channels()[0]
            t = ck.CkksEngine.rotate_ciphertext(cta, key)
            t0, t1 = t.c0.channels()[0], t.c1.channels()[0]
        else:
            ob = orc.Basis(n, moduli)
            m0, m1 = ob.mul_ciphertexts_gadget(a0, a1, b0, b1, ka, kb)
            r0, r1, _ = ob.rescale_ciphertext(m0, m1)
            t0, t1 = ob.rotate_ciphertext(a0, a1, ka, kb, 1)
        d = os.path.join(out, name)
        os.makedirs(d, exist_ok=True)
        for nm, arr, mods in (("ct1_c0", a0, moduli), ("ct1_c1", a1, moduli), ("ct2_c0", b0, moduli), ("ct2_c1", b1, moduli),
                              ("key_a", ka, moduli), ("key_b", kb, moduli), ("mul_rescale_c0", r0, moduli[:-1]),
                              ("mul_rescale_c1", r1, moduli[:-1]), ("rotate1_c0", t0, moduli), ("rotate1_c1", t1, moduli)):
            limbdump.write(os.path.join(d, nm), arr, mods, note=f"{backend}; key_* are [digit][limb][N] with digit as the batch index")
        print("wrote", d)


if __name__ == "__main__":
    main()
