"""Regenerates tests/golden/*.json from the oracle (python tests/golden/make_golden.py).

The Rust reference cannot be built in this image (no cargo/rustc), so these vectors are outputs of
the CPU oracle, which tests/test_oracle_kats.py pins against every known-answer test and identity
of the reference's own unit tests.  They freeze the oracle's behaviour so that a later change to
either the oracle or the CUDA path is caught by both `-m "not gpu"` and `-m gpu` suites."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle as orc  # noqa: E402


def uni(rng, moduli, n, *lead):
    q = np.array(moduli, dtype=np.uint64)
    return (rng.integers(0, 1 << 63, size=(*lead, len(moduli), n), dtype=np.uint64) % q[:, None]).astype(np.uint64)


def L(a):
    return np.asarray(a).astype(object).tolist() if np.asarray(a).dtype == object else [[int(v) for v in row] for row in np.asarray(a).reshape(-1, np.asarray(a).shape[-1])]


def main():
    out = {}
    # 1. N = 8 transforms over {17, 97, 113} (SURVEY.md 8c constants)
    b = orc.Basis(8, [17, 97, 113])
    co = [1, -2, 3, 4, -5, 6, 7, -8]
    x = b.from_coeffs(co)
    out["ntt_n8"] = {"moduli": [17, 97, 113], "coeffs": co, "channels": L(x), "ntt": L(b.to_ntt(x)), "psi": [b.psi(i) for i in range(3)]}
    # 2. examples/encrypt_mul shape: N = 16, generate_primes(31, 4, 16)
    n, l = 16, 4
    moduli = orc.generate_primes(31, l, n)
    b = orc.Basis(n, moduli)
    rng = np.random.default_rng(2024)
    a0, a1, b0, b1 = (uni(rng, moduli, n) for _ in range(4))
    ka, kb = uni(rng, moduli, n, l), uni(rng, moduli, n, l)
    m0, m1 = b.mul_ciphertexts_gadget(a0, a1, b0, b1, ka, kb)
    r0, r1, bits = b.rescale_ciphertext(m0, m1)
    q0, q1 = b.rotate_ciphertext(a0, a1, ka, kb, 3)
    n0, n1 = b.rotate_ciphertext(a0, a1, ka, kb, -2)
    out["engine_n16"] = {
        "moduli": moduli, "a0": L(a0), "a1": L(a1), "b0": L(b0), "b1": L(b1), "key_a": L(ka.reshape(l * l, n)), "key_b": L(kb.reshape(l * l, n)),
        "mul0": L(m0), "mul1": L(m1), "rescaled0": L(r0), "rescaled1": L(r1), "bits_dropped": bits,
        "rot3_0": L(q0), "rot3_1": L(q1), "rotm2_0": L(n0), "rotm2_1": L(n1),
        "automorphism_5": L(b.automorphism(a0, 5)[0]), "automorphism_6": L(b.automorphism(a0, 6)[0]),
    }
    # 3. a four-step size: N = 256, 40-bit primes
    n, l = 256, 2
    moduli = orc.generate_primes(40, l, n)
    b = orc.Basis(n, moduli)
    x, y = uni(rng, moduli, n), uni(rng, moduli, n)
    out["poly_n256"] = {"moduli": moduli, "x": L(x), "y": L(y), "ntt_x": L(b.to_ntt(x)), "mul": L(b.mul(x, y)), "rescale": L(b.rescale(x))}
    with open(os.path.join(HERE, "golden_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "golden_vectors.json"), os.path.getsize(os.path.join(HERE, "golden_vectors.json")), "bytes")


if __name__ == "__main__":
    main()
