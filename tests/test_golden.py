"""Golden vectors (tests/golden/golden_vectors.json, made by tests/golden/make_golden.py): the oracle
must reproduce them on CPU, the CUDA path must reproduce them on the GPU."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden_vectors.json")) as f:
    G = json.load(f)


def A(x, *shape):
    a = np.array(x, dtype=np.uint64)
    return a.reshape(*shape) if shape else a


def test_oracle_reproduces_golden(orc):
    g = G["ntt_n8"]
    b = orc.Basis(8, g["moduli"])
    assert [b.psi(i) for i in range(3)] == g["psi"]
    assert np.array_equal(b.from_coeffs(g["coeffs"]), A(g["channels"]))
    assert np.array_equal(b.to_ntt(A(g["channels"])), A(g["ntt"]))
    # forward NTT of [1,-2,3,4,-5,6,7,-8]
    assert g["ntt"][0] == [10, 2, 9, 9, 7, 6, 7, 9] and g["ntt"][2] == [107, 7, 84, 80, 102, 59, 11, 10]
    g = G["engine_n16"]
    n, l = 16, 4
    b = orc.Basis(n, g["moduli"])
    ka, kb = A(g["key_a"], l, l, n), A(g["key_b"], l, l, n)
    m0, m1 = b.mul_ciphertexts_gadget(A(g["a0"]), A(g["a1"]), A(g["b0"]), A(g["b1"]), ka, kb)
    assert np.array_equal(m0, A(g["mul0"])) and np.array_equal(m1, A(g["mul1"]))
    r0, r1, bits = b.rescale_ciphertext(m0, m1)
    assert np.array_equal(r0, A(g["rescaled0"])) and np.array_equal(r1, A(g["rescaled1"])) and bits == g["bits_dropped"]
    q0, q1 = b.rotate_ciphertext(A(g["a0"]), A(g["a1"]), ka, kb, 3)
    assert np.array_equal(q0, A(g["rot3_0"])) and np.array_equal(q1, A(g["rot3_1"]))
    assert np.array_equal(b.automorphism(A(g["a0"]), 6)[0], A(g["automorphism_6"]))
    g = G["poly_n256"]
    b = orc.Basis(256, g["moduli"])
    assert np.array_equal(b.to_ntt(A(g["x"])), A(g["ntt_x"]))
    assert np.array_equal(b.mul(A(g["x"]), A(g["y"])), A(g["mul"]))
    assert np.array_equal(b.rescale(A(g["x"])), A(g["rescale"]))


@pytest.mark.gpu
def test_gpu_reproduces_golden(gpu):
    g = G["ntt_n8"]
    b = gpu.RnsBasis(8, g["moduli"])
    p = gpu.RnsPoly.from_coeffs(g["coeffs"], b)
    assert np.array_equal(p.channels()[0], A(g["channels"]))
    p.to_ntt_domain()
    assert np.array_equal(p.channels()[0], A(g["ntt"]))
    g = G["engine_n16"]
    n, l = 16, 4
    b = gpu.RnsBasis(n, g["moduli"])
    P = lambda k: gpu.RnsPoly.from_channels(A(g[k]), b)
    key = gpu.GadgetKey.upload(b, A(g["key_a"], l, l, n), A(g["key_b"], l, l, n), rotation=3)
    cta, ctb = gpu.Ciphertext(P("a0"), P("a1"), 30, 124), gpu.Ciphertext(P("b0"), P("b1"), 30, 124)
    m = gpu.CkksEngine.mul_ciphertexts_gadget(cta, ctb, key)
    assert np.array_equal(m.c0.channels()[0], A(g["mul0"])) and np.array_equal(m.c1.channels()[0], A(g["mul1"]))
    r = gpu.CkksEngine.rescale_ciphertext(m)
    assert np.array_equal(r.c0.channels()[0], A(g["rescaled0"])) and np.array_equal(r.c1.channels()[0], A(g["rescaled1"]))
    assert r.logp == 60 - g["bits_dropped"]
    q = gpu.CkksEngine.rotate_ciphertext(cta, key)
    assert np.array_equal(q.c0.channels()[0], A(g["rot3_0"])) and np.array_equal(q.c1.channels()[0], A(g["rot3_1"]))
    key.rotation = -2
    q = gpu.CkksEngine.rotate_ciphertext(cta, key)
    assert np.array_equal(q.c0.channels()[0], A(g["rotm2_0"])) and np.array_equal(q.c1.channels()[0], A(g["rotm2_1"]))
    assert np.array_equal(P("a0").automorphism(5).channels()[0], A(g["automorphism_5"]))
    assert np.array_equal(P("a0").automorphism(6).channels()[0], A(g["automorphism_6"]))
    g = G["poly_n256"]
    b = gpu.RnsBasis(256, g["moduli"])
    x, y = gpu.RnsPoly.from_channels(A(g["x"]), b), gpu.RnsPoly.from_channels(A(g["y"]), b)
    assert np.array_equal(x.rescale().channels()[0], A(g["rescale"]))
    xn = x.clone()
    xn.to_ntt_domain()
    assert np.array_equal(xn.channels()[0], A(g["ntt_x"]))
    x *= y
    assert np.array_equal(x.channels()[0], A(g["mul"]))


def test_limb_dump_round_trip_and_replay_bundle(tmp_path, orc):
    """The limb dump format round-trips, rejects corruption and non-reduced words, and the
    replay bundle written from the oracle is self-consistent."""
    import importlib.util
    import subprocess
    import sys

    root = os.path.dirname(HERE)
    spec = importlib.util.spec_from_file_location("limbdump", os.path.join(root, "toy-heaan-ckks_b200", "limbdump.py"))
    ld = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ld)
    g = G["engine_n16"]
    a0 = A(g["a0"])
    hdr = ld.write(str(tmp_path / "x"), a0, g["moduli"], note="t")
    back, h2 = ld.read(str(tmp_path / "x"))
    assert np.array_equal(back[0], a0) and h2 == hdr and hdr["limbs"] == 4 and hdr["degree"] == 16
    with pytest.raises(ValueError):
        bad = a0.copy()
        bad[0, 0] = g["moduli"][0]
        ld.write(str(tmp_path / "y"), bad, g["moduli"])
    with open(str(tmp_path / "x.u64"), "r+b") as f:
        f.write(b"\x01")
    with pytest.raises(ValueError):
        ld.read(str(tmp_path / "x"))
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "make_replay_bundle.py"), str(tmp_path / "bundle"), "oracle"],
                          stdout=subprocess.DEVNULL)
    r0, h = ld.read(str(tmp_path / "bundle" / "cfg1_n16" / "mul_rescale_c0"))
    assert h["limbs"] == 3 and h["moduli"] == orc.generate_primes(31, 4, 16)[:3]
    b = orc.Basis(16, orc.generate_primes(31, 4, 16))
    rd = lambda n: ld.read(str(tmp_path / "bundle" / "cfg1_n16" / n))[0]
    m0, m1 = b.mul_ciphertexts_gadget(rd("ct1_c0")[0], rd("ct1_c1")[0], rd("ct2_c0")[0], rd("ct2_c1")[0], rd("key_a"), rd("key_b"))
    q0, _, _ = b.rescale_ciphertext(m0, m1)
    assert np.array_equal(q0, r0[0])
