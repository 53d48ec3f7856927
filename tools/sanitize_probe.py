"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ck = importlib.import_module("toy-heaan-ckks_b200")
rng = np.random.default_rng(1)


def uni(moduli, n, *lead):
    q = np.array(moduli, dtype=np.uint64)
    return (rng.integers(0, 1 << 62, size=(*lead, len(moduli), n), dtype=np.uint64) % q[:, None]).astype(np.uint64)


for n, bits, l in ((16, 31, 4), (256, 30, 3), (4096, 40, 3), (8192, 30, 3), (65536, 61, 2)):
    moduli = ck.generate_primes(bits, l, n)
    b = ck.RnsBasis(n, moduli)
    a0, a1, b0, b1 = (uni(moduli, n, 2) for _ in range(4))
    key = ck.GadgetKey.upload(b, uni(moduli, n, l), uni(moduli, n, l), rotation=-1)
    cta = ck.Ciphertext(ck.RnsPoly.from_channels(a0, b), ck.RnsPoly.from_channels(a1, b), 30, 90)
    ctb = ck.Ciphertext(ck.RnsPoly.from_channels(b0, b), ck.RnsPoly.from_channels(b1, b), 30, 90)
    for tma in (True, False):
        ck.set_tma(tma)
        out = ck.CkksEngine.mul_relin_rescale(cta, ctb, key)
        rot = ck.CkksEngine.rotate_ciphertext(cta, key)
        out.c0.channels(), rot.c1.channels()
    ck.set_tma(True)
    p = cta.c0.clone()
    p.to_ntt_domain()
    p.channels()
    p.to_coeff_domain()
    p.automorphism(6).channels()
    cta.c0.rescale().channels()
    o0 = np.zeros((2, l - 1, n), dtype=np.uint64)
    o1 = np.zeros_like(o0)
    ck.mul_relin_rescale_host(b, b.drop_last(1), key, a0, a1, b0, b1, o0, o1)
    assert np.array_equal(o0, out.c0.channels())
    print("ok", n, bits, l, flush=True)
print(ck.launch_table())
