"""CkksEncoder on the device: O(N log N) canonical embedding, tolerance-checked against the
oracle's restatement of the reference's O(N^2) Vandermonde encoder (ckks_encoder.rs:65-156)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,sb,nv", [(8, 30, 3), (8, 30, 4), (64, 40, 32), (1024, 30, 512), (1024, 50, 100), (8192, 40, 4096)])
def test_encode_decode_match_oracle(gpu, orc, n, sb, nv):
    moduli = orc.generate_primes(40, 3, n)
    gb = gpu.RnsBasis(n, moduli)
    ob = orc.Basis(n, moduli)
    rng = np.random.default_rng(n + sb)
    v = rng.uniform(-1, 1, (2, nv)) + 1j * rng.uniform(-1, 1, (2, nv))
    enc = gpu.CkksEncoder(n, sb)
    pt = enc.encode_complex(v, gb)
    assert pt.slots == nv and pt.scale_bits == sb and not pt.poly.is_ntt_domain()
    got = pt.poly.to_coeffs()
    for b in range(2):
        ref = orc.encode(n, sb, v[b])
        # f64 summation order differs: the reference (and the oracle) build root powers by repeated
        # multiplication inside an O(N^2) sum, so ITS error grows like N^1.5 * 2^-52 * Delta; the FFT's is smaller
        tol = 2 + 2.0 ** (sb - 52) * n ** 1.5
        assert np.max(np.abs(got[b] - ref)) <= tol, (np.max(np.abs(got[b] - ref)), tol)
        # limbs are the rem_euclid residues of exactly those integer coefficients
        assert np.array_equal(pt.poly.channels()[b], ob.from_coeffs(got[b]))
    dec = enc.decode_complex(pt)
    for b in range(2):
        refd = orc.decode(n, sb, got[b], nv)
        assert np.max(np.abs(dec[b] - refd)) < 1e-9 * max(1.0, n / 512) ** 1.5  # the O(N^2) reference sum is the noisier side
        assert np.max(np.abs(dec[b] - v[b])) < n * 2.0 ** -(sb - 1)  # rounding noise of encoding, about 1/Delta per coefficient
    # real-valued front end and the reference's encode/decode unit test (ckks_encoder.rs:161-228: eps = 0.1 at N = 8)
    r = enc.decode(enc.encode([1.5, -2.0, 0.25][:nv], gb))
    assert np.max(np.abs(r[0] - np.array([1.5, -2.0, 0.25][:nv]))) < 0.1


def test_encoder_argument_checks(gpu, orc):
    gb = gpu.RnsBasis(8, [17, 97, 113])
    enc = gpu.CkksEncoder(8, 10)
    with pytest.raises(gpu.RnsNttError):  # more values than slots: the reference asserts (ckks_encoder.rs:70-75)
        enc.encode([1.0] * 5, gb)
    with pytest.raises(gpu.RnsNttError):
        gpu.CkksEncoder(12, 10)
    assert enc.max_slots() == 4 and enc.scale_factor() == 1024.0


def test_encrypt_mul_decrypt_entirely_on_device(gpu, orc):
    """examples/encrypt_mul.rs with encoder, encryption, multiplication, rescale, decryption and decoding all on
    the device (only sampling on the host): error <= 1e-4 (:149)."""
    n, l, sb = 16, 4, 30
    moduli = orc.generate_primes(31, l, n)
    gb = gpu.RnsBasis(n, moduli)
    rng = np.random.default_rng(42)
    sig = 3.2 ** 0.5

    def ternary():
        v = np.zeros(n, dtype=np.int64)
        idx = rng.permutation(n)[: n // 2]
        v[idx] = rng.choice([-1, 1], size=n // 2)
        return v

    gauss = lambda *lead: np.rint(rng.normal(0, sig, size=(*lead, n))).astype(np.int64)
    uni = lambda *lead: (rng.integers(0, 1 << 62, size=(*lead, l, n), dtype=np.uint64) % np.array(moduli, dtype=np.uint64)[:, None]).astype(np.uint64)
    s = gpu.RnsPoly.from_coeffs(ternary(), gb)
    pk_a = gpu.RnsPoly.from_channels(uni(), gb)
    pk_b = pk_a.clone()
    pk_b *= s
    pk_b = -pk_b
    pk_b += gpu.RnsPoly.from_coeffs(gauss(), gb)
    ka = gpu.RnsPoly.from_channels(uni(l), gb)
    s2 = s.clone()
    s2 *= s
    rlk = gpu.GadgetKey.from_polys(ka, gpu.CkksEngine.gadget_key_b(s, s2, ka, gpu.RnsPoly.from_coeffs(gauss(l), gb)))
    enc = gpu.CkksEncoder(n, sb)
    va, vb = np.array([1.0, 2.0, 3.0, 4.0]), np.array([0.5, 1.0, 1.5, 2.0])
    cts = []
    for v in (va, vb):
        m = enc.encode(v, gb).poly
        cts.append(gpu.CkksEngine.encrypt(pk_b, pk_a, gpu.RnsPoly.from_coeffs(ternary(), gb), gpu.RnsPoly.from_coeffs(gauss(), gb),
                                          gpu.RnsPoly.from_coeffs(gauss(), gb), m, sb, 124))
    out = gpu.CkksEngine.mul_relin_rescale(cts[0], cts[1], rlk)
    dec = gpu.CkksEngine.decrypt(out, s.mod_drop_last(1, out.c0.basis()))
    got = enc.decode(gpu.Plaintext(dec, out.logp, 4))[0]
    assert np.max(np.abs(got - va * vb)) <= 1e-4
