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


def _centred(x, q):
    x %= q
    return x - q if x > q // 2 else x


def _wrap64(x):
    return ((x + (1 << 63)) % (1 << 64)) - (1 << 63)


def test_wide_crt_equals_reference_semantics_below_2_128(gpu, orc):
    """ckks_poly_to_coeffs_wide vs RnsBasis::reconstruct_centered_coeff (basis.rs:158-180) while Q < 2^128: the i64
    words are identical on uniform limbs -- including the reference's `as i64` truncation of values that do not fit --
    and equal Python's exact centred CRT; KATs of basis.rs:310-324."""
    for n, bits, l in ((256, 62, 2), (1024, 40, 3), (256, 30, 4)):
        moduli = orc.generate_primes(bits, l, n)
        gb, ob = gpu.RnsBasis(n, moduli), orc.Basis(n, moduli)
        rng = np.random.default_rng(n + bits)
        q = np.array(moduli, dtype=np.uint64)
        ch = (rng.integers(0, 1 << 63, size=(2, l, n), dtype=np.uint64) % q[:, None]).astype(np.uint64)
        p = gpu.RnsPoly.from_channels(ch, gb)
        wi, wf, ov = p.to_coeffs_wide()
        assert np.array_equal(wi, p.to_coeffs())  # the product's u128 path = the reference's arithmetic
        Q = 1
        for m in moduli:
            Q *= m
        for b, k in ((0, 0), (1, 17), (1, n - 1)):
            res = [int(ch[b, i, k]) for i in range(l)]
            x = sum(r * (Q // m) * pow(Q // m, -1, m) for r, m in zip(res, moduli))
            c = _centred(x, Q)
            assert int(wi[b, k]) == _wrap64(c) == ob.reconstruct_centered_coeff(res)
            assert abs(wf[b, k] - float(c)) <= abs(float(c)) * 1e-14
        assert ov == (bits * l > 64)
    gb = gpu.RnsBasis(8, [17, 97])
    kat = np.zeros((1, 2, 8), dtype=np.uint64)
    kat[0, :, 0] = [3, 3]
    kat[0, :, 1] = [10, 90]
    wi, _, ov = gpu.RnsPoly.from_channels(kat, gb).to_coeffs_wide()
    assert wi[0, 0] == 3 and wi[0, 1] == -7 and not ov


@pytest.mark.parametrize("n,bits,l", [(256, 61, 24), (512, 30, 8), (256, 61, 3)])
def test_wide_crt_beyond_2_128_matches_python_big_integers(gpu, orc, n, bits, l):
    """Q >= 2^128 (the reference's u128 product overflows there: basis.rs:152-160): Garner's mixed radix on the device
    against Python's arbitrary-precision centred CRT, for planted values from +-1 to +-Q/2 and for uniform limbs."""
    moduli = orc.generate_primes(bits, l, n)
    gb = gpu.RnsBasis(n, moduli)
    Q = 1
    for m in moduli:
        Q *= m
    rng = np.random.default_rng(5)
    planted = [0, 1, -1, 2, -2, (1 << 40) + 3, -(1 << 40) - 3, (1 << 62) + 12345, -(1 << 62) - 12345, (1 << 63) - 1, -(1 << 63), (1 << 63), -(1 << 63) - 1,
               (1 << 100) + 7, -(1 << 100) - 7, Q // 2, -(Q // 2), Q // 2 - 1, Q // 3, -(Q // 3), moduli[0], -moduli[0], moduli[0] * moduli[1], -moduli[0] * moduli[1] - 1]
    planted = [v for v in planted if abs(v) <= Q // 2]
    vals = planted + [int(rng.integers(-(1 << 62), 1 << 62)) * int(rng.integers(1, 1 << 62)) % Q - Q // 2 for _ in range(n - len(planted))]
    ch = np.zeros((1, l, n), dtype=np.uint64)
    for k, v in enumerate(vals):
        for i, m in enumerate(moduli):
            ch[0, i, k] = v % m
    p = gpu.RnsPoly.from_channels(ch, gb)
    wi, wf, ov = p.to_coeffs_wide()
    for k, v in enumerate(vals):
        c = _centred(v, Q)
        assert int(wi[0, k]) == _wrap64(c), (k, v)
        if abs(c) < (1 << 1000):
            assert abs(wf[0, k] - float(c)) <= abs(float(c)) * l * 2.0 ** -52, (k, v)
        else:  # beyond the range of a double (24 x 61 bits = 2^1464): signed infinity
            assert np.isinf(wf[0, k]) and (wf[0, k] > 0) == (c > 0), (k, v)
    assert ov  # values beyond 2^63 were planted
    small = np.zeros((1, l, n), dtype=np.uint64)
    for k in range(n):
        v = int(rng.integers(-(1 << 62), 1 << 62))
        for i, m in enumerate(moduli):
            small[0, i, k] = v % m
    _, _, ov = gpu.RnsPoly.from_channels(small, gb).to_coeffs_wide()
    assert not ov
    # the NTT-domain polynomial gives the same values (poly.rs:405-411: to_coeffs transforms a clone first)
    p.to_ntt_domain()
    wi2, _, _ = p.to_coeffs_wide()
    assert np.array_equal(wi, wi2)


def test_decode_at_full_level_without_mod_drop(gpu, orc):
    """A ciphertext at a level where Q >= 2^128 (5 x 61 bits here; 24 x 61 at cfg4) decrypts AND decodes directly:
    the reference has to mod_drop_last down to two primes first (horner_chain.rs:21-35, basis.rs:152-160).  Checked
    against the plain values within north_star's 2^-(scale_bits-10) and against the decode after mod_drop_last."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_gpu_engine import Party

    n, l, sb = 256, 5, 40
    moduli = orc.generate_primes(61, l, n)
    party = Party(orc, n, moduli, seed=11, hw=16)
    gb = gpu.RnsBasis(n, moduli)
    rng = np.random.default_rng(12)
    vals = rng.uniform(-0.9, 0.9, n // 2)
    c0, c1 = party.encrypt(vals, sb)
    ct = gpu.Ciphertext(gpu.RnsPoly.from_channels(c0, gb), gpu.RnsPoly.from_channels(c1, gb), sb, gb.total_bits())
    s = gpu.RnsPoly.from_channels(party.s, gb)
    dec = gpu.CkksEngine.decrypt(ct, s)
    enc = gpu.CkksEncoder(n, sb)
    got = enc.decode(gpu.Plaintext(dec, sb, n // 2))[0]  # 305-bit Q: the wide CRT
    assert np.max(np.abs(got - vals)) <= 2.0 ** -(sb - 10)  # north_star's bound
    assert np.max(np.abs(got - vals)) <= (10 * 3.2 * (16 * n) ** 0.5 + 4) / 2.0**sb  # the reference's own (encrypt_add.rs:122-131)
    low = enc.decode(gpu.Plaintext(dec.mod_drop_last(l - 2), sb, n // 2))[0]  # the reference's route: two primes left
    assert np.max(np.abs(got - low)) <= 1e-9
    wi, _, ov = dec.to_coeffs_wide()
    assert not ov and np.array_equal(wi, dec.mod_drop_last(l - 2).to_coeffs())


def test_wide_crt_and_decode_at_cfg4_size(gpu, orc):
    """N=2^16, L=24, generate_primes(61, 24, 65536) (BASELINE configs[3]): Q has 1464 bits, where the reference can
    neither run reconstruct_centered_coeff (u128 product, basis.rs:152-160) nor decode.  from_coeffs -> wide CRT returns
    the coefficients bit for bit (from either domain), and encode -> decode at the full level stays within
    2^-(scale_bits - 10) without any mod_drop_last."""
    n, l, sb = 65536, 24, 50
    moduli = orc.generate_primes(61, l, n)
    gb = gpu.RnsBasis(n, moduli)
    rng = np.random.default_rng(2)
    coeffs = rng.integers(-(1 << 62), 1 << 62, size=(2, n), dtype=np.int64)
    coeffs[0, :4] = [0, -1, (1 << 63) - 1, -(1 << 63)]
    p = gpu.RnsPoly.from_coeffs(coeffs, gb)
    wi, wf, ov = p.to_coeffs_wide()
    assert not ov and np.array_equal(wi, coeffs)
    assert np.all(np.abs(wf - coeffs.astype(np.float64)) <= np.abs(coeffs.astype(np.float64)) * 2.0 ** -50)
    p.to_ntt_domain()
    wi2, _, _ = p.to_coeffs_wide()
    assert np.array_equal(wi2, coeffs)
    enc = gpu.CkksEncoder(n, sb)
    vals = rng.uniform(-0.9, 0.9, (2, n // 2))
    pt = enc.encode(vals, gb)
    got = enc.decode(pt)
    assert np.max(np.abs(got - vals)) <= 2.0 ** -(sb - 10)
