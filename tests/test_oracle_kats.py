"""Pins the oracle against every known-answer test and identity the reference's own unit tests
hold for the hot path.  Each test cites the reference test it restates."""
import numpy as np
import pytest

from conftest import uniform_limbs

Q8 = [17, 97, 113]


def test_prime_search_kats(orc):
    # primes.rs:407-410
    assert orc.get_first_prime_up(30, 1024) == 1073750017
    # basis.rs:273-282: 19 is prime but not == 1 mod 16
    assert orc.is_prime(19) and not orc.is_ntt_friendly_prime(19, 8)
    # Carmichael numbers are composite (primes.rs:300-320)
    for c in (561, 1105, 1729, 2465, 2821, 6601):
        assert not orc.is_prime(c)
    for p in (2, 3, 5, 7, 97, 7681, 12289, 1152921504606584833):
        assert orc.is_prime(p)


def test_generate_primes_chains(orc):
    # derived constants listed in SURVEY.md 8c (descending chains from 2^bits in steps of 2N)
    assert orc.generate_primes(31, 4, 16) == [2147483489, 2147483137, 2147482817, 2147482273]
    assert orc.generate_primes(30, 3, 32) == [1073741441, 1073740609, 1073739649]
    assert orc.generate_primes(62, 2, 1024) == [4611686018427365377, 4611686018427322369]
    assert orc.generate_primes(40, 3, 1024) == [1099511592961, 1099511590913, 1099511560193]
    for bits, cnt, n in ((31, 4, 16), (40, 3, 4096), (61, 5, 8192)):
        ps = orc.generate_primes(bits, cnt, n)
        assert all(orc.is_ntt_friendly_prime(p, n) and (1 << (bits - 1)) <= p < (1 << bits) for p in ps)
        assert ps == sorted(ps, reverse=True) and len(set(ps)) == cnt
    with pytest.raises(orc.OracleError):  # utils.rs:98-104: not enough primes -> panic
        orc.generate_primes(5, 50, 8)


def test_basis_errors(orc):
    with pytest.raises(orc.OracleError) as e:
        orc.Basis(8, [])
    assert e.value.kind == "EmptyBasis"  # basis.rs:285-289
    with pytest.raises(orc.OracleError) as e:
        orc.Basis(8, [19])
    assert e.value.kind == "NonNttFriendlyModulus"  # basis.rs:273-282
    with pytest.raises(orc.OracleError) as e:
        orc.Basis(12, [17])
    assert e.value.kind == "InvalidDegree"
    b = orc.Basis(8, Q8)
    assert b.drop_last(1).moduli == [17, 97]  # basis.rs:292-300
    with pytest.raises(orc.OracleError) as e:
        b.drop_last(3)
    assert e.value.kind == "InvalidModDrop"


def test_psi_and_ntt_constants(orc):
    b = orc.Basis(8, Q8)
    assert [b.psi(i) for i in range(3)] == [3, 8, 40]
    got = b.to_ntt(b.from_coeffs([1, -2, 3, 4, -5, 6, 7, -8]))
    assert got[0].tolist() == [10, 2, 9, 9, 7, 6, 7, 9]
    assert got[1].tolist() == [30, 16, 43, 0, 60, 44, 13, 93]
    assert got[2].tolist() == [107, 7, 84, 80, 102, 59, 11, 10]
    # slot k = p(psi^(2k+1)), natural order
    for ch, q in enumerate(Q8):
        psi = b.psi(ch)
        co = [1, -2, 3, 4, -5, 6, 7, -8]
        for k in range(8):
            assert int(got[ch][k]) == sum(c * pow(psi, j * (2 * k + 1), q) for j, c in enumerate(co)) % q


def test_centered_crt_kats(orc):
    # basis.rs:310-324
    assert orc.Basis(8, [17, 97]).reconstruct_centered_coeff([3, 3]) == 3
    assert orc.Basis(8, [17, 97]).reconstruct_centered_coeff([10, 90]) == -7
    assert orc.Basis(8, [97]).reconstruct_centered_coeff([96]) == -1


def test_from_coeffs_and_reducedness(orc):
    b = orc.Basis(8, Q8)
    ch = b.from_coeffs([1, -1, 18, -18, 0, 113, -113, 114])  # poly.rs:683-692
    assert ch[0].tolist() == [1, 16, 1, 16, 0, 11, 6, 12]
    assert ch[2].tolist() == [1, 112, 18, 95, 0, 0, 0, 1]
    bad = ch.copy()
    bad[0, 0] = 17
    with pytest.raises(orc.OracleError) as e:
        b.from_channels_check(bad)  # poly.rs:704-714
    assert e.value.kind == "NonReducedCoefficient"
    with pytest.raises(orc.OracleError) as e:
        b.from_channels_check(ch[:2])
    assert e.value.kind == "ChannelCountMismatch"
    with pytest.raises(orc.OracleError):
        b.from_coeffs([1, 2, 3])  # poly.rs:50-54


def test_add_neg_mul_kats(orc):
    b = orc.Basis(8, Q8)
    x = b.from_coeffs([16, 0, 0, 0, 0, 0, 0, 0])
    y = b.from_coeffs([2, 0, 0, 0, 0, 0, 0, 0])
    assert int(b.add(x, y)[0][0]) == 1  # 16 + 2 == 1 mod 17, poly.rs:767-775
    assert int(b.neg(b.from_coeffs([3] + [0] * 7))[0][0]) == 14  # poly.rs:778-786
    one_x = b.from_coeffs([1, 1, 0, 0, 0, 0, 0, 0])
    sq = b.mul(one_x, one_x)  # (1+x)^2, poly.rs:789-802
    assert np.array_equal(sq, b.from_coeffs([1, 2, 1, 0, 0, 0, 0, 0]))
    x7 = b.from_coeffs([0] * 7 + [1])
    x1 = b.from_coeffs([0, 1] + [0] * 6)
    assert np.array_equal(b.mul(x7, x1), b.from_coeffs([-1] + [0] * 7))  # x^7 * x = -1, poly.rs:805-815


def test_ntt_identities(orc):
    rng = np.random.default_rng(11)
    b = orc.Basis(8, Q8)
    x, y = uniform_limbs(rng, Q8, 8), uniform_limbs(rng, Q8, 8)
    assert np.array_equal(b.to_coeff(b.to_ntt(x)), x)  # poly.rs:717-729
    ntt_prod = b.to_coeff(b.mul(b.to_ntt(x), b.to_ntt(y), in_ntt=True))
    assert np.array_equal(ntt_prod, b.mul(x, y))  # poly.rs:854-877
    assert np.array_equal(b.mul(x, y), b.mul_naive(x, y))  # poly.rs:960-975
    n = 64
    qs = orc.generate_primes(40, 2, n)
    b2 = orc.Basis(n, qs)
    x, y = uniform_limbs(rng, qs, n), uniform_limbs(rng, qs, n)
    assert np.array_equal(b2.mul(x, y), b2.mul_naive(x, y))


def test_automorphism_kats(orc):
    b = orc.Basis(8, Q8)
    rng = np.random.default_rng(5)
    x = uniform_limbs(rng, Q8, 8)
    for e in (1, 17):  # identity exponents, poly.rs:880-892
        out, dom = b.automorphism(x, e)
        assert np.array_equal(out, x) and dom is False
    out, dom = b.automorphism(b.to_ntt(x), 16, in_ntt=True)  # e % 2N == 0 keeps the flag (quirk)
    assert dom is True
    one_x = b.from_coeffs([1, 1, 0, 0, 0, 0, 0, 0])
    out, _ = b.automorphism(one_x, 9)  # X -> X^9 = -X, poly.rs:895-912
    assert np.array_equal(out, b.from_coeffs([1, -1, 0, 0, 0, 0, 0, 0]))
    out, dom = b.automorphism(b.to_ntt(x), 3, in_ntt=True)  # poly.rs:944-957
    ref, _ = b.automorphism(x, 3)
    assert dom is False and np.array_equal(out, ref)


def test_rescale_kats(orc):
    b = orc.Basis(8, Q8)
    assert b.rescale(b.from_coeffs([226] + [0] * 7)).shape == (2, 8)  # poly.rs:992-1000
    with pytest.raises(orc.OracleError) as e:
        orc.Basis(8, [17]).rescale(np.zeros((1, 8), dtype=np.uint64))  # poly.rs:1003-1009
    assert e.value.kind == "InvalidModDrop"
    out = b.rescale(b.from_coeffs([226] + [0] * 7))  # 226 / 113 = 2, poly.rs:1012-1033
    assert out[0].tolist() == [2] + [0] * 7 and out[1].tolist() == [2] + [0] * 7
    rng = np.random.default_rng(3)
    x = uniform_limbs(rng, Q8, 8)
    assert np.array_equal(b.rescale(b.to_ntt(x), in_ntt=True), b.rescale(x))  # poly.rs:1036-1049


def test_encrypt_mul_example_decodes(orc):
    """examples/encrypt_mul.rs as shipped: N=16, generate_primes(31,4,16), scale 2^30; error <= 1e-4 (:149)."""
    n, l, sb = 16, 4, 30
    primes = orc.generate_primes(31, l, n)
    b = orc.Basis(n, primes)
    rng = np.random.default_rng(42)
    s = b.from_coeffs(rng.permutation([1] * 4 + [-1] * 4 + [0] * 8))
    gauss = lambda *lead: np.stack([b.from_coeffs(np.rint(rng.normal(0, 3.2, n)).astype(np.int64)) for _ in range(int(np.prod(lead)))]).reshape(*lead, l, n) if lead else b.from_coeffs(np.rint(rng.normal(0, 3.2, n)).astype(np.int64))
    pk_a = uniform_limbs(rng, primes, n)
    pk_b = b.gen_public_key(s, pk_a, gauss())
    ka = uniform_limbs(rng, primes, n, l)
    kb = b.gen_gadget_relin_key(s, ka, gauss(l))
    va, vb = [1.0, 2.0, 3.0, 4.0], [0.5, 1.0, 1.5, 2.0]
    cts = []
    for v in (va, vb):
        m = b.from_coeffs(orc.encode(n, sb, v))
        u = b.from_coeffs(rng.permutation([1] * 4 + [-1] * 4 + [0] * 8))
        cts.append(b.encrypt(pk_b, pk_a, u, gauss(), gauss(), m))
    m0, m1 = b.mul_ciphertexts_gadget(*cts[0], *cts[1], ka, kb)
    r0, r1, bits = b.rescale_ciphertext(m0, m1)
    assert bits == 31
    b3 = b.drop_last(1)
    dec = b3.decrypt(r0, r1, s[:3])
    b2 = b3.drop_last(1)  # Q < 2^128 for the CRT (basis.rs:152-160): 3 x 31 bits is fine, keep 3
    vals = orc.decode(n, 2 * sb - bits, b3.to_coeffs(dec), 4)
    expect = np.array(va) * np.array(vb)
    assert np.max(np.abs(vals.real - expect)) <= 1e-4
