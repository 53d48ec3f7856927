"""GPU parity of the RnsPoly / RnsBasis surface against the oracle: bit-exact limbs.
Mirrors the reference's unit tests (poly.rs:657-1050, basis.rs:261-325) through the C ABI."""
import numpy as np
import pytest

from conftest import uniform_limbs

pytestmark = pytest.mark.gpu
Q8 = [17, 97, 113]


def _bases(gpu, orc, n, moduli, path=0):
    gpu.set_ntt_path(path)
    try:
        b = gpu.RnsBasis(n, moduli)
    finally:
        gpu.set_ntt_path(0)
    return b, orc.Basis(n, moduli)


@pytest.mark.parametrize("path,n,bits,l", [
    (1, 1, 20, 2), (1, 2, 20, 2), (1, 8, 0, 3), (1, 16, 31, 4), (1, 32, 30, 3), (1, 256, 40, 2), (1, 1024, 62, 2), (1, 2048, 63, 2),
    (2, 256, 30, 3), (2, 512, 40, 2), (2, 1024, 62, 2), (2, 1024, 40, 3), (2, 2048, 63, 2), (2, 4096, 40, 3), (2, 8192, 61, 3),
    (2, 16384, 30, 8), (2, 32768, 61, 2), (2, 65536, 61, 3), (2, 65536, 63, 2), (2, 65536, 30, 2),
])
def test_ntt_matches_oracle(gpu, orc, path, n, bits, l):
    """to_ntt_domain / to_coeff_domain (poly.rs:136-166): natural-order NTT words equal the oracle's,
    round trip is the identity (poly.rs:717-729), idempotent on the flag (poly.rs:732-752)."""
    moduli = Q8 if bits == 0 else orc.generate_primes(bits, l, n)
    gb, ob = _bases(gpu, orc, n, moduli, path)
    assert [gb.psi(i) for i in range(len(moduli))] == [ob.psi(i) for i in range(len(moduli))]
    rng = np.random.default_rng(n + bits)
    batch = 3 if n <= 16384 else 2
    x = uniform_limbs(rng, moduli, n, batch)
    x[0, :, 0] = np.array(moduli, dtype=np.uint64) - 1  # extreme residue
    p = gpu.RnsPoly.from_channels(x, gb)
    p.to_ntt_domain()
    assert p.is_ntt_domain()
    got = p.channels()
    for i in range(batch):
        assert np.array_equal(got[i], ob.to_ntt(x[i])), f"forward NTT differs (batch {i})"
    p.to_ntt_domain()  # no-op
    assert np.array_equal(p.channels(), got)
    p.to_coeff_domain()
    assert not p.is_ntt_domain()
    assert np.array_equal(p.channels(), x)
    # NTT-domain upload in the reference's natural order
    pn = gpu.RnsPoly.from_channels(got, gb, is_ntt_domain=True)
    pn.to_coeff_domain()
    assert np.array_equal(pn.channels(), x)


@pytest.mark.parametrize("n,bits,l", [(4096, 30, 3), (4096, 61, 2), (8192, 31, 2), (8192, 40, 3), (16384, 30, 4), (16384, 62, 2), (16384, 63, 2)])
def test_fused_and_two_pass_transforms_agree(gpu, orc, n, bits, l):
    """2^12 <= N <= 2^14: the single-kernel transform (limb resident in shared memory) and the two-pass
    transform produce the oracle's words; round trips are exact."""
    moduli = orc.generate_primes(bits, l, n)
    gb, ob = _bases(gpu, orc, n, moduli)
    rng = np.random.default_rng(n + bits + 7)
    x = uniform_limbs(rng, moduli, n, 3)
    ref = ob.to_ntt(x[2])
    for fused in (True, False):
        gpu.set_fused_ntt(fused)
        try:
            p = gpu.RnsPoly.from_channels(x, gb)
            p.to_ntt_domain()
            assert np.array_equal(p.channels()[2], ref), f"fused={fused}"
            p.to_coeff_domain()
            assert np.array_equal(p.channels(), x), f"fused={fused}"
        finally:
            gpu.set_fused_ntt(True)


def test_basis_surface(gpu, orc):
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.RnsBasis(8, [])
    assert e.value.kind == "EmptyBasis"
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.RnsBasis(8, [19])
    assert e.value.kind == "NonNttFriendlyModulus"
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.RnsBasis(12, [17])
    assert e.value.kind == "InvalidDegree"
    b = gpu.RnsBasis(8, Q8)
    assert b.moduli() == Q8 and b.channel_count() == 3
    assert b.total_bits() == orc.Basis(8, Q8).total_bits()
    c = b.drop_last(1)
    assert c.moduli() == [17, 97]
    with pytest.raises(gpu.RnsNttError) as e:
        b.drop_last(3)
    assert e.value.kind == "InvalidModDrop"
    assert gpu.RnsBasis(8, [17, 97]).reconstruct_centered_coeff([10, 90]) == -7
    assert gpu.RnsBasis(8, [97]).reconstruct_centered_coeff([96]) == -1
    assert gpu.generate_primes(31, 4, 16) == orc.generate_primes(31, 4, 16)
    assert gpu.generate_primes(61, 24, 65536) == orc.generate_primes(61, 24, 65536)


def test_constructors_and_errors(gpu, orc):
    gb, ob = _bases(gpu, orc, 8, Q8)
    co = [1, -1, 18, -18, 0, 113, -113, 114]
    assert np.array_equal(gpu.RnsPoly.from_coeffs(co, gb).channels()[0], ob.from_coeffs(co))
    big = np.array([[2**63 - 1, -(2**63), 5, -5, 0, 1, -1, 2**40]], dtype=np.int64)
    assert np.array_equal(gpu.RnsPoly.from_coeffs(big, gb).channels()[0], ob.from_coeffs(big[0]))
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.RnsPoly.from_coeffs([1, 2, 3], gb)
    assert e.value.kind == "ShortInput"
    ch = ob.from_coeffs(co)
    bad = ch.copy()
    bad[0, 0] = 17
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.RnsPoly.from_channels(bad, gb)
    assert e.value.kind == "NonReducedCoefficient"
    with pytest.raises(gpu.RnsNttError) as e:
        gpu.RnsPoly.from_channels(ch[:2], gb)
    assert e.value.kind == "ChannelCountMismatch"
    z = gpu.RnsPoly.zero(gb, 2)
    assert not z.channels().any() and not z.is_ntt_domain()
    # empty batch
    e0 = gpu.RnsPoly.from_channels(np.zeros((0, 3, 8), dtype=np.uint64), gb)
    e0.to_ntt_domain()
    assert e0.channels().shape == (0, 3, 8)
    # domain / basis mismatch are loud
    a = gpu.RnsPoly.from_channels(ch, gb)
    b = a.clone()
    b.to_ntt_domain()
    with pytest.raises(gpu.RnsNttError) as e:
        a += b
    assert e.value.kind == "DomainMismatch"
    other = gpu.RnsPoly.zero(gpu.RnsBasis(8, [17, 97]))
    with pytest.raises(gpu.RnsNttError) as e:
        a += other
    assert e.value.kind == "BasisMismatch"


def test_add_neg_mul_kats(gpu, orc):
    gb, ob = _bases(gpu, orc, 8, Q8)
    P = lambda c: gpu.RnsPoly.from_coeffs(c, gb)
    x = P([16, 0, 0, 0, 0, 0, 0, 0])
    x += P([2, 0, 0, 0, 0, 0, 0, 0])
    assert int(x.channels()[0, 0, 0]) == 1
    assert int((-P([3] + [0] * 7)).channels()[0, 0, 0]) == 14
    s = P([1, 1, 0, 0, 0, 0, 0, 0])
    s *= P([1, 1, 0, 0, 0, 0, 0, 0])
    assert np.array_equal(s.channels()[0], ob.from_coeffs([1, 2, 1, 0, 0, 0, 0, 0])) and not s.is_ntt_domain()
    m = P([0] * 7 + [1])
    m *= P([0, 1] + [0] * 6)
    assert np.array_equal(m.channels()[0], ob.from_coeffs([-1] + [0] * 7))


@pytest.mark.parametrize("n,bits,l", [(8, 0, 3), (64, 40, 2), (1024, 62, 2), (4096, 40, 3), (16384, 30, 4)])
def test_arithmetic_matches_oracle(gpu, orc, n, bits, l):
    moduli = Q8 if bits == 0 else orc.generate_primes(bits, l, n)
    gb, ob = _bases(gpu, orc, n, moduli)
    rng = np.random.default_rng(n)
    x, y = uniform_limbs(rng, moduli, n, 2), uniform_limbs(rng, moduli, n, 2)
    a = gpu.RnsPoly.from_channels(x, gb)
    a += gpu.RnsPoly.from_channels(y, gb)
    assert np.array_equal(a.channels()[1], ob.add(x[1], y[1]))
    assert np.array_equal((-gpu.RnsPoly.from_channels(x, gb)).channels()[0], ob.neg(x[0]))
    a = gpu.RnsPoly.from_channels(x, gb)
    a -= gpu.RnsPoly.from_channels(y, gb)
    assert np.array_equal(a.channels()[0], ob.add(x[0], ob.neg(y[0])))
    # coefficient-domain multiply == oracle == (for small n) schoolbook, poly.rs:960-975
    a = gpu.RnsPoly.from_channels(x, gb)
    a *= gpu.RnsPoly.from_channels(y, gb)
    for i in range(2):
        assert np.array_equal(a.channels()[i], ob.mul(x[i], y[i]))
    if n <= 64:
        assert np.array_equal(a.channels()[0], ob.mul_naive(x[0], y[0]))
    # NTT-domain multiply stays in the NTT domain and equals the coefficient-domain product (poly.rs:854-877)
    u, v = gpu.RnsPoly.from_channels(x, gb), gpu.RnsPoly.from_channels(y, gb)
    u.to_ntt_domain()
    v.to_ntt_domain()
    u *= v
    assert u.is_ntt_domain()
    u.to_coeff_domain()
    assert np.array_equal(u.channels(), a.channels())
    # broadcast of a batch-1 right-hand side
    w = gpu.RnsPoly.from_channels(x, gb)
    w *= gpu.RnsPoly.from_channels(y[0], gb)
    assert np.array_equal(w.channels()[1], ob.mul(x[1], y[0]))


@pytest.mark.parametrize("n,bits,l", [(8, 0, 3), (16, 31, 4), (1024, 40, 3), (16384, 30, 3)])
def test_automorphism_matches_oracle(gpu, orc, n, bits, l):
    moduli = Q8 if bits == 0 else orc.generate_primes(bits, l, n)
    gb, ob = _bases(gpu, orc, n, moduli)
    rng = np.random.default_rng(n + 1)
    x = uniform_limbs(rng, moduli, n, 2)
    x[0, :, ::3] = 0  # zero coefficients exercise the reference's skip (poly.rs:523-525)
    p = gpu.RnsPoly.from_channels(x, gb)
    exps = [1, 3, 5, 9, 2 * n - 1, 2 * n + 1, 25 % (2 * n), 2, 4, 6, n, 2 * n, 0, 4 * n]
    for e in exps:
        out = p.automorphism(e)
        ref, dom = ob.automorphism(x[0], e)
        assert np.array_equal(out.channels()[0], ref), f"automorphism exponent {e}"
        assert out.is_ntt_domain() == dom
    pn = p.clone()
    pn.to_ntt_domain()
    out = pn.automorphism(3)  # NTT-domain input is converted first (poly.rs:494-503, 944-957)
    assert not out.is_ntt_domain() and np.array_equal(out.channels()[1], ob.automorphism(x[1], 3)[0])
    assert pn.automorphism(2 * n).is_ntt_domain()  # clone quirk keeps the flag
    for k in (0, 1, 2, 7, -1, -3, n // 2 - 1 if n > 8 else 3):
        out = p.rotate_slots(k)
        assert np.array_equal(out.channels()[1], ob.rotate_slots(x[1], k)[0]), f"rotate_slots {k}"
    one_x = gpu.RnsPoly.from_coeffs([1, 1] + [0] * (n - 2), gb).automorphism(2 * n + 1 if n == 8 else n + 1)
    if n == 8:
        assert np.array_equal(one_x.channels()[0], ob.from_coeffs([1, 1] + [0] * 6))


@pytest.mark.parametrize("n,bits,l", [(8, 0, 3), (1024, 40, 3), (4096, 61, 4), (65536, 61, 3)])
def test_rescale_and_mod_drop(gpu, orc, n, bits, l):
    moduli = Q8 if bits == 0 else orc.generate_primes(bits, l, n)
    gb, ob = _bases(gpu, orc, n, moduli)
    rng = np.random.default_rng(n + 2)
    x = uniform_limbs(rng, moduli, n, 2)
    p = gpu.RnsPoly.from_channels(x, gb)
    child = gb.drop_last(1)
    r = p.rescale_into(child)
    assert r.channel_count() == len(moduli) - 1 and not r.is_ntt_domain()
    for i in range(2):
        assert np.array_equal(r.channels()[i], ob.rescale(x[i]))
    pn = p.clone()
    pn.to_ntt_domain()
    assert np.array_equal(pn.rescale_into(child).channels(), r.channels())  # poly.rs:1036-1049
    assert np.array_equal(p.rescale().channels(), r.channels())
    d = p.mod_drop_last(1)
    assert np.array_equal(d.channels(), x[:, :-1, :])
    if n == 8:
        k = gpu.RnsPoly.from_coeffs([226] + [0] * 7, gb).rescale()
        assert k.channels()[0].tolist() == [[2] + [0] * 7, [2] + [0] * 7]
        with pytest.raises(gpu.RnsNttError) as e:
            gpu.RnsPoly.zero(gpu.RnsBasis(8, [17])).rescale()
        assert e.value.kind == "InvalidModDrop"


def test_to_coeffs_centered(gpu, orc):
    n = 1024
    moduli = orc.generate_primes(40, 3, n)
    gb, ob = _bases(gpu, orc, n, moduli)
    rng = np.random.default_rng(9)
    co = rng.integers(-(2**50), 2**50, size=(2, n), dtype=np.int64)
    p = gpu.RnsPoly.from_coeffs(co, gb)
    assert np.array_equal(p.to_coeffs(), co)
    p.to_ntt_domain()
    assert np.array_equal(p.to_coeffs(), co)


@pytest.mark.parametrize("n,bits,l", [(8, 0, 3), (64, 40, 2), (256, 61, 3), (1024, 30, 2)])
def test_mul_assign_naive_matches_ntt_multiply_and_oracle(gpu, orc, n, bits, l):
    """poly.rs:960-975 (`ntt_mul_matches_naive`): the NTT-based `*=` equals the O(N^2) schoolbook product, here both
    on the device and against the oracle's restatement of mul_assign_naive (poly.rs:339-367)."""
    moduli = [17, 97, 113] if bits == 0 else orc.generate_primes(bits, l, n)
    gb, ob = gpu.RnsBasis(n, moduli), orc.Basis(n, moduli)
    rng = np.random.default_rng(n)
    a, b = uniform_limbs(rng, moduli, n, 3), uniform_limbs(rng, moduli, n, 3)
    fast = gpu.RnsPoly.from_channels(a, gb)
    fast *= gpu.RnsPoly.from_channels(b, gb)
    slow = gpu.RnsPoly.from_channels(a, gb)
    slow.mul_assign_naive(gpu.RnsPoly.from_channels(b, gb))
    assert not slow.is_ntt_domain()
    got = slow.channels()
    assert np.array_equal(got, fast.channels())
    for i in range(3):
        assert np.array_equal(got[i], ob.mul_naive(a[i], b[i]))
    # rhs broadcast over the batch, and the domain rule of the reference's debug_assert
    one = gpu.RnsPoly.from_channels(a, gb)
    one.mul_assign_naive(gpu.RnsPoly.from_channels(b[:1], gb))
    assert np.array_equal(one.channels()[2], ob.mul_naive(a[2], b[0]))
    ntt = gpu.RnsPoly.from_channels(a, gb)
    ntt.to_ntt_domain()
    with pytest.raises(gpu.RnsNttError) as e:
        ntt.mul_assign_naive(ntt)
    assert e.value.kind == "DomainMismatch"


@pytest.mark.parametrize("n,moduli", [(8, [17, 97, 113]), (1024, None)])
def test_ntt_table_accessor_matches_reference_layout(gpu, orc, n, moduli):
    """RnsBasis::ntt_table (basis.rs:112-114): forward/inverse roots, twist/untwist factors and n_inv exactly as
    NttTable::new builds them (basis.rs:21-84), although the device keeps its own table layout."""
    moduli = moduli or orc.generate_primes(40, 2, n)
    gb, ob = gpu.RnsBasis(n, moduli), orc.Basis(n, moduli)
    for ch, q in enumerate(moduli):
        t = gb.ntt_table(ch)
        assert t["modulus"] == q and t["n_inv"] == int(ob.table(ch, "n_inv")[0]) and (t["n_inv"] * n) % q == 1
        assert np.array_equal(t["forward_roots"], ob.table(ch, "forward_roots"))
        assert np.array_equal(t["inverse_roots"], ob.table(ch, "inverse_roots"))
        assert np.array_equal(t["twist_factors"], ob.table(ch, "twist"))
        assert np.array_equal(t["untwist_factors"], ob.table(ch, "untwist"))
        assert int(t["twist_factors"][1]) == gb.psi(ch)
