"""CPU emulation of the device transform code (tests/emul/ntt_emul.cpp compiles csrc/ntt_tile.cuh,
csrc/modarith.cuh and csrc/tables_host.hpp with g++): the four-step index math, twiddle tables and
lazy-range invariants are checked against the oracle without a GPU."""
import ctypes as C
import os
import random

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul(built):
    lib = C.CDLL(os.path.join(HERE, "emul", "libntt_emul.so"))
    lib.emul_ntt_4step.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int]
    lib.emul_ntt_4step32.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int]
    lib.emul_ntt_pos.argtypes = [C.c_uint32, C.c_int, C.c_int]
    lib.emul_ntt_pos.restype = C.c_uint32
    lib.emul_psi.argtypes = [C.c_uint64, C.c_uint64]
    lib.emul_psi.restype = C.c_uint64
    for name, nargs in (("emul_mulmod", 3), ("emul_mulmod_add", 4), ("emul_barrett_word", 2), ("emul_shoup_lazy8", 3)):
        f = getattr(lib, name)
        f.argtypes = [C.c_uint64] * nargs
        f.restype = C.c_uint64
    return lib


# mode: 0 = what the table builder selects (lazy8 for q < 2^61, Harvey for q < 2^62, strict above),
#       1 = strict, 2 = Harvey lazy
@pytest.mark.parametrize("logn,bits,strict", [(8, 30, 0), (9, 61, 0), (9, 61, 2), (10, 62, 0), (11, 63, 0), (12, 40, 0), (12, 40, 1), (12, 40, 2),
                                                (13, 61, 0), (14, 30, 0), (14, 60, 0), (16, 61, 0), (16, 61, 2)])
def test_four_step_matches_oracle(emul, orc, logn, bits, strict):
    n = 1 << logn
    q = orc.generate_primes(bits, 1, n)[0]
    ob = orc.Basis(n, [q])
    assert emul.emul_psi(n, q) == ob.psi(0)  # same root selection as basis.rs:217-237
    rng = np.random.default_rng(logn * 100 + bits)
    x = rng.integers(0, q, n, dtype=np.uint64)
    x[:3] = [q - 1, 0, 1]
    ref = ob.to_ntt(x[None, :])[0]
    d = x.copy()
    assert emul.emul_ntt_4step(n, q, d.ctypes.data_as(C.POINTER(C.c_uint64)), 0, strict) == 0, "lazy range violated"
    a1 = (logn + 1) // 2
    pos = np.array([emul.emul_ntt_pos(k, a1, logn - a1) for k in range(n)])
    assert np.array_equal(d[pos], ref)
    assert emul.emul_ntt_4step(n, q, d.ctypes.data_as(C.POINTER(C.c_uint64)), 1, strict) == 0
    assert np.array_equal(d, x)


@pytest.mark.parametrize("logn,bits,strict", [(8, 30, 0), (9, 31, 0), (10, 20, 0), (12, 30, 1), (14, 30, 0), (16, 30, 0), (13, 31, 0)])
def test_four_step_32bit_words_match_oracle(emul, orc, logn, bits, strict):
    """The 32-bit word path (all q < 2^31; lazy for q < 2^30, strict for 31-bit primes)."""
    n = 1 << logn
    q = orc.generate_primes(bits, 1, n)[0]
    ob = orc.Basis(n, [q])
    rng = np.random.default_rng(logn * 100 + bits)
    x = rng.integers(0, q, n, dtype=np.uint64)
    x[:3] = [q - 1, 0, 1]
    ref = ob.to_ntt(x[None, :])[0]
    d = x.copy()
    assert emul.emul_ntt_4step32(n, q, d.ctypes.data_as(C.POINTER(C.c_uint64)), 0, strict) == 0
    a1 = (logn + 1) // 2
    pos = np.array([emul.emul_ntt_pos(k, a1, logn - a1) for k in range(n)])
    assert np.array_equal(d[pos], ref)
    assert emul.emul_ntt_4step32(n, q, d.ctypes.data_as(C.POINTER(C.c_uint64)), 1, strict) == 0
    assert np.array_equal(d, x)
    big = orc.generate_primes(40, 1, n)[0]
    assert emul.emul_ntt_4step32(n, big, d.ctypes.data_as(C.POINTER(C.c_uint64)), 0, 0) == -2  # not selected for q >= 2^31


def test_modarith_against_python_integers(emul):
    rnd = random.Random(5)
    qs = [17, 97, 1073741441, 1099511592961, 2305843009211596801, 4611686018427365377, 9223372036854744577]
    for q in qs:
        cases = [(0, 0, 0), (q - 1, q - 1, q - 1), (1, q - 1, 2**64 - 1), (2**63 - 1, 2**63 - 1, 2**64 - 1)]
        cases += [(rnd.randrange(2**63), rnd.randrange(2**63), rnd.randrange(2**64)) for _ in range(300)]
        for a, b, c in cases:
            assert emul.emul_mulmod(a, b, q) == (a * b) % q
            assert emul.emul_mulmod_add(a, b, c, q) == (a * b + c) % q
            assert emul.emul_barrett_word(c, q) == c % q


def test_approximate_shoup_range_and_value(emul):
    """shoup_lazy8 (three partial products for the quotient): congruent to x*w and below 4q for any word x."""
    rnd = random.Random(9)
    for q in (2305843009211596801, 2305843009132953601, 1152921504606584833, 1099511592961, 1073741441):
        edge = [0, 1, q - 1, q, 2 * q, 2**64 - 1, 2**63, 2**32 - 1, 2**32]
        for x in edge + [rnd.randrange(2**64) for _ in range(400)]:
            for w in (1, q - 1, rnd.randrange(q)):
                r = emul.emul_shoup_lazy8(x, w, q)
                assert r % q == (x * w) % q and r < 4 * q


def _wavefronts(addrs_bytes, size):
    """Shared-memory wavefronts of one warp-wide access: 32 banks of 4 bytes; 64-bit accesses go half a warp at a time."""
    groups = [addrs_bytes] if size == 4 else [addrs_bytes[:16], addrs_bytes[16:]]
    total = 0
    for grp in groups:
        banks = {}
        for a in grp:
            for w in range(size // 4):
                banks.setdefault((a // 4 + w) % 32, set()).add(a // 4 + w)
        total += max(len(v) for v in banks.values())
    return total


@pytest.mark.parametrize("size", [8, 4])
def test_ks_pass2_exchange_layout_is_conflict_free(emul, size):
    """The bit-weighted exchange tile of ks_pass2's digit loop (csrc/ntt_tile.cuh tile_addr, E = 3, C = 4): injective
    within the [2^A][C+1] allocation, and every register-window access of every transform size costs the minimum
    number of shared-memory wavefronts (1 per warp for u32, 2 for u64), where the padded layout needs twice as
    many (ncu: 2.4x / 3.3x excess wavefronts before the change).  The padded layout stays conflict-free for the
    transposing store's row-lane read, which is why the epilogue keeps it."""
    emul.emul_tile_addr.argtypes = [C.c_int] * 3
    e, c_cols, ideal = 3, 4, size // 4
    swz = 1
    for a in range(6, 9):
        rows = 1 << a
        addr = [[emul.emul_tile_addr(swz, r, c) for c in range(c_cols)] for r in range(rows)]
        pad = [[emul.emul_tile_addr(0, r, c) for c in range(c_cols)] for r in range(rows)]
        flat = [x for row in addr for x in row]
        assert len(set(flat)) == rows * c_cols and 0 <= min(flat) and max(flat) < rows * (c_cols + 1)
        assert len({x for row in pad for x in row}) == rows * c_cols and max(max(r) for r in pad) < rows * (c_cols + 1)
        nthreads = c_cols << (a - e)
        ns = (a + e - 1) // e
        worse = 0
        for lo in sorted({max(a - (t + 1) * e, 0) for t in range(ns)}):
            for w0 in range(0, nthreads, 32):
                for k in range(1 << e):
                    lanes, lanes_pad = [], []
                    for tid in range(w0, w0 + 32):
                        c, g = tid % c_cols, tid // c_cols
                        r = ((g >> lo) << (lo + e)) | (k << lo) | (g & ((1 << lo) - 1))
                        lanes.append(addr[r][c] * size)
                        lanes_pad.append(pad[r][c] * size)
                    assert _wavefronts(lanes, size) == ideal, (a, lo, w0, k)
                    worse += _wavefronts(lanes_pad, size) > ideal
        assert worse > 0  # the padded layout does conflict on these accesses
        for e0 in range(0, c_cols << a, 32):  # transposed read on the padded layout: lane = consecutive rows of one column
            lanes = [pad[x & (rows - 1)][x >> a] * size for x in range(e0, e0 + 32)]
            assert _wavefronts(lanes, size) == ideal, (a, "transposed", e0)


def test_key_row_permutation_matches_ks_pass2_indexing(emul):
    """ksk_finalize stores row (g * 2^E + k) of every key limb at (k * G + g), G = 2^(a2-E): exactly the index
    `(k * GM::G + g)` ks_pass2 uses for its key tiles, and a bijection on the rows of a limb."""
    emul.emul_perm_row.argtypes = [C.c_uint64, C.c_int, C.c_int]
    emul.emul_perm_row.restype = C.c_uint64
    e = 3
    for a2 in range(4, 9):
        rows, groups = 1 << a2, 1 << (a2 - e)
        seen = set()
        for g in range(groups):
            for k in range(1 << e):
                p = emul.emul_perm_row(g * (1 << e) + k, a2, e)
                assert p == k * groups + g
                seen.add(p)
        assert seen == set(range(rows))
    assert emul.emul_perm_row(5, 8, -1) == 5  # natural order when the key is not permuted (small-N path)


def test_device_arena_bookkeeping(emul):
    """csrc/arena.hpp (the allocator behind every polynomial, key and scratch buffer): random take / give stress on
    fake segments -- live ranges disjoint, aligned, inside a segment; free + live == total; free ranges coalesce;
    double frees refused; everything returns to the 'device' at the end -- with and without a memory cap, and a
    horner_chain-like schedule that must not ask the driver for memory after its first pass."""
    emul.emul_arena_stress.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    emul.emul_arena_chain.argtypes = [C.c_int, C.c_int, C.c_uint64]
    for seed in range(6):
        peak, late = C.c_uint64(0), C.c_int(0)
        assert emul.emul_arena_stress(seed, 6000, 3200, 0, C.byref(peak), C.byref(late)) == 0
        assert peak.value <= 80 << 30  # at most 24 live ranges of <= 3.2 GiB: fragmentation stays bounded
        assert emul.emul_arena_stress(seed, 4000, 3200, 24, C.byref(peak), C.byref(late)) == 0  # 24 GiB device: refusals handled
        assert peak.value <= 24 << 30
        assert emul.emul_arena_stress(seed, 3000, 8, 0, C.byref(peak), C.byref(late)) == 0  # small polynomials: a small arena
        assert peak.value <= 512 << 20
    assert emul.emul_arena_chain(24, 4, 128 << 20) == 0
    assert emul.emul_arena_chain(24, 4, (128 << 20) + 4096) == 0


def _negacyclic(a, b, n):
    """Product of two integer polynomials modulo X^n + 1 (Python integers, schoolbook)."""
    out = [0] * n
    for i, ai in enumerate(a):
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            k = i + j
            if k < n:
                out[k] += ai * bj
            else:
                out[k - n] -= ai * bj
    return out


@pytest.mark.parametrize("n,bits,l,extreme", [(256, 61, 3, False), (256, 61, 3, True), (256, 40, 4, False), (512, 62, 2, True)])
def test_auxiliary_basis_gadget_product_is_exact(emul, orc, n, bits, l, extreme):
    """The algorithm of csrc/aux_ks.cuh on the CPU, with the library's own constants and device arithmetic
    (csrc/aux_crt.cuh compiled by g++) and the emulated 32-bit four-step transforms: sum_i alpha_i (*) key[i][j] computed
    in the auxiliary 30-bit primes and reconstructed by Garner equals the integer negacyclic sums reduced mod q_j --
    including the largest words the reference admits (every digit and key word = q - 1), which sit at the edge of the
    bound the number of auxiliary primes is derived from."""
    emul.emul_aux_primes.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint64)]
    emul.emul_aux_crt.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_int)]
    moduli = [int(q) for q in orc.generate_primes(bits, l, n)]
    mod_arr = (C.c_uint64 * l)(*moduli)
    primes_arr = (C.c_uint64 * 8)()
    K = emul.emul_aux_primes(n, mod_arr, l, primes_arr)
    primes = [int(primes_arr[k]) for k in range(K)]
    assert K >= 2 and all((1 << 29) < p < (1 << 30) and (p - 1) % (2 * n) == 0 for p in primes)
    big_p = 1
    for p in primes:
        big_p *= p
    assert big_p > 2 * l * n * max(moduli) ** 2, "the centred range of P covers every coefficient of the integer sums"
    rng = random.Random(77 + n + bits)
    if extreme:
        digits = [[moduli[i] - 1] * n for i in range(l)]
        keys = [[[moduli[j] - 1] * n for j in range(l)] for _ in range(l)]
        if l > 1:  # mixed signs of the wrapped terms too
            keys[1] = [[(moduli[j] - 1) if (t % 2 == 0) else 0 for t in range(n)] for j in range(l)]
    else:
        digits = [[rng.randrange(moduli[i]) for _ in range(n)] for i in range(l)]
        keys = [[[rng.randrange(moduli[j]) for _ in range(n)] for j in range(l)] for _ in range(l)]
    # the integer sums S_j and what the reference's per-limb arithmetic gives: S_j mod q_j
    sums = []
    for j in range(l):
        s = [0] * n
        for i in range(l):
            s = [x + y for x, y in zip(s, _negacyclic(digits[i], keys[i][j], n))]
        sums.append(s)
    assert max(abs(c) for s in sums for c in s) < big_p // 2

    def ntt(words, p, inverse):
        buf = np.array(words, dtype=np.uint64)
        rc = emul.emul_ntt_4step32(n, p, buf.ctypes.data_as(C.POINTER(C.c_uint64)), int(inverse), 0)
        assert rc == 0
        return buf

    res = np.zeros((l, K, n), dtype=np.uint64)
    for k, p in enumerate(primes):
        x = [ntt([d % p for d in digits[i]], p, False) for i in range(l)]
        for j in range(l):
            acc = np.zeros(n, dtype=np.uint64)
            for i in range(l):
                kk = ntt([w % p for w in keys[i][j]], p, False)
                acc = (acc + (x[i] * kk) % np.uint64(p)) % np.uint64(p)  # products below 2^60
            res[j, k] = ntt(acc.tolist(), p, True)
            assert all(int(res[j, k, e]) == sums[j][e] % p for e in range(0, n, 37)), "residues of the integer sum"
    out = np.zeros((l, n), dtype=np.uint64)
    negs = np.zeros((l, n), dtype=np.int32)
    got_k = emul.emul_aux_crt(n, mod_arr, l, n, res.ctypes.data_as(C.POINTER(C.c_uint64)), out.ctypes.data_as(C.POINTER(C.c_uint64)),
                              negs.ctypes.data_as(C.POINTER(C.c_int)))
    assert got_k == K
    for j in range(l):
        assert [int(v) for v in out[j]] == [c % moduli[j] for c in sums[j]], f"limb {j}"
        assert [int(v) for v in negs[j]] == [1 if c < 0 else 0 for c in sums[j]]
    # and the same words come out of the oracle's NTT-domain arithmetic mod q_j
    ob = orc.Basis(n, moduli)
    for j in range(l):
        want = np.zeros(n, dtype=np.uint64)
        for i in range(l):
            a = np.array([[d % q for d in digits[i]] for q in moduli], dtype=np.uint64)
            b = np.array([[w if jj == j else 0 for w in keys[i][jj]] for jj, q in enumerate(moduli)], dtype=np.uint64)
            prod = ob.mul(a, b)
            want = (want.astype(object) + prod[j].astype(object)) % moduli[j]
        assert [int(v) for v in out[j]] == [int(v) for v in want], f"oracle, limb {j}"


def test_auxiliary_sum_reduction_is_exact(emul):
    """aux_reduce_sum: s mod p for any 64-bit s and every auxiliary prime range the library uses, with the quotient taken
    from a double-precision estimate: exact at the multiples of p (where an estimate off by one either way must be
    corrected), at the ends of the 64-bit range and on random words."""
    emul.emul_aux_reduce_sum.argtypes = [C.c_uint64, C.c_uint32]
    emul.emul_aux_reduce_sum.restype = C.c_uint32
    rng = random.Random(4242)
    primes = [536903681, 759250061, 759169033, 663224321, (1 << 29) + 11, 759250123, (1 << 30) - 35]
    for p in primes:
        cases = [0, 1, p - 1, p, p + 1, 2 * p - 1, 2 * p, (1 << 64) - 1, (1 << 64) - p, (1 << 63), (1 << 53) - 1, (1 << 53), (1 << 53) + 1]
        top = ((1 << 64) - 1) // p
        for _ in range(2000):
            m = rng.randrange(top + 1)
            cases += [m * p, m * p + p - 1, max(0, m * p - 1)]
        cases += [rng.randrange(1 << 64) for _ in range(4000)]
        cases += [32 * (p - 1) * (p - 1), 24 * (p - 1) * (p - 1)]  # the largest sums of 32 / 24 products
        for s in cases:
            if s < (1 << 64):
                assert emul.emul_aux_reduce_sum(s, p) == s % p, (s, p)
