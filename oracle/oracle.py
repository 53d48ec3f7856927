"""ctypes binding of the CPU oracle (oracle/libckks_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Polynomials are numpy uint64 arrays shaped [L, N] (the reference's `Vec<[u64; N]>`,
src/rings/backends/rns_ntt/poly.rs:26-30); keys are [L(digit), L(limb), N].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libckks_oracle.so")

ERRORS = {
    1: "InvalidDegree",
    2: "EmptyBasis",
    3: "NonNttFriendlyModulus",
    4: "InvalidModDrop",
    5: "ChannelCountMismatch",
    6: "NonReducedCoefficient",
    100: "Panic",
}


class OracleError(Exception):
    def __init__(self, code: int):
        self.code = code
        self.kind = ERRORS.get(code, f"code {code}")
        super().__init__(self.kind)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no external deps)."""
    src = os.path.join(_HERE, "ckks_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libckks_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_get_first_prime_up.restype = C.c_uint64
        L.orc_get_first_prime_up.argtypes = [C.c_uint32, C.c_uint64]
        L.orc_get_first_prime_down.restype = C.c_uint64
        L.orc_get_first_prime_down.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_is_ntt_friendly_prime.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_generate_primes.argtypes = [C.c_int, C.c_int, C.c_uint64, _u64p]
        L.orc_basis_new.argtypes = [C.c_uint64, _u64p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.orc_basis_free.argtypes = [C.c_void_p]
        L.orc_basis_free.restype = None
        L.orc_basis_drop_last.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.orc_basis_degree.restype = C.c_uint64
        L.orc_basis_degree.argtypes = [C.c_void_p]
        L.orc_basis_channel_count.restype = C.c_size_t
        L.orc_basis_channel_count.argtypes = [C.c_void_p]
        L.orc_basis_moduli.argtypes = [C.c_void_p, _u64p]
        L.orc_basis_moduli.restype = None
        L.orc_basis_total_bits.restype = C.c_uint32
        L.orc_basis_total_bits.argtypes = [C.c_void_p]
        L.orc_basis_psi.restype = C.c_uint64
        L.orc_basis_psi.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_basis_table.argtypes = [C.c_void_p, C.c_size_t, C.c_int, _u64p]
        L.orc_basis_table.restype = None
        L.orc_reconstruct_centered_coeff.restype = C.c_int64
        L.orc_reconstruct_centered_coeff.argtypes = [C.c_void_p, _u64p]
        L.orc_from_coeffs.argtypes = [C.c_void_p, _i64p, _u64p]
        L.orc_from_coeffs.restype = None
        L.orc_from_channels_check.argtypes = [C.c_void_p, _u64p, C.c_size_t]
        for name in ("orc_to_ntt_domain", "orc_to_coeff_domain", "orc_neg"):
            getattr(L, name).argtypes = [C.c_void_p, _u64p]
            getattr(L, name).restype = None
        for name in ("orc_add_assign", "orc_mul_assign_naive"):
            getattr(L, name).argtypes = [C.c_void_p, _u64p, _u64p]
            getattr(L, name).restype = None
        L.orc_mul_assign.argtypes = [C.c_void_p, _u64p, _u64p, C.c_int]
        L.orc_mul_assign.restype = None
        L.orc_rescale.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
        L.orc_automorphism.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_uint64, _u64p]
        L.orc_rotate_slots.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int32, _u64p]
        L.orc_to_coeffs.argtypes = [C.c_void_p, _u64p, C.c_int, _i64p]
        L.orc_to_coeffs.restype = None
        L.orc_encrypt.argtypes = [C.c_void_p] + [_u64p] * 8
        L.orc_encrypt.restype = None
        L.orc_decrypt.argtypes = [C.c_void_p] + [_u64p] * 4
        L.orc_decrypt.restype = None
        L.orc_add_ciphertexts.argtypes = [C.c_void_p] + [_u64p] * 6
        L.orc_add_ciphertexts.restype = None
        L.orc_mul_ciphertexts_gadget.argtypes = [C.c_void_p] + [_u64p] * 8
        L.orc_mul_ciphertexts_gadget.restype = None
        L.orc_rescale_ciphertext.argtypes = [C.c_void_p] + [_u64p] * 4 + [C.POINTER(C.c_uint32)]
        L.orc_rotate_ciphertext.argtypes = [C.c_void_p] + [_u64p] * 4 + [C.c_int32, _u64p, _u64p]
        L.orc_rotate_ciphertext.restype = None
        L.orc_gen_public_key.argtypes = [C.c_void_p] + [_u64p] * 4
        L.orc_gen_public_key.restype = None
        L.orc_gen_gadget_relin_key.argtypes = [C.c_void_p] + [_u64p] * 4
        L.orc_gen_gadget_relin_key.restype = None
        L.orc_gen_gadget_rotation_key.argtypes = [C.c_void_p, _u64p, C.c_int32, _u64p, _u64p, _u64p]
        L.orc_gen_gadget_rotation_key.restype = None
        L.orc_encode.argtypes = [C.c_uint64, C.c_uint32, _f64p, C.c_size_t, _i64p]
        L.orc_encode.restype = None
        L.orc_decode.argtypes = [C.c_uint64, C.c_uint32, _i64p, C.c_size_t, _f64p]
        L.orc_decode.restype = None
        L.orc_bench_mul_rescale.restype = C.c_double
        L.orc_bench_mul_rescale.argtypes = [C.c_void_p, C.c_size_t, C.c_int] + [_u64p] * 8
        L.orc_bench_mul_gadget.restype = C.c_double
        L.orc_bench_mul_gadget.argtypes = [C.c_void_p, C.c_size_t, C.c_int] + [_u64p] * 8
        L.orc_bench_rotate.restype = C.c_double
        L.orc_bench_rotate.argtypes = [C.c_void_p, C.c_size_t, C.c_int] + [_u64p] * 4 + [C.c_int32, _u64p, _u64p]
        L.orc_bench_ntt.restype = C.c_double
        L.orc_bench_ntt.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, _u64p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _u(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


# ── src/math ────────────────────────────────────────────────────────────────────────────────────
def is_prime(n: int) -> bool:
    return bool(lib().orc_is_prime(n))


def is_ntt_friendly_prime(p: int, n: int) -> bool:
    return bool(lib().orc_is_ntt_friendly_prime(p, n))


def get_first_prime_up(logq: int, n: int) -> int:
    return int(lib().orc_get_first_prime_up(logq, n))


def get_first_prime_down(bound: int, n: int):
    r = int(lib().orc_get_first_prime_down(bound, n))
    return r if r else None


def generate_primes(bit_size: int, count: int, degree: int) -> list[int]:
    out = np.zeros(max(count, 1), dtype=np.uint64)
    rc = lib().orc_generate_primes(bit_size, count, degree, _p(out))
    if rc:
        raise OracleError(rc)
    return [int(x) for x in out[:count]]


# ── RnsBasis / RnsPoly ──────────────────────────────────────────────────────────────────────────
class Basis:
    """Mirror of RnsBasis<N> (basis.rs:91-181)."""

    def __init__(self, n: int, moduli, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            m = _u(list(moduli))
            rc = lib().orc_basis_new(n, _p(m) if len(m) else None, len(m), C.byref(self._h))
            if rc:
                raise OracleError(rc)
        self.n = int(lib().orc_basis_degree(self._h))
        self.l = int(lib().orc_basis_channel_count(self._h))
        mm = np.zeros(self.l, dtype=np.uint64)
        lib().orc_basis_moduli(self._h, _p(mm))
        self.moduli = [int(x) for x in mm]

    def __del__(self):
        try:
            if self._h:
                lib().orc_basis_free(self._h)
        except Exception:
            pass

    def drop_last(self, k: int) -> "Basis":
        h = C.c_void_p()
        rc = lib().orc_basis_drop_last(self._h, k, C.byref(h))
        if rc:
            raise OracleError(rc)
        return Basis(0, [], _handle=h)

    def total_bits(self) -> int:
        return int(lib().orc_basis_total_bits(self._h))

    def psi(self, ch: int) -> int:
        return int(lib().orc_basis_psi(self._h, ch))

    def table(self, ch: int, which: str) -> np.ndarray:
        idx = {"forward_roots": 0, "inverse_roots": 1, "twist": 2, "untwist": 3, "n_inv": 4}[which]
        out = np.zeros(self.n, dtype=np.uint64)
        lib().orc_basis_table(self._h, ch, idx, _p(out))
        return out[:1] if idx == 4 else out

    def reconstruct_centered_coeff(self, residues) -> int:
        r = _u(residues)
        return int(lib().orc_reconstruct_centered_coeff(self._h, _p(r)))

    # poly.rs
    def from_coeffs(self, coeffs) -> np.ndarray:
        c = np.ascontiguousarray(coeffs, dtype=np.int64)
        if c.shape[0] < self.n:
            raise OracleError(100)  # the reference asserts (poly.rs:50-54)
        out = np.zeros((self.l, self.n), dtype=np.uint64)
        lib().orc_from_coeffs(self._h, c.ctypes.data_as(_i64p), _p(out))
        return out

    def from_channels_check(self, ch: np.ndarray) -> None:
        ch = _u(ch)
        rc = lib().orc_from_channels_check(self._h, _p(ch), ch.shape[0])
        if rc:
            raise OracleError(rc)

    def to_ntt(self, ch) -> np.ndarray:
        o = _u(ch).copy()
        lib().orc_to_ntt_domain(self._h, _p(o))
        return o

    def to_coeff(self, ch) -> np.ndarray:
        o = _u(ch).copy()
        lib().orc_to_coeff_domain(self._h, _p(o))
        return o

    def add(self, a, b) -> np.ndarray:
        o = _u(a).copy()
        lib().orc_add_assign(self._h, _p(o), _p(_u(b)))
        return o

    def neg(self, a) -> np.ndarray:
        o = _u(a).copy()
        lib().orc_neg(self._h, _p(o))
        return o

    def mul(self, a, b, in_ntt: bool = False) -> np.ndarray:
        o = _u(a).copy()
        lib().orc_mul_assign(self._h, _p(o), _p(_u(b)), int(in_ntt))
        return o

    def mul_naive(self, a, b) -> np.ndarray:
        o = _u(a).copy()
        lib().orc_mul_assign_naive(self._h, _p(o), _p(_u(b)))
        return o

    def rescale(self, ch, in_ntt: bool = False) -> np.ndarray:
        out = np.zeros((max(self.l - 1, 0), self.n), dtype=np.uint64)
        rc = lib().orc_rescale(self._h, _p(_u(ch)), int(in_ntt), _p(out) if out.size else None)
        if rc:
            raise OracleError(rc)
        return out

    def automorphism(self, ch, exponent: int, in_ntt: bool = False):
        out = np.zeros((self.l, self.n), dtype=np.uint64)
        d = lib().orc_automorphism(self._h, _p(_u(ch)), int(in_ntt), exponent, _p(out))
        return out, bool(d)

    def rotate_slots(self, ch, k: int, in_ntt: bool = False):
        out = np.zeros((self.l, self.n), dtype=np.uint64)
        d = lib().orc_rotate_slots(self._h, _p(_u(ch)), int(in_ntt), k, _p(out))
        return out, bool(d)

    def to_coeffs(self, ch, in_ntt: bool = False) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.int64)
        lib().orc_to_coeffs(self._h, _p(_u(ch)), int(in_ntt), out.ctypes.data_as(_i64p))
        return out

    # engine.rs
    def encrypt(self, pk_b, pk_a, u, e0, e1, m):
        c0 = np.zeros((self.l, self.n), dtype=np.uint64)
        c1 = np.zeros_like(c0)
        lib().orc_encrypt(self._h, _p(_u(pk_b)), _p(_u(pk_a)), _p(_u(u)), _p(_u(e0)), _p(_u(e1)), _p(_u(m)),
                          _p(c0), _p(c1))
        return c0, c1

    def decrypt(self, c0, c1, s) -> np.ndarray:
        out = np.zeros((self.l, self.n), dtype=np.uint64)
        lib().orc_decrypt(self._h, _p(_u(c0)), _p(_u(c1)), _p(_u(s)), _p(out))
        return out

    def add_ciphertexts(self, a0, a1, b0, b1):
        c0 = np.zeros((self.l, self.n), dtype=np.uint64)
        c1 = np.zeros_like(c0)
        lib().orc_add_ciphertexts(self._h, _p(_u(a0)), _p(_u(a1)), _p(_u(b0)), _p(_u(b1)), _p(c0), _p(c1))
        return c0, c1

    def mul_ciphertexts_gadget(self, a0, a1, b0, b1, rlk_a, rlk_b):
        c0 = np.zeros((self.l, self.n), dtype=np.uint64)
        c1 = np.zeros_like(c0)
        lib().orc_mul_ciphertexts_gadget(self._h, _p(_u(a0)), _p(_u(a1)), _p(_u(b0)), _p(_u(b1)),
                                         _p(_u(rlk_a)), _p(_u(rlk_b)), _p(c0), _p(c1))
        return c0, c1

    def rescale_ciphertext(self, c0, c1):
        o0 = np.zeros((max(self.l - 1, 0), self.n), dtype=np.uint64)
        o1 = np.zeros_like(o0)
        bits = C.c_uint32(0)
        rc = lib().orc_rescale_ciphertext(self._h, _p(_u(c0)), _p(_u(c1)), _p(o0) if o0.size else None,
                                          _p(o1) if o1.size else None, C.byref(bits))
        if rc:
            raise OracleError(rc)
        return o0, o1, int(bits.value)

    def rotate_ciphertext(self, c0, c1, rotk_a, rotk_b, rotation: int):
        o0 = np.zeros((self.l, self.n), dtype=np.uint64)
        o1 = np.zeros_like(o0)
        lib().orc_rotate_ciphertext(self._h, _p(_u(c0)), _p(_u(c1)), _p(_u(rotk_a)), _p(_u(rotk_b)), rotation,
                                    _p(o0), _p(o1))
        return o0, o1

    def gen_public_key(self, s, a, e) -> np.ndarray:
        b = np.zeros((self.l, self.n), dtype=np.uint64)
        lib().orc_gen_public_key(self._h, _p(_u(s)), _p(_u(a)), _p(_u(e)), _p(b))
        return b

    def gen_gadget_relin_key(self, s, a, e) -> np.ndarray:
        b = np.zeros((self.l, self.l, self.n), dtype=np.uint64)
        lib().orc_gen_gadget_relin_key(self._h, _p(_u(s)), _p(_u(a)), _p(_u(e)), _p(b))
        return b

    def gen_gadget_rotation_key(self, s, rotation: int, a, e) -> np.ndarray:
        b = np.zeros((self.l, self.l, self.n), dtype=np.uint64)
        lib().orc_gen_gadget_rotation_key(self._h, _p(_u(s)), rotation, _p(_u(a)), _p(_u(e)), _p(b))
        return b

    # CPU-baseline drivers (seconds)
    def bench_mul_rescale(self, threads, a0, a1, b0, b1, rlk_a, rlk_b):
        a0 = _u(a0)
        count = a0.shape[0]
        o0 = np.zeros((count, self.l - 1, self.n), dtype=np.uint64)
        o1 = np.zeros_like(o0)
        sec = lib().orc_bench_mul_rescale(self._h, count, threads, _p(a0), _p(_u(a1)), _p(_u(b0)), _p(_u(b1)),
                                          _p(_u(rlk_a)), _p(_u(rlk_b)), _p(o0), _p(o1))
        return float(sec), o0, o1

    def bench_mul_gadget(self, threads, a0, a1, b0, b1, rlk_a, rlk_b):
        """mul_ciphertexts_gadget of `count` pairs, unrescaled outputs [count][L][N]."""
        a0 = _u(a0)
        count = a0.shape[0]
        o0 = np.zeros((count, self.l, self.n), dtype=np.uint64)
        o1 = np.zeros_like(o0)
        sec = lib().orc_bench_mul_gadget(self._h, count, threads, _p(a0), _p(_u(a1)), _p(_u(b0)), _p(_u(b1)),
                                         _p(_u(rlk_a)), _p(_u(rlk_b)), _p(o0), _p(o1))
        return float(sec), o0, o1

    def bench_rotate(self, threads, c0, c1, rotk_a, rotk_b, rotation):
        c0 = _u(c0)
        count = c0.shape[0]
        o0 = np.zeros((count, self.l, self.n), dtype=np.uint64)
        o1 = np.zeros_like(o0)
        sec = lib().orc_bench_rotate(self._h, count, threads, _p(c0), _p(_u(c1)), _p(_u(rotk_a)),
                                     _p(_u(rotk_b)), rotation, _p(o0), _p(o1))
        return float(sec), o0, o1

    def bench_ntt(self, threads, polys, inverse=False):
        p = _u(polys).copy()
        sec = lib().orc_bench_ntt(self._h, p.shape[0], threads, int(inverse), _p(p))
        return float(sec), p


# ── encoder (f64) ───────────────────────────────────────────────────────────────────────────────
def encode(n: int, scale_bits: int, values) -> np.ndarray:
    """CkksEncoder::encode / encode_complex -> rounded integer coefficients (ckks_encoder.rs:65-122)."""
    v = np.asarray(values, dtype=np.complex128)
    if v.shape[0] > n // 2:
        raise OracleError(100)
    flat = np.ascontiguousarray(np.stack([v.real, v.imag], axis=-1).reshape(-1), dtype=np.float64)
    out = np.zeros(n, dtype=np.int64)
    lib().orc_encode(n, scale_bits, flat.ctypes.data_as(_f64p), v.shape[0], out.ctypes.data_as(_i64p))
    return out


def decode(n: int, scale_bits: int, coeffs, slots: int) -> np.ndarray:
    """CkksEncoder::decode_complex (ckks_encoder.rs:134-156)."""
    c = np.ascontiguousarray(coeffs, dtype=np.int64)
    out = np.zeros(2 * slots, dtype=np.float64)
    lib().orc_decode(n, scale_bits, c.ctypes.data_as(_i64p), slots, out.ctypes.data_as(_f64p))
    return out[0::2] + 1j * out[1::2]
