"""ckks-b200: host-side mirror of the reference's RNS-NTT backend over libckks_b200.so.

The classes keep the reference's names and semantics so tests read like the reference's own:

    RnsBasis        src/rings/backends/rns_ntt/basis.rs:91-181
    RnsPoly         src/rings/backends/rns_ntt/poly.rs:26-570 (a *batch* of polynomials, device resident)
    RnsNttError     src/rings/backends/rns_ntt/errors.rs:3-22
    Ciphertext      src/crypto/types.rs:22-35
    CkksEngine      src/crypto/engine.rs (add / mul_ciphertexts_gadget / rescale_ciphertext /
                    rotate_ciphertext / encrypt / decrypt, gadget keys)

Everything is computed by the CUDA library through its C ABI (include/ckks_b200.h); there is no
CPU fallback and nothing here imports the oracle.  Import fails loudly if the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CKKS_B200_LIB: load another build of the same library (kernel variants under development, tools/variants.sh)
LIB_PATH = os.environ.get("CKKS_B200_LIB") or os.path.join(_HERE, "libckks_b200.so")

_ERR_NAMES = {
    1: "InvalidDegree",
    2: "EmptyBasis",
    3: "NonNttFriendlyModulus",
    4: "InvalidModDrop",
    5: "ChannelCountMismatch",
    6: "NonReducedCoefficient",
    20: "BasisMismatch",
    21: "DomainMismatch",
    22: "BatchMismatch",
    23: "LevelMismatch",
    24: "ShortInput",
    30: "BadHandle",
    31: "BadArgument",
    32: "Unsupported",
    40: "CudaError",
    41: "NcclError",
}


class RnsNttError(Exception):
    """Mirrors `RnsNttError` (errors.rs:3-22) plus the ABI's own failure codes."""

    def __init__(self, code: int, detail: str = ""):
        self.code = code
        self.kind = _ERR_NAMES.get(code, f"code {code}")
        super().__init__(self.kind + (": " + detail if detail else ""))


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a).  This package has no CPU fallback."
    )

_lib = C.CDLL(LIB_PATH)
_vp = C.c_void_p
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_pp = C.POINTER(C.c_void_p)


def _sig(name, restype, *argtypes):
    f = getattr(_lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("ckks_status_str", C.c_char_p, C.c_int)
_sig("ckks_last_error", C.c_char_p)
_sig("ckks_device_count", C.c_int)
_sig("ckks_is_prime", C.c_int, C.c_uint64)
_sig("ckks_is_ntt_friendly_prime", C.c_int, C.c_uint64, C.c_uint64)
_sig("ckks_generate_primes", C.c_int, C.c_int, C.c_int, C.c_uint64, _u64p)
_sig("ckks_ctx_create", C.c_int, C.c_uint64, _u64p, C.c_size_t, C.c_int, _pp)
_sig("ckks_ctx_drop_last", C.c_int, _vp, C.c_size_t, _pp)
_sig("ckks_ctx_destroy", C.c_int, _vp)
_sig("ckks_ctx_sync", C.c_int, _vp)
_sig("ckks_ctx_trim", C.c_int, _vp)
_sig("ckks_ctx_set_stream", C.c_int, _vp, _vp)
_sig("ckks_ctx_degree", C.c_uint64, _vp)
_sig("ckks_ctx_channel_count", C.c_size_t, _vp)
_sig("ckks_ctx_moduli", C.c_int, _vp, _u64p)
_sig("ckks_ctx_total_bits", C.c_uint32, _vp)
_sig("ckks_ctx_psi", C.c_uint64, _vp, C.c_size_t)
_sig("ckks_ctx_reconstruct_centered_coeff", C.c_int, _vp, _u64p, _i64p)
_sig("ckks_ctx_ntt_table", C.c_int, _vp, C.c_size_t, C.c_int, _u64p)
_sig("ckks_set_ntt_path", C.c_int, C.c_int)
_sig("ckks_set_unfused", C.c_int, C.c_int)
_sig("ckks_set_word32", C.c_int, C.c_int)
_sig("ckks_set_lazy8", C.c_int, C.c_int)
_sig("ckks_set_tma", C.c_int, C.c_int)
_sig("ckks_set_fused_ntt", C.c_int, C.c_int)
_sig("ckks_set_host_chunk_mib", C.c_int, C.c_int)
_sig("ckks_set_ks_scratch_mib", C.c_int, C.c_int)
_sig("ckks_set_ks_aux", C.c_int, C.c_int)
_sig("ckks_ks_chunk", C.c_size_t, _vp, C.c_size_t)
_sig("ckks_prof_enable", C.c_int, C.c_int)
_sig("ckks_set_nvtx", C.c_int, C.c_int)
_sig("ckks_alloc_stats", C.c_int, _u64p, _u64p, _u64p, _u64p, C.c_int)
_sig("ckks_prof_collect", C.c_size_t, C.c_char_p, C.c_size_t)
_sig("ckks_poly_from_device", C.c_int, _vp, C.c_size_t, _u64p, C.c_int, _pp)
_sig("ckks_poly_device_ptr", C.c_int, _vp, C.POINTER(_u64p))
_sig("ckks_poly_alloc", C.c_int, _vp, C.c_size_t, _pp)
_sig("ckks_poly_from_coeffs", C.c_int, _vp, C.c_size_t, _i64p, C.c_size_t, _pp)
_sig("ckks_poly_from_channels", C.c_int, _vp, C.c_size_t, _u64p, C.c_size_t, C.c_int, _pp)
_sig("ckks_poly_download", C.c_int, _vp, _u64p)
_sig("ckks_poly_clone", C.c_int, _vp, _pp)
_sig("ckks_poly_free", C.c_int, _vp)
_sig("ckks_poly_batch", C.c_size_t, _vp)
_sig("ckks_poly_channel_count", C.c_size_t, _vp)
_sig("ckks_poly_is_ntt_domain", C.c_int, _vp)
_sig("ckks_poly_to_ntt_domain", C.c_int, _vp)
_sig("ckks_poly_to_coeff_domain", C.c_int, _vp)
_sig("ckks_poly_add_assign", C.c_int, _vp, _vp)
_sig("ckks_poly_sub_assign", C.c_int, _vp, _vp)
_sig("ckks_poly_neg", C.c_int, _vp)
_sig("ckks_poly_mul_assign", C.c_int, _vp, _vp)
_sig("ckks_poly_mul_assign_naive", C.c_int, _vp, _vp)
_sig("ckks_poly_mod_drop_last", C.c_int, _vp, _vp, _pp)
_sig("ckks_poly_rescale_into", C.c_int, _vp, _vp, _pp)
_sig("ckks_poly_automorphism", C.c_int, _vp, C.c_uint64, _pp)
_sig("ckks_poly_rotate_slots", C.c_int, _vp, C.c_int32, _pp)
_sig("ckks_poly_to_coeffs", C.c_int, _vp, _i64p)
_sig("ckks_poly_to_coeffs_wide", C.c_int, _vp, _i64p, C.POINTER(C.c_double), C.POINTER(C.c_int))
_sig("ckks_ksk_upload", C.c_int, _vp, _u64p, _u64p, _pp)
_sig("ckks_ksk_from_polys", C.c_int, _vp, _vp, _pp)
_sig("ckks_ksk_free", C.c_int, _vp)
_sig("ckks_gen_gadget_key_b", C.c_int, _vp, _vp, _vp, _vp, _pp)
_sig("ckks_ct_add", C.c_int, _vp, _vp, _vp, _vp, _pp, _pp)
_sig("ckks_ct_mul_relin", C.c_int, _vp, _vp, _vp, _vp, _vp, _pp, _pp)
_sig("ckks_ct_rescale", C.c_int, _vp, _vp, _vp, _pp, _pp, C.POINTER(C.c_uint32))
_sig("ckks_ct_mul_relin_rescale", C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _pp, _pp)
_sig("ckks_ct_rotate", C.c_int, _vp, _vp, _vp, C.c_int32, _pp, _pp)
_sig("ckks_ct_encrypt", C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _pp, _pp)
_sig("ckks_ct_decrypt", C.c_int, _vp, _vp, _vp, _pp)
_sig("ckks_encode", C.c_int, _vp, C.c_uint32, C.c_size_t, C.POINTER(C.c_double), C.c_size_t, _pp)
_sig("ckks_decode", C.c_int, _vp, C.c_uint32, C.c_size_t, C.POINTER(C.c_double))
_sig("ckks_ct_mul_relin_rescale_host", C.c_int, _vp, _vp, _vp, C.c_size_t, _u64p, _u64p, _u64p, _u64p, _u64p, _u64p)
_sig("ckks_ct_rotate_host", C.c_int, _vp, _vp, C.c_int32, C.c_size_t, _u64p, _u64p, _u64p, _u64p)
_sig("ckks_host_alloc", C.c_int, C.c_size_t, _pp)
_sig("ckks_host_free", C.c_int, _vp)
_sig("ckks_launch_count", C.c_uint64)
_sig("ckks_launch_table", C.c_size_t, C.c_char_p, C.c_size_t)
_sig("ckks_bench_modmul_peak", C.c_double, C.c_int, C.c_int)
_sig("ckks_bench_mac32_peak", C.c_double, C.c_int, C.c_int)
_sig("ckks_bench_mac32_peak_ex", C.c_double, C.c_int, C.c_int, C.c_int)
_sig("ckks_bench_host_copy", C.c_double, C.c_int, _vp, C.c_size_t, C.c_int, _vp, C.c_size_t, C.c_int, C.c_int)
_sig("ckks_comm_init", C.c_int, C.c_int, C.POINTER(C.c_int), C.c_uint64, _u64p, C.c_size_t, _pp)
_sig("ckks_comm_destroy", C.c_int, _vp)
_sig("ckks_comm_drop_last", C.c_int, _vp, C.c_size_t, _pp)
_sig("ckks_comm_size", C.c_int, _vp)
_sig("ckks_comm_ctx", _vp, _vp, C.c_int)
_sig("ckks_comm_ksk_upload", C.c_int, _vp, _u64p, _u64p, _pp)
_sig("ckks_comm_ksk_free", C.c_int, _vp)
_sig("ckks_comm_ct_mul_relin_rescale_host", C.c_int, _vp, _vp, C.c_size_t, _u64p, _u64p, _u64p, _u64p, _u64p, _u64p)
_sig("ckks_comm_ct_rotate_host", C.c_int, _vp, _vp, C.c_int32, C.c_size_t, _u64p, _u64p, _u64p, _u64p)
_sig("ckks_lshard_create", C.c_int, C.c_uint64, _u64p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_size_t, _pp)
_sig("ckks_lshard_destroy", C.c_int, _vp)
_sig("ckks_lshard_drop_last", C.c_int, _vp, _pp)
_sig("ckks_lshard_local_ctx", _vp, _vp)
_sig("ckks_lshard_channel_count", C.c_size_t, _vp)
_sig("ckks_lshard_chunk", C.c_size_t, _vp)
_sig("ckks_lshard_ipc_size", C.c_size_t)
_sig("ckks_lshard_ipc_export", C.c_int, _vp, _vp)
_sig("ckks_lshard_ipc_import", C.c_int, _vp, _vp)
_sig("ckks_lshard_connect_local", C.c_int, _pp, C.c_int)
_sig("ckks_lshard_ksk_upload", C.c_int, _vp, _u64p, _u64p, _pp)
_sig("ckks_lshard_ct_mul_relin_rescale", C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pp, _pp)
_sig("ckks_lshard_mul_phase", C.c_int, _vp, C.c_int, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int)
_sig("ckks_lshard_barrier", C.c_int, _vp)
_sig("ckks_lshard_ct_rotate", C.c_int, _vp, _vp, _vp, _vp, C.c_int32, _pp, _pp)
_sig("ckks_lshard_ks_phase", C.c_int, _vp, C.c_int, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp, C.c_int)
_sig("ckks_lshard_barrier_local", C.c_int, _pp, C.c_int)
_sig("ckks_lshard_buffers", C.c_int, _vp, C.POINTER(_u64p), C.POINTER(C.c_size_t), C.POINTER(_u64p), C.POINTER(C.c_size_t))
_sig("ckks_lshard_check", C.c_int, _vp)
_sig("ckks_lshard_set_timeout_ms", C.c_int, _vp, C.c_uint64)
_sig("ckks_lshard_set_exchange", C.c_int, _vp, C.c_int)

# Every symbol include/ckks_b200.h declares (tests/test_abi.py checks the header against this list).
ABI_SYMBOLS = sorted(n for n in dir(_lib) if n.startswith("ckks_")) or []


def _check(rc: int):
    if rc != 0:
        detail = _lib.ckks_last_error().decode() if rc in (40, 41) else ""
        raise RnsNttError(rc, detail)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u64p)


def device_count() -> int:
    return int(_lib.ckks_device_count())


def launch_count() -> int:
    return int(_lib.ckks_launch_count())


def alloc_stats(reset: bool = False) -> dict:
    """Block-cache hits, driver-pool allocations and the host time the latter took (requests >= 1 MiB)."""
    v = [C.c_uint64(0) for _ in range(4)]
    _check(_lib.ckks_alloc_stats(C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), 1 if reset else 0))
    return {"cache_hits": v[0].value, "pool_allocs": v[1].value, "pool_ms": v[2].value / 1e3, "pool_max_ms": v[3].value / 1e3}


def launch_table() -> dict:
    n = _lib.ckks_launch_table(None, 0)
    buf = C.create_string_buffer(int(n))
    _lib.ckks_launch_table(buf, n)
    out = {}
    for line in buf.value.decode().splitlines():
        k, v = line.split("=")
        out[k] = int(v)
    return out


def set_ntt_path(path: int):
    """0 = automatic, 1 = small single-CTA NTT (N <= 2048), 2 = four-step (N >= 256)."""
    _check(_lib.ckks_set_ntt_path(path))


def set_word32(on: bool):
    """Test hook: allow (default) or forbid the 32-bit word path for contexts created afterwards."""
    _check(_lib.ckks_set_word32(int(on)))


def set_fused_ntt(on: bool):
    """Test hook: single-kernel (default) or two-pass transforms for 2^12 <= N <= 2^14."""
    _check(_lib.ckks_set_fused_ntt(int(on)))


def set_lazy8(on: bool):
    """Test hook: approximate-quotient butterflies (default, q < 2^61) or Harvey butterflies."""
    _check(_lib.ckks_set_lazy8(int(on)))


def set_tma(on: bool):
    """Test hook: TMA (default) or cp.async staging in the fused key-switch kernel."""
    _check(_lib.ckks_set_tma(int(on)))


def set_ks_aux(mode: int):
    """Gadget product through auxiliary 30-bit NTT primes (csrc/aux_ks.cuh): 0 never, 1 automatic (default: 64-bit
    four-step path, 11 limbs and more), 2 whenever possible (test hook).  Read at key upload and at every product."""
    _check(_lib.ckks_set_ks_aux(int(mode)))


def set_unfused(on: bool):
    """Test hook: run the key-switch from its unfused building blocks instead of the fused kernels."""
    _check(_lib.ckks_set_unfused(int(on)))


def modmul_peak(device: int = 0, iters: int = 4096) -> float:
    return float(_lib.ckks_bench_modmul_peak(device, iters))


def mac32_peak(device: int = 0, iters: int = 4096) -> float:
    """32 x 32 -> 64-bit multiply-accumulates per second on the whole device (the roof of aux_mac)."""
    return float(_lib.ckks_bench_mac32_peak(device, iters))


def mac32_peak_ex(device: int = 0, iters: int = 4096, vary: bool = True) -> float:
    return float(_lib.ckks_bench_mac32_peak_ex(device, iters, int(vary)))


# ── src/math ─────────────────────────────────────────────────────────────────────────────────────
def is_prime(n: int) -> bool:
    return bool(_lib.ckks_is_prime(n))


def is_ntt_friendly_prime(p: int, n: int) -> bool:
    return bool(_lib.ckks_is_ntt_friendly_prime(p, n))


def generate_primes(bit_size: int, count: int, degree: int) -> list:
    """utils.rs:47-80.  Raises where the reference panics."""
    out = np.zeros(max(count, 1), dtype=np.uint64)
    rc = _lib.ckks_generate_primes(bit_size, count, degree, _ptr(out))
    if rc:
        raise RnsNttError(rc, "generate_primes: not enough primes / bad arguments")
    return [int(x) for x in out[:count]]


# ── RnsBasis ─────────────────────────────────────────────────────────────────────────────────────
class RnsBasis:
    """`Arc<RnsBasis<N>>` (basis.rs:91-181) with its device tables."""

    def __init__(self, degree: int, moduli, device: int = 0, _handle=None, _owned: bool = True):
        self._owned = _owned  # False: a handle borrowed from a LimbShard
        if _handle is not None:
            self._h = _handle
        else:
            m = _u64(list(moduli))
            h = _vp()
            _check(_lib.ckks_ctx_create(degree, _ptr(m) if len(m) else None, len(m), device, C.byref(h)))
            self._h = h
        self.degree = int(_lib.ckks_ctx_degree(self._h))
        self.device = device

    @classmethod
    def new(cls, degree: int, moduli, device: int = 0) -> "RnsBasis":
        return cls(degree, moduli, device)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                if self._owned:
                    _lib.ckks_ctx_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def moduli(self) -> list:
        out = np.zeros(self.channel_count(), dtype=np.uint64)
        _check(_lib.ckks_ctx_moduli(self._h, _ptr(out)))
        return [int(x) for x in out]

    def channel_count(self) -> int:
        return int(_lib.ckks_ctx_channel_count(self._h))

    def drop_last(self, drop_count: int) -> "RnsBasis":
        h = _vp()
        _check(_lib.ckks_ctx_drop_last(self._h, drop_count, C.byref(h)))
        return RnsBasis(0, [], self.device, _handle=h)

    def total_bits(self) -> int:
        return int(_lib.ckks_ctx_total_bits(self._h))

    def psi(self, channel: int) -> int:
        return int(_lib.ckks_ctx_psi(self._h, channel))

    def ntt_table(self, channel: int) -> dict:
        """`NttTable<N>` of one channel in the reference's layout (basis.rs:6-17)."""
        out = {}
        for idx, name in enumerate(("forward_roots", "inverse_roots", "twist_factors", "untwist_factors")):
            a = np.zeros(self.degree, dtype=np.uint64)
            _check(_lib.ckks_ctx_ntt_table(self._h, channel, idx, _ptr(a)))
            out[name] = a
        a = np.zeros(1, dtype=np.uint64)
        _check(_lib.ckks_ctx_ntt_table(self._h, channel, 4, _ptr(a)))
        out["n_inv"] = int(a[0])
        out["modulus"] = self.moduli()[channel]
        return out

    def reconstruct_centered_coeff(self, residues) -> int:
        r = _u64(residues)
        out = C.c_int64(0)
        _check(_lib.ckks_ctx_reconstruct_centered_coeff(self._h, _ptr(r), C.byref(out)))
        return int(out.value)

    def sync(self):
        _check(_lib.ckks_ctx_sync(self._h))

    def trim(self):
        """Return the scratch cached in the context's private memory pool to the device."""
        _check(_lib.ckks_ctx_trim(self._h))

    def set_stream(self, cuda_stream: int):
        _check(_lib.ckks_ctx_set_stream(self._h, _vp(cuda_stream)))


# ── RnsPoly ──────────────────────────────────────────────────────────────────────────────────────
class RnsPoly:
    """A batch of `RnsPoly<N>` (poly.rs:26-30) resident in HBM.

    Host arrays are [batch, L, N] uint64 (a 2-D [L, N] array is a batch of one)."""

    def __init__(self, handle, basis: RnsBasis):
        self._h = handle
        self._basis = basis

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.ckks_poly_free(self._h)
                self._h = None
        except Exception:
            pass

    # constructors ------------------------------------------------------------------------------
    @classmethod
    def zero(cls, basis: RnsBasis, batch: int = 1) -> "RnsPoly":
        h = _vp()
        _check(_lib.ckks_poly_alloc(basis._h, batch, C.byref(h)))
        return cls(h, basis)

    @classmethod
    def from_coeffs(cls, coeffs, basis: RnsBasis) -> "RnsPoly":
        c = np.ascontiguousarray(coeffs, dtype=np.int64)
        if c.ndim == 1:
            c = c[None, :]
        h = _vp()
        _check(_lib.ckks_poly_from_coeffs(basis._h, c.shape[0], c.ctypes.data_as(_i64p), c.shape[1], C.byref(h)))
        return cls(h, basis)

    @classmethod
    def from_channels(cls, channels, basis: RnsBasis, is_ntt_domain: bool = False) -> "RnsPoly":
        ch = _u64(channels)
        if ch.ndim == 2:
            ch = ch[None, :, :]
        if ch.shape[2] != basis.degree:
            raise RnsNttError(1, "channel length != N")
        h = _vp()
        _check(_lib.ckks_poly_from_channels(basis._h, ch.shape[0], _ptr(ch), ch.shape[1], int(is_ntt_domain), C.byref(h)))
        return cls(h, basis)

    # accessors ---------------------------------------------------------------------------------
    def channels(self) -> np.ndarray:
        out = np.zeros((self.batch(), self.channel_count(), self._basis.degree), dtype=np.uint64)
        _check(_lib.ckks_poly_download(self._h, _ptr(out)))
        return out

    def basis(self) -> RnsBasis:
        return self._basis

    def context(self) -> RnsBasis:
        return self._basis

    def batch(self) -> int:
        return int(_lib.ckks_poly_batch(self._h))

    def channel_count(self) -> int:
        return int(_lib.ckks_poly_channel_count(self._h))

    def is_ntt_domain(self) -> bool:
        return bool(_lib.ckks_poly_is_ntt_domain(self._h))

    def clone(self) -> "RnsPoly":
        h = _vp()
        _check(_lib.ckks_poly_clone(self._h, C.byref(h)))
        return RnsPoly(h, self._basis)

    # transforms and arithmetic -----------------------------------------------------------------
    def to_ntt_domain(self):
        _check(_lib.ckks_poly_to_ntt_domain(self._h))

    def to_coeff_domain(self):
        _check(_lib.ckks_poly_to_coeff_domain(self._h))

    def __iadd__(self, rhs: "RnsPoly"):
        _check(_lib.ckks_poly_add_assign(self._h, rhs._h))
        return self

    def __isub__(self, rhs: "RnsPoly"):
        _check(_lib.ckks_poly_sub_assign(self._h, rhs._h))
        return self

    def __imul__(self, rhs: "RnsPoly"):
        _check(_lib.ckks_poly_mul_assign(self._h, rhs._h))
        return self

    def mul_assign_naive(self, rhs: "RnsPoly"):
        """Schoolbook O(N^2) product (poly.rs:339-367), coefficient domain only."""
        _check(_lib.ckks_poly_mul_assign_naive(self._h, rhs._h))

    def __neg__(self) -> "RnsPoly":
        r = self.clone()
        _check(_lib.ckks_poly_neg(r._h))
        return r

    def mod_drop_last(self, drop_count: int = 1, basis: RnsBasis | None = None) -> "RnsPoly":
        child = basis if basis is not None else self._basis.drop_last(drop_count)
        h = _vp()
        _check(_lib.ckks_poly_mod_drop_last(self._h, child._h, C.byref(h)))
        return RnsPoly(h, child)

    def rescale_into(self, new_basis: RnsBasis) -> "RnsPoly":
        h = _vp()
        _check(_lib.ckks_poly_rescale_into(self._h, new_basis._h, C.byref(h)))
        return RnsPoly(h, new_basis)

    def rescale(self) -> "RnsPoly":
        if self.channel_count() < 2:
            raise RnsNttError(4)
        return self.rescale_into(self._basis.drop_last(1))

    def automorphism(self, exponent: int) -> "RnsPoly":
        h = _vp()
        _check(_lib.ckks_poly_automorphism(self._h, exponent, C.byref(h)))
        return RnsPoly(h, self._basis)

    def rotate_slots(self, k: int) -> "RnsPoly":
        h = _vp()
        _check(_lib.ckks_poly_rotate_slots(self._h, k, C.byref(h)))
        return RnsPoly(h, self._basis)

    def to_coeffs(self) -> np.ndarray:
        out = np.zeros((self.batch(), self._basis.degree), dtype=np.int64)
        _check(_lib.ckks_poly_to_coeffs(self._h, out.ctypes.data_as(_i64p)))
        return out


# ── keys and ciphertexts ─────────────────────────────────────────────────────────────────────────
def _to_coeffs_wide(self):
    """Centred CRT for a basis of any size (Garner mixed radix on the device): (i64 [batch, N], f64 [batch, N],
    overflow).  i64 equals `to_coeffs` (basis.rs:158-180) bit for bit while Q < 2^128."""
    n = self._basis.degree
    oi = np.zeros((self.batch(), n), dtype=np.int64)
    of = np.zeros((self.batch(), n), dtype=np.float64)
    ov = C.c_int(0)
    _check(_lib.ckks_poly_to_coeffs_wide(self._h, oi.ctypes.data_as(_i64p), of.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ov)))
    return oi, of, bool(ov.value)


RnsPoly.to_coeffs_wide = _to_coeffs_wide


class GadgetKey:
    """`RnsGadgetRelinKey` / `RnsGadgetRotationKey` (engine.rs:225-253), transformed and resident."""

    def __init__(self, handle, basis: RnsBasis, rotation: int = 0):
        self._h = handle
        self._basis = basis
        self.rotation = rotation

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.ckks_ksk_free(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def upload(cls, basis: RnsBasis, a, b, rotation: int = 0) -> "GadgetKey":
        a = _u64(a)
        b = _u64(b)
        l, n = basis.channel_count(), basis.degree
        if a.shape != (l, l, n) or b.shape != (l, l, n):
            raise RnsNttError(5, "gadget key must be [L, L, N]")
        h = _vp()
        _check(_lib.ckks_ksk_upload(basis._h, _ptr(a), _ptr(b), C.byref(h)))
        return cls(h, basis, rotation)

    @classmethod
    def from_polys(cls, a: RnsPoly, b: RnsPoly, rotation: int = 0) -> "GadgetKey":
        h = _vp()
        _check(_lib.ckks_ksk_from_polys(a._h, b._h, C.byref(h)))
        return cls(h, a.basis(), rotation)


class Ciphertext:
    """`Ciphertext` (types.rs:22-35): (c0, c1, logp, logq); c0/c1 are batched RnsPoly."""

    def __init__(self, c0: RnsPoly, c1: RnsPoly, logp: int, logq: int):
        self.c0, self.c1, self.logp, self.logq = c0, c1, logp, logq


class CkksEngine:
    """RnsPoly-specific operations of `CkksEngine` (engine.rs:84-151, 255-540) on device batches.

    Randomness stays on the host (north_star): the sampled polynomials are arguments."""

    @staticmethod
    def encrypt(pk_b: RnsPoly, pk_a: RnsPoly, u: RnsPoly, e0: RnsPoly, e1: RnsPoly, m: RnsPoly, logp: int, logq: int):
        h0, h1 = _vp(), _vp()
        _check(_lib.ckks_ct_encrypt(pk_b._h, pk_a._h, u._h, e0._h, e1._h, m._h, C.byref(h0), C.byref(h1)))
        b = u.basis()
        return Ciphertext(RnsPoly(h0, b), RnsPoly(h1, b), logp, logq)

    @staticmethod
    def decrypt(ct: Ciphertext, s: RnsPoly) -> RnsPoly:
        h = _vp()
        _check(_lib.ckks_ct_decrypt(ct.c0._h, ct.c1._h, s._h, C.byref(h)))
        return RnsPoly(h, ct.c0.basis())

    @staticmethod
    def add_ciphertexts(a: Ciphertext, b: Ciphertext) -> Ciphertext:
        if a.logp != b.logp or a.logq != b.logq:  # engine.rs:135-136
            raise RnsNttError(23)
        h0, h1 = _vp(), _vp()
        _check(_lib.ckks_ct_add(a.c0._h, a.c1._h, b.c0._h, b.c1._h, C.byref(h0), C.byref(h1)))
        bs = a.c0.basis()
        return Ciphertext(RnsPoly(h0, bs), RnsPoly(h1, bs), a.logp, a.logq)

    @staticmethod
    def mul_ciphertexts_gadget(a: Ciphertext, b: Ciphertext, rlk: GadgetKey) -> Ciphertext:
        if a.logq != b.logq:  # engine.rs:478
            raise RnsNttError(23)
        h0, h1 = _vp(), _vp()
        _check(_lib.ckks_ct_mul_relin(a.c0._h, a.c1._h, b.c0._h, b.c1._h, rlk._h, C.byref(h0), C.byref(h1)))
        bs = a.c0.basis()
        return Ciphertext(RnsPoly(h0, bs), RnsPoly(h1, bs), a.logp + b.logp, a.logq)

    @staticmethod
    def rescale_ciphertext(ct: Ciphertext, new_basis: RnsBasis | None = None) -> Ciphertext:
        if ct.c0.channel_count() < 2:
            raise RnsNttError(4)
        child = new_basis if new_basis is not None else ct.c0.basis().drop_last(1)
        h0, h1 = _vp(), _vp()
        bits = C.c_uint32(0)
        _check(_lib.ckks_ct_rescale(ct.c0._h, ct.c1._h, child._h, C.byref(h0), C.byref(h1), C.byref(bits)))
        return Ciphertext(RnsPoly(h0, child), RnsPoly(h1, child), ct.logp - bits.value, ct.logq - bits.value)

    @staticmethod
    def mul_relin_rescale(a: Ciphertext, b: Ciphertext, rlk: GadgetKey, new_basis: RnsBasis | None = None) -> Ciphertext:
        """mul_ciphertexts_gadget followed by rescale_ciphertext in one device call."""
        if a.logq != b.logq:
            raise RnsNttError(23)
        child = new_basis if new_basis is not None else a.c0.basis().drop_last(1)
        h0, h1 = _vp(), _vp()
        _check(_lib.ckks_ct_mul_relin_rescale(a.c0._h, a.c1._h, b.c0._h, b.c1._h, rlk._h, child._h, C.byref(h0), C.byref(h1)))
        bits = a.c0.basis().moduli()[-1].bit_length()
        return Ciphertext(RnsPoly(h0, child), RnsPoly(h1, child), a.logp + b.logp - bits, a.logq - bits)

    @staticmethod
    def rotate_ciphertext(ct: Ciphertext, rotk: GadgetKey) -> Ciphertext:
        h0, h1 = _vp(), _vp()
        _check(_lib.ckks_ct_rotate(ct.c0._h, ct.c1._h, rotk._h, rotk.rotation, C.byref(h0), C.byref(h1)))
        bs = ct.c0.basis()
        return Ciphertext(RnsPoly(h0, bs), RnsPoly(h1, bs), ct.logp, ct.logq)

    @staticmethod
    def gadget_key_b(s: RnsPoly, target: RnsPoly, a: RnsPoly, e: RnsPoly) -> RnsPoly:
        """b_i = -(a_i s) + e_i + [target in limb i] (engine.rs:318-327 / :378-387)."""
        h = _vp()
        _check(_lib.ckks_gen_gadget_key_b(s._h, target._h, a._h, e._h, C.byref(h)))
        return RnsPoly(h, s.basis())


class Plaintext:
    """`Plaintext` (types.rs:4-20): (poly, scale_bits, slots)."""

    def __init__(self, poly: RnsPoly, scale_bits: int, slots: int):
        self.poly, self.scale_bits, self.slots = poly, scale_bits, slots


class CkksEncoder:
    """`CkksEncoder<N>` (ckks_encoder.rs:32-157) evaluated on the device in O(N log N)."""

    def __init__(self, degree: int, scale_bits: int):
        if degree & (degree - 1) or scale_bits <= 0:
            raise RnsNttError(31, "CkksEncoder: DEGREE must be a power of two and scale_bits positive")
        self.degree, self.scale_bits = degree, scale_bits

    def scale_factor(self) -> float:
        return 2.0 ** self.scale_bits

    def max_slots(self) -> int:
        return self.degree // 2

    def encode_complex(self, values, basis: RnsBasis) -> Plaintext:
        v = np.ascontiguousarray(values, dtype=np.complex128)
        if v.ndim == 1:
            v = v[None, :]
        flat = np.ascontiguousarray(v.view(np.float64))
        h = _vp()
        _check(_lib.ckks_encode(basis._h, self.scale_bits, v.shape[0], flat.ctypes.data_as(C.POINTER(C.c_double)), v.shape[1], C.byref(h)))
        return Plaintext(RnsPoly(h, basis), self.scale_bits, v.shape[1])

    def encode(self, values, basis: RnsBasis) -> Plaintext:
        return self.encode_complex(np.asarray(values, dtype=np.float64).astype(np.complex128), basis)

    def decode_complex(self, pt: Plaintext) -> np.ndarray:
        out = np.zeros((pt.poly.batch(), pt.slots), dtype=np.complex128)
        _check(_lib.ckks_decode(pt.poly._h, pt.scale_bits, pt.slots, out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def decode(self, pt: Plaintext) -> np.ndarray:
        return self.decode_complex(pt).real


# ── host-buffer entry points ─────────────────────────────────────────────────────────────────────
class PinnedBuffer:
    """Page-locked host memory viewed as a numpy uint64 array."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        n = int(np.prod(self.shape))
        p = _vp()
        _check(_lib.ckks_host_alloc(n * 8, C.byref(p)))
        self._p = p
        buf = (C.c_uint64 * n).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.uint64).reshape(self.shape)

    def __del__(self):
        try:
            if getattr(self, "_p", None):
                self.array = None
                _lib.ckks_host_free(self._p)
                self._p = None
        except Exception:
            pass


def mul_relin_rescale_host(basis: RnsBasis, child: RnsBasis, rlk: GadgetKey, a0, a1, b0, b1, o0, o1):
    """Host [batch, L, N] arrays in, host [batch, L-1, N] arrays out (copies inside)."""
    batch = a0.shape[0]
    _check(_lib.ckks_ct_mul_relin_rescale_host(basis._h, child._h, rlk._h, batch, _ptr(a0), _ptr(a1), _ptr(b0), _ptr(b1),
                                               _ptr(o0), _ptr(o1)))


def rotate_host(basis: RnsBasis, rotk: GadgetKey, c0, c1, o0, o1):
    _check(_lib.ckks_ct_rotate_host(basis._h, rotk._h, rotk.rotation, c0.shape[0], _ptr(c0), _ptr(c1), _ptr(o0), _ptr(o1)))


class BatchShard:
    """The batch-sharded multi-GPU group (`ckks_comm_*`): one process, one context per device, a host batch cut into
    contiguous shares.  Replaces the reference's serial loop over a Vec<Ciphertext> (horner_chain.rs:211-278)."""

    def __init__(self, degree: int, moduli, devices=None, _handle=None, _parent=None):
        self._parent = _parent
        if _handle is not None:
            self._h = _handle
            return
        if devices is None:
            devices = list(range(max(1, device_count())))
        devs = (C.c_int * len(devices))(*devices)
        arr = (C.c_uint64 * len(moduli))(*moduli)
        h = _vp()
        _check(_lib.ckks_comm_init(len(devices), devs, degree, arr, len(moduli), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.ckks_comm_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def size(self) -> int:
        return int(_lib.ckks_comm_size(self._h))

    def basis(self, i: int = 0) -> "RnsBasis":
        """Borrowed view of device slot i's context."""
        h = _lib.ckks_comm_ctx(self._h, i)
        if not h:
            raise RnsNttError(30)
        b = RnsBasis(int(_lib.ckks_ctx_degree(h)), [], _handle=_vp(h), _owned=False)
        b._keepalive = self
        return b

    def drop_last(self, k: int = 1) -> "BatchShard":
        h = _vp()
        _check(_lib.ckks_comm_drop_last(self._h, k, C.byref(h)))
        return BatchShard(0, [], _handle=h, _parent=self)

    def upload_key(self, a, b, rotation: int = 0) -> "BatchShardKey":
        a, b = _u64(a), _u64(b)
        h = _vp()
        _check(_lib.ckks_comm_ksk_upload(self._h, _ptr(a), _ptr(b), C.byref(h)))
        return BatchShardKey(h, self, rotation)

    def mul_relin_rescale_host(self, rlk: "BatchShardKey", a0, a1, b0, b1, o0, o1):
        _check(_lib.ckks_comm_ct_mul_relin_rescale_host(self._h, rlk._h, a0.shape[0], _ptr(a0), _ptr(a1), _ptr(b0), _ptr(b1), _ptr(o0), _ptr(o1)))

    def rotate_host(self, rotk: "BatchShardKey", c0, c1, o0, o1):
        _check(_lib.ckks_comm_ct_rotate_host(self._h, rotk._h, rotk.rotation, c0.shape[0], _ptr(c0), _ptr(c1), _ptr(o0), _ptr(o1)))


class BatchShardKey:
    def __init__(self, handle, comm: BatchShard, rotation: int = 0):
        self._h, self._comm, self.rotation = handle, comm, rotation

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.ckks_comm_ksk_free(self._h)
                self._h = None
        except Exception:
            pass


# ── optional limb-sharded mode ────────────────────────────────────────────────────
def owned_limbs(channel_count: int, rank: int, world: int) -> list:
    """Basis limbs GPU `rank` of `world` holds: j with j mod world == rank (balanced under drop_last)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, channel_count, world))


class LimbShard:
    """One GPU's share of a limb-sharded batch: every ciphertext's limbs j = rank (mod world) and the
    matching slices of the gadget keys.  `mul_relin_rescale` is mul_ciphertexts_gadget + rescale_ciphertext
    (engine.rs:473-539, 263-282); the digits are all-gathered and the dropped limb broadcast by stores into
    peer HBM from the producing kernels.  Every rank of the group makes the same calls."""

    def __init__(self, degree: int, moduli, rank: int, world: int, device: int = 0, chunk: int = 0, _handle=None, _parent=None,
                 _moduli=None):
        self.rank, self.world, self.device, self.degree = rank, world, device, degree
        self._parent = _parent  # children share the parent's exchange buffers: keep it alive
        if _handle is not None:
            self._h, self._moduli = _handle, list(_moduli)
        else:
            m = _u64(list(moduli))
            h = _vp()
            _check(_lib.ckks_lshard_create(degree, _ptr(m) if len(m) else None, len(m), rank, world, device, chunk, C.byref(h)))
            self._h, self._moduli = h, [int(x) for x in m]
        self._basis = RnsBasis(0, [], device, _handle=_vp(_lib.ckks_lshard_local_ctx(self._h)), _owned=False)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._basis._h = None
                _lib.ckks_lshard_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def moduli(self) -> list:
        """The whole basis at this level."""
        return list(self._moduli)

    def channel_count(self) -> int:
        return int(_lib.ckks_lshard_channel_count(self._h))

    def owned(self) -> list:
        return owned_limbs(self.channel_count(), self.rank, self.world)

    def local_basis(self) -> RnsBasis:
        return self._basis

    def chunk(self) -> int:
        return int(_lib.ckks_lshard_chunk(self._h))

    def drop_last(self) -> "LimbShard":
        h = _vp()
        _check(_lib.ckks_lshard_drop_last(self._h, C.byref(h)))
        return LimbShard(self.degree, None, self.rank, self.world, self.device, _handle=h, _parent=self, _moduli=self._moduli[:-1])

    # wiring -------------------------------------------------------------------------------------
    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(int(_lib.ckks_lshard_ipc_size()))
        _check(_lib.ckks_lshard_ipc_export(self._h, buf))
        return buf.raw

    def connect(self, handles):
        """handles: every rank's `ipc_handle()` in rank order (one process per GPU)."""
        blob = b"".join(handles)
        if len(blob) != self.world * int(_lib.ckks_lshard_ipc_size()):
            raise RnsNttError(31, "connect: need one handle per rank")
        _check(_lib.ckks_lshard_ipc_import(self._h, C.create_string_buffer(blob, len(blob))))

    def connect_process_group(self, group=None):
        """Exchange the handles over torch.distributed (plumbing only) and connect."""
        import torch.distributed as dist

        handles = [None] * self.world
        dist.all_gather_object(handles, self.ipc_handle(), group=group)
        self.connect(handles)

    @staticmethod
    def connect_local(shards):
        """All ranks live in this process (several GPUs, or several ranks on one GPU)."""
        arr = (C.c_void_p * len(shards))(*[s._h for s in shards])
        _check(_lib.ckks_lshard_connect_local(arr, len(shards)))

    # data ---------------------------------------------------------------------------------------
    def scatter(self, channels, is_ntt_domain: bool = False) -> RnsPoly:
        """This GPU's limbs of host polynomials [batch, L, N] (reference layout) -> device."""
        ch = _u64(channels)
        if ch.ndim == 2:
            ch = ch[None]
        if ch.shape[1] != self.channel_count():
            raise RnsNttError(5)
        return RnsPoly.from_channels(np.ascontiguousarray(ch[:, self.rank :: self.world]), self._basis, is_ntt_domain)

    def upload_key(self, a, b, rotation: int = 0) -> GadgetKey:
        """a, b: the whole gadget key [L, L, N] (digit, limb, N) or already this GPU's slice [L, L_own, N]."""
        a, b = _u64(a), _u64(b)
        l, own, n = self.channel_count(), len(self.owned()), self.degree
        if a.shape == (l, l, n):
            a, b = a[:, self.rank :: self.world], b[:, self.rank :: self.world]
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        if a.shape != (l, own, n) or b.shape != (l, own, n):
            raise RnsNttError(5, "gadget key slice must be [L, L_own, N]")
        h = _vp()
        _check(_lib.ckks_lshard_ksk_upload(self._h, _ptr(a), _ptr(b), C.byref(h)))
        return GadgetKey(h, self._basis, rotation)

    # operations ---------------------------------------------------------------------------------
    def mul_relin_rescale(self, a: Ciphertext, b: Ciphertext, rlk: GadgetKey, child: "LimbShard | None" = None) -> Ciphertext:
        """mul_ciphertexts_gadget, then rescale_ciphertext into `child`'s level when given."""
        if a.logq != b.logq:  # engine.rs:478
            raise RnsNttError(23)
        h0, h1 = _vp(), _vp()
        _check(_lib.ckks_lshard_ct_mul_relin_rescale(self._h, a.c0._h, a.c1._h, b.c0._h, b.c1._h, rlk._h, child._h if child else None,
                                                     C.byref(h0), C.byref(h1)))
        bs = child._basis if child else self._basis
        bits = self._moduli[-1].bit_length() if child else 0
        return Ciphertext(RnsPoly(h0, bs), RnsPoly(h1, bs), a.logp + b.logp - bits, a.logq - bits)

    def rotate_ciphertext(self, ct: Ciphertext, rotk: GadgetKey) -> Ciphertext:
        """rotate_ciphertext (engine.rs:412-463) by the key's `rotation`."""
        h0, h1 = _vp(), _vp()
        _check(_lib.ckks_lshard_ct_rotate(self._h, ct.c0._h, ct.c1._h, rotk._h, rotk.rotation, C.byref(h0), C.byref(h1)))
        return Ciphertext(RnsPoly(h0, self._basis), RnsPoly(h1, self._basis), ct.logp, ct.logq)

    @staticmethod
    def group_rotate(shards, cts, keys):
        """rotate_ciphertext for a group living in this process, driven in lockstep (see group_mul_relin_rescale)."""
        world = len(shards)
        arr = (C.c_void_p * world)(*[s._h for s in shards])
        batch = cts[0].c0.batch()
        r0 = [ct.c0.rotate_slots(keys[r].rotation) for r, ct in enumerate(cts)]
        r1 = [ct.c1.rotate_slots(keys[r].rotation) for r, ct in enumerate(cts)]
        k0 = [RnsPoly.zero(s._basis, batch) for s in shards]
        k1 = [RnsPoly.zero(s._basis, batch) for s in shards]
        step = shards[0].chunk()
        for s0 in range(0, batch, step):
            cs = min(step, batch - s0)
            for phase in range(2):
                for r, s in enumerate(shards):
                    _check(_lib.ckks_lshard_ks_phase(s._h, phase, s0, cs, r1[r]._h, keys[r]._h, k0[r]._h, k1[r]._h, 1))
                _check(_lib.ckks_lshard_barrier_local(arr, world))
        outs = []
        for r, ct in enumerate(cts):
            r0[r] += k0[r]
            outs.append(Ciphertext(r0[r], k1[r], ct.logp, ct.logq))
        return outs

    def mul_phase(self, phase: int, s0: int, cs: int, a: Ciphertext, b: Ciphertext, rlk: GadgetKey, child, out: Ciphertext,
                  peer_stores: bool):
        _check(_lib.ckks_lshard_mul_phase(self._h, phase, s0, cs, a.c0._h, a.c1._h, b.c0._h, b.c1._h, rlk._h,
                                          child._h if child else None, out.c0._h, out.c1._h, int(peer_stores)))

    def barrier(self):
        _check(_lib.ckks_lshard_barrier(self._h))

    @staticmethod
    def group_mul_relin_rescale(shards, a, b, keys, kids=None):
        """A whole group living in this process, driven in lockstep: per chunk, phase A on every rank, barrier,
        phase B, barrier, phase C.  a, b, keys (and kids) are per-rank lists; returns the per-rank results."""
        world = len(shards)
        arr = (C.c_void_p * world)(*[s._h for s in shards])
        batch = a[0].c0.batch()
        outs = []
        for r, s in enumerate(shards):
            if a[r].logq != b[r].logq:  # engine.rs:478
                raise RnsNttError(23)
            bs = kids[r]._basis if kids else s._basis
            bits = s._moduli[-1].bit_length() if kids else 0
            outs.append(Ciphertext(RnsPoly.zero(bs, batch), RnsPoly.zero(bs, batch), a[r].logp + b[r].logp - bits, a[r].logq - bits))
        step = shards[0].chunk()
        for s0 in range(0, batch, step):
            cs = min(step, batch - s0)
            for phase in range(3):
                for r, s in enumerate(shards):
                    s.mul_phase(phase, s0, cs, a[r], b[r], keys[r], kids[r] if kids else None, outs[r], True)
                if phase < 2:
                    _check(_lib.ckks_lshard_barrier_local(arr, world))
        return outs

    def buffers(self):
        """(gather_ptr, gather_words, last_ptr, last_words): device pointers of the exchange buffers."""
        g, l = _u64p(), _u64p()
        gw, lw = C.c_size_t(0), C.c_size_t(0)
        _check(_lib.ckks_lshard_buffers(self._h, C.byref(g), C.byref(gw), C.byref(l), C.byref(lw)))
        return C.cast(g, C.c_void_p).value, int(gw.value), C.cast(l, C.c_void_p).value, int(lw.value)

    def check(self):
        """Synchronise; raises if a barrier gave up waiting for a peer."""
        _check(_lib.ckks_lshard_check(self._h))

    def set_exchange(self, mode: int):
        """0: digits stored into peer HBM by the producing kernel; 1: pushed by the copy engines."""
        _check(_lib.ckks_lshard_set_exchange(self._h, mode))

    def set_timeout_ms(self, ms: int):
        _check(_lib.ckks_lshard_set_timeout_ms(self._h, ms))
