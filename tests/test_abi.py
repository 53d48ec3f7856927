"""The C-ABI library loads on a CPU-only box, exports every symbol include/ckks_b200.h declares,
validates arguments like the reference, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ckks_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ckks_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(ck):
    lib = C.CDLL(ck.LIB_PATH)
    names = header_functions()
    assert len(names) >= 55
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/ckks_b200.h but not exported: {missing}"


def test_host_number_theory_matches_oracle(ck, orc):
    for args in ((31, 4, 16), (30, 3, 32), (62, 2, 1024), (40, 3, 4096), (61, 24, 65536), (30, 8, 16384), (63, 2, 2048)):
        assert ck.generate_primes(*args) == orc.generate_primes(*args)
    with pytest.raises(ck.RnsNttError):
        ck.generate_primes(5, 50, 8)
    for v in (0, 1, 2, 3, 4, 17, 19, 561, 1105, 7681, 1073750017, 2305843009211596801, 2305843009211596803):
        assert ck.is_prime(v) == orc.is_prime(v)
    assert not ck.is_ntt_friendly_prime(19, 8) and ck.is_ntt_friendly_prime(17, 8)


def test_validation_precedes_device_and_there_is_no_cpu_fallback(ck):
    for n, moduli, kind in ((8, [], "EmptyBasis"), (8, [19], "NonNttFriendlyModulus"), (12, [17], "InvalidDegree"), (0, [17], "InvalidDegree")):
        with pytest.raises(ck.RnsNttError) as e:
            ck.RnsBasis(n, moduli)
        assert e.value.kind == kind
    if ck.device_count() == 0:
        with pytest.raises(ck.RnsNttError) as e:
            ck.RnsBasis(8, [17, 97, 113])
        assert e.value.kind == "CudaError" and "no CPU fallback" in str(e.value)
        assert ck.modmul_peak() == 0.0


def test_status_strings(ck):
    lib = C.CDLL(ck.LIB_PATH)
    lib.ckks_status_str.restype = C.c_char_p
    names = {1: b"InvalidDegree", 2: b"EmptyBasis", 3: b"NonNttFriendlyModulus", 4: b"InvalidModDrop", 5: b"ChannelCountMismatch", 6: b"NonReducedCoefficient"}
    for code, name in names.items():
        assert lib.ckks_status_str(code) == name  # 1:1 with RnsNttError (errors.rs:3-22)


def test_product_never_touches_the_oracle():
    """The package and the CUDA sources must not import, link or read anything under oracle/."""
    pkg = os.path.join(ROOT, "toy-heaan-ckks_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "libckks_oracle" not in text and "ckks_oracle.h" not in text, f


def test_cpp_host_mirror_links_and_maps_errors(built):
    """toy-heaan-ckks_b200/host/rns_poly.hpp (the compiled-language mirror of RnsBasis / RnsPoly /
    CkksEngine) builds against the C ABI; validation errors map onto RnsNttError; on a GPU box it also
    runs the (1+x)^2 KAT (poly.rs:789-802)."""
    import subprocess

    exe = os.path.join(ROOT, "tests", "emul", "host_mirror_check")
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
