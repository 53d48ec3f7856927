"""Limb dump format: the reference has no wire format, so polynomials cross machines as a
raw little-endian u64 file `[batch][limb][N]` next to a JSON header.  A Rust-equipped machine can replay
the same inputs through the real crate (`RnsPoly::from_channels`) and diff against these outputs.

    <name>.u64   little-endian words, canonical representatives, reference layout (limb-major);
                 NTT-domain data in the reference's natural order
    <name>.json  {"format": "ckks-b200-limbs-v1", "degree": N, "moduli": [...], "batch": B, "limbs": L,
                  "ntt_domain": bool, "sha256": "...", "note": "..."}
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

FORMAT = "ckks-b200-limbs-v1"


def write(path_base: str, channels, moduli, ntt_domain: bool = False, note: str = "") -> dict:
    a = np.ascontiguousarray(channels, dtype="<u8")
    if a.ndim == 2:
        a = a[None, :, :]
    if a.ndim != 3 or a.shape[1] != len(moduli):
        raise ValueError("channels must be [batch, L, N] with L == len(moduli)")
    q = np.array(moduli, dtype=np.uint64)[None, :, None]
    if (a >= q).any():
        raise ValueError("non-reduced coefficient (poly.rs:83-93)")
    raw = a.tobytes()
    hdr = {
        "format": FORMAT,
        "degree": int(a.shape[2]),
        "moduli": [int(m) for m in moduli],
        "batch": int(a.shape[0]),
        "limbs": int(a.shape[1]),
        "ntt_domain": bool(ntt_domain),
        "sha256": hashlib.sha256(raw).hexdigest(),
        "note": note,
    }
    with open(path_base + ".u64", "wb") as f:
        f.write(raw)
    with open(path_base + ".json", "w") as f:
        json.dump(hdr, f, indent=1)
    return hdr


def read(path_base: str):
    with open(path_base + ".json") as f:
        hdr = json.load(f)
    if hdr.get("format") != FORMAT:
        raise ValueError("not a ckks-b200 limb dump")
    raw = open(path_base + ".u64", "rb").read()
    if hashlib.sha256(raw).hexdigest() != hdr["sha256"]:
        raise ValueError("limb dump checksum mismatch")
    a = np.frombuffer(raw, dtype="<u8").reshape(hdr["batch"], hdr["limbs"], hdr["degree"]).astype(np.uint64)
    return a, hdr
