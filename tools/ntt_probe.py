"""One forward + inverse limb-batched transform at a chosen size (ncu target)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ck = importlib.import_module("toy-heaan-ckks_b200")
bits, logn, l, batch = (int(x) for x in sys.argv[1:5])
n = 1 << logn
moduli = ck.generate_primes(bits, l, n)
b = ck.RnsBasis(n, moduli)
rng = np.random.default_rng(0)
q = np.array(moduli, dtype=np.uint64)
x = (rng.integers(0, 1 << 62, size=(batch, l, n), dtype=np.uint64) % q[:, None]).astype(np.uint64)
p = ck.RnsPoly.from_channels(x, b)
for _ in range(3):
    p.to_ntt_domain()
    p.to_coeff_domain()
b.sync()
assert np.array_equal(p.channels(), x)
print("ok", ck.launch_table())
