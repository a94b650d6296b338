"""Golden digit slices of the radix-256 slicing (csrc/ozaki_slice_kernels.cuh slice256_kernel), generated with the
top-down `rint` formulation of round 2 (commit 193b87a) compiled for the host by tests/emu/.  The byte-parallel
formulation that replaced it must reproduce these digits bit for bit (tests/test_emu_ozaki.py).
Inputs are seeded and include the rounding ties, carry chains and range ends.  Run from the repo root:
    python tests/golden/make_slice256_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def inputs():
    rng = np.random.default_rng(20261017)
    K = 96
    cols = []
    cols.append(rng.standard_normal(K))
    cols.append(rng.standard_normal(K) * np.exp(3.0 * rng.standard_normal(K)))
    c = rng.standard_normal(K)
    c[0] = 127.0                                   # column maximum exactly at the range end 127/128 * 2^e
    c[1:] *= 0.5
    cols.append(c)
    # exact ties at every digit level: (integer + 1/2) * 256^-s, both signs, odd and even integer parts
    ties = []
    for s in range(8):
        for n in (0, 1, 2, 3, 126, 127, -1, -2, -127, -128, 255, 256):
            ties.append((n + 0.5) * 256.0 ** -s)
            ties.append(-(n + 0.5) * 256.0 ** -s)
    ties = np.array(ties[:K - 1] + [100.0])
    cols.append(ties)
    # carry chains: digits that round up to +128 repeatedly (0x7f.80 80 80 ... patterns) and all-0xff mantissas
    chain = [sum(128.0 * 256.0 ** -t for t in range(1, n + 1)) for n in range(1, 8)]
    chain += [-x for x in chain] + [np.nextafter(1.0, 0.0) * 2.0 ** -j for j in range(20)] + \
             [np.nextafter(1.0, 2.0) * 2.0 ** -j for j in range(20)]
    chain = np.array((chain + [0.0] * K)[:K - 1] + [64.0])
    cols.append(chain)
    cols.append(np.zeros(K))                       # all-zero column
    cols.append(2.0 ** -rng.integers(0, 60, K) * rng.choice([-1.0, 1.0], K))
    cols.append(rng.standard_normal(K) * 1e-300)   # tiny magnitudes (scale 2^(7 - e) large)
    cols.append(rng.standard_normal(K) * 1e300)
    return np.asfortranarray(np.stack(cols, axis=1))


if __name__ == "__main__":
    from test_emu_ozaki import slice_emu
    A = inputs()
    out = {"A": A}
    for ns in (1, 3, 5, 7, 8):
        D, expo, dscale, ldd = slice_emu(A, 256, ns)
        out[f"D{ns}"] = D
        out[f"expo{ns}"] = expo
    np.savez_compressed(os.path.join(HERE, "slice256_digits.npz"), **out)
    print({k: v.shape for k, v in out.items()})
