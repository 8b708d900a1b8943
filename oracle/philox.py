"""CPU restatement (numpy) of the counter-based dropout masks of the training kernels (synchformer_b200/csrc/philox.cuh).

TEST INFRASTRUCTURE, NOT PRODUCT (same rules as synchformer_oracle.py).

Philox4x32-10 is the published generator of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
also shipped as Random123 `philox4x32` and cuRAND's `curandStatePhilox4_32_10_t`; it is not part of the reference repository (the
reference calls `nn.Dropout`, whose CUDA stream is a torch implementation detail).  Pinned by the Random123 known-answer vectors in
`tests/test_train_oracle_cpu.py`.

    keep(seed, site, e) = philox4x32_10(counter = (lo32(e >> 2), hi32(e >> 2), site, 0), key = (lo32(seed), hi32(seed)))[e & 3] >= thr
    thr = min(floor(p * 2^32), 2^32 - 1);  multiplier = keep / (1 - p)                                  (nn.Dropout semantics)
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised over uint32 counter arrays c0..c3 (broadcastable); key words are Python ints.  Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3))
    for r in range(10):
        kk0 = np.uint64((k0 + r * W0) & 0xFFFFFFFF)
        kk1 = np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        p0 = M0 * c0                      # 32 x 32 -> 64 bit products (no overflow in uint64)
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ kk0, lo1, hi0 ^ c3 ^ kk1, lo0
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def threshold(p: float) -> int:
    if p <= 0.0:
        return 0
    return int(min(np.floor(np.float64(np.float32(p)) * 4294967296.0), 4294967295.0))


def dropout_multiplier(shape, p: float, seed: int, site: int) -> np.ndarray:
    """float32 array of `shape`: 0 where element e (C-order linear index) is dropped, 1 / (1 - p) where it is kept."""
    n = int(np.prod(shape))
    thr = threshold(p)
    if thr == 0:
        return np.ones(shape, dtype=np.float32)
    ctr = np.arange((n + 3) // 4, dtype=np.uint64)
    r = philox4x32_10(ctr & MASK32, ctr >> np.uint64(32), np.uint64(site), np.uint64(0), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.stack(r, axis=1).reshape(-1)[:n]
    inv_keep = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(u >= np.uint32(thr), inv_keep, np.float32(0.0)).astype(np.float32).reshape(shape)
