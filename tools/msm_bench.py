"""KZG10 commit (bucket MSM) timing at several sizes; run under ncu for the per-kernel launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kzg_setup_powersoftau_b200 as kz

ctx = kz.Context()
tau = 0x1234567890ABCDEF1234567890ABCDEF
L = kz._ffi.lib()
sizes = [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22]
nmax = 1 << max(sizes)
pw = ctx.convert(kz.G1, kz.FMT_ZCASH_UNCOMPRESSED, ctx.generate(kz.G1, kz.FMT_ZCASH_UNCOMPRESSED, 1, tau, 0, nmax),
                 kz.FMT_ARK_MONT_LIMBS, 0)
sc = np.random.default_rng(1).integers(0, 256, size=nmax * 32, dtype=np.uint8)
sc.reshape(nmax, 32)[:, 31] &= 0x3F
out = np.zeros(104, dtype=np.uint8)
for lg in sizes:
    n = 1 << lg
    best = 1e9
    for _ in range(3):
        assert L.ptau_kzg_commit(ctx._h, pw.ctypes.data, sc.ctypes.data, n, out.ctypes.data) == 0
        best = min(best, ctx.timing()["kernel_ms"][0])
    print("n=2^%d  %.2f ms  %.1f M points/s" % (lg, best, n / best / 1e3), flush=True)
