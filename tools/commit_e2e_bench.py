"""KZG10 commit end to end (wall clock of the Python call): host arrays vs device-resident powers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kzg_setup_powersoftau_b200 as kz

ctx = kz.Context(1)
tau = 0x1234567890ABCDEF1234567890ABCDEF
ZU, ML = kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ARK_MONT_LIMBS
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
pw = ctx.convert(kz.G1, ZU, ctx.generate(kz.G1, ZU, 1, tau, 0, n), ML, 0).reshape(n, 104)
sc = np.random.default_rng(1).integers(0, 256, size=n * 32, dtype=np.uint8)
sc.reshape(n, 32)[:, 31] &= 0x3F
L = kz._ffi.lib()
out = np.zeros(104, dtype=np.uint8)
res = kz.ResidentPoints(ctx, pw)
for name, call in (("host arrays", lambda: L.ptau_kzg_commit(ctx._h, pw.ctypes.data, sc.ctypes.data, n, out.ctypes.data)),
                   ("resident powers", lambda: L.ptau_kzg_commit_resident(ctx._h, res._h, sc.ctypes.data, n, out.ctypes.data))):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        assert call() == 0
        best = min(best, time.perf_counter() - t0)
    print("2^%d terms, %-16s: %.1f ms wall (kernels %.1f ms)  %s" % (lg, name, best * 1e3, ctx.timing()["kernel_ms"][0], out[:8].tobytes().hex()), flush=True)
