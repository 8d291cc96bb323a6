"""KZG10 commit (MSM) and check sharded over 1..N GPUs of one context: python tools/consumer_scale.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kzg_setup_powersoftau_b200 as kz
from kzg_setup_powersoftau_b200 import _ffi

ng = _ffi.lib().ptau_device_count()
tau = 0x1234567890ABCDEF1234567890ABCDEF
ZU, ML = kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ARK_MONT_LIMBS
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << lg
with kz.Context(1) as c1:
    pw = c1.convert(kz.G1, ZU, c1.generate(kz.G1, ZU, 1, tau, 0, n), ML, 0)
    g2p = c1.convert(kz.G2, ZU, c1.generate(kz.G2, ZU, 1, tau, 0, 2), ML, 0).reshape(2, 200)
sc = np.random.default_rng(1).integers(0, 256, size=n * 32, dtype=np.uint8)
sc.reshape(n, 32)[:, 31] &= 0x3F
pwk = pw.reshape(-1, 104)
vk = kz.VerifierKey(g=pwk[0], gamma_g=pwk[5], h=g2p[0], beta_h=g2p[1])
pws = kz.Powers(powers_of_g=pwk[:32], powers_of_gamma_g=pwk[:1])
poly = [int(x) for x in np.random.default_rng(2).integers(1, 1 << 62, size=16)]
ref = None
G = 1
while G <= ng:
    with kz.Context(G) as ctx:
        out = np.zeros(104, dtype=np.uint8)
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            assert _ffi.lib().ptau_kzg_commit(ctx._h, pw.ctypes.data, sc.ctypes.data, n, out.ctypes.data) == 0
            wall = time.perf_counter() - t0
            best = min(best, max(ctx.timing()["kernel_ms"][:G]))
        ref = ref or out.tobytes()
        assert out.tobytes() == ref
        comm = kz.KZG10.commit(pws, poly, ctx=ctx)
        val, prf, _ = kz.KZG10.open(pws, poly, 12345, ctx=ctx)
        nk = 148 * 256 * G
        for _ in range(2):
            ok = kz.KZG10.check_many(vk, np.tile(comm, (nk, 1)), [12345] * nk, [val] * nk, np.tile(prf, (nk, 1)), ctx=ctx)
        assert ok.all()
        kms = max(ctx.timing()["kernel_ms"][:G])
        print("GPUs=%d  MSM 2^%d: %.2f ms kernels (max over GPUs, %.1f M points/s; wall incl. H2D from pageable memory %.0f ms)"
              "   check %d openings: %.2f ms (%.0f k openings/s)" % (G, lg, best, n / best / 1e3, wall * 1e3, nk, kms, nk / kms), flush=True)
    G *= 2
