"""KZG10::check throughput at several batch sizes: python tools/kzg_check_bench.py [n ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kzg_setup_powersoftau_b200 as kz

ctx = kz.Context(1)
tau = 0x1234567890ABCDEF1234567890ABCDEF
ZU, ML = kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ARK_MONT_LIMBS
pwk = ctx.convert(kz.G1, ZU, ctx.generate(kz.G1, ZU, 1, tau, 0, 32), ML, 0).reshape(-1, 104)
g2p = ctx.convert(kz.G2, ZU, ctx.generate(kz.G2, ZU, 1, tau, 0, 2), ML, 0).reshape(2, 200)
vk = kz.VerifierKey(g=pwk[0], gamma_g=pwk[5], h=g2p[0], beta_h=g2p[1])
pws = kz.Powers(powers_of_g=pwk, powers_of_gamma_g=pwk[:1])
poly = [int(x) for x in np.random.default_rng(2).integers(1, 1 << 62, size=16)]
comm = kz.KZG10.commit(pws, poly, ctx=ctx)
val, prf, _ = kz.KZG10.open(pws, poly, 12345, ctx=ctx)
for nk in [int(a) for a in sys.argv[1:]] or [1, 1024, 148 * 256, 4 * 148 * 256]:
    best = 1e9
    for _ in range(2):
        ok = kz.KZG10.check_many(vk, np.tile(comm, (nk, 1)), [12345] * nk, [val] * nk, np.tile(prf, (nk, 1)), ctx=ctx)
        assert ok.all()
        best = min(best, ctx.timing()["kernel_ms"][0])
    print("n=%d  %.2f ms  %.0f openings/s" % (nk, best, nk / best * 1e3), flush=True)
