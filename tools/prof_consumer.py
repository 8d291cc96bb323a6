"""One bucket-MSM at 2^20 and one KZG10 check batch of one full wave, 37,888 openings (for ncu): python tools/prof_consumer.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kzg_setup_powersoftau_b200 as kz

ctx = kz.Context(1)
tau = 0x1234567890ABCDEF1234567890ABCDEF
ZU, ML = kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ARK_MONT_LIMBS
L = kz._ffi.lib()
n = 1 << 20
pw = ctx.convert(kz.G1, ZU, ctx.generate(kz.G1, ZU, 1, tau, 0, n), ML, 0)
sc = np.random.default_rng(1).integers(0, 256, size=n * 32, dtype=np.uint8)
sc.reshape(n, 32)[:, 31] &= 0x3F
out = np.zeros(104, dtype=np.uint8)
assert L.ptau_kzg_commit(ctx._h, pw.ctypes.data, sc.ctypes.data, n, out.ctypes.data) == 0
print("msm ms", ctx.timing()["kernel_ms"][0])
nk = int(os.environ.get("NK", 148 * 256))  # 296 blocks of 128 = two per SM
pwk = pw.reshape(-1, 104)[:32]
g2p = ctx.convert(kz.G2, ZU, ctx.generate(kz.G2, ZU, 1, tau, 0, 2), ML, 0).reshape(2, 200)
vk = kz.VerifierKey(g=pwk[0], gamma_g=pwk[5], h=g2p[0], beta_h=g2p[1])
pws = kz.Powers(powers_of_g=pwk, powers_of_gamma_g=pwk[:1])
poly = [int(x) for x in np.random.default_rng(2).integers(1, 1 << 62, size=16)]
comm = kz.KZG10.commit(pws, poly, ctx=ctx)
val, prf, _ = kz.KZG10.open(pws, poly, 12345, ctx=ctx)
ok = kz.KZG10.check_many(vk, np.tile(comm, (nk, 1)), [12345] * nk, [val] * nk, np.tile(prf, (nk, 1)), ctx=ctx)
assert ok.all()
print("check ms", ctx.timing()["kernel_ms"][0])
