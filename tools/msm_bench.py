"""KZG10 commit (bucket MSM) timing at several sizes; run under ncu for the per-kernel launch list."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kzg_setup_powersoftau_b200 as kz

ctx = kz.Context()
tau = 0x1234567890ABCDEF1234567890ABCDEF
L = kz._ffi.lib()
verify = "--verify" in sys.argv  # compare with [sum c_i tau^i]G computed by Horner on the host (slow for 2^24)
sizes = [int(a) for a in sys.argv[1:] if a != "--verify"] or [16, 18, 20, 22]
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
    if verify:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        import ptau_oracle as o
        R = o.R_ORDER
        acc = 0
        for i in range(n - 1, -1, -1):
            acc = (acc * tau + int.from_bytes(sc[32 * i:32 * i + 32].tobytes(), "little")) % R
        q = o.g1_mul(o.G1_GEN, acc)
        assert out.tobytes() == o.g1_mont_record(q[0], q[1], False), "MSM result differs from [p(tau)]G"
        print("   verified against [p(tau)]G", flush=True)
