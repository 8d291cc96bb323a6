"""Quick A/B kernel timing: PTAU_LIB=... python tools/ab_bench.py [log2n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kzg_setup_powersoftau_b200 as kz
ctx = kz.Context(1)
N = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
tau = 0x1234567890ABCDEF1234567890ABCDEF
status = torch.full((1,), -1, dtype=torch.int64, device="cuda")
def bench(group, in_fmt, out_fmt, checks, name, n=N, reps=5):
    L = kz._ffi.lib()
    ri, ro = L.ptau_record_size(group, in_fmt), L.ptau_record_size(group, out_fmt)
    d_in = torch.empty(n * ri, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n * ro, dtype=torch.uint8, device="cuda")
    ctx.generate_device(group, in_fmt, 1, tau, 0, n, d_in.data_ptr()); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.convert_device(group, in_fmt, d_in.data_ptr(), out_fmt, d_out.data_ptr(), n, checks, status.data_ptr()); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    assert int(status.item()) == -1
    print("%-18s %8.3f ms  %7.2f Mpts/s" % (name, best, n / best / 1e3), flush=True)
print(os.environ.get("PTAU_LIB", "default"))
S = kz.CHECKS_STRICT
bench(kz.G1, 1, 3, S, "g1_unc_strict")
bench(kz.G1, 2, 3, S, "g1_comp_strict")
bench(kz.G2, 1, 3, S, "g2_unc_strict")
bench(kz.G2, 2, 3, S, "g2_comp_strict")
bench(kz.G1, 2, 1, 0, "g1_decompress")
bench(kz.G2, 2, 1, 0, "g2_decompress")
for kind, name in ((0, "imad"), (1, "imad_wide"), (2, "fq_mul")):
    ms, ops = ctx.microbench(kind, 2000); print(name, ops / ms / 1e6, "G/s")
for kind, name in ((3, "g1 dbl loop, calls"), (4, "g1 dbl loop, inlined")):
    ms, ops = ctx.microbench(kind, 500); print(name, "%.2f G dbl/s  (%.1f Mpts/s-equivalent at 126 dbl/pt)" % (ops / ms / 1e6, ops / ms / 1e3 / 126))
