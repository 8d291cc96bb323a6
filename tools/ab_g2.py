"""A/B of the G2 point kernels (and the G1 compressed one as the noise reference) at 2^20 points:
PTAU_LIB=.../libptau_b200_NAME.so python tools/ab_g2.py   ->  one line per kernel, best of 5 launches"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kzg_setup_powersoftau_b200 as kz
ctx = kz.Context(1)
N = 1 << 20
tau = 0x1234567890ABCDEF1234567890ABCDEF
status = torch.full((1,), -1, dtype=torch.int64, device="cuda")
L = kz._ffi.lib()
print(os.environ.get("PTAU_LIB", "default"))
for group, in_fmt, name in ((kz.G2, 2, "g2_comp_strict"), (kz.G2, 1, "g2_unc_strict"), (kz.G1, 2, "g1_comp_strict")):
    ri, ro = L.ptau_record_size(group, in_fmt), L.ptau_record_size(group, 3)
    d_in = torch.empty(N * ri, dtype=torch.uint8, device="cuda"); d_out = torch.empty(N * ro, dtype=torch.uint8, device="cuda")
    ctx.generate_device(group, in_fmt, 1, tau, 0, N, d_in.data_ptr()); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.convert_device(group, in_fmt, d_in.data_ptr(), 3, d_out.data_ptr(), N, kz.CHECKS_STRICT, status.data_ptr()); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    assert int(status.item()) == -1, "a valid point was rejected"
    print("%-16s %8.3f ms  %7.2f Mpts/s" % (name, best, N / best / 1e3), flush=True)
