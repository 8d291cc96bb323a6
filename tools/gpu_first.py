"""First GPU bring-up: microbenchmarks, small parity vs the oracle, raw kernel throughput."""
import json, os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import kzg_setup_powersoftau_b200 as kz
import ptau_oracle as o

res = {}
ctx = kz.Context(1)
prop = torch.cuda.get_device_properties(0)
res["gpu"] = prop.name; res["sms"] = prop.multi_processor_count
for kind, name in ((0, "imad32"), (1, "imad_wide_chain"), (2, "fq_mul")):
    ms, ops = ctx.microbench(kind, 2000)
    res["mb_" + name] = {"ms": ms, "ops": ops, "Gops_s": ops / ms / 1e6}
    print(name, ms, "ms", ops / ms / 1e6, "Gops/s", flush=True)

rnd = random.Random(11)
g1 = [o.g1_mul(o.G1_GEN, rnd.randrange(1, o.R_ORDER)) for _ in range(300)]
g2 = [o.g2_mul(o.G2_GEN, rnd.randrange(1, o.R_ORDER)) for _ in range(150)]
def chk(name, got, want):
    ok = bytes(got) == want
    print(name, "OK" if ok else "MISMATCH", flush=True)
    res["parity_" + name] = ok
zu1 = b"".join(o.zcash_g1_uncompressed_encode(q) for q in g1); zc1 = b"".join(o.zcash_g1_compressed_encode(q) for q in g1)
au1 = b"".join(o.ark_g1_serialize_uncompressed(q) for q in g1); ml1 = b"".join(o.g1_mont_record(q[0], q[1], False) for q in g1)
zu2 = b"".join(o.zcash_g2_uncompressed_encode(q) for q in g2); zc2 = b"".join(o.zcash_g2_compressed_encode(q) for q in g2)
au2 = b"".join(o.ark_g2_serialize_uncompressed(q) for q in g2); ml2 = b"".join(o.g2_mont_record(q[0], q[1], False) for q in g2)
S = kz.CHECKS_STRICT
chk("g1_zu_au", ctx.convert(kz.G1, kz.FMT_ZCASH_UNCOMPRESSED, zu1, kz.FMT_ARK_UNCOMPRESSED, S), au1)
chk("g1_zc_au", ctx.convert(kz.G1, kz.FMT_ZCASH_COMPRESSED, zc1, kz.FMT_ARK_UNCOMPRESSED, S), au1)
chk("g1_zc_zu", ctx.convert(kz.G1, kz.FMT_ZCASH_COMPRESSED, zc1, kz.FMT_ZCASH_UNCOMPRESSED, 0), zu1)
chk("g1_au_ml", ctx.convert(kz.G1, kz.FMT_ARK_UNCOMPRESSED, au1, kz.FMT_ARK_MONT_LIMBS, 0), ml1)
chk("g1_zu_ml", ctx.convert(kz.G1, kz.FMT_ZCASH_UNCOMPRESSED, zu1, kz.FMT_ARK_MONT_LIMBS, S), ml1)
chk("g2_zu_au", ctx.convert(kz.G2, kz.FMT_ZCASH_UNCOMPRESSED, zu2, kz.FMT_ARK_UNCOMPRESSED, S), au2)
chk("g2_zc_au", ctx.convert(kz.G2, kz.FMT_ZCASH_COMPRESSED, zc2, kz.FMT_ARK_UNCOMPRESSED, S), au2)
chk("g2_zc_zu", ctx.convert(kz.G2, kz.FMT_ZCASH_COMPRESSED, zc2, kz.FMT_ZCASH_UNCOMPRESSED, 0), zu2)
chk("g2_au_ml", ctx.convert(kz.G2, kz.FMT_ARK_UNCOMPRESSED, au2, kz.FMT_ARK_MONT_LIMBS, 0), ml2)
# bad points
bad = []
while len(bad) < 5:
    x = rnd.randrange(o.P); y = o.fq_sqrt((x ** 3 + 4) % o.P)
    if y is not None: bad.append((x, y))
data = bytearray(zu1); data[96 * 77:96 * 78] = o.zcash_g1_uncompressed_encode(bad[0]); data[96 * 200:96 * 201] = o.zcash_g1_uncompressed_encode(bad[1])
try:
    ctx.convert(kz.G1, kz.FMT_ZCASH_UNCOMPRESSED, bytes(data), kz.FMT_ARK_UNCOMPRESSED, S); res["parity_bad_g1"] = False
except kz.PtauError as e:
    res["parity_bad_g1"] = (e.index == 77 and e.code == kz.BAD_NOT_IN_SUBGROUP); print("bad g1:", e, flush=True)
# generator vs oracle
tau = rnd.randrange(1, o.R_ORDER); alpha = rnd.randrange(1, o.R_ORDER)
gen = ctx.generate(kz.G1, kz.FMT_ZCASH_COMPRESSED, alpha, tau, 3, 40)
want = b"".join(o.zcash_g1_compressed_encode(o.g1_mul(o.G1_GEN, alpha * pow(tau, 3 + i, o.R_ORDER) % o.R_ORDER)) for i in range(40))
chk("gen_g1", gen, want)
gen = ctx.generate(kz.G2, kz.FMT_ZCASH_COMPRESSED, 1, tau, 0, 20)
want = b"".join(o.zcash_g2_compressed_encode(o.g2_mul(o.G2_GEN, pow(tau, i, o.R_ORDER))) for i in range(20))
chk("gen_g2", gen, want)
gen = ctx.generate(kz.G2, kz.FMT_ZCASH_UNCOMPRESSED, 1, tau, 5, 20)
want = b"".join(o.zcash_g2_uncompressed_encode(o.g2_mul(o.G2_GEN, pow(tau, 5 + i, o.R_ORDER))) for i in range(20))
chk("gen_g2_unc", gen, want)

# raw kernel throughput, device-resident
N = 1 << int(os.environ.get("LOGN", "18"))
status = torch.full((1,), -1, dtype=torch.int64, device="cuda")
def bench(group, in_fmt, out_fmt, checks, name, reps=3):
    ri = {1: {1: 96, 2: 48, 3: 96}, 2: {1: 192, 2: 96, 3: 192}}[group][in_fmt]
    ro = {1: {1: 96, 3: 96, 4: 104}, 2: {1: 192, 3: 192, 4: 200}}[group][out_fmt]
    gen_fmt = in_fmt if in_fmt != kz.FMT_ARK_UNCOMPRESSED else kz.FMT_ZCASH_UNCOMPRESSED
    d_in = torch.empty(N * ri, dtype=torch.uint8, device="cuda"); d_out = torch.empty(N * ro, dtype=torch.uint8, device="cuda")
    t0 = time.time()
    ctx.generate_device(group, gen_fmt, 1, tau, 0, N, d_in.data_ptr()); torch.cuda.synchronize()
    tgen = time.time() - t0
    if in_fmt == kz.FMT_ARK_UNCOMPRESSED:
        tmp = torch.empty_like(d_in)
        ctx.convert_device(group, gen_fmt, d_in.data_ptr(), kz.FMT_ARK_UNCOMPRESSED, tmp.data_ptr(), N, 0, status.data_ptr()); torch.cuda.synchronize(); d_in = tmp
    status.fill_(-1)
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.convert_device(group, in_fmt, d_in.data_ptr(), out_fmt, d_out.data_ptr(), N, checks, status.data_ptr(), stream=0)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = int(status.item())
    r = {"ms": best, "Mpts_s": N / best / 1e3, "status": st, "gen_s": tgen}
    print(name, r, flush=True)
    res["tp_" + name] = r
bench(kz.G1, kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ARK_UNCOMPRESSED, S, "g1_unc_strict")
bench(kz.G1, kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ARK_UNCOMPRESSED, 0, "g1_unc_nocheck")
bench(kz.G1, kz.FMT_ZCASH_COMPRESSED, kz.FMT_ARK_UNCOMPRESSED, S, "g1_comp_strict")
bench(kz.G1, kz.FMT_ZCASH_COMPRESSED, kz.FMT_ZCASH_UNCOMPRESSED, 0, "g1_decompress")
bench(kz.G1, kz.FMT_ARK_UNCOMPRESSED, kz.FMT_ARK_MONT_LIMBS, 0, "g1_load")
bench(kz.G2, kz.FMT_ZCASH_UNCOMPRESSED, kz.FMT_ARK_UNCOMPRESSED, S, "g2_unc_strict")
bench(kz.G2, kz.FMT_ZCASH_COMPRESSED, kz.FMT_ARK_UNCOMPRESSED, S, "g2_comp_strict")
bench(kz.G2, kz.FMT_ZCASH_COMPRESSED, kz.FMT_ZCASH_UNCOMPRESSED, 0, "g2_decompress")
bench(kz.G2, kz.FMT_ARK_UNCOMPRESSED, kz.FMT_ARK_MONT_LIMBS, 0, "g2_load")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "first.json"), "w"), indent=1)
print("DONE", flush=True)
