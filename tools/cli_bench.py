"""Wall clock of the drop-in binaries and of load_kzg_setup on a synthetic 2^K-power ceremony file:
   python tools/cli_bench.py [K=21] [runs=5] [gpus=1]     (PTAU_TRACE=1 adds the stage marks of csrc/files.cu)"""
import hashlib, os, shutil, statistics, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kzg_setup_powersoftau_b200 as kz

k = int(sys.argv[1]) if len(sys.argv) > 1 else 21
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 1 << k
tau = 0x1234567890ABCDEF1234567890ABCDEF
L = kz._ffi.lib()
ctx = kz.Context(n_gpus=gpus)
ZC = kz.FMT_ZCASH_COMPRESSED
resp = kz.PinnedBuffer(L.ptau_response_size(n))
off = 64
for g, s0, cnt in ((kz.G1, 1, 2 * n - 1), (kz.G2, 1, n), (kz.G1, 7, n), (kz.G1, 11, n), (kz.G2, 11, 1)):
    ln = cnt * L.ptau_record_size(g, ZC)
    ctx.generate(g, ZC, s0, tau, 0, cnt, out=resp.array[off:off + ln])
    off += ln
tmpd = tempfile.mkdtemp(prefix="ptau_cli_", dir=os.environ.get("PTAU_BENCH_DIR"))
resp.array.tofile(os.path.join(tmpd, "powersoftau"))
hexd = hashlib.blake2b(memoryview(resp.array), digest_size=64).hexdigest()
setup = kz.PinnedBuffer(L.ptau_setup_size(kz.VARIANT_KGZ, n))  # pinned output: a pageable one halves the copy rate
for _ in range(2):
    ctx.preprocess(kz.VARIANT_KGZ, resp, n, out=setup)
t0 = time.perf_counter(); ctx.preprocess(kz.VARIANT_KGZ, resp, n, out=setup); t_lib = time.perf_counter() - t0
want = setup.array
print("library host->host (pinned in, pinned out): %.3f s" % t_lib, flush=True)
exe = os.path.join(ROOT, "kzg_setup_powersoftau_b200", "bin", "preprocess-kgz")
for tag, flags in (("like_reference", ["--expect-digest", hexd]), ("no_digest_no_intermediate", ["--skip-digest", "--no-uncompressed"])):
    walls = []
    for i in range(runs):
        for f in ("powersoftau_uncompressed", "kzg_setup"):
            if os.path.exists(os.path.join(tmpd, f)):
                os.remove(os.path.join(tmpd, f))
        t0 = time.perf_counter()
        r = subprocess.run([exe, "--dir", tmpd, "--log2-powers", str(k), "--gpus", str(gpus)] + flags, capture_output=True, text=True)
        walls.append(time.perf_counter() - t0)
        assert r.returncode == 0, r.stderr
        if i == runs - 1 and r.stderr:
            print(r.stderr, end="")
    same = open(os.path.join(tmpd, "kzg_setup"), "rb").read() == want.tobytes()
    print("%-28s min %.3f median %.3f all %s  output==library:%s" % (tag, min(walls), statistics.median(walls), ["%.2f" % w for w in walls], same), flush=True)
for tag, checks in (("load_kzg_setup", kz.CHECKS_LOAD), ("load_kzg_setup validated", kz.CHECKS_STRICT)):
    walls = []
    for _ in range(runs):
        t0 = time.perf_counter()
        kz.load_kzg_setup(os.path.join(tmpd, "kzg_setup"), ctx=ctx, checks=checks)
        walls.append(time.perf_counter() - t0)
    print("%-28s min %.3f median %.3f" % (tag, min(walls), statistics.median(walls)), flush=True)
shutil.rmtree(tmpd, ignore_errors=True)
