import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kzg_setup_powersoftau_b200 as kz
ctx = kz.Context(1)
print(os.environ.get("PTAU_LIB", "default"))
for kind, name in ((0, "imad32"), (1, "imad.wide.x chains"), (5, "imad.wide plain"), (2, "fq_mul chain"), (3, "g1 dbl loop, calls"), (4, "g1 dbl loop, inlined")):
    best = 0
    for _ in range(3):
        ms, ops = ctx.microbench(kind, 500); best = max(best, ops / ms / 1e6)
    print("  %-24s %.3f G/s" % (name, best))
