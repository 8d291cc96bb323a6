# Exact integer model of the FP64 (DFMA) Montgomery multiplier: 8 x 48-bit limbs, R = 2^384.
# Every double is modelled as a Python int (all values are integers, some logically scaled);
# "exact" ops assert that the result fits 53 significant bits.
import random
p = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
W = 48
M48 = (1 << 48) - 1
R = 1 << 384
PL = [(p >> (48 * i)) & M48 for i in range(8)]
PINV = (-pow(p, -1, 1 << 48)) % (1 << 48)
C1 = 1 << 100

def sig_ok(x):
    if x == 0: return True
    x = abs(x)
    tz = (x & -x).bit_length() - 1
    return x.bit_length() - tz <= 53
def ex(x):
    assert sig_ok(x), hex(x)
    return x
def rz(x):
    if x == 0: return 0
    s = -1 if x < 0 else 1
    x = abs(x)
    bl = x.bit_length()
    if bl <= 53: return s * x
    sh = bl - 53
    return s * ((x >> sh) << sh)
def fma_rz(a, b, c): return rz(a * b + c)
def fma(a, b, c): return ex(a * b + c)
def add(a, b): return ex(a + b)
def add_rz(a, b): return rz(a + b)

stats = {"ops": 0}
def prod(a, b, H, L):
    hn = fma_rz(a, b, H)
    d = add(H, -hn)
    lo = fma(a, b, d)
    assert 0 <= lo < (1 << 48)
    L = add(L, lo)
    stats["ops"] += 4
    return hn, L

def split(V):
    hq = add_rz(V, C1)
    qv = add(hq, -C1)
    t = add(V, -qv)
    assert 0 <= t < (1 << 48) and qv % (1 << 48) == 0
    stats["ops"] += 3
    return qv, t

def scale_hp(H):  # (H - C1) / 2^48, exact: fma(H, 2^-48, -2^52)
    assert (H - C1) % (1 << 48) == 0
    stats["ops"] += 1
    return ex((H - C1) >> 48)

def fd_mul(a, b):
    H = [C1] * 16
    L = [0] * 16
    for i in range(8):
        for j in range(8):
            H[i + j], L[i + j] = prod(a[i], b[j], H[i + j], L[i + j])
    carry = 0  # scaled by 2^48 (multiple of 2^48)
    for i in range(8):
        V = L[i]
        if i > 0:
            V = add(V, scale_hp(H[i - 1])); stats["ops"] += 1
            assert carry % (1 << 48) == 0
            V = ex(V + (carry >> 48)); stats["ops"] += 1   # fma(carry, 2^-48, V)
        qv, t = split(V)
        hm = fma_rz(t, PINV, C1); dm = add(C1, -hm); m = fma(t, PINV, dm); stats["ops"] += 3
        assert m == (t * PINV) & M48
        for j in range(8):
            if j == 0:
                hn = fma_rz(m, PL[0], H[i]); d = add(H[i], -hn); lo = fma(m, PL[0], d); H[i] = hn
                V = ex(V + lo); stats["ops"] += 4
                assert V % (1 << 48) == 0
            else:
                H[i + j], L[i + j] = prod(m, PL[j], H[i + j], L[i + j])
        carry = V
    r = [0] * 8
    for k in range(8, 16):
        V = add(L[k], scale_hp(H[k - 1])); stats["ops"] += 1
        V = ex(V + (carry >> 48)); stats["ops"] += 1
        if k < 15:
            qv, t = split(V)
            r[k - 8] = t
            carry = qv
        else:
            r[7] = V
    return r

def to_limbs(x): return [(x >> (48 * i)) & M48 for i in range(8)]
def from_limbs(l): return sum(v << (48 * i) for i, v in enumerate(l))

if __name__ == "__main__":
    random.seed(1)
    for it in range(2000):
        if it == 0: x, y = 2 * p - 1, 2 * p - 1
        elif it == 1: x, y = 0, 5
        elif it == 2: x, y = (1 << 383) - 1, (1<<382)  # loose
        else: x, y = random.randrange(2 * p), random.randrange(2 * p)
        stats["ops"] = 0
        r = fd_mul(to_limbs(x), to_limbs(y))
        v = from_limbs(r)
        assert v % p == (x * y * pow(R, -1, p)) % p, it
        assert v < 2 * p, it
        assert all(0 <= t < (1 << 48) for t in r)
    print("ok; FP64 ops per mul:", stats["ops"])
