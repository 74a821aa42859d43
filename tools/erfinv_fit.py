import mpmath as mp, numpy as np, sys
mp.mp.dps = 50

def target(w):
    # w = -log(1 - z^2)  ->  z = sqrt(1 - exp(-w)),  f = erfinv(z) / z
    w = mp.mpf(w)
    if w == 0:
        return mp.sqrt(mp.pi) / 2
    z = mp.sqrt(-mp.expm1(-w))
    return mp.erfinv(z) / z

def cheb_fit(fun, a, b, n):
    # Chebyshev interpolation of degree n on [a, b] -> monomial coefficients in t = x - c (c = midpoint)
    k = np.arange(n + 1)
    nodes = [mp.cos(mp.pi * (2 * i + 1) / (2 * (n + 1))) for i in range(n + 1)]
    half = (mp.mpf(b) - mp.mpf(a)) / 2
    c = (mp.mpf(b) + mp.mpf(a)) / 2
    fx = [fun(c + half * x) for x in nodes]
    # Chebyshev coefficients
    coef = []
    for j in range(n + 1):
        s = mp.mpf(0)
        for i in range(n + 1):
            s += fx[i] * mp.cos(mp.pi * j * (2 * i + 1) / (2 * (n + 1)))
        coef.append(2 * s / (n + 1))
    coef[0] /= 2
    # convert Chebyshev series in x to monomials in x, then x = t / half
    # T_0 = 1, T_1 = x, T_{k+1} = 2 x T_k - T_{k-1}
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for kk in range(2, n + 1):
        prev, prev2 = T[-1], T[-2]
        new = [mp.mpf(0)] + [2 * v for v in prev]
        for i, v in enumerate(prev2):
            new[i] -= v
        T.append(new)
    mono = [mp.mpf(0)] * (n + 1)
    for j in range(n + 1):
        for i, v in enumerate(T[j]):
            mono[i] += coef[j] * v
    mono_t = [mono[i] / half**i for i in range(n + 1)]
    return c, mono_t

def f_central(w):
    return target(w)

def f_tail(s_center):
    def f(s):
        w = (mp.mpf(s) + s_center) ** 2
        return target(w)
    return f

out = {}
# central: w in [0, 6.25], variable t = w - 3.125
c, m = cheb_fit(f_central, 0, 6.25, int(sys.argv[1]) if len(sys.argv) > 1 else 24)
out['central'] = (c, m)
# mid: w in [6.25, 16] -> s = sqrt(w) in [2.5, 4], variable t = s - 3.25
c2, m2 = cheb_fit(lambda s: target(mp.mpf(s) ** 2), 2.5, 4.0, int(sys.argv[2]) if len(sys.argv) > 2 else 20)
out['mid'] = (c2, m2)
# far: w in [16, 22] -> s in [4, 4.7], t = s - 4.35
c3, m3 = cheb_fit(lambda s: target(mp.mpf(s) ** 2), 4.0, 4.7, int(sys.argv[3]) if len(sys.argv) > 3 else 14)
out['far'] = (c3, m3)

def evalp(m, t):
    t = np.float64(t)
    acc = np.float64(float(m[-1]))
    for v in m[-2::-1]:
        acc = acc * t + np.float64(float(v))
    return acc

# validate in double arithmetic (Horner, as the device will do with FMA ~ slightly better)
rng = np.random.default_rng(1)
worst = 0
for region, (lo, hi) in (('central', (0, 6.25)), ('mid', (6.25, 16)), ('far', (16, 21.5))):
    c_, m_ = out[region]
    err = 0
    for w in np.concatenate([rng.uniform(lo, hi, 3000), [lo + 1e-9, hi - 1e-9]]):
        z = np.sqrt(-np.expm1(-w))
        z = np.float64(z)
        # device formula: w from z
        wd = -np.log((1.0 - z) * (1.0 + z))
        if region == 'central':
            t = wd - float(c_)
        else:
            t = np.sqrt(wd) - float(c_)
        val = evalp(m_, t) * z
        ref = mp.erfinv(mp.mpf(float(z)))
        e = abs((mp.mpf(float(val)) - ref) / ref)
        err = max(err, float(e))
    print(region, 'degree', len(m_) - 1, 'max rel err (double eval)', err)
    worst = max(worst, err)
import pickle
pickle.dump({k: (float(c), [mp.nstr(v, 25) for v in m]) for k, (c, m) in out.items()}, open('/tmp/erf/coef.pkl', 'wb'))
