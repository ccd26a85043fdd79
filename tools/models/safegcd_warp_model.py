#!/usr/bin/env python3
"""Lane-level model of sg_modinv_warp (csrc/modinv.cuh): one modular inverse computed by a warp.

Lanes 0..NL-1 hold limb j of (f, g), lanes 16..16+NL-1 limb j of (d, e), as signed 30-bit limbs kept only
PARTIALLY normalised (one carry pass per batch of 30 divsteps).  The model checks the 32/64-bit ranges
the device code relies on and the result against pow(a, -1, p).  Run: python tools/models/safegcd_warp_model.py
"""
import random
import sys

M30 = (1 << 30) - 1


def s32(x):
    assert -(1 << 31) <= x < (1 << 31), x
    return x


def s64(x):
    assert -(1 << 63) <= x < (1 << 63), x
    return x


def u32(x):
    return x & 0xFFFFFFFF


def divsteps30_var(zeta, f0, g0):
    u, v, q, r, f, g = 1, 0, 0, 1, u32(f0), u32(g0)
    i = 30
    while True:
        x = g | (1 << i)
        z = (x & -x).bit_length() - 1
        g >>= z
        u = u32(u << z)
        v = u32(v << z)
        zeta -= z
        i -= z
        if i == 0:
            break
        if zeta < 0:
            zeta = -zeta - 1
            f, g = g, u32(-f)
            u, q = q, u32(-u)
            v, r = r, u32(-v)
        limit = min(zeta + 1, i, 4)
        m = (1 << limit) - 1
        w = u32(f * g * u32(f * f - 2)) & m
        g = u32(g + f * w)
        q = u32(q + u * w)
        r = u32(r + v * w)
    sg = lambda t: t - (1 << 32) if t >> 31 else t
    return zeta, sg(u), sg(v), sg(q), sg(r)


def limbs(x, nl):
    return [(x >> (30 * i)) & M30 for i in range(nl)]


def modinv_warp(a, p, nl, maxb):
    X = [0] * 32   # f (lanes < 16) / d (lanes >= 16)
    Y = [0] * 32   # g / e
    ml = limbs(p, nl)
    for j in range(nl):
        X[j], Y[j] = ml[j], limbs(a, nl)[j]
    Y[16] = 1      # e = 1
    pinv = pow(p, -1, 1 << 30)
    zeta = -1
    batches = 0
    for _ in range(maxb):
        f0, g0 = X[0], Y[0]
        d0, e0, dt, et = X[16], Y[16], X[16 + nl - 1], Y[16 + nl - 1]
        zeta, u, v, q, r = divsteps30_var(zeta, f0, g0)
        batches += 1
        sd, se = (-1 if dt < 0 else 0), (-1 if et < 0 else 0)
        md = (u & sd) + (v & se)
        me = (q & sd) + (r & se)
        cd = u32(u * d0 + v * e0)
        ce = u32(q * d0 + r * e0)
        md -= (pinv * cd + md) & M30
        me -= (pinv * ce + me) & M30
        s32(md); s32(me)
        tx, ty = [0] * 32, [0] * 32
        for l in range(32):
            j, de = l & 15, l >= 16
            if j >= nl:
                continue
            mj = ml[j] if de else 0
            tx[l] = s64(u * X[l] + v * Y[l] + mj * md)
            ty[l] = s64(q * X[l] + r * Y[l] + mj * me)
        assert tx[0] & M30 == 0 and ty[0] & M30 == 0 and tx[16] & M30 == 0 and ty[16] & M30 == 0
        nX, nY, cX, cY = [0] * 32, [0] * 32, [0] * 32, [0] * 32
        for l in range(32):
            j = l & 15
            if j >= nl:
                continue
            for t, n, c in ((tx, nX, cX), (ty, nY, cY)):
                lo_next = (t[l + 1] & M30) if j < nl - 1 else 0
                nx = s64((t[l] >> 30) + lo_next)
                if j < nl - 1:
                    c[l] = s32(nx >> 30)
                    n[l] = nx & M30
                else:
                    c[l] = 0
                    n[l] = s32(nx)
        for l in range(32):
            j = l & 15
            if j >= nl:
                continue
            X[l] = s32(nX[l] + (cX[l - 1] if j > 0 else 0))
            Y[l] = s32(nY[l] + (cY[l - 1] if j > 0 else 0))
            assert -8 <= X[l] <= (1 << 30) + 8 or j == nl - 1, (j, X[l])
        if all(Y[j] == 0 for j in range(nl)):
            break
    # every lane gathers d and the low limb of f, then finishes serially
    d = sum(X[16 + j] << (30 * j) for j in range(nl))
    f = sum(X[j] << (30 * j) for j in range(nl))
    g = sum(Y[j] << (30 * j) for j in range(nl))
    assert g == 0 and (f in (1, -1) or a == 0), (f, g)
    fneg = X[0] != 1   # f = +-1, and the low limb is exact mod 2^30 (a = 0: f = p, d = 0, the sign is irrelevant)
    assert a == 0 or (fneg == (f == -1) and X[0] in (1, M30))
    assert -3 * p < d < 2 * p
    if fneg:
        d = -d
    for _ in range(3):
        if d < 0:
            d += p
    for _ in range(2):
        if d >= p:
            d -= p
    return d, batches


def main():
    rnd = random.Random(1)
    cases = [((1 << 255) - 19, 9, 22), (0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff, 9, 22),
             ((1 << 384) - (1 << 128) - (1 << 96) + (1 << 32) - 1, 13, 32),
             ((1 << 448) - (1 << 224) - 1, 15, 37),
             (0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 13, 32)]
    for p, nl, maxb in cases:
        worst = 0
        edge = [1, 2, p - 1, p - 2, (p + 1) // 2, 3, 1 << 200, (1 << 200) - 1]
        for t in range(3000):
            a = edge[t] if t < len(edge) else rnd.randrange(1, p)
            r, b = modinv_warp(a, p, nl, maxb)
            assert r == pow(a, -1, p), (hex(a), hex(p))
            worst = max(worst, b)
        r, _ = modinv_warp(0, p, nl, maxb)
        assert r == 0
        print("p %d bits: ok, max batches %d (bound %d)" % (p.bit_length(), worst, maxb))


if __name__ == "__main__":
    main()
