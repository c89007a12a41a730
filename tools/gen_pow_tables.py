"""Generate simplemath_b200/csrc/smb_pow_tables.h: the two small lookup tables of
the f32 pow kernel (see smb_math.cuh, "table-driven f32 pow core").

    python tools/gen_pow_tables.py [output path]

f32 LOG table, 128 entries indexed by the top 7 mantissa bits of |x| = 2^E * m, m in [1, 2):
    invc   k/256, an 8-bit reciprocal of the entry's centre.  With only 8 significant bits
           r = m*invc - 1 is EXACT in one f32 FMA (|r| <= 2^-7, so M*k - 2^31 fits 24 bits);
           exactly 1 for the first entry and exactly 1/2 for the last two, so values close to a
           power of two -- in particular close to 1 -- keep full RELATIVE accuracy
           (log2 x = E + L + log2(1 + r) with E + L == 0)
    L_hi   -log2(invc) rounded to a multiple of 2^-15, MINUS 127 (the exponent bias) MINUS j/128:
           the kernel converts the upper half-word of x, [sign | biased exponent | j], in one
           instruction; float(that)/128 + L_hi is exact in f32 and equals E + L_hi
    L_lo   -log2(invc) - (L_hi + 127)             (f32)
  and the coefficients of log2(1 + r) = C1*r + r^2*(C2 + C3 r [+ C4 r^2 + C5 r^3]) fitted on the
  table's actual range of r (Chebyshev-node interpolation, near-minimax).
f64 LOG table: {c, L_hi - 1023, L_lo} with p = (m - c)/(m + c) (see pow_f64_fast).
EXP table, 64 entries: 2^(j/64) as T_hi * (1 + T_rel) (two f32 / two f64): the relative form folds the
           correction into the exp2 polynomial's last fma.
"""
import os
import struct

import mpmath as mp

mp.mp.prec = 200
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "simplemath_b200", "csrc", "smb_pow_tables.h")
NL, NE = 128, 64


def f32(x) -> float:
    return struct.unpack("f", struct.pack("f", float(x)))[0]


def from_bits(b: int) -> float:
    return struct.unpack("f", struct.pack("I", b))[0]


def from_bits_of(x: float) -> int:
    return struct.unpack("I", struct.pack("f", x))[0]


def main():
    log_rows, worst_r = [], 0.0
    for j in range(NL):
        lo_b, hi_b = 0x3F800000 + (j << 16), 0x3F800000 + (j << 16) + 0xFFFF
        lo, hi = from_bits(lo_b), from_bits(hi_b)
        if j == 0:
            k = 256
        elif j >= NL - 2:
            k = 128
        else:  # the 8-bit reciprocal with the smallest worst-case |r| over the entry
            k = min(range(128, 257), key=lambda kk: max(abs(lo * kk / 256 - 1), abs(hi * kk / 256 - 1)))
        invc = k / 256.0
        for mb in (lo_b, hi_b):  # exactness of r = fma(m, invc, -1): (M*k - 2^31) * 2^-31 must fit 24 bits
            M = (mb & 0x7FFFFF) | 0x800000
            assert abs(M * k - 2 ** 31) <= 2 ** 24, (j, k)
        rmax = max(abs(lo * invc - 1), abs(hi * invc - 1))
        worst_r = max(worst_r, rmax)
        L = -mp.log(mp.mpf(invc), 2)
        L_hi = float(mp.nint(L * 2 ** 15) / 2 ** 15)
        L_lo = f32(L - mp.mpf(L_hi))
        # fast two-sum precondition in the kernel: |E + L_hi| >= |C1 r| or E + L_hi == 0, for every integer E
        for E in range(-3, 4):
            h1 = E + L_hi
            assert h1 == 0 or abs(h1) >= 1.4427 * rmax, (j, E, h1, rmax)
        log_rows.append((invc, L_hi - 127.0 - j / 128.0, L_lo))
    # log2(1 + r) = r/ln2 + r^2 g(r); fit g on [-R, R] at Chebyshev nodes
    R = mp.mpf(worst_r) * mp.mpf("1.0001")

    def g(r):
        return (mp.log(1 + r, 2) - r / mp.log(2)) / (r * r)

    def cheb_fit(deg):
        n = deg + 1
        xs = [R * mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
        A = mp.matrix(n, n)
        b = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for d in range(n):
                A[i, d] = x ** d
            b[i] = g(x)
        sol = mp.lu_solve(A, b)
        co = [f32(sol[d]) for d in range(n)]
        err = max(abs((sum(mp.mpf(co[d]) * x ** d for d in range(n)) - g(x)) * x * x)
                  for x in [R * (mp.mpf(i) / 500 - 1) for i in range(1001)] if x != 0)
        return co, float(err)

    poly_small, err_small = cheb_fit(1)
    poly_large, err_large = cheb_fit(3)
    c1 = 1 / mp.log(2)
    c1h = f32(c1)
    c1l = f32(c1 - mp.mpf(c1h))
    # large-y variant: {c, log2 c} with p = (m - c)/(m + c); c = 1 / c = 2 unit entries
    logc_rows, worst_unit, worst_reg = [], 0.0, 0.0
    for j in range(NL):
        lo, hi = 1.0 + j / NL, from_bits(from_bits_of(1.0 + (j + 1) / NL) - 1)
        if hi <= 1.012:
            c, unit = 1.0, True
        elif lo >= 2 * 0.988:
            c, unit = 2.0, True
        else:
            c, unit = f32((mp.mpf(lo) + mp.mpf(hi)) / 2), False
        L = mp.log(mp.mpf(c), 2)
        L_hi = float(mp.nint(L * 2 ** 15) / 2 ** 15)
        L_lo = f32(L - mp.mpf(L_hi))
        pm = max(abs((lo - c) / (lo + c)), abs((hi - c) / (hi + c)))
        if unit:
            worst_unit = max(worst_unit, pm)
        else:
            worst_reg = max(worst_reg, pm)
        logc_rows.append((c, L_hi - 127.0 - j / 128.0, L_lo))
    exp_rows = []
    for j in range(NE):
        t = mp.power(2, mp.mpf(j) / NE)
        t_hi = f32(t)
        t_rel = f32((t - mp.mpf(t_hi)) / mp.mpf(t_hi))
        exp_rows.append((t_hi, t_rel))
    # ---- f64 tables: same indexing, L_hi a multiple of 2^-40 minus the double bias 1023
    log64, exp64 = [], []
    for j in range(NL):
        lo, hi = 1.0 + j / NL, 1.0 + (j + 1) / NL
        if hi <= 1.012:
            c = 1.0
        elif lo >= 2 * 0.988:
            c = 2.0
        else:
            c = float((mp.mpf(lo) + mp.mpf(hi)) / 2)
        L = mp.log(mp.mpf(c), 2)
        L_hi = float(mp.nint(L * 2 ** 40) / 2 ** 40)
        L_lo = float(L - mp.mpf(L_hi))
        log64.append((c, L_hi - 1023.0, L_lo))
    for j in range(NE):
        t = mp.power(2, mp.mpf(j) / NE)
        t_hi = float(t)
        exp64.append((t_hi, float((t - mp.mpf(t_hi)) / mp.mpf(t_hi))))
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/gen_pow_tables.py -- do not edit.\n")
        f.write("// Lookup tables of the table-driven f32 / f64 pow cores (smb_math.cuh).\n")
        f.write(f"// f32 log table: max |r| = {worst_r:.6f}; log2(1+r) fit error: small-y {err_small:.3e}, large-y {err_large:.3e}\n")
        f.write("#pragma once\n\n")
        f.write("#define SMB_POW_LOG_ENTRIES 128\n#define SMB_POW_EXP_ENTRIES 64\n\n")
        f.write("// {invc, L_hi - 127 - j/128, L_lo, 0}: -log2(invc) = L_hi + L_lo, L_hi a multiple of 2^-15, invc = k/256\n")
        f.write("#define SMB_POW_LOG_TABLE_INIT { \\\n")
        for c, lh, ll in log_rows:
            f.write(f"    {{{c!r}f, {lh!r}f, {ll!r}f, 0.0f}}, \\\n")
        f.write("}\n\n// large-y variant {c, L_hi - 127 - j/128, L_lo, 0}: log2(c) = L_hi + L_lo; p = (m-c)/(m+c), max |p| "
                f"{max(worst_unit, worst_reg):.6f}\n#define SMB_POW_LOGC_TABLE_INIT {{ \\\n")
        for c, lh, ll in logc_rows:
            f.write(f"    {{{c!r}f, {lh!r}f, {ll!r}f, 0.0f}}, \\\n")
        f.write("}\n\n// log2(1 + r) = (C1H + C1L) r + r^2 (C2 + C3 r [+ C4 r^2 + C5 r^3])\n")
        f.write(f"#define SMB_POW_C1H {c1h!r}f\n#define SMB_POW_C1L {c1l!r}f\n")
        f.write(f"#define SMB_POW_S_C2 {poly_small[0]!r}f\n#define SMB_POW_S_C3 {poly_small[1]!r}f\n")
        for d in range(4):
            f.write(f"#define SMB_POW_L_C{d + 2} {poly_large[d]!r}f\n")
        f.write("\n// {T_hi, T_rel}: 2^(j/64) = T_hi * (1 + T_rel)\n#define SMB_POW_EXP_TABLE_INIT { \\\n")
        for th, tl in exp_rows:
            f.write(f"    {{{th!r}f, {tl!r}f}}, \\\n")
        f.write("}\n\n// f64: {c, L_hi - 1023, L_lo, 0}: log2(c) = L_hi + L_lo, L_hi a multiple of 2^-40\n")
        f.write("#define SMB_POW64_LOG_TABLE_INIT { \\\n")
        for c, lh, ll in log64:
            f.write(f"    {{{c!r}, {lh!r}, {ll!r}, 0.0}}, \\\n")
        f.write("}\n\n// f64: {T_hi, T_rel}: 2^(j/64) = T_hi * (1 + T_rel)\n#define SMB_POW64_EXP_TABLE_INIT { \\\n")
        for th, tl in exp64:
            f.write(f"    {{{th!r}, {tl!r}}}, \\\n")
        f.write("}\n")
    print(f"wrote {OUT}; f32 max |r| {worst_r:.6f}, fit errors {err_small:.3e} / {err_large:.3e}")


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1:
        OUT = sys.argv[1]   # regenerate somewhere else (tests compare it with the committed header)
    main()
