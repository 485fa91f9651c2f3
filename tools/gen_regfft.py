#!/usr/bin/env python3
"""Generates multiview-simulation_b200/csrc/fft/regfft_gen.cuh: straight-line in-register complex
FFTs (natural order in, natural order out) for the sub-transform sizes the two-level line FFT uses.

Every size is decomposed recursively (radix 4, 2, 3, 5, decimation in time) with the twiddles
folded in as float literals (computed in double); trivial twiddles (1, -i, ...) emit no multiplies.
The emitted functions are __host__ __device__ so the CPU emulation tests run the same code.
"""
import math
import sys

SIZES = [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 18, 20, 24, 25, 27, 30, 32, 36, 40]


class Gen:
    def __init__(self, sign):
        self.sign = sign
        self.lines = []
        self.n = 0

    def tmp(self):
        self.n += 1
        return f"t{self.n}"

    def emit(self, expr):
        v = self.tmp()
        self.lines.append(f"const float {v} = {expr};")
        return v

    @staticmethod
    def lit(x):
        import numpy as np
        return repr(float(np.float32(x))) + "f"

    def add(self, a, b):
        return (self.emit(f"{a[0]} + {b[0]}"), self.emit(f"{a[1]} + {b[1]}"))

    def sub(self, a, b):
        return (self.emit(f"{a[0]} - {b[0]}"), self.emit(f"{a[1]} - {b[1]}"))

    def scale(self, a, c):
        return (self.emit(f"{a[0]} * {self.lit(c)}"), self.emit(f"{a[1]} * {self.lit(c)}"))

    def add_i(self, a, b, s):
        """a + s*i*b  (s = +1/-1)"""
        if s > 0:
            return (self.emit(f"{a[0]} - {b[1]}"), self.emit(f"{a[1]} + {b[0]}"))
        return (self.emit(f"{a[0]} + {b[1]}"), self.emit(f"{a[1]} - {b[0]}"))

    def fma2(self, a, b, cb, c=None, cc=None):
        """a + cb*b (+ cc*c), component-wise with real constants"""
        if c is None:
            return (self.emit(f"fmaf({self.lit(cb)}, {b[0]}, {a[0]})"),
                    self.emit(f"fmaf({self.lit(cb)}, {b[1]}, {a[1]})"))
        return (self.emit(f"fmaf({self.lit(cc)}, {c[0]}, fmaf({self.lit(cb)}, {b[0]}, {a[0]}))"),
                self.emit(f"fmaf({self.lit(cc)}, {c[1]}, fmaf({self.lit(cb)}, {b[1]}, {a[1]}))"))

    def lin2(self, b, cb, c, cc):
        """cb*b + cc*c"""
        return (self.emit(f"fmaf({self.lit(cc)}, {c[0]}, {b[0]} * {self.lit(cb)})"),
                self.emit(f"fmaf({self.lit(cc)}, {c[1]}, {b[1]} * {self.lit(cb)})"))

    def twiddle(self, a, num, den):
        """a * exp(sign * 2 pi i num / den)"""
        num %= den
        if num == 0:
            return a
        g = math.gcd(num, den)
        num //= g
        den //= g
        if den == 2:       # -1
            return (self.emit(f"-{a[0]}"), self.emit(f"-{a[1]}"))
        if den == 4:       # +-i
            s = self.sign if num == 1 else -self.sign
            # a * (s*i) = (-s*a.im, s*a.re)
            if s > 0:
                return (self.emit(f"-{a[1]}"), a[0])
            return (a[1], self.emit(f"-{a[0]}"))
        ang = self.sign * 2.0 * math.pi * num / den
        c, s = math.cos(ang), math.sin(ang)
        if den == 8:
            r = math.sqrt(0.5)
            sc = 1 if c > 0 else -1
            ss = 1 if s > 0 else -1
            # (a.re*c - a.im*s, a.re*s + a.im*c) with |c|=|s|=r
            re = self.emit(f"({'' if sc > 0 else '-'}{a[0]} {'-' if ss > 0 else '+'} {a[1]}) * {self.lit(r)}")
            im = self.emit(f"({'' if ss > 0 else '-'}{a[0]} {'+' if sc > 0 else '-'} {a[1]}) * {self.lit(r)}")
            return (re, im)
        re = self.emit(f"fmaf({a[0]}, {self.lit(c)}, -{a[1]} * {self.lit(s)})")
        im = self.emit(f"fmaf({a[0]}, {self.lit(s)}, {a[1]} * {self.lit(c)})")
        return (re, im)

    def butterfly(self, r, a):
        sg = self.sign
        if r == 2:
            return [self.add(a[0], a[1]), self.sub(a[0], a[1])]
        if r == 4:
            t0 = self.add(a[0], a[2])
            t1 = self.sub(a[0], a[2])
            t2 = self.add(a[1], a[3])
            t3 = self.sub(a[1], a[3])
            return [self.add(t0, t2), self.add_i(t1, t3, sg), self.sub(t0, t2), self.add_i(t1, t3, -sg)]
        if r == 3:
            t1 = self.add(a[1], a[2])
            t2 = self.fma2(a[0], t1, -0.5)
            d = self.scale(self.sub(a[1], a[2]), math.sqrt(3.0) / 2.0)
            return [self.add(a[0], t1), self.add_i(t2, d, sg), self.add_i(t2, d, -sg)]
        if r == 5:
            c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
            s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
            t1 = self.add(a[1], a[4])
            t2 = self.add(a[2], a[3])
            t3 = self.sub(a[1], a[4])
            t4 = self.sub(a[2], a[3])
            x0 = self.add(a[0], self.add(t1, t2))
            m1 = self.fma2(a[0], t1, c1, t2, c2)
            m2 = self.fma2(a[0], t1, c2, t2, c1)
            n1 = self.lin2(t3, s1, t4, s2)
            n2 = self.lin2(t3, s2, t4, -s1)
            return [x0, self.add_i(m1, n1, sg), self.add_i(m2, n2, sg), self.add_i(m2, n2, -sg), self.add_i(m1, n1, -sg)]
        raise ValueError(r)

    def fft(self, n, x):
        if n == 1:
            return x
        if n in (2, 3, 4, 5):
            return self.butterfly(n, x)
        for r in (4, 2, 3, 5):
            if n % r == 0:
                break
        else:
            raise ValueError(n)
        m = n // r
        subs = [self.fft(m, x[q::r]) for q in range(r)]
        out = [None] * n
        for k in range(m):
            t = [self.twiddle(subs[q][k], q * k, n) for q in range(r)]
            u = self.butterfly(r, t)
            for j in range(r):
                out[k + m * j] = u[j]
        return out


class GenPacked(Gen):
    """Same dataflow, but every complex value is one float2 and the arithmetic uses Blackwell's packed FP32x2
    instructions (add2 / mul2 / fma2 = __fadd2_rn / __fmul2_rn / __ffma2_rn -> FADD2 / FMUL2 / FFMA2): one instruction per complex add."""

    def emit(self, expr):
        v = self.tmp()
        self.lines.append(f"const float2 {v} = {expr};")
        return v

    def pair(self, a, b):
        return f"make_float2({self.lit(a)}, {self.lit(b)})"

    def swap(self, a):
        return f"make_float2({a}.y, {a}.x)"

    def add(self, a, b):
        return self.emit(f"add2({a}, {b})")

    def sub(self, a, b):
        return self.emit(f"fma2({b}, {self.pair(-1.0, -1.0)}, {a})")

    def scale(self, a, c):
        return self.emit(f"mul2({a}, {self.pair(c, c)})")

    def add_i(self, a, b, s):
        # a + s*i*b = (a.x - s b.y, a.y + s b.x)
        return self.emit(f"fma2({self.swap(b)}, {self.pair(-s, s)}, {a})")

    def fma2(self, a, b, cb, c=None, cc=None):
        first = f"fma2({b}, {self.pair(cb, cb)}, {a})"
        if c is None:
            return self.emit(first)
        return self.emit(f"fma2({c}, {self.pair(cc, cc)}, {first})")

    def lin2(self, b, cb, c, cc):
        return self.emit(f"fma2({c}, {self.pair(cc, cc)}, mul2({b}, {self.pair(cb, cb)}))")

    def twiddle(self, a, num, den):
        num %= den
        if num == 0:
            return a
        g = math.gcd(num, den)
        num //= g
        den //= g
        if den == 2:
            return self.emit(f"mul2({a}, {self.pair(-1.0, -1.0)})")
        if den == 4:
            s = self.sign if num == 1 else -self.sign
            return self.emit(f"mul2({self.swap(a)}, {self.pair(-s, s)})")
        ang = self.sign * 2.0 * math.pi * num / den
        c, sn = math.cos(ang), math.sin(ang)
        return self.emit(f"fma2({self.swap(a)}, {self.pair(-sn, sn)}, mul2({a}, {self.pair(c, c)}))")


class GenPackedZ(GenPacked):
    """GenPacked with zero propagation: a value that is known to be zero is None and emits nothing.  Used for the forward
    transform of a zero-extended line (the PSF along z): only the first K inputs of the sub-transform carry data."""

    def neg(self, a):
        return self.emit(f"mul2({a}, {self.pair(-1.0, -1.0)})")

    def add(self, a, b):
        if a is None:
            return b
        if b is None:
            return a
        return super().add(a, b)

    def sub(self, a, b):
        if b is None:
            return a
        if a is None:
            return self.neg(b)
        return super().sub(a, b)

    def scale(self, a, c):
        return None if a is None else super().scale(a, c)

    def add_i(self, a, b, s):
        if b is None:
            return a
        if a is None:
            return self.emit(f"mul2({self.swap(b)}, {self.pair(-s, s)})")
        return super().add_i(a, b, s)

    def fma2(self, a, b, cb, c=None, cc=None):
        terms = [(b, cb)] + ([(c, cc)] if c is not None or cc is not None else [])
        terms = [(v, k) for v, k in terms if v is not None]
        acc = a
        for v, k in terms:
            acc = self.emit(f"mul2({v}, {self.pair(k, k)})") if acc is None else self.emit(f"fma2({v}, {self.pair(k, k)}, {acc})")
        return acc

    def lin2(self, b, cb, c, cc):
        return self.fma2(None, b, cb, c, cc)

    def twiddle(self, a, num, den):
        return None if a is None else super().twiddle(a, num, den)


def pruned_inputs(n):
    """number of leading inputs the pruned forward variant reads: ceil(n / 5)"""
    return (n + 4) // 5


def gen_size_packed_pruned(n):
    g = GenPackedZ(-1)
    k = pruned_inputs(n)
    x = [f"x[{i}]" if i < k else None for i in range(n)]
    out = g.fft(n, x)
    body = ["    " + l for l in g.lines]
    for i, v in enumerate(out):
        body.append(f"    x[{i}] = {v if v is not None else 'make_float2(0.f, 0.f)'};")
    return (f"template <> struct RegFFTPZ<{n}> {{\n  static constexpr int K = {k};\n"
            f"  static MVSIM_HD void run(float2 (&x)[{n}]) {{\n" + "\n".join(body) + "\n  }\n};\n")


# ---- inverse transforms of which only the outputs o = r (mod inc) are wanted (the kept planes of extractSlices fall on one
# residue class of the second-level index when inc divides B): full dataflow generated, dead code removed ----------------
DECIM = [(n, inc) for inc in (3, 5) for n in SIZES if n % inc == 0 and n // inc >= 2]


def gen_size_packed_decimated(n, inc, r):
    import re
    g = GenPacked(1)
    x = [f"x[{i}]" for i in range(n)]
    out = g.fft(n, x)
    wanted = [out[r + inc * j] for j in range(n // inc)]
    defs = {}
    for idx, line in enumerate(g.lines):
        m = re.match(r"const float2 (t\d+) = (.*);$", line)
        defs[m.group(1)] = (idx, set(re.findall(r"\bt\d+\b", m.group(2))))
    live, stack = set(), [w for w in wanted if w in defs]
    while stack:
        v = stack.pop()
        if v in live:
            continue
        live.add(v)
        stack.extend(d for d in defs[v][1] if d not in live)
    body = ["    " + line for line in g.lines if re.match(r"const float2 (t\d+) =", line).group(1) in live]
    for j, v in enumerate(wanted):
        body.append(f"    o[{j}] = {v};")
    return (f"template <> struct RegFFTPD<{n}, {inc}, {r}> {{\n"
            f"  static MVSIM_HD void run(const float2 (&x)[{n}], float2 (&o)[{n // inc}]) {{\n" + "\n".join(body) + "\n  }\n};\n"), len(body)


def gen_size_packed(n, sign):
    g = GenPacked(sign)
    x = [f"x[{i}]" for i in range(n)]
    out = g.fft(n, x)
    body = ["    " + l for l in g.lines]
    for i, v in enumerate(out):
        body.append(f"    x[{i}] = {v};")
    name = "-1" if sign < 0 else "1"
    return (f"template <> struct RegFFTP<{n}, {name}> {{\n"
            f"  static MVSIM_HD void run(float2 (&x)[{n}]) {{\n" + "\n".join(body) + "\n  }\n};\n")


def gen_size(n, sign):
    g = Gen(sign)
    x = [(f"x[{i}].x", f"x[{i}].y") for i in range(n)]
    out = g.fft(n, x)
    body = ["    " + l for l in g.lines]
    for i, (re, im) in enumerate(out):
        body.append(f"    x[{i}].x = {re}; x[{i}].y = {im};")
    name = "-1" if sign < 0 else "1"
    return (f"template <> struct RegFFT<{n}, {name}> {{\n"
            f"  static MVSIM_HD void run(float2 (&x)[{n}]) {{\n" + "\n".join(body) + "\n  }\n};\n")


def main_packed(path):
    parts = ["// GENERATED by tools/gen_regfft.py -- do not edit.\n"
             "// Packed variant: complex values are float2, arithmetic is Blackwell FP32x2 (FADD2 / FMUL2 / FFMA2).\n"
             "#pragma once\n#include \"fft_defs.cuh\"\n\nnamespace mvsim {\n\n"
             "template <int N, int DIR> struct RegFFTP;\n\n"
             "template <int DIR> struct RegFFTP<1, DIR> { static MVSIM_HD void run(float2 (&)[1]) {} };\n\n"]
    for n in SIZES:
        for sign in (-1, 1):
            parts.append(gen_size_packed(n, sign))
            parts.append("\n")
    parts.append("// Forward transforms of lines whose inputs x[K..N) are zero (K = ceil(N/5)): zero terms propagated away at generation time.\n"
                 "// x[K..N) is not read; all N outputs are written.\n"
                 "template <int N> struct RegFFTPZ;\n\n")
    for n in SIZES:
        parts.append(gen_size_packed_pruned(n))
        parts.append("\n")
    parts.append("// Inverse transforms restricted to the outputs o = R (mod INC): o[j] = X[R + INC*j]; dead code removed at generation time.\n"
                 "template <int N, int INC, int R> struct RegFFTPD;\n\n")
    for n, inc in DECIM:
        for r in range(inc):
            parts.append(gen_size_packed_decimated(n, inc, r)[0])
            parts.append("\n")
    parts.append("}  // namespace mvsim\n")
    with open(path, "w") as f:
        f.write("".join(parts))


def main(path):
    parts = ["// GENERATED by tools/gen_regfft.py -- do not edit.\n"
             "// In-register complex FFTs, natural order in/out. DIR=-1 forward (exp(-2 pi i nk/N)), +1 inverse (unscaled).\n"
             "#pragma once\n#include \"fft_defs.cuh\"\n\nnamespace mvsim {\n\n"
             "template <int N, int DIR> struct RegFFT;\n\n"
             "template <int DIR> struct RegFFT<1, DIR> { static MVSIM_HD void run(float2 (&)[1]) {} };\n\n"]
    for n in SIZES:
        for sign in (-1, 1):
            parts.append(gen_size(n, sign))
            parts.append("\n")
    parts.append("}  // namespace mvsim\n")
    with open(path, "w") as f:
        f.write("".join(parts))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "multiview-simulation_b200/csrc/fft/regfft_gen.cuh")
    main_packed(sys.argv[2] if len(sys.argv) > 2 else "multiview-simulation_b200/csrc/fft/regfft_gen_packed.cuh")
