#!/usr/bin/env python3
"""Emit zikkurat_algebra_b200/csrc/curve_params.cuh: 32-bit-limb constants for BN254 / BLS12-381.

Primes are the reference's (codegen/src/Zikkurat/CodeGen/Curve/Params.hs:154-166,189-204);
everything else (R mod p, R^2 mod p, -p^-1 mod 2^32, ...) is derived here with Python ints.
"""
import os

CURVES = {
    "Bn254": dict(
        sym="bn128",
        p=21888242871839275222246405745257275088696311157297823662689037894645226208583,
        r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
        b=3,
        gen=(1, 2),
    ),
    "Bls12381": dict(
        sym="bls12_381",
        p=4002409555221667393417789825735904156556882819939007885332058136124031650490837864442687629129015664037894272559787,
        r=52435875175126190479447740508185965837690552500527637822603658699938581184513,
        b=4,
        gen=(3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
             1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569),
    ),
}


def limbs(x, n):
    return ", ".join(f"0x{(x >> (32 * i)) & 0xFFFFFFFF:08x}u" for i in range(n))


def field(name, mod, n):
    R = 1 << (32 * n)
    inv = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    out = [f"struct {name} {{"]
    out.append(f"  static constexpr int L = {n};")
    out.append(f"  static constexpr int BITS = {mod.bit_length()};")
    out.append(f"  static constexpr uint32_t INV = 0x{inv:08x}u;   // -mod^-1 mod 2^32")
    out.append(f"  static constexpr bool THREE_MOD_FITS = {'true' if 3 * mod < R else 'false'};   // 3*mod < 2^(32L): row bound of the fused / squaring products")
    for nm, v in (("MOD", mod), ("ONE", R % mod), ("R2", R * R % mod), ("R3", R * R * R % mod), ("MOD2", 2 * mod),
                  ("HALF", (mod + 1) // 2 * R % mod)):
        out.append(f"  ZK_HD static constexpr uint32_t {nm.lower()}(int i) {{")
        out.append(f"    constexpr uint32_t t[{n}] = {{{limbs(v, n)}}};")
        out.append("    return t[i];")
        out.append("  }")
    out.append("};")
    return "\n".join(out)


def _sqrt_mod(a, p):
    """Tonelli-Shanks."""
    a %= p
    assert pow(a, (p - 1) // 2, p) == 1
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c, t, r = i, b * b % p, t * b * b % p, r * b % p
    return r


def _affine_mul(k, P, p):
    """double-and-add on y^2 = x^3 + b (a = 0), affine big ints; only used to match beta with lambda"""
    def add(A, B):
        if A is None:
            return B
        if B is None:
            return A
        if A[0] == B[0]:
            if (A[1] + B[1]) % p == 0:
                return None
            l = 3 * A[0] * A[0] * pow(2 * A[1], -1, p) % p
        else:
            l = (B[1] - A[1]) * pow(B[0] - A[0], -1, p) % p
        x = (l * l - A[0] - B[0]) % p
        return (x, (l * (A[0] - x) - A[1]) % p)
    R = None
    while k:
        if k & 1:
            R = add(R, P)
        P = add(P, P)
        k >>= 1
    return R


GLV_T = 320   # fixed-point position of the rounded quotients


def glv(cname, c):
    """GLV constants: the endomorphism phi(x, y) = (beta x, y) acts as multiplication by lambda on the r-torsion
    (lambda^2 + lambda + 1 = 0 mod r; the reference lists (beta, lambda) in Params.hs:162-165,200-203 but never uses them).
    A reduced basis (a1, b1), (a2, b2) of the lattice {(x, y): x + y lambda = 0 mod r} with determinant +r gives
      c1 = round(b2 k / r), c2 = round(-b1 k / r),  k1 = k - c1 a1 - c2 a2,  k2 = -c1 b1 - c2 b2,  k = k1 + k2 lambda (mod r)
    with |k1|, |k2| < 2^127 for every k < 2^256 (checked by tests/test_host_emul.py against this very derivation)."""
    import math
    p, r = c["p"], c["r"]
    s = _sqrt_mod(-3, r)
    lam = min((-1 + s) * pow(2, -1, r) % r, (-1 - s) * pow(2, -1, r) % r)
    sp = _sqrt_mod(-3, p)
    G = c["gen"]
    beta = None
    for b in ((-1 + sp) * pow(2, -1, p) % p, (-1 - sp) * pow(2, -1, p) % p):
        if _affine_mul(lam, G, p) == (G[0] * b % p, G[1]):
            beta = b
    assert beta is not None
    rows = [(1, 0, r), (0, 1, lam)]
    while rows[-1][2] != 0:
        q = rows[-2][2] // rows[-1][2]
        rows.append(tuple(a - q * b for a, b in zip(rows[-2], rows[-1])))
    sq = math.isqrt(r)
    l = max(i for i, row in enumerate(rows) if row[2] >= sq)
    a1, b1 = rows[l + 1][2], -rows[l + 1][1]
    a2, b2 = min((rows[l][2], -rows[l][1]), (rows[l + 2][2], -rows[l + 2][1]), key=lambda v: v[0] * v[0] + v[1] * v[1])
    if a1 * b2 - a2 * b1 == -r:
        a1, b1, a2, b2 = a2, b2, a1, b1
    assert a1 * b2 - a2 * b1 == r and (a1 + b1 * lam) % r == 0 and (a2 + b2 * lam) % r == 0
    sgn = lambda v: -1 if v < 0 else 1
    sc1, sc2 = sgn(b2), sgn(-b1)
    g1 = ((abs(b2) << GLV_T) + r // 2) // r
    g2 = ((abs(b1) << GLV_T) + r // 2) // r
    assert max(g1, g2) < 1 << 224 and max(abs(a1), abs(a2), abs(b1), abs(b2)) < 1 << 128
    lp = (p.bit_length() + 63) // 64 * 2
    R = 1 << (32 * lp)
    out = [f"// endomorphism phi(x, y) = (beta x, y) = [lambda](x, y), lambda = 0x{lam:x}",
           f"struct {cname}Glv {{",
           f"  static constexpr int SHIFT = {GLV_T};       // c_j = (k * G_j + 2^(SHIFT-1)) >> SHIFT",
           f"  static constexpr int BITS = 127;        // |k1|, |k2| < 2^127",
           f"  static constexpr int S11 = {-sc1 * sgn(a1)}, S12 = {-sc2 * sgn(a2)};   // k1 = k + S11 |c1||a1| + S12 |c2||a2|",
           f"  static constexpr int S21 = {-sc1 * sgn(b1)}, S22 = {-sc2 * sgn(b2)};   // k2 =     S21 |c1||b1| + S22 |c2||b2|"]
    for nm, v, n in (("BETA", beta * R % p, lp), ("G1", g1, 7), ("G2", g2, 7), ("A1", abs(a1), 4), ("B1", abs(b1), 4), ("A2", abs(a2), 4),
                     ("B2", abs(b2), 4)):
        out.append(f"  ZK_HD static constexpr uint32_t {nm.lower()}(int i) {{")
        out.append(f"    constexpr uint32_t t[{n}] = {{{limbs(v, n)}}};")
        out.append("    return t[i];")
        out.append("  }")
    out.append("};")
    return "\n".join(out)


def main():
    parts = [
        "// GENERATED by tools/gen_params.py -- do not edit.",
        "// 32-bit little-endian limb constants for the two supported curves.",
        "#pragma once",
        "#include <stdint.h>",
        '#include "hd.cuh"',
        "",
        "namespace zk {",
        "",
    ]
    for cname, c in CURVES.items():
        lp = (c["p"].bit_length() + 63) // 64 * 2
        parts.append(field(f"{cname}Fp", c["p"], lp))
        parts.append("")
        parts.append(field(f"{cname}Fr", c["r"], 8))
        parts.append("")
        parts.append(f"struct {cname} {{")
        parts.append(f"  using Fp = {cname}Fp;")
        parts.append(f"  using Fr = {cname}Fr;")
        parts.append(f"  static constexpr int B = {c['b']};")
        parts.append(f"  static constexpr const char* name() {{ return \"{c['sym']}\"; }}")
        parts.append("};")
        parts.append("")
        parts.append(glv(cname, c))
        parts.append("")
    parts.append("}  // namespace zk")
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "zikkurat_algebra_b200", "csrc", "curve_params.cuh")
    with open(dst, "w") as f:
        f.write("\n".join(parts) + "\n")
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
