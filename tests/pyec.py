"""Independent pure-Python (big-int, affine) elliptic-curve arithmetic for BN254 / BLS12-381 G1.

TEST INFRASTRUCTURE ONLY.  This is the *third* implementation used to pin the two CPU oracles
(`oracle/_ref/libzk_ref.so` = unmodified reference C, `oracle/libzk_oracle.so` = our plain-C
restatement): it shares no code and no algorithm with them (affine chord/tangent law on Python
ints, double-and-add), so agreement on canonical affine bytes pins all three.

Curve constants: /root/reference/codegen/src/Zikkurat/CodeGen/Curve/Params.hs:154-166 (BN128),
:189-204 (BLS12-381).  Memory layouts: SURVEY.md section 8a (a1-a3).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

Point = Optional[Tuple[int, int]]  # None = point at infinity


@dataclass(frozen=True)
class Curve:
    name: str          # C symbol prefix used by the reference ("bn128" / "bls12_381")
    p: int
    r: int
    b: int
    gen: Tuple[int, int]
    nlimbs_p: int      # 64-bit limbs per Fp element
    nlimbs_r: int = 4

    @property
    def R(self) -> int:               # Montgomery radix for Fp
        return 1 << (64 * self.nlimbs_p)

    @property
    def Rr(self) -> int:              # Montgomery radix for Fr
        return 1 << (64 * self.nlimbs_r)

    @property
    def fp_bytes(self) -> int:
        return 8 * self.nlimbs_p

    @property
    def affine_bytes(self) -> int:
        return 2 * self.fp_bytes

    @property
    def proj_bytes(self) -> int:
        return 3 * self.fp_bytes

    # ---- group law (affine, a = 0) -------------------------------------------------------
    def is_on_curve(self, P: Point) -> bool:
        if P is None:
            return True
        x, y = P
        return (y * y - x * x * x - self.b) % self.p == 0

    def neg(self, P: Point) -> Point:
        if P is None:
            return None
        return (P[0], (-P[1]) % self.p)

    def add(self, P: Point, Q: Point) -> Point:
        if P is None:
            return Q
        if Q is None:
            return P
        p = self.p
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = (3 * x1 * x1) * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        y3 = (lam * (x1 - x3) - y1) % p
        return (x3, y3)

    def mul(self, k: int, P: Point) -> Point:
        if k < 0:
            return self.mul(-k, self.neg(P))
        acc: Point = None
        while k:
            if k & 1:
                acc = self.add(acc, P)
            P = self.add(P, P)
            k >>= 1
        return acc

    def msm(self, ks: Sequence[int], Ps: Sequence[Point]) -> Point:
        acc: Point = None
        for k, P in zip(ks, Ps):
            acc = self.add(acc, self.mul(k, P))
        return acc

    # ---- encodings (reference memory layout) --------------------------------------------
    def fp_to_bytes(self, x: int) -> bytes:
        """canonical integer -> Montgomery form, little-endian 64-bit limbs."""
        return ((x * self.R) % self.p).to_bytes(self.fp_bytes, "little")

    def fp_from_bytes(self, bs: bytes) -> int:
        return (int.from_bytes(bs, "little") * pow(self.R, -1, self.p)) % self.p

    def affine_to_bytes(self, P: Point) -> bytes:
        if P is None:                      # bn128_G1_affine.c:43-49,88-91
            return b"\xff" * self.affine_bytes
        return self.fp_to_bytes(P[0]) + self.fp_to_bytes(P[1])

    def affine_from_bytes(self, bs: bytes) -> Point:
        assert len(bs) == self.affine_bytes
        if bs == b"\xff" * self.affine_bytes:
            return None
        n = self.fp_bytes
        return (self.fp_from_bytes(bs[:n]), self.fp_from_bytes(bs[n:]))

    def proj_from_bytes(self, bs: bytes) -> Point:
        """(X:Y:Z) homogeneous projective, Montgomery limbs -> affine point."""
        n = self.fp_bytes
        X, Y, Z = (self.fp_from_bytes(bs[i * n:(i + 1) * n]) for i in range(3))
        if Z == 0:
            return None
        zi = pow(Z, -1, self.p)
        return (X * zi % self.p, Y * zi % self.p)

    def jac_from_bytes(self, bs: bytes) -> Point:
        n = self.fp_bytes
        X, Y, Z = (self.fp_from_bytes(bs[i * n:(i + 1) * n]) for i in range(3))
        if Z == 0:
            return None
        zi = pow(Z, -1, self.p)
        return (X * zi * zi % self.p, Y * zi * zi * zi % self.p)

    def scalar_std_bytes(self, k: int) -> bytes:
        return k.to_bytes(8 * self.nlimbs_r, "little")

    def scalar_mont_bytes(self, k: int) -> bytes:
        return ((k * self.Rr) % self.r).to_bytes(8 * self.nlimbs_r, "little")

    def points_to_bytes(self, Ps: Sequence[Point]) -> bytes:
        return b"".join(self.affine_to_bytes(P) for P in Ps)


BN254 = Curve(
    name="bn128",
    p=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    b=3,
    gen=(1, 2),
    nlimbs_p=4,
)

BLS12_381 = Curve(
    name="bls12_381",
    p=4002409555221667393417789825735904156556882819939007885332058136124031650490837864442687629129015664037894272559787,
    r=52435875175126190479447740508185965837690552500527637822603658699938581184513,
    b=4,
    gen=(
        3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
        1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569,
    ),
    nlimbs_p=6,
)

CURVES = {"bn128": BN254, "bls12_381": BLS12_381}


def chain_points(curve: Curve, n: int, s0: int = 0x1234567, s1: int = 0x7654321) -> List[Point]:
    """P_i = (s0 + i*s1)*G built as an affine chain."""
    G = curve.gen
    P = curve.mul(s0, G)
    D = curve.mul(s1, G)
    out = []
    for _ in range(n):
        out.append(P)
        P = curve.add(P, D)
    return out
