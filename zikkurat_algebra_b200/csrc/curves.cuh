// The four groups served by the library: G1 and G2 of BN254 ("bn128") and BLS12-381.
// A curve type provides Fp (coordinate field of the group: the base field for G1, Fp2 for G2) and Fr (scalars).
#pragma once
#include "curve_params.cuh"
#include "fp.cuh"
#include "fp2.cuh"

namespace zk {

ZK_DEFINE_EXT2(Bn254Fp)
ZK_DEFINE_EXT2(Bls12381Fp)

// G2: y^2 = x^3 + b' over Fp2 (a = 0, so the constant never enters the XYZZ formulas);
// reference: lib/cbits/curves/g2/{affine,proj}/<curve>_G2_*.c
struct Bn254G2 {
  using Fp = Ext2<Bn254Fp>;
  using Fr = Bn254Fr;
};
struct Bls12381G2 {
  using Fp = Ext2<Bls12381Fp>;
  using Fr = Bls12381Fr;
};

}  // namespace zk
