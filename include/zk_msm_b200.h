/* zk_msm_b200.h -- C ABI of libzkmsm_b200.so: B200 (sm_100a) G1 multi-scalar multiplication for
 * BN254 ("bn128") and BLS12-381, a drop-in for the MSM entry points of bkomuves/zikkurat-algebra's
 * generated C library.  Plain pointers and sizes only; there is NO CPU fallback: every entry point
 * needs a CUDA device and aborts (message on stderr + abort(), the reference's own convention for
 * malloc failure: lib/cbits/curves/g1/proj/bn128_G1_proj.c:516-517,630-631) when CUDA fails.
 *
 * Memory layouts are the reference's:
 *   scalars  npoints x expo_nlimbs little-endian uint64 limbs; "std" = plain integer (any value
 *            < 2^(64*expo_nlimbs) is accepted), "mont" = k * 2^256 mod r (expo_nlimbs must be 4)
 *   points   npoints x (x || y), Fp coordinates in Montgomery form, 4 (bn128) or 6 (bls12_381)
 *            uint64 limbs each; the point at infinity is the all-0xFF record
 *   result   proj  (X:Y:Z)  3 coordinates, infinity = (0, R mod p, 0)
 *            jac   (X:Y:Z)  3 coordinates, x = X/Z^2, y = Y/Z^3, infinity = (R, R, 0) mod p
 *            affine (x || y), canonical, infinity = all 0xFF   <- bit-identical to the reference
 * proj/jac results are valid representatives accepted by the reference's *_to_affine / *_is_infinity
 * (they are not the same representative the reference's operation order would produce).
 * expo_nlimbs: 1..4 supported (the Haskell bindings always pass 4: .../BN128/G1/Proj.hs:245,262).
 */
#ifndef ZK_MSM_B200_H
#define ZK_MSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the reference's symbols (same names, same signatures) -------------------------------------
 * replaces lib/cbits/curves/g1/proj/bn128_G1_proj.h:43-46  (definitions bn128_G1_proj.c:596-669) */
void bn128_G1_proj_MSM_std_coeff_proj_out   (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G1_proj_MSM_mont_coeff_proj_out  (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G1_proj_MSM_std_coeff_affine_out (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G1_proj_MSM_mont_coeff_affine_out(int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
/* replaces lib/cbits/curves/g1/jac/bn128_G1_jac.h:43-46  (definitions bn128_G1_jac.c:645-718) */
void bn128_G1_jac_MSM_std_coeff_jac_out     (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G1_jac_MSM_mont_coeff_jac_out    (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G1_jac_MSM_std_coeff_affine_out  (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G1_jac_MSM_mont_coeff_affine_out (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
/* replaces lib/cbits/curves/g1/proj/bls12_381_G1_proj.h:43-46 (definitions bls12_381_G1_proj.c:597-670) */
void bls12_381_G1_proj_MSM_std_coeff_proj_out   (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_proj_MSM_mont_coeff_proj_out  (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_proj_MSM_std_coeff_affine_out (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_proj_MSM_mont_coeff_affine_out(int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
/* replaces lib/cbits/curves/g1/jac/bls12_381_G1_jac.h:43-46 (definitions bls12_381_G1_jac.c:646-719) */
void bls12_381_G1_jac_MSM_std_coeff_jac_out     (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_jac_MSM_mont_coeff_jac_out    (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_jac_MSM_std_coeff_affine_out  (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_jac_MSM_mont_coeff_affine_out (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);

/* The un-prototyped but exported "_variable" definitions (bn128_G1_proj.c:506, bn128_G1_jac.c:555 and
 * twins).  window_size is honoured as OUR signed-digit window width c (clamped to [1,24]); the
 * reference asserts 1 <= window_size <= 64 (bn128_G1_proj.c:508). */
void bn128_G1_proj_MSM_std_coeff_proj_out_variable    (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs, int window_size);
void bn128_G1_jac_MSM_std_coeff_jac_out_variable      (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs, int window_size);
void bls12_381_G1_proj_MSM_std_coeff_proj_out_variable(int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs, int window_size);
void bls12_381_G1_jac_MSM_std_coeff_jac_out_variable  (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs, int window_size);

/* ---- next row of the scope table: batch conversions, same names and signatures as the
 * reference (lib/cbits/curves/g1/proj/bn128_G1_proj.h:9-10, .../jac/bn128_G1_jac.h:9-10 and twins).  They feed
 * Proj.msmProj = msm cs (batchToAffine gs)  (lib/src/ZK/Algebra/Curves/BN128/G1/Proj.hs:222-223).  The
 * reference performs N separate inversions; here one inversion per 4 points on the GPU.  Bit-identical output. */
void bn128_G1_proj_batch_to_affine       (int N, const uint64_t *src, uint64_t *tgt);
void bn128_G1_proj_batch_from_affine     (int N, const uint64_t *src, uint64_t *tgt);
void bn128_G1_jac_batch_to_affine        (int N, const uint64_t *src, uint64_t *tgt);
void bn128_G1_jac_batch_from_affine      (int N, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_proj_batch_to_affine   (int N, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_proj_batch_from_affine (int N, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_jac_batch_to_affine    (int N, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_jac_batch_from_affine  (int N, const uint64_t *src, uint64_t *tgt);

/* ---- scope row 8f.2: Fr number-theoretic transform, same names and signatures as the reference
 * (lib/cbits/curves/poly/mont/bn128_poly_mont.h:27-28, definitions bn128_poly_mont.c:418-525 and the bls12_381
 * twin): src, tgt = 2^m canonical Montgomery Fr elements (natural order in and out), gen = generator of the
 * order-2^m subgroup.  forward: tgt[k] = sum_j src[j] gen^(jk); inverse: tgt[j] = 2^-m sum_k src[k] gen^(-jk).
 * Bit-identical output.  It is the step right before the MSM in a KZG commitment (examples/KZG.hs:96,139). */
void bn128_poly_mont_ntt_forward    (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bn128_poly_mont_ntt_inverse    (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_poly_mont_ntt_forward(int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_poly_mont_ntt_inverse(int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);

/* ---- scope row 8f.3: G2 multi-scalar multiplication (points over Fp2 = Fp[u]/(u^2+1), coordinates c0 || c1),
 * same names and signatures as the reference (lib/cbits/curves/g2/proj/bn128_G2_proj.h:43-46, definitions
 * bn128_G2_proj.c:498-660 and the bls12_381 twin).  Same pipeline as G1 with Fp2 arithmetic; affine output is
 * bit-identical, infinity = all 0xFF (2 x 2 coordinates), proj infinity = (0, 1, 0). */
void bn128_G2_proj_MSM_std_coeff_proj_out      (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G2_proj_MSM_mont_coeff_proj_out     (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G2_proj_MSM_std_coeff_affine_out    (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G2_proj_MSM_mont_coeff_affine_out   (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G2_proj_MSM_std_coeff_proj_out  (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G2_proj_MSM_mont_coeff_proj_out (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G2_proj_MSM_std_coeff_affine_out(int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G2_proj_MSM_mont_coeff_affine_out(int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);

/* ---- scope row 8f.4: FFT of G1 group elements (KZG setup, examples/KZG.hs:55), same names and signatures as the
 * reference (lib/cbits/curves/g1/proj/bn128_G1_proj.h:48-49, definitions bn128_G1_proj.c:678-789 and twin):
 * src, tgt = 2^m projective points, gen = Montgomery-form generator of the order-2^m subgroup of Fr.
 * forward: tgt[k] = sum_j gen^(jk) * src[j]; inverse: tgt[j] = 2^-m sum_k gen^(-jk) * src[k]; the results are
 * normalised ((x, y, 1) / (0, 1, 0)) like the reference's, hence bit-identical.
 * The G1 twiddle products use the same endomorphism split as the MSM (k = k1 + k2 lambda, 128 instead of 256 doublings) and
 * therefore share its precondition and its switch: input points in the prime-order subgroup, zkb200_set_glv(0) / $ZKB200_GLV=0
 * for arbitrary curve points (see zkb200_set_glv below).  G2 is not affected. */
void bn128_G1_proj_fft_forward    (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bn128_G1_proj_fft_inverse    (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_proj_fft_forward(int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_proj_fft_inverse(int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
/* the Jacobian-input and G2 twins (lib/cbits/curves/g1/jac/bn128_G1_jac.h:48-49, .../g2/proj/bn128_G2_proj.h:48-49) */
void bn128_G1_jac_fft_forward     (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bn128_G1_jac_fft_inverse     (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_jac_fft_forward (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_G1_jac_fft_inverse (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bn128_G2_proj_fft_forward    (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bn128_G2_proj_fft_inverse    (int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_G2_proj_fft_forward(int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
void bls12_381_G2_proj_fft_inverse(int m, const uint64_t *gen, const uint64_t *src, uint64_t *tgt);
/* G2 batch conversions (lib/cbits/curves/g2/proj/bn128_G2_proj.h:9-10) */
void bn128_G2_proj_batch_to_affine      (int N, const uint64_t *src, uint64_t *tgt);
void bn128_G2_proj_batch_from_affine    (int N, const uint64_t *src, uint64_t *tgt);
void bls12_381_G2_proj_batch_to_affine  (int N, const uint64_t *src, uint64_t *tgt);
void bls12_381_G2_proj_batch_from_affine(int N, const uint64_t *src, uint64_t *tgt);
/* the exported-but-unused slow MSM variants (bn128_G1_proj.c:610-619; header name misspelt at bn128_G1_proj.h:47) */
void bn128_G1_proj_MSM_std_coeff_proj_out_slow_reference    (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G1_jac_MSM_std_coeff_jac_out_slow_reference      (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bn128_G2_proj_MSM_std_coeff_proj_out_slow_reference    (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_proj_MSM_std_coeff_proj_out_slow_reference(int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G1_jac_MSM_std_coeff_jac_out_slow_reference  (int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);
void bls12_381_G2_proj_MSM_std_coeff_proj_out_slow_reference(int npoints, const uint64_t *expos, const uint64_t *grps, uint64_t *tgt, int expo_nlimbs);

/* ---- extensions (not in the reference) -------------------------------------------------------- */
enum { ZKB200_BN128 = 0, ZKB200_BLS12_381 = 1, ZKB200_BN128_G2 = 2, ZKB200_BLS12_381_G2 = 3 };
enum { ZKB200_OUT_PROJ = 0, ZKB200_OUT_JAC = 1, ZKB200_OUT_AFFINE = 2, ZKB200_OUT_XYZZ = 3 };
enum { ZKB200_HOST = 0, ZKB200_DEVICE = 1 };

/* Generic entry point behind all of the above.  nmsm independent MSMs of npoints each over ONE shared
 * point array (SURVEY.md section 8b "batch" extension; KZG commitments over a shared SRS):
 *   scalars  nmsm x npoints x expo_nlimbs limbs,  out  nmsm records of (2|3|4) coordinates.
 * scalars_loc / points_loc say whether the pointer is host or device (current device) memory; `out` is
 * always host memory.  window = 0 picks the tuned window width. */
void zkb200_msm(int curve, int nmsm, long npoints, const uint64_t *scalars, int scalars_loc,
                const uint64_t *points, int points_loc, int expo_nlimbs, int mont_coeff, int out_mode,
                int window, uint64_t *out);

/* Sum of k group elements (multi-GPU combine of partial MSM results): in_mode / out_mode as above
 * (in_mode: PROJ, JAC or XYZZ records in host memory). */
void zkb200_sum_points(int curve, int k, const uint64_t *in, int in_mode, int out_mode, uint64_t *out);

/* The same two calls with the result / the operands in DEVICE memory of the current device, so that a multi-GPU
 * run keeps its partial points on the GPUs: zkb200_msm_ex writes the nmsm result records to `out` (out_loc =
 * ZKB200_DEVICE: device-to-device, nothing but the completion wait touches the host), a collective (NCCL all-gather over
 * NVLink, see zikkurat_algebra_b200/distributed.py) moves them, zkb200_sum_points_ex reads the k records from device
 * memory (in_loc = ZKB200_DEVICE) and returns the sum in host memory.  With *_loc = ZKB200_HOST they are zkb200_msm /
 * zkb200_sum_points. */
void zkb200_msm_ex(int curve, int nmsm, long npoints, const uint64_t *scalars, int scalars_loc,
                   const uint64_t *points, int points_loc, int expo_nlimbs, int mont_coeff, int out_mode,
                   int window, uint64_t *out, int out_loc);
void zkb200_sum_points_ex(int curve, int k, const uint64_t *in, int in_loc, int in_mode, int out_mode, uint64_t *out);

/* Workload synthesis on the GPU (not part of the MSM path): out[i] = P0 + (start + i) * D for
 * i < n as canonical affine Montgomery points, written to host (out_loc = ZKB200_HOST) or device memory.
 * Used by bench.py and the tests to build large synthetic point arrays. */
void zkb200_gen_chain(int curve, unsigned long long start, long n, const uint64_t *p0_affine,
                      const uint64_t *d_affine, uint64_t *out, int out_loc);

/* Fr NTT with explicit buffer locations (ZKB200_HOST / ZKB200_DEVICE): with device buffers the result can feed
 * zkb200_msm's device-resident scalars directly -- a KZG commitment from evaluations (examples/KZG.hs:90-97:
 * inverse NTT, then MSM over the SRS) without the coefficients ever leaving the GPU.  gen: host pointer. */
void zkb200_ntt(int curve, int m, const uint64_t *gen, const uint64_t *src, int src_loc, uint64_t *tgt, int tgt_loc,
                int inverse);

/* Resident inputs without writing any CUDA code: copy a host buffer (e.g. the SRS of a KZG prover, which is the same
 * for every commitment: examples/KZG.hs:77-88) to the selected device once and pass the returned pointer to
 * zkb200_msm / zkb200_ntt with location ZKB200_DEVICE.  Free it with zkb200_device_free. */
void *zkb200_device_upload(const void *host, size_t bytes);
void zkb200_device_free(void *device_ptr);

/* Resident point arrays ("SRS cache").  Every entry point that takes a HOST point array keeps a copy of it on the
 * device after the first call, keyed by (host pointer, byte count, curve) and guarded by a 64-bit fingerprint of a strided
 * sample of its words; later calls over the same array (the SRS of a KZG prover: examples/KZG.hs:77-88,110-116) move only
 * the scalars.  Budget per device: $ZKB200_SRS_CACHE_MB (default 16384, 0 = off), least recently used arrays go first.
 * The reference's point arrays are immutable (lib/src/ZK/Algebra/Class/Flat.hs:81-90); a C caller that rewrites a few
 * points of an array IN PLACE between calls must disable the cache (or pass a fresh buffer), because the fingerprint
 * samples about 1500 words, not all of them.  zkb200_last_srs_hit: 1 when the last MSM on the current device used a
 * resident copy. */
int zkb200_last_srs_hit(void);
/* forget every resident point array on every device (the next call over an array uploads it again) */
void zkb200_srs_cache_drop(void);

/* Endomorphism (GLV) split of the scalars, on by default ($ZKB200_GLV=0 or zkb200_set_glv(0) switch it off): every scalar
 * longer than 160 bits is split as k = k1 + k2*lambda (mod r) with 127-bit halves and the MSM runs over the 2n points
 * P_i, phi(P_i), phi(x, y) = (beta*x, y) -- half the windows for the same number of bucket insertions.
 * PRECONDITION that the reference does not have: phi(P) = lambda*P holds for points of the prime-order subgroup only.
 * BN254 G1 has cofactor 1, so every point on the curve qualifies; BLS12-381 G1 has cofactor 0x396c8c005555e1568c00aaab0000aaab,
 * so a caller that feeds curve points OUTSIDE the subgroup (the reference multiplies them by the integer scalar like
 * any other point: it validates nothing) must switch the split off to get the reference's answer.
 * G1 elements of an SRS, commitments, proofs -- everything a KZG prover handles -- are subgroup points.
 * tests/test_configs_gpu.py::test_glv_precondition_and_switch shows both sides.  The switch also governs the twiddle products
 * of the G1 group FFT (scope row 8f.4 above). */
void zkb200_set_glv(int on);

/* Give back all device memory this library holds on every device it has used (work arrays that only grow otherwise,
 * the pinned result buffer, the resident point arrays).  The next call allocates again.  A work-array allocation
 * that fails first drops the resident point arrays and retries; batches that would not fit are processed in halves;
 * only then does the library give up (message + abort(), the reference's convention for malloc failure). */
void zkb200_release_workspaces(void);

/* Element-wise self-tests of the device primitives, used by tests/test_device_primitives.py to compare the compiled
 * PTX carry chains with the reference's field and group functions operand by operand (not part of the MSM path).
 * field: 0 = bn128 Fp, 1 = bls12_381 Fp, 2 = bn128 Fr, 3 = bls12_381 Fr;  op: 0 mul, 1 sqr, 2 a*b+c*d, 3 add, 4 sub, 5 neg,
 * 6 inv (0 -> 0), 7..9 = 0..2 through the out-of-line multiplication, 10 double, 11 Montgomery -> standard form.
 * n elements of 4/6 uint64 limbs per operand (b, c, d may be NULL where unused).
 * group op: 0 XYZZ += affine, 1 the same with out-of-line multiplications, 2 XYZZ + XYZZ, 3 the same out of line,
 * 4 double, 5 double of an affine point.  Operand k is the affine point p_k (all 0xFF = infinity) lifted with the
 * scale z_k to (x z^2, y z^3, z^2, z^3); results are canonical affine records. */
void zkb200_selftest_field(int field, int op, long n, const uint64_t *a, const uint64_t *b, const uint64_t *c,
                           const uint64_t *d, uint64_t *out);
void zkb200_selftest_group(int curve, int op, long n, const uint64_t *p1, const uint64_t *z1, const uint64_t *p2,
                           const uint64_t *z2, uint64_t *out_affine);

/* Device time (ms, CUDA events on the launching stream) of the kernels of the most recent batch conversion, NTT or
 * group-FFT call on the current device, copies excluded (the roofline numerator's time for bench.py --row). */
float zkb200_last_op_ms(void);

/* Number of kernels this library has launched since it was loaded (bench.py's "gpu_launches"). */
long long zkb200_launch_count(void);

/* Device used by subsequent calls of the calling process (default: $ZKB200_DEVICE or 0). */
void zkb200_set_device(int device);

/* Several devices: every host-buffer MSM call is then sharded over them from inside the library (one host
 * thread per device; contiguous slices of one MSM, or whole MSMs of a batch), the partial points are summed
 * on devices[0].  count = 0 or 1 restores single-device operation.  Same as $ZKB200_DEVICES="0,1,.."|"all". */
void zkb200_set_devices(const int *devices, int count);

/* Phase timings (CUDA events, ms) of the most recent zkb200_msm on this thread's device:
 *  [0] h2d scalars  [1] recode  [2] sort  [3] wait for points h2d  [4] accumulate  [5] fixup
 *  [6] reduce  [7] tail + d2h   [8] total on the compute stream.
 * The accumulation runs on several streams (lanes); the phases are what the caller's stream sees:
 *  [4] = until every lane has finished (affine pre-reduction, XYZZ accumulation AND the fix-up tree, so [5] ~ 0),
 *  [6] = what is left of the per-group bucket reduction / window combination after that point (the rest ran under
 *  the accumulation), [7] = sum of the groups' shares, output conversion, D2H.
 * Also: window width c, number of windows W, insertions n*W, of the same call. */
void zkb200_last_stats(float phase_ms[9], int *window_c, int *nwindows, long long *insertions);
/* levels of the batched-affine pre-reduction the last MSM on the current device used (0 = plain XYZZ accumulation) */
int zkb200_last_affine_levels(void);

/* Register-resident integer-multiply throughput probe: returns 32x32-bit products per second of the
 * whole GPU for kind 0 = mad.lo/madc.hi carry chains (as used by the field code), 1 = mad.wide.u32 with a 64-bit
 * accumulator and no carry, 2 = 32-bit mad.lo.u32 only, 3 = mul.wide.u32 (product only).  Every repetition uses its own
 * multiplier so that ptxas cannot reuse a product.  Used for the IMAD roofline denominator. */
double zkb200_imad_peak(int kind, int iters);

const char *zkb200_version(void);

#ifdef __cplusplus
}
#endif
#endif
