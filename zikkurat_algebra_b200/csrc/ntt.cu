// Fr number-theoretic transform on the GPU -- scope row 8f.2 (the step right before the MSM in a KZG
// commitment: examples/KZG.hs:96,139).  Same semantics as the reference's recursive radix-2 routines
//   <curve>_poly_mont_ntt_forward / _inverse      lib/cbits/curves/poly/mont/bn128_poly_mont.c:418-525
// i.e. natural order in and out,
//   forward:  tgt[k] = sum_j src[j] * gen^(j*k)            inverse:  tgt[j] = N^-1 * sum_k src[k] * gen^(-j*k)
// with N = 2^m and `gen` a generator of the order-N subgroup; elements are canonical Montgomery Fr values
// (4 x u64), so the output bytes are identical to the reference's whatever the order of operations.
//
// Algorithm: decimation in time.  The bit reversal is folded into the first pass's gather (one 32-byte
// element = one DRAM sector, so element-granular scatter/gather costs no extra traffic); every pass stages a
// tile of 512 elements in shared memory and runs up to 9 butterfly stages on it, one butterfly per thread and
// stage.  Twiddles come from a table w^i, i < N/2, built on the device at the start of every call (one kernel, N/2 Fr multiplications: about the cost of one butterfly pass; not cached across calls); the inverse uses
// w^-i = -w^(N/2-i) from the same table and multiplies by N^-1 = (1/2)^m in its last pass.
// HBM traffic: ceil(m/9) passes x 64 B per element (+ table reads); 1 Fr multiplication per butterfly.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "curve_params.cuh"
#include "fp.cuh"
#include "ntt.cuh"

namespace zk {

constexpr int NTT_TILE_LOG = 9;
constexpr int NTT_TILE = 1 << NTT_TILE_LOG;   // elements staged per CTA
constexpr int NTT_THREADS = NTT_TILE / 2;     // one butterfly per thread and stage

template <class F>
__device__ __forceinline__ Fe<F> ld_fe8(const uint32_t* p) {
  Fe<F> r;
  uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 4);
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  return r;
}
template <class F>
__device__ __forceinline__ void st_fe8(uint32_t* p, const Fe<F>& v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  *reinterpret_cast<uint4*>(p + 4) = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// table[i] = gen^i for i < half (half = N/2, at least 1); table[half] = (1/2)^m = N^-1
template <class F>
__global__ void __launch_bounds__(128) k_ntt_table(const uint32_t* __restrict__ gen, size_t half, int m, uint32_t* __restrict__ table) {
  constexpr int CH = 32;  // consecutive powers per thread
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t first = t * CH;
  if (first == 0 && t == 0) {
    Fe<F> h, acc = fe_one<F>();
    for (int i = 0; i < 8; i++) h.l[i] = F::half(i);
    for (int i = 0; i < m; i++) acc = fe_mul<F>(acc, h);
    st_fe8<F>(table + half * 8, acc);
  }
  if (first >= half) return;
  Fe<F> g = ld_fe8<F>(gen);
  // g^first by square-and-multiply
  Fe<F> p = fe_one<F>(), b = g;
  for (size_t e = first; e; e >>= 1) {
    if (e & 1) p = fe_mul<F>(p, b);
    b = fe_sqr<F>(b);
  }
  for (int k = 0; k < CH && first + k < half; k++) {
    st_fe8<F>(table + (first + k) * 8, p);
    p = fe_mul<F>(p, g);
  }
}

__device__ __forceinline__ size_t bitrev(size_t x, int bits) {
  return bits == 0 ? 0 : (size_t)(__brevll((unsigned long long)x) >> (64 - bits));
}

// One pass: stages s0+1 .. s0+S on tiles of (2^S x LO) elements, LO = tile/2^S consecutive low indices.
//   global index p = hi * 2^(s0+S) + mid * 2^s0 + lo,  mid in [0, 2^S) varies inside the tile.
// first pass (s0 = 0) gathers from the bit-reversed source position; last pass of the inverse scales by N^-1.
template <class F>
__global__ void __launch_bounds__(NTT_THREADS)
k_ntt_pass(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ table, int m, int s0, int S,
           int inverse, int scale) {
  __shared__ __align__(16) uint32_t sm[NTT_TILE * 8];
  const size_t N = (size_t)1 << m;
  const int tile_log = (m < NTT_TILE_LOG) ? m : NTT_TILE_LOG;
  const int tile = 1 << tile_log;
  const int lo_log = tile_log - S;                 // LO = 2^lo_log
  const int LO = 1 << lo_log;
  const size_t tile_id = blockIdx.x;
  const size_t lo_tiles = ((size_t)1 << s0) >> lo_log;     // tiles along the low index
  const size_t hi = tile_id / lo_tiles, lo_base = (tile_id % lo_tiles) << lo_log;
  const size_t base = (hi << (s0 + S)) + lo_base;
  // ---- load (element e of the tile: mid = e >> lo_log, lo_off = e & (LO-1)) ----
  for (int e = threadIdx.x; e < tile; e += NTT_THREADS) {
    size_t p = base + ((size_t)(e >> lo_log) << s0) + (e & (LO - 1));
    size_t src = s0 == 0 ? bitrev(p, m) : p;
    const uint4* q = reinterpret_cast<const uint4*>(in + src * 8);
    uint4 a = q[0], b = q[1];
    *reinterpret_cast<uint4*>(sm + e * 8) = a;
    *reinterpret_cast<uint4*>(sm + e * 8 + 4) = b;
  }
  __syncthreads();
  const size_t half = N >> 1;
  for (int q = 1; q <= S; q++) {
    const int t = threadIdx.x;
    if (t < tile / 2) {
      const int lo_off = t & (LO - 1);
      const int k = t >> lo_log;                                   // butterfly index along mid
      const int mid_lo = k & ((1 << (q - 1)) - 1);
      const int mid_a = ((k >> (q - 1)) << q) + mid_lo;
      const int ea = (mid_a << lo_log) + lo_off, eb = ea + ((1 << (q - 1)) << lo_log);
      const int s = s0 + q;                                        // global stage
      const size_t j = ((size_t)mid_lo << s0) + lo_base + lo_off;   // position inside the half block, < 2^(s-1)
      const size_t idx = j << (m - s);                             // exponent of w = gen: j * N / 2^s, < N/2
      Fe<F> w;
      if (!inverse) {
        w = ld_fe8<F>(table + idx * 8);
      } else if (idx == 0) {
        w = fe_one<F>();
      } else {
        w = fe_neg<F>(ld_fe8<F>(table + (half - idx) * 8));         // w^-idx = -w^(N/2 - idx)
      }
      Fe<F> a = ld_fe8<F>(sm + ea * 8);
      Fe<F> b = fe_mul<F>(ld_fe8<F>(sm + eb * 8), w);
      st_fe8<F>(sm + ea * 8, fe_add<F>(a, b));
      st_fe8<F>(sm + eb * 8, fe_sub<F>(a, b));
    }
    __syncthreads();
  }
  Fe<F> ninv;
  if (scale) ninv = ld_fe8<F>(table + half * 8);
  for (int e = threadIdx.x; e < tile; e += NTT_THREADS) {
    size_t p = base + ((size_t)(e >> lo_log) << s0) + (e & (LO - 1));
    if (scale) {
      st_fe8<F>(out + p * 8, fe_mul<F>(ld_fe8<F>(sm + e * 8), ninv));
    } else {
      *reinterpret_cast<uint4*>(out + p * 8) = *reinterpret_cast<const uint4*>(sm + e * 8);
      *reinterpret_cast<uint4*>(out + p * 8 + 4) = *reinterpret_cast<const uint4*>(sm + e * 8 + 4);
    }
  }
}

template <class F>
void ntt_build_table(cudaStream_t s, const uint32_t* d_gen, size_t half, int m, uint32_t* d_table) {
  size_t tthreads = ((half ? half : 1) + 31) / 32;
  k_ntt_table<F><<<(unsigned)((tthreads + 127) / 128), 128, 0, s>>>(d_gen, half, m, d_table);
}
template void ntt_build_table<Bn254Fr>(cudaStream_t, const uint32_t*, size_t, int, uint32_t*);
template void ntt_build_table<Bls12381Fr>(cudaStream_t, const uint32_t*, size_t, int, uint32_t*);

template <class F>
void ntt_device(cudaStream_t s, int m, const uint32_t* d_gen, const uint32_t* d_src, uint32_t* d_tmp, uint32_t* d_dst,
                uint32_t* d_table, int inverse) {
  const size_t N = (size_t)1 << m;
  const size_t half = N >> 1;
  ntt_build_table<F>(s, d_gen, half, m, d_table);
  if (m == 0) {
    cudaMemcpyAsync(d_dst, d_src, 32, cudaMemcpyDeviceToDevice, s);
    return;
  }
  const int tile_log = m < NTT_TILE_LOG ? m : NTT_TILE_LOG;
  const size_t tiles = N >> tile_log;
  // plan the passes so that the LAST one writes d_dst
  int npass = (m + NTT_TILE_LOG - 1) / NTT_TILE_LOG;
  const uint32_t* in = d_src;
  int s0 = 0;
  for (int pass = 0; pass < npass; pass++) {
    int S = m - s0 < NTT_TILE_LOG ? m - s0 : NTT_TILE_LOG;
    // a pass with s0 > 0 needs LO = tile/2^S <= 2^s0 low indices: always true (s0 >= 9 >= tile_log - S)
    bool last = pass == npass - 1;
    uint32_t* out = last ? d_dst : (((npass - 1 - pass) & 1) ? d_tmp : d_dst);
    if (out == in) out = (out == d_tmp) ? d_dst : d_tmp;
    k_ntt_pass<F><<<(unsigned)tiles, NTT_THREADS, 0, s>>>(in, out, d_table, m, s0, S, inverse, inverse && last);
    in = out;
    s0 += S;
  }
}

template void ntt_device<Bn254Fr>(cudaStream_t, int, const uint32_t*, const uint32_t*, uint32_t*, uint32_t*, uint32_t*, int);
template void ntt_device<Bls12381Fr>(cudaStream_t, int, const uint32_t*, const uint32_t*, uint32_t*, uint32_t*, uint32_t*, int);

}  // namespace zk
