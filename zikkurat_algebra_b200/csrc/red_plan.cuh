// Index logic of the low-latency bucket reduction (kernels_red.cuh, K5'): pure integer functions shared by the device
// kernels and by the host-side model in tests/host_emul (which runs the same three steps over integers mod a prime and
// compares with  sum_k (k+1) B_k ).
//
// The NB = 2^(c-1) buckets of a window form a 2^hr x 2^hc matrix, bucket k = k1 * 2^hc + k0:
//     sum_k (k+1) B_k = 2^hc * sum_k1 k1 Row_k1 + sum_k0 k0 Col_k0 + sum_k B_k
//   step 1 (k_red_rowcol)   Row_k1 and Col_k0: every sum is shared by `tpo` teams (interleaved entries), 32 teams per block
//   step 2 (k_red_bits)     T_j = sum of the columns (j < hc) resp. rows (hc <= j < hc + hr) whose index has bit j resp.
//                           j - hc set;  T_{hc+hr} = sum of all rows
//   step 3 (k_tail_group_bits)  sum_j 2^j T_j + T_{hc+hr}  by Horner over the bit positions, cut into NCH pieces
#pragma once
#include <stdint.h>

#include "hd.cuh"

namespace zk {

struct RedPlan {
  int hr, hc;                  // log2 of the number of rows / columns, hr + hc = c - 1, hc >= hr
  int tpo_r, tpo_c;            // teams per row sum / per column sum (powers of two, 1 .. 32)
  unsigned row_blocks, col_blocks;
};
ZK_HD int red_teams_per_output(uint32_t entries) {   // about 8 entries per team
  uint32_t t = entries / 8;
  return (int)(t < 1 ? 1 : (t > 32 ? 32 : t));
}
ZK_HD RedPlan red_plan(int c) {
  RedPlan p;
  p.hc = c / 2;
  p.hr = c - 1 - p.hc;
  p.tpo_r = red_teams_per_output(1u << p.hc);   // a row has 2^hc entries
  p.tpo_c = red_teams_per_output(1u << p.hr);
  p.row_blocks = (unsigned)((((uint64_t)1 << p.hr) * (uint64_t)p.tpo_r + 31) / 32);
  p.col_blocks = (unsigned)((((uint64_t)1 << p.hc) * (uint64_t)p.tpo_c + 31) / 32);
  return p;
}
// What team `tq` (0..31) of block `block` of step 1 does: which sum, which share of it.
struct RedTask {
  bool rows, valid;
  uint32_t out;        // row or column number
  int part, tpo;       // this team takes entries part, part + tpo, ...
  uint32_t entries;    // entries of the sum
  uint32_t first;      // bucket index of entry 0
  uint32_t stride;     // bucket index distance between entries
};
ZK_HD RedTask red_rowcol_task(const RedPlan& p, unsigned block, int tq) {
  RedTask t;
  const uint32_t NR = 1u << p.hr, NC = 1u << p.hc;
  t.rows = block < p.row_blocks;
  t.tpo = t.rows ? p.tpo_r : p.tpo_c;
  t.out = (t.rows ? block : block - p.row_blocks) * (uint32_t)(32 / t.tpo) + (uint32_t)(tq / t.tpo);
  t.part = tq % t.tpo;
  t.entries = t.rows ? NC : NR;
  t.valid = t.out < (t.rows ? NR : NC);        // a block may have more teams than there are sums (tiny windows)
  t.first = t.rows ? t.out * NC : t.out;
  t.stride = t.rows ? 1u : NC;
  return t;
}
// Step 2, sum j of a window: where its entries live in RC[rows | columns] and which ones they are.
struct RedBits {
  uint32_t base;       // offset into the window's RC record array (rows first, then columns)
  uint32_t entries;
  int bit;             // -1: all entries 0 .. entries-1; else the entries-th ... see red_bit_member
};
ZK_HD RedBits red_bits_task(const RedPlan& p, int j) {
  RedBits b;
  const uint32_t NR = 1u << p.hr, NC = 1u << p.hc;
  if (j < p.hc) { b.base = NR; b.entries = NC >> 1; b.bit = j; }
  else if (j < p.hc + p.hr) { b.base = 0; b.entries = NR >> 1; b.bit = j - p.hc; }
  else { b.base = 0; b.entries = NR; b.bit = -1; }
  return b;
}
// the e-th (ascending) index that has bit `bit` set
ZK_HD uint32_t red_bit_member(uint32_t e, int bit) {
  return bit < 0 ? e : (((e >> bit) << (bit + 1)) | (1u << bit) | (e & ((1u << bit) - 1u)));
}
// Step 3: piece q of NCH covers the bit positions [lo, hi) of nb = c - 1
ZK_HD void red_piece(int nb, int nch, int q, int& lo, int& hi) {
  lo = (nb * q) / nch;
  hi = (nb * (q + 1)) / nch;
}

}  // namespace zk
