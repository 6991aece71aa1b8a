// libzkmsm_b200.so -- thin C host layer + kernel launches for the B200 G1 MSM.
//
// Boundary: include/zk_msm_b200.h (the reference's 16 MSM symbols + extensions).  The host layer only
// moves bytes (H2D of scalars and points, D2H of one point per MSM) and launches kernels; every field
// and group operation of the path runs on the GPU.  There is no CPU fallback: CUDA errors abort.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/zk_msm_b200.h"
#include "msm_common.cuh"
#include "gfft.cuh"
#include "glv.cuh"
#include "ntt.cuh"
#include "selftest.cuh"
#include "sort.cuh"

using namespace zk;

#define CK(x)                                                                                        \
  do {                                                                                               \
    cudaError_t e__ = (x);                                                                           \
    if (e__ != cudaSuccess) {                                                                        \
      fprintf(stderr, "[zkmsm_b200] fatal: %s failed at %s:%d: %s\n", #x, __FILE__, __LINE__,        \
              cudaGetErrorString(e__));                                                              \
      abort();                                                                                       \
    }                                                                                                \
  } while (0)

namespace {

enum Buf {
  B_SCALARS, B_POINTS, B_KEYS0, B_KEYS1, B_VALS0, B_VALS1, B_CNT, B_ROWSUM, B_BUCKETS, B_HEADS, B_HEADKEYS, B_HEADS2, B_HEADKEYS2,
  B_U0, B_V0, B_U1, B_V1, B_OUT, B_NTT_TABLE, B_AFF_TMP, B_AFF_PRE, B_AFF_BINV, B_AFF_ST0, B_AFF_ST1, B_AFF_KEYS,
  B_AFF_VALS, B_PARTIALS, B_GLV_POINTS, B_RED_RC, B_RED_T, B_COUNT
};
constexpr int N_EV = 9;
constexpr int MAX_DEV = 16;

struct Stats {
  float ms[9] = {0};
  int c = 0, W = 0, slices = 1, aff_levels = 0;
  bool have_phases = false;
  bool srs_hit = false;        // the point array came from the resident-copy cache (no points crossed PCIe)
  bool slice_ran[8] = {false};
  long long insertions = 0;
};
constexpr int MAX_SLICES = 8;

struct DeviceCtx {
  std::atomic<bool> ready{false};
  int dev = 0;
  cudaStream_t s_main = nullptr, s_copy = nullptr;
  struct Pipeline {           // lanes and post-processing streams of the window-group pipeline (run_msm)
    cudaStream_t s_big[2] = {nullptr, nullptr}, s_chain[2] = {nullptr, nullptr};
    cudaStream_t big[2] = {nullptr, nullptr}, chain[2] = {nullptr, nullptr};   // the ones in use for the current call
    cudaStream_t post[8] = {nullptr};
    cudaStream_t stag_pool[4][4] = {{nullptr}};                        // staggered mode: [lane][priority level] streams
    cudaStream_t stag_big[4] = {nullptr}, stag_chain[4] = {nullptr};   // the ones in use: lanes of descending priority
    cudaEvent_t stag_a[4] = {nullptr}, stag_c[4] = {nullptr}, stag_lane[4] = {nullptr};
    cudaEvent_t ev_a[2] = {nullptr}, ev_c[2] = {nullptr}, ev_lane[2] = {nullptr}, ev_sorted = nullptr;
    cudaEvent_t ev_fix[8] = {nullptr}, ev_post[8] = {nullptr};
  } pl;
  cudaEvent_t ev[N_EV + 1] = {nullptr};
  cudaEvent_t gev[6 * 8] = {nullptr};  // per input slice: accumulate start/end, fix-up end, recode start, sort end, recode end
  cudaEvent_t ev_sc[8] = {nullptr}, ev_pt[8] = {nullptr};   // per input slice: scalars / points have arrived
  void* buf[B_COUNT] = {nullptr};
  size_t cap[B_COUNT] = {0};
  uint32_t* h_out = nullptr;  // pinned staging for results
  size_t h_out_cap = 0;
  // pinned staging ring for pageable host inputs (see host_to_device)
  static constexpr int STAGE_SLOTS = 16;                 // two ring slots per staging thread (at most 8 threads)
  static constexpr size_t STAGE_BYTES = (size_t)2 << 20;
  uint8_t* stage = nullptr;
  cudaEvent_t stage_ev[STAGE_SLOTS] = {nullptr};
  bool stage_used[STAGE_SLOTS] = {false};
  std::mutex mu;              // one MSM at a time per device (workspaces are shared)
  Stats stats;
  int resident[4] = {0, 0, 0, 0};   // resident k_accumulate threads of this device, per curve id (filled under mu)
  float last_op_ms = 0.f;           // device time of the kernels of the last conversion / NTT / group FFT call (no copies)
  struct StagePool* pool = nullptr; // persistent staging threads for pageable sources (created on first use)
  // Device-resident copies of point arrays the caller keeps passing (the SRS of a KZG prover), see srs_lookup
  struct SrsEntry { const void* host; size_t bytes; uint64_t fp; int curve; void* dev; uint64_t last_use; };
  std::vector<SrsEntry> srs;
  size_t srs_bytes = 0;
  uint64_t srs_clock = 0;

  void srs_drop_all() {
    for (auto& e : srs) CK(cudaFree(e.dev));
    srs.clear();
    srs_bytes = 0;
  }
  void* ensure(int which, size_t bytes) {
    if (bytes > cap[which]) {
      if (buf[which]) CK(cudaFree(buf[which]));
      buf[which] = nullptr;
      cap[which] = 0;
      size_t want = bytes + bytes / 8 + 256;
      cudaError_t e = cudaMalloc(&buf[which], want);
      if (e == cudaErrorMemoryAllocation) {
        // out of device memory: give back what is only a cache (resident point arrays), then ask for the exact size
        (void)cudaGetLastError();
        srs_drop_all();
        want = bytes + 256;
        e = cudaMalloc(&buf[which], want);
      }
      if (e != cudaSuccess) {
        size_t fr = 0, tot = 0;
        (void)cudaMemGetInfo(&fr, &tot);
        fprintf(stderr, "[zkmsm_b200] fatal: device %d cannot provide a %zu-byte workspace (%zu of %zu bytes free): %s\n", dev, want,
                fr, tot, cudaGetErrorString(e));
        abort();
      }
      cap[which] = want;
    }
    return buf[which];
  }
  void release_workspaces() {   // under mu
    for (int i = 0; i < B_COUNT; i++) {
      if (buf[i]) CK(cudaFree(buf[i]));
      buf[i] = nullptr;
      cap[i] = 0;
    }
    srs_drop_all();
    if (h_out) { CK(cudaFreeHost(h_out)); h_out = nullptr; h_out_cap = 0; }
  }
  uint32_t* ensure_host(size_t bytes) {
    if (bytes > h_out_cap) {
      if (h_out) CK(cudaFreeHost(h_out));
      CK(cudaMallocHost((void**)&h_out, bytes));
      h_out_cap = bytes;
    }
    return h_out;
  }
};

DeviceCtx g_ctx[MAX_DEV];
std::mutex g_init_mu;
std::atomic<int> g_device{-1};
std::atomic<long long> g_launches{0};  // kernels launched by this library since load

std::vector<int> device_list();

// GLV split on/off: zkb200_set_glv, else $ZKB200_GLV, else on
std::atomic<int> g_glv{-1};
bool glv_enabled() {
  int v = g_glv.load();
  if (v < 0) {
    const char* e = getenv("ZKB200_GLV");
    v = e ? (atoi(e) != 0) : 1;
    g_glv.store(v);
  }
  return v != 0;
}

// The ONE place that decides which GPU a call without an explicit device runs on: zkb200_set_device, else a
// one-element device list (zkb200_set_devices / $ZKB200_DEVICES="2"), else $ZKB200_DEVICE, else 0.  Uploads
// (zkb200_device_upload), statistics and MSM calls therefore always agree.
int current_device_choice() {
  int d = g_device.load();
  if (d >= 0) return d;
  std::vector<int> devs = device_list();
  if (devs.size() == 1) return devs[0];
  const char* e = getenv("ZKB200_DEVICE");
  return e ? atoi(e) : 0;
}

DeviceCtx& get_ctx(int d = -1) {
  if (d < 0) d = current_device_choice();
  if (d < 0 || d >= MAX_DEV) { fprintf(stderr, "[zkmsm_b200] fatal: bad device index %d\n", d); abort(); }
  DeviceCtx& cx = g_ctx[d];
  if (!cx.ready) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (!cx.ready) {
      int count = 0;
      cudaError_t e = cudaGetDeviceCount(&count);
      if (e != cudaSuccess || count == 0) {
        fprintf(stderr, "[zkmsm_b200] fatal: no CUDA device available (%s); this library has no CPU path\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        abort();
      }
      if (d >= count) { fprintf(stderr, "[zkmsm_b200] fatal: device %d requested, %d visible\n", d, count); abort(); }
      int prev_dev = -1;
      if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
      CK(cudaSetDevice(d));
      cx.dev = d;
      CK(cudaStreamCreateWithFlags(&cx.s_main, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&cx.s_copy, cudaStreamNonBlocking));
      {
        int lo_pri = 0, hi_pri = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        for (int g = 0; g < 2; g++) {
          CK(cudaStreamCreateWithPriority(&cx.pl.s_big[g], cudaStreamNonBlocking, lo_pri));
          CK(cudaStreamCreateWithPriority(&cx.pl.s_chain[g], cudaStreamNonBlocking, hi_pri));
          CK(cudaEventCreateWithFlags(&cx.pl.ev_a[g], cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&cx.pl.ev_c[g], cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&cx.pl.ev_lane[g], cudaEventDisableTiming));
        }
        for (int g = 0; g < 8; g++) {
          CK(cudaStreamCreateWithPriority(&cx.pl.post[g], cudaStreamNonBlocking, hi_pri));
          CK(cudaEventCreate(&cx.pl.ev_fix[g]));    // timing enabled: the $ZKB200_TRACE timeline reads them
          CK(cudaEventCreate(&cx.pl.ev_post[g]));
        }
        CK(cudaEventCreateWithFlags(&cx.pl.ev_sorted, cudaEventDisableTiming));
        for (int g = 0; g < 4; g++) {
          // lane 0 (top windows) gets the highest of the "big" priorities; chains and post-processing stay above all
          for (int lv = 0; lv < 4; lv++) {   // level 0 = lowest priority
            int pri = lo_pri - lv;
            if (pri < hi_pri + 1) pri = hi_pri + 1 < lo_pri ? hi_pri + 1 : lo_pri;
            CK(cudaStreamCreateWithPriority(&cx.pl.stag_pool[g][lv], cudaStreamNonBlocking, pri));
          }
          cx.pl.stag_big[g] = cx.pl.stag_pool[g][3 - g];
          CK(cudaStreamCreateWithPriority(&cx.pl.stag_chain[g], cudaStreamNonBlocking, hi_pri));
          CK(cudaEventCreateWithFlags(&cx.pl.stag_a[g], cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&cx.pl.stag_c[g], cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&cx.pl.stag_lane[g], cudaEventDisableTiming));
        }
      }
      for (int i = 0; i < 6 * 8; i++) CK(cudaEventCreate(&cx.gev[i]));
      for (int i = 0; i <= N_EV; i++) CK(cudaEventCreate(&cx.ev[i]));
      for (int i = 0; i < 8; i++) {
        CK(cudaEventCreateWithFlags(&cx.ev_sc[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&cx.ev_pt[i], cudaEventDisableTiming));
      }
      if (prev_dev >= 0) cudaSetDevice(prev_dev);
      cx.ready = true;
    }
  }
  return cx;
}

// Host -> device copy that is fast for ORDINARY (pageable) memory too.  The reference's callers hand over
// GHC-heap buffers (mallocForeignPtrBytes, lib/src/ZK/Algebra/Class/Flat.hs:186-194), which the driver copies
// through its own bounce buffer at ~10 GB/s on one thread.  Here 4 host threads copy 4 MiB chunks into a
// ring of pinned buffers and each chunk is sent with an async DMA as soon as it is staged, so the host-side
// copy runs at several threads' memory bandwidth and overlaps with the PCIe transfer.  Pinned (registered)
// source memory skips all that and is sent directly.  Returns when every chunk has been queued on `stream`.
bool host_is_pinned(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
  (void)cudaGetLastError();
  return false;
}

// Copy into the pinned ring with streaming (non-temporal) stores: the destination is written once and read by the
// DMA engine only, so it should neither be read for ownership nor displace the caller's data from the caches --
// with 8 processes staging at once the host's memory bandwidth is the limit (profiles/r2_notes.md).  dst is 64-byte
// aligned (ring slots are multiples of 2 MiB inside a page-aligned allocation); src has the caller's alignment.
#if defined(__x86_64__) && defined(__SSE2__)
#include <emmintrin.h>
inline void stage_copy(uint8_t* dst, const uint8_t* src, size_t n, bool streaming) {
  if (!streaming) { memcpy(dst, src, n); return; }
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
    const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
    _mm_stream_si128((__m128i*)(dst + i), a);
    _mm_stream_si128((__m128i*)(dst + i + 16), b);
    _mm_stream_si128((__m128i*)(dst + i + 32), c);
    _mm_stream_si128((__m128i*)(dst + i + 48), d);
  }
  if (i < n) memcpy(dst + i, src + i, n - i);
  _mm_sfence();      // the stores are globally visible before the DMA is queued
}
#else
inline void stage_copy(uint8_t* dst, const uint8_t* src, size_t n, bool) { memcpy(dst, src, n); }
#endif

// Persistent staging threads (created on the first pageable copy of a device, never joined: they sleep on a condition
// variable between copies).  A job = one host_to_device call; worker w stages chunks w, w+T, ... into its ring slots.
struct StagePool {
  // Threads: the host cores this process can count on -- online cores / visible GPUs (one process per GPU is the usual
  // multi-GPU arrangement, and 8 x 4 staging threads on a 32-core box fight each other) -- between 2 and 8
  // ($ZKB200_STAGE_THREADS overrides).  A slot is always reused by the same thread: chunk i uses slot i % (2T) and
  // belongs to thread i % T.
  int T = 4;        // threads of the pool
  int Tjob = 4;     // threads working on the current job (small copies use at most 4, see copy())
  bool forced = false;
  size_t chunk = DeviceCtx::STAGE_BYTES;   // bytes per ring slot in use ($ZKB200_STAGE_CHUNK_KB, at most the slot size)
  bool streaming = true;                   // non-temporal stores into the ring ($ZKB200_STAGE_NT=0: ordinary stores)
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  uint64_t gen = 0;
  int pending = 0;
  DeviceCtx* cx = nullptr;
  uint8_t* dst = nullptr;
  const uint8_t* src = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;

  explicit StagePool(DeviceCtx* c) : cx(c) {
    const char* e = getenv("ZKB200_STAGE_THREADS");
    if (e && atoi(e) > 0) { T = atoi(e); forced = true; }
    else {
      int gpus = 1;
      if (cudaGetDeviceCount(&gpus) != cudaSuccess || gpus < 1) gpus = 1;
      const unsigned hc = std::thread::hardware_concurrency();
      T = (int)(hc ? hc : 8) / gpus;
      // measured (B200 box, 16 cores): 32 MB of scalars per call: 2 threads 8.0 ms, 4: 7.7, 8: 8.3; 128 MB (scalars and
      // points of BLS12-381 2^20, no resident copy): 2: 11.1 ms, 4: 9.5, 8: 9.1, 12: 9.1; 384 MB (BN254 2^22): 4: 26.4, 8: 17.9
      if (T > 8) T = 8;
    }
    if (T < 2) T = 2;
    if (T > DeviceCtx::STAGE_SLOTS / 2) T = DeviceCtx::STAGE_SLOTS / 2;
    if (const char* q = getenv("ZKB200_STAGE_CHUNK_KB")) {
      const size_t kb = (size_t)atoll(q);
      if (kb >= 16 && (kb << 10) <= DeviceCtx::STAGE_BYTES) chunk = (kb << 10) & ~(size_t)63;
    }
    if (const char* q = getenv("ZKB200_STAGE_NT")) streaming = atoi(q) != 0;
    for (int w = 0; w < T; w++) std::thread([this, w] { worker(w); }).detach();
  }
  void worker(int w) {
    if (cudaSetDevice(cx->dev) != cudaSuccess) abort();
    uint64_t seen = 0;
    for (;;) {
      int TJ;
      {
        std::unique_lock<std::mutex> lk(m);
        cv_work.wait(lk, [&] { return gen != seen; });
        seen = gen;
        TJ = Tjob;                                // the job this generation belongs to (a thread that is not part of a
      }                                           // job is not waited for, so it must not look at the job any later)
      if (w >= TJ) continue;
      const int R = 2 * TJ;
      const size_t CH = chunk;
      const size_t nchunks = (bytes + CH - 1) / CH;
      for (size_t i = w; i < nchunks; i += TJ) {
        const int slot = (int)(i % R);
        if (cx->stage_used[slot]) CK(cudaEventSynchronize(cx->stage_ev[slot]));   // previous DMA out of this slot finished
        const size_t off = i * CH, len = bytes - off < CH ? bytes - off : CH;
        uint8_t* ring = cx->stage + (size_t)slot * DeviceCtx::STAGE_BYTES;
        stage_copy(ring, src + off, len, streaming);
        CK(cudaMemcpyAsync(dst + off, ring, len, cudaMemcpyHostToDevice, stream));
        CK(cudaEventRecord(cx->stage_ev[slot], stream));
        cx->stage_used[slot] = true;
      }
      {
        std::lock_guard<std::mutex> lk(m);
        if (--pending == 0) cv_done.notify_one();
      }
    }
  }
  void copy(uint8_t* d, const uint8_t* s, size_t n, cudaStream_t st) {   // returns when every chunk has been queued
    std::unique_lock<std::mutex> lk(m);
    dst = d; src = s; bytes = n; stream = st;
    Tjob = (forced || n >= ((size_t)48 << 20) || T < 4) ? T : 4;   // more than 4 threads only pay for big copies
    pending = Tjob;
    gen++;
    cv_work.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};

void host_to_device(DeviceCtx& cx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return;
  const bool pinned = host_is_pinned(src);
  static const bool no_staging = getenv("ZKB200_NO_STAGING") != nullptr;
  if (pinned || bytes < ((size_t)2 << 20) || no_staging) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return;
  }
  if (!cx.stage) {
    CK(cudaMallocHost((void**)&cx.stage, DeviceCtx::STAGE_SLOTS * DeviceCtx::STAGE_BYTES));
    for (int i = 0; i < DeviceCtx::STAGE_SLOTS; i++) CK(cudaEventCreateWithFlags(&cx.stage_ev[i], cudaEventDisableTiming));
  }
  if (!cx.pool) cx.pool = new StagePool(&cx);
  cx.pool->copy((uint8_t*)dst, (const uint8_t*)src, bytes, stream);
}

// ---- resident point arrays ("SRS cache") --------------------------------------------------------------------------------
// The reference's callers commit again and again over ONE point array (examples/KZG.hs:77-88,110-116: every commitPoly /
// openingProof is `msm coeffs tauG1s`), and points are 2/3..3/4 of the input bytes.  A host point array is therefore kept
// on the device after its first use, keyed by (host pointer, byte count, curve) and guarded by a 64-bit fingerprint of
// a strided sample of its words (first and last 256 bytes in full): a later call with the same key and fingerprint uses
// the resident copy and moves only the scalars.  A different fingerprint under the same key (the allocator handed the
// address to another array) replaces the entry.  Least recently used entries go when the budget
// ($ZKB200_SRS_CACHE_MB, default 16384; 0 disables the cache) is exceeded or when a workspace allocation fails.
// Contract: the fingerprint samples ~1500 words, so a caller that rewrites a FEW points of an array in place between
// calls must disable the cache or use a fresh buffer (the reference's FlatArrays are immutable: Class/Flat.hs:81-90).
size_t srs_budget_bytes() {
  static const size_t b = [] {
    const char* e = getenv("ZKB200_SRS_CACHE_MB");
    return (size_t)(e ? atoll(e) : 16384) << 20;
  }();
  return b;
}
uint64_t srs_fingerprint(const void* p, size_t bytes) {
  const uint64_t* w = (const uint64_t*)p;
  const size_t nw = bytes / 8;
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)bytes;
  auto mix = [&](uint64_t v) { h = (h ^ v) * 0x100000001B3ull; h ^= h >> 29; };
  const size_t edge = nw < 32 ? nw : 32;
  for (size_t i = 0; i < edge; i++) { mix(w[i]); mix(w[nw - 1 - i]); }
  const size_t S = 1024;
  if (nw > 64)
    for (size_t i = 0; i < S; i++) mix(w[(size_t)(((unsigned __int128)i * nw) / S)]);
  return h;
}
// under cx.mu.  Returns the resident copy or nullptr; *fp_out receives the fingerprint for srs_insert.
const uint32_t* srs_lookup(DeviceCtx& cx, int curve, const void* host, size_t bytes, uint64_t* fp_out) {
  if (srs_budget_bytes() == 0 || bytes < ((size_t)64 << 10) || bytes > srs_budget_bytes()) return nullptr;
  const uint64_t fp = srs_fingerprint(host, bytes);
  *fp_out = fp;
  for (size_t i = 0; i < cx.srs.size(); i++) {
    auto& e = cx.srs[i];
    if (e.host == host && e.bytes == bytes && e.curve == curve) {
      if (e.fp == fp) { e.last_use = ++cx.srs_clock; return (const uint32_t*)e.dev; }
      CK(cudaFree(e.dev));                       // same address, other content: stale
      cx.srs_bytes -= e.bytes;
      cx.srs.erase(cx.srs.begin() + i);
      return nullptr;
    }
  }
  return nullptr;
}
// under cx.mu, after a call that uploaded `bytes` of points to d_src (device): keep a copy.  Failure to allocate is
// not an error (the cache is an optimisation).
void srs_insert(DeviceCtx& cx, int curve, const void* host, size_t bytes, uint64_t fp, const void* d_src, cudaStream_t s) {
  if (srs_budget_bytes() == 0 || bytes < ((size_t)64 << 10) || bytes > srs_budget_bytes()) return;
  while (cx.srs_bytes + bytes > srs_budget_bytes() && !cx.srs.empty()) {
    size_t lru = 0;
    for (size_t i = 1; i < cx.srs.size(); i++) if (cx.srs[i].last_use < cx.srs[lru].last_use) lru = i;
    CK(cudaFree(cx.srs[lru].dev));
    cx.srs_bytes -= cx.srs[lru].bytes;
    cx.srs.erase(cx.srs.begin() + lru);
  }
  void* d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) { (void)cudaGetLastError(); return; }
  CK(cudaMemcpyAsync(d, d_src, bytes, cudaMemcpyDeviceToDevice, s));
  CK(cudaStreamSynchronize(s));
  cx.srs.push_back({host, bytes, fp, curve, d, ++cx.srs_clock});
  cx.srs_bytes += bytes;
}

struct DeviceGuard {  // run on our device, then give the caller its own current device back
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    CK(cudaSetDevice(dev));
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <class C> struct CurveId;
template <> struct CurveId<Bn254> { static constexpr int value = ZKB200_BN128; };
template <> struct CurveId<Bls12381> { static constexpr int value = ZKB200_BLS12_381; };
template <> struct CurveId<Bn254G2> { static constexpr int value = ZKB200_BN128_G2; };
template <> struct CurveId<Bls12381G2> { static constexpr int value = ZKB200_BLS12_381_G2; };

// curves whose affine pre-reduction kernels are built (kernels_aff.cuh; <curve>_aff.cu)
template <class C> struct HasAffineTree { static constexpr bool value = false; };
template <> struct HasAffineTree<Bn254> { static constexpr bool value = true; };
template <> struct HasAffineTree<Bls12381> { static constexpr bool value = true; };
template <> struct HasAffineTree<Bn254G2> { static constexpr bool value = true; };
template <> struct HasAffineTree<Bls12381G2> { static constexpr bool value = true; };

int ilog2_floor(size_t x) { int r = 0; while (x > 1) { x >>= 1; r++; } return r; }

// Window width c for signed digits: minimise the Fp-multiplication count of the two throughput phases,
//   W(c) * ( n * 10  +  2^(c-1) * 2 * 14 ),   W(c) = ceil((nbits+1)/c)
// (n*W bucket insertions at 10 mul, 2 additions of 14 mul per bucket in the reduction).  Widths whose
// top window would be nearly empty lose automatically because they pay a whole extra window.
// Override: $ZKB200_WINDOW.
int pick_window(size_t n, int nmsm, int nbits) {
  const char* e = getenv("ZKB200_WINDOW");
  if (e && atoi(e) > 0) return atoi(e);
  (void)nmsm;
  int best = 1;
  double best_cost = 1e300;
  for (int c = 1; c <= 24; c++) {
    double W = (double)signed_windows(nbits, c);
    double cost = W * ((double)n * 10.0 + (double)(1ull << (c - 1)) * 28.0);
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

template <class C>
void run_msm(DeviceCtx& cx, int nmsm, size_t n, const uint64_t* scalars, int sloc, const uint64_t* points, int ploc,
             int nl, int mont, int out_mode, int window, uint64_t* out, int out_loc = ZKB200_HOST) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  using Mem = XyzzMem<P>;
  if (nl < 1 || nl > 4) { fprintf(stderr, "[zkmsm_b200] fatal: expo_nlimbs = %d unsupported (1..4)\n", nl); abort(); }
  if (mont && nl != 4) { fprintf(stderr, "[zkmsm_b200] fatal: Montgomery scalars need expo_nlimbs = 4\n"); abort(); }
  if (nmsm <= 0) return;
  const int out_coords = out_mode == OUT_AFFINE ? 2 : (out_mode == OUT_XYZZ ? 4 : 3);
  cudaStream_t s = cx.s_main;
  Stats st;
  uint32_t* d_out = (uint32_t*)cx.ensure(B_OUT, (size_t)nmsm * 4 * L * 4);
  uint32_t* h_out = cx.ensure_host((size_t)nmsm * 4 * L * 4);

  int nbits = mont ? C::Fr::BITS : 64 * nl;
  // GLV split (glv.cuh): 2n points (P_i, phi(P_i)) with 127-bit scalars -> half the windows.  Only for scalars that are
  // longer than the split halves in the first place ($ZKB200_GLV=0 switches it off).
  const bool glv = GlvOf<C>::available && glv_enabled() && nbits > 160 && n > 0;
  const size_t F = glv ? 2 : 1;        // pairs per point and window
  if (glv) nbits = 127;
  int c = 0, W = 0, K = 1;
  int trace_groups = 0, trace_g0[9] = {0};
  const auto host_t0 = std::chrono::steady_clock::now();
  double host_ms[4] = {0, 0, 0, 0};   // $ZKB200_TRACE: host time until the work arrays exist / all launches queued / the stream is done
  auto host_now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
  CK(cudaEventRecord(cx.ev[0], s));
  if (n == 0) {
    g_launches++;
    launch_tail<C>(s, nullptr, nmsm, 0, 0, out_mode, d_out);
    CK(cudaGetLastError());
  } else {
    c = window > 0 ? window : pick_window(n * F, nmsm, nbits);
    if (c < 1) c = 1;
    if (c > 24) c = 24;
    W = signed_windows(nbits, c);
    if ((long long)nmsm * W > 65535) {
      // segments are a grid dimension (<= 65535): very large batches are processed in sub-batches
      int sub = 65535 / W;
      const int out_c = out_mode == OUT_AFFINE ? 2 : (out_mode == OUT_XYZZ ? 4 : 3);
      for (int m0 = 0; m0 < nmsm; m0 += sub) {
        int cntm = nmsm - m0 < sub ? nmsm - m0 : sub;
        run_msm<C>(cx, cntm, n, scalars + (size_t)m0 * n * nl, sloc, points, ploc, nl, mont, out_mode, c,
                   out + (size_t)m0 * out_c * (L / 2), out_loc);
      }
      return;
    }
    const uint32_t NB = 1u << (c - 1);
    if (nmsm > 1) {
      // Batches are bounded by memory as well as by the grid limit: pairs (ping-pong keys + values) cost 16 bytes per
      // insertion, buckets and chunk heads about two XYZZ records per bucket.  A batch that would not fit in what is
      // free (plus the workspaces this device already holds) is processed in halves instead of aborting in cudaMalloc.
      size_t fr = 0, tot = 0, held = 0;
      const double need = (double)nmsm * W * ((double)(n * F) * 24.0 + (double)NB * 2.5 * sizeof(Mem)) + (double)nmsm * n * nl * 8.0;
      // (the driver query costs 0.2 .. 3 ms: only asked when the batch is big enough to matter)
      if (need > 8e9) {
        CK(cudaMemGetInfo(&fr, &tot));
        for (int i = 0; i < B_COUNT; i++) held += cx.cap[i];
      }
      if (need > 8e9 && need > 0.8 * (double)(fr + held)) {
        const int half = nmsm / 2;
        const int out_c = out_mode == OUT_AFFINE ? 2 : (out_mode == OUT_XYZZ ? 4 : 3);
        run_msm<C>(cx, half, n, scalars, sloc, points, ploc, nl, mont, out_mode, c, out, out_loc);
        run_msm<C>(cx, nmsm - half, n, scalars + (size_t)half * n * nl, sloc, points, ploc, nl, mont, out_mode, c,
                   out + (size_t)half * out_c * (L / 2), out_loc);
        return;
      }
    }
    const int nseg = nmsm * W;
    st.c = c; st.W = W; st.insertions = (long long)nseg * (long long)(n * F);

    // ---- point slices -----------------------------------------------------------------------------------
    // With host buffers the input vectors are cut into K contiguous slices that go through
    // H2D -> recode -> sort -> accumulate one after the other, each into its OWN bucket array (the K arrays
    // are summed on the fly by the first reduction level), so that the PCIe transfer of slice k+1 runs under
    // the accumulation of slice k instead of in front of everything.  Costs one extra bucket addition per
    // bucket and slice in k_reduce_first.  (Adding into shared buckets inside k_accumulate was tried and
    // loses: the rare per-run addition diverges and is paid by the whole warp on almost every step.)
    // (only when the points travel too: with resident points -- SRS cache -- the 32 bytes per scalar are not worth the
    //  smaller kernels of two slices: measured 8.5 -> 9.4 ms pageable, 7.7 -> 8.5 ms pinned at BLS12-381 2^20)
    if (nmsm == 1 && sloc == ZKB200_HOST && ploc == ZKB200_HOST) {
      const char* e = getenv("ZKB200_SLICES");
      K = e ? atoi(e) : (n >= ((size_t)1 << 19) ? 2 : 1);   // measured: 2 slices -4 % (2^20) .. -14 % (2^24) end to end
      if (K > MAX_SLICES) K = MAX_SLICES;
      if (K < 1) K = 1;
      if ((size_t)K > n) K = 1;
    }
    st.slices = K;
    size_t lo[MAX_SLICES + 1];
    size_t nmax = 0;
    for (int k = 0; k <= K; k++) lo[k] = (n * (size_t)k / K) & ~(size_t)3;   // multiples of 4 keep 16-byte alignment
    if (K == 2) {
      // the first slice is what nothing can hide: make it the smaller one, as long as its accumulation still
      // covers the transfer of the rest ($ZKB200_SLICE0 = percent of the points in slice 0)
      const char* e = getenv("ZKB200_SLICE0");
      // accumulate time of slice 0 must cover the PCIe time of slice 1: f >= t/(a+t) with a = accumulate and
      // t = transfer time per point (measured: BN254 ~1.9/1.9 us per 1000 points, BLS12-381 5.4/2.6, G2 heavier)
      // Ordinary (pageable) memory goes through the staging ring at about half the PCIe rate, so t doubles.
      const bool pinned = host_is_pinned(points);
      int pct = e ? atoi(e) : (pinned ? (L <= 8 ? 50 : 25) : (L <= 12 ? 50 : 35));   // measured (profiles/r1_notes.md)
      if (pct < 5) pct = 5;
      if (pct > 95) pct = 95;
      lo[1] = (n * (size_t)pct / 100) & ~(size_t)3;
    }
    lo[K] = n;
    for (int k = 0; k < K; k++) if (lo[k + 1] - lo[k] > nmax) nmax = lo[k + 1] - lo[k];
    const size_t pmax = nmax * F;        // sorted pairs per segment of the longest slice

    // ---- inputs -> device: copy stream, queued slice by slice just before the work that needs them ----
    const uint64_t* d_scalars = scalars;
    const uint32_t* d_points = (const uint32_t*)points;
    if (sloc == ZKB200_HOST) d_scalars = (const uint64_t*)cx.ensure(B_SCALARS, (size_t)nmsm * n * nl * 8);
    if (ploc == ZKB200_HOST) d_points = (const uint32_t*)cx.ensure(B_POINTS, n * (size_t)(2 * L) * 4);
    // GLV: slice k's points are expanded to [P ; phi(P)] (2 nk records) at record offset 2 lo[k] of this array
    constexpr int OS = own_stride<P>();           // record stride (words) of the library's own point arrays
    const int pstride = glv ? OS : 2 * L;         // stride of the array the pairs' indices refer to
    uint32_t* glv_points = glv ? (uint32_t*)cx.ensure(B_GLV_POINTS, 2 * n * (size_t)OS * 4) : nullptr;

    // ---- work arrays (sized for the longest slice) ----
    const size_t pairs_max = (size_t)nseg * pmax;
    // packed (key, value) pairs, ping-pong between the sorting passes; the last pass writes the separate sorted arrays
    uint2* packed[2] = {(uint2*)cx.ensure(B_KEYS0, pairs_max * 8), (uint2*)cx.ensure(B_KEYS1, pairs_max * 8)};
    uint32_t* keys_sorted = (uint32_t*)cx.ensure(B_VALS0, pairs_max * 4);
    uint32_t* vals_sorted = (uint32_t*)cx.ensure(B_VALS1, pairs_max * 4);
    const int tiles_max = (int)((pmax + SORT_TILE - 1) / SORT_TILE);
    uint32_t* cnt = (uint32_t*)cx.ensure(B_CNT, sort_cnt_bytes(nseg, tiles_max));
    uint32_t* rowsum = (uint32_t*)cx.ensure(B_ROWSUM, (size_t)nseg * SORT_RADIX * 4);
    const size_t slice_stride = (size_t)nseg * NB;   // buckets per slice
    Mem* buckets = (Mem*)cx.ensure(B_BUCKETS, (size_t)K * slice_stride * sizeof(Mem));
    CK(cudaMemsetAsync(buckets, 0, (size_t)K * slice_stride * sizeof(Mem), s));  // ZZ = 0: every bucket starts at infinity
    int& resident = cx.resident[CurveId<C>::value];   // per device and curve (cx.mu is held)
    if (resident == 0) resident = accumulate_resident_threads<C>();
    // Sorted pairs per thread: every accumulate grid is a whole number of waves of resident threads so that
    // all SMs drain together; about 48 insertions per thread for small problems (several waves), up to 192
    // for big ones (fewer chunk heads to fold afterwards, still >= 12 waves).
    // ---- affine pre-reduction (kernels_aff.cuh): R levels of pairwise affine sums in front of k_accumulate ----
    int R = 0;
    if constexpr (HasAffineTree<C>::value) {
      // Levels: the tree pays while runs of equal keys are long (avg run = n / NB) and the problem is big enough
      // to hide the per-level inversion chains; measured on B200 (profiles/r1_notes.md).  Override: $ZKB200_AFFINE.
      const char* e = getenv("ZKB200_AFFINE");
      if (e) {
        R = atoi(e);
      } else if (L > 8 ? nmax >= ((size_t)1 << 19) : (nmax >= ((size_t)1 << 20) && nmax < ((size_t)1 << 22))) {   // (points, not pairs)
        // 12 limbs: -7 % (2^19), -15 % (2^20), -22 % (2^21), -17 % (2^22, 2^24) of the whole MSM; 8 limbs: -7 % at 2^20,
        // -2 % at 2^21, nothing at 2^22 and a loss at 2^24 (the cheaper multiplication leaves the tree's extra
        // memory traffic exposed)
        R = ilog2_floor(pmax / NB + 1) - 2;
        if (pairs_max >= ((size_t)1 << 26) || L >= 16) R++;   // G2: heavier additions, one more level pays
        if (R > 5) R = 5;
      }
      if (R < 0) R = 0;
      if (R > 12) R = 12;
      while (R > 0 && (pmax >> R) < 2) R--;
      if (pairs_max >= ((size_t)1 << 30) || n * F >= ((size_t)1 << 30)) R = 0;   // 30-bit refs
      if (R > 0) {   // workspace guard: temporary points + running products
        AffSizes z = aff_sizes(pmax, nseg, R);
        if ((z.tmp_points * OS + (z.pre_elems + z.pre2_elems) * L) * (size_t)4 > ((size_t)48 << 30)) R = 0;
      }
    }
    st.aff_levels = R;
    // ---- window groups and lanes (see the slice loop) ----
    // split_tail: a single MSM's windows are cut into NG groups, top windows first; a group's bucket reduction and
    // its share of the window combination start as soon as its buckets are complete, under the accumulation of the
    // lower groups.  Two lanes (streams) work on two groups at a time so that the latency-bound kernels of one lane
    // (inversion chains of the affine tree) run under the other lane's additions.
    int NG = 1;
    static const int min_split_w = [] { const char* e = getenv("ZKB200_MIN_SPLIT_W"); return e ? atoi(e) : 8; }();
    if (nmsm == 1 && W >= min_split_w) {
      const char* e = getenv("ZKB200_WGROUPS");
      NG = e ? atoi(e) : 2;   // measured: 2 groups -1.3 % (BLS12-381 2^20) .. -2 % (2^22); 4 and 8 lose (smaller kernels)
      if (NG < 1) NG = 1;
      if (NG > 8) NG = 8;
    }
    // Staggered mode (needs the tree's short-lived blocks): every group gets its own lane, upper windows at higher
    // stream priority, so the upper groups finish first -- their reduction and their long doubling chains then run
    // under the lower groups' work -- while each lane's inversion chains hide behind the lower-priority lanes.
    // Groups from the bottom: [0, 3W/16), [3W/16, 9W/16), [9W/16, W)  ($ZKB200_STAGGER = number of groups 2..4,
    // 0 = off; $ZKB200_STAGGER_SPLIT = the inner boundaries in windows).
    int stagger = 0;
    if (nmsm == 1 && W >= min_split_w && R > 0) {
      const char* e = getenv("ZKB200_STAGGER");
      stagger = e ? atoi(e) : 3;
      if (stagger < 2) stagger = 0;
      if (stagger > 4) stagger = 4;
      if (stagger) NG = stagger;
    }
    const bool split_tail = NG > 1;
    static const bool reduce_team = [] { const char* e = getenv("ZKB200_REDUCE_TEAM"); return e ? atoi(e) != 0 : true; }();
    int nlanes = 1;
    {
      const char* e = getenv("ZKB200_AFF_GROUPS");
      if ((e ? atoi(e) : 2) == 2 && nseg >= 2 && (R > 0 || split_tail)) nlanes = 2;
    }
    if (stagger) {
      nlanes = 1;
      // priority level of every lane (lane 0 = top group), one digit each; default: the upper two groups above the
      // bottom one, whose kernels then fill the holes the others' inversion chains leave and which finishes last -- it has
      // the shortest post-processing chain (no trailing doublings).  Measured at BLS12-381 2^20, three runs each on one
      // box: median call 6.81-6.84 ms with "1000", 6.61-6.65 with "1100", 6.70 with "2100"; best calls equal
      // (6.5-6.6): "1000" lets the middle group end up alone on the GPU in half of the calls (profiles/r2_notes.md).
      const char* e = getenv("ZKB200_STAGGER_PRI");
      const char* dflt = "1100";
      for (int j = 0; j < 4; j++) {
        int lv = (e && strlen(e) > (size_t)j ? e[j] : dflt[j]) - '0';
        if (lv < 0) lv = 0;
        if (lv > 3) lv = 3;
        cx.pl.stag_big[j] = cx.pl.stag_pool[j][lv];
      }
    }
    if (!split_tail) NG = nlanes;   // without the split the lanes simply halve the segments
    int grp0[9];                    // group g covers segments [grp0[g], grp0[g+1]), bottom windows first
    for (int g = 0; g <= NG; g++) {
      if (stagger == 3) grp0[g] = g == 0 ? 0 : (g == 3 ? W : ((g == 1 ? 3 : 9) * W + 8) / 16);   // measured best: 3/16, 9/16
      else if (stagger) grp0[g] = g == 0 ? 0 : (g == NG ? W : (W >> (NG - g)));
      else grp0[g] = (int)((long long)nseg * g / NG);
    }
    if (stagger) {   // $ZKB200_STAGGER_SPLIT = "a,b,..": inner group boundaries in windows, bottom first (NG - 1 values)
      const char* e = getenv("ZKB200_STAGGER_SPLIT");
      if (e) {
        int g = 1;
        for (const char* q = e; *q && g < NG; g++) {
          int v = atoi(q);
          if (v > grp0[g - 1] && v < W) grp0[g] = v;
          while (*q && *q != ',') q++;
          if (*q == ',') q++;
        }
        for (g = 1; g < NG; g++) if (grp0[g] <= grp0[g - 1]) grp0[g] = grp0[g - 1] + 1;
      }
    }
    trace_groups = split_tail ? NG : 0;
    for (int g = 0; g <= NG; g++) trace_g0[g] = grp0[g];
    const int conc_segs = stagger ? nseg : (nseg + NG - 1) / NG * nlanes;   // segments in flight at a time
    // (records of the affine tree: half of the slots are empty and a lane's k_accumulate_rec runs next to the other
    // lanes' kernels, so short chunks -- about 2.5 resident waves of threads over all lanes, at least 16 slots each --
    // keep its latency down: BLS12-381 2^20 median call 6.53 -> 6.45 ms with 16 instead of 48 slots, 2^22 unchanged)
    auto pick_chunk = [&](size_t per_seg, bool records = false) -> int {
      const char* e = getenv("ZKB200_CHUNK");
      if (e && atoi(e) > 0) return atoi(e);
      const size_t pairs_total = per_seg * (size_t)conc_segs;
      double target = (double)pairs_total / ((double)resident * (records ? 2.5 : 12.0));
      if (target < (records ? 16.0 : 48.0)) target = records ? 16.0 : 48.0;
      if (target > 192.0) target = 192.0;
      double waves = (double)pairs_total / ((double)resident * target);
      size_t nw = waves < 1.0 ? 1 : (size_t)(waves + 0.5);
      size_t per_seg_threads = ((size_t)resident * nw) / (size_t)conc_segs;  // threads available to one segment
      if (per_seg_threads < 1) per_seg_threads = 1;
      size_t ch = (per_seg + per_seg_threads - 1) / per_seg_threads;
      if (ch < 8) ch = 8;
      if (ch > 1024) ch = 1024;
      return (int)ch;
    };
    const int chunk = pick_chunk(pmax);
    AffSizes az{};
    AffWork aw{};
    int chunk_rec = chunk;
    size_t binv_stride = 0;
    if (R > 0) {
      az = aff_sizes(pmax, nseg, R);
      binv_stride = az.binv_elems + 64;
      aw.tmp = (uint32_t*)cx.ensure(B_AFF_TMP, az.tmp_points * (size_t)OS * 4);
      aw.pre = (uint32_t*)cx.ensure(B_AFF_PRE, (az.pre_elems + az.pre2_elems) * (size_t)L * 4);
      aw.pre2 = aw.pre + az.pre_elems * (size_t)L;
      aw.binv = (uint32_t*)cx.ensure(B_AFF_BINV, 4 * binv_stride * (size_t)L * 4);
      aw.st[0] = (uint4*)cx.ensure(B_AFF_ST0, az.st0 * 16 + 16);
      aw.st[1] = (uint4*)cx.ensure(B_AFF_ST1, az.st1 * 16 + 16);
      aw.keys_out = (uint32_t*)cx.ensure(B_AFF_KEYS, az.rec * 4 + 16);
      aw.vals_out = (uint32_t*)cx.ensure(B_AFF_VALS, az.rec * 4 + 16);
      chunk_rec = pick_chunk(az.nrec, true);
    }
    uint32_t cps_max = (uint32_t)((pmax + chunk - 1) / chunk);
    if (R > 0) {
      uint32_t c2 = (uint32_t)((az.nrec + chunk_rec - 1) / chunk_rec);
      if (c2 > cps_max) cps_max = c2;
    }
    Mem* heads = (Mem*)cx.ensure(B_HEADS, (size_t)nseg * cps_max * sizeof(Mem));
    uint32_t* head_keys = (uint32_t*)cx.ensure(B_HEADKEYS, (size_t)nseg * cps_max * 4);
    const size_t cap2_per_seg = (cps_max + FIXUP_FAN - 1) / FIXUP_FAN;
    Mem* heads2 = (Mem*)cx.ensure(B_HEADS2, (size_t)nseg * cap2_per_seg * sizeof(Mem));
    uint32_t* head_keys2 = (uint32_t*)cx.ensure(B_HEADKEYS2, (size_t)nseg * cap2_per_seg * 4);
    int log_m1 = 0;
    if (c - 1 > 0) {
      const char* e = getenv("ZKB200_LOGM1");
      // about one wave of threads per launch (measured); with window groups the biggest launch has half the segments
      log_m1 = e ? atoi(e) : ilog2_floor(((size_t)(split_tail ? (nseg + 1) / 2 : nseg) << (c - 1)) / 32768 + 1);
      if (log_m1 < 1) log_m1 = 1;
      if (log_m1 > 5) log_m1 = 5;
      if (log_m1 > c - 1) log_m1 = c - 1;
    }
    // The window group that finishes last (the bottom one) uses the team version of the first reduction level with a
    // short chain (4 buckets per task); the output arrays are spaced for the smaller of the two group sizes.
    const int log_mt = c - 1 < 2 ? c - 1 : 2;
    const int log_mmin = log_mt < log_m1 ? log_mt : log_m1;
    const size_t red_out = (size_t)nseg << (c - 1 - log_mmin);
    Mem* Ub[2] = {(Mem*)cx.ensure(B_U0, red_out * sizeof(Mem)), (Mem*)cx.ensure(B_U1, red_out * sizeof(Mem))};
    Mem* Vb[2] = {(Mem*)cx.ensure(B_V0, red_out * sizeof(Mem)), (Mem*)cx.ensure(B_V1, red_out * sizeof(Mem))};
    Mem* partials = (Mem*)cx.ensure(B_PARTIALS, 16 * sizeof(Mem));
    // low-latency reduction of the window groups (kernels_red.cuh, K5'): row / column sums and bit sums per window
    const bool red2d_on = [] { const char* e = getenv("ZKB200_RED2D"); return e ? atoi(e) != 0 : true; }();   // per call: the tests switch it
    const bool red2d = split_tail && red2d_on && c - 1 >= RED2D_MIN_BITS;
    Mem* red_rc = nullptr;
    Mem* red_t = nullptr;
    if (red2d) {
      red_rc = (Mem*)cx.ensure(B_RED_RC, (size_t)nseg * (((size_t)1 << (c - 1 - c / 2)) + ((size_t)1 << (c / 2))) * sizeof(Mem));
      red_t = (Mem*)cx.ensure(B_RED_T, (size_t)nseg * c * sizeof(Mem));
    }
    if (nlanes == 1) { cx.pl.big[0] = s; cx.pl.chain[0] = s; } else { cx.pl.big[0] = cx.pl.s_big[0]; cx.pl.chain[0] = cx.pl.s_chain[0]; }
    cx.pl.big[1] = cx.pl.s_big[1];
    cx.pl.chain[1] = cx.pl.s_chain[1];

    host_ms[0] = host_now();
    for (int k = 0; k < K; k++) {
      const size_t nk = lo[k + 1] - lo[k];
      if (nk == 0) continue;
      st.slice_ran[k] = true;
      cudaEvent_t* ge = cx.gev + 6 * k;
      // ---- recode + sort this slice's pairs (segment-major with stride nk) ----
      if (sloc == ZKB200_HOST) {
        size_t off = (K == 1 ? 0 : lo[k]) * nl, cnt64 = (K == 1 ? (size_t)nmsm * n : nk) * nl;
        host_to_device(cx, (uint64_t*)d_scalars + off, scalars + off, cnt64 * 8, cx.s_copy);
        CK(cudaEventRecord(cx.ev_sc[k], cx.s_copy));
        CK(cudaStreamWaitEvent(s, cx.ev_sc[k], 0));
      }
      CK(cudaEventRecord(ge[3], s));
      g_launches++;
      const size_t pk = nk * F;          // pairs per segment of this slice
      const int tiles = (int)((pk + SORT_TILE - 1) / SORT_TILE);
      if constexpr (GlvOf<C>::available) {
        if (glv) launch_recode_glv<C>(s, d_scalars + (K == 1 ? 0 : lo[k]) * nl, nl, nk, nmsm, mont, c, W, packed[0]);
        else launch_recode<C>(s, d_scalars + (K == 1 ? 0 : lo[k]) * nl, nl, nk, nmsm, mont, nbits, c, W, packed[0]);
      } else {
        launch_recode<C>(s, d_scalars + (K == 1 ? 0 : lo[k]) * nl, nl, nk, nmsm, mont, nbits, c, W, packed[0]);
      }
      CK(cudaGetLastError());
      CK(cudaEventRecord(ge[5], s));
      // keys are 1 .. 2^(c-1) (0 = no insertion): c bits cover them
      g_launches += sort_pairs(s, packed[0], packed[1], keys_sorted, vals_sorted, pk, nseg, c, cnt, rowsum, tiles);
      CK(cudaGetLastError());
      uint32_t* const keys[1] = {keys_sorted};
      uint32_t* const vals[1] = {vals_sorted};
      const int cur = 0;
      CK(cudaEventRecord(ge[4], s));
      // ---- bucket accumulation of this slice ----
      if (ploc == ZKB200_HOST) {
        size_t off = lo[k] * (size_t)(2 * L), cnt32 = nk * (size_t)(2 * L);
        host_to_device(cx, (uint32_t*)d_points + off, (const uint32_t*)points + off, cnt32 * 4, cx.s_copy);
      }
      const uint32_t* pts_k = d_points + lo[k] * (size_t)(2 * L);
      if constexpr (GlvOf<C>::available) {
        if (glv) {   // on the copy stream: behind this slice's upload, under the recoding and the sort of the pairs
          uint32_t* ext = glv_points + 2 * lo[k] * (size_t)OS;
          g_launches++;
          launch_glv_points<C>(cx.s_copy, pts_k, nk, ext);
          CK(cudaGetLastError());
          pts_k = ext;
        }
      }
      if (ploc == ZKB200_HOST || glv) {
        CK(cudaEventRecord(cx.ev_pt[k], cx.s_copy));
        CK(cudaStreamWaitEvent(s, cx.ev_pt[k], 0));
      }
      CK(cudaEventRecord(ge[0], s));
      Mem* kb_ = buckets + (size_t)k * slice_stride;
      AffSizes zk_ = aff_sizes(pk, nseg, R > 0 ? R : 1);
      const uint32_t cps = R > 0 ? (uint32_t)((zk_.nrec + chunk_rec - 1) / chunk_rec) : (uint32_t)((pk + chunk - 1) / chunk);
      const bool last_slice = k == K - 1;
      // Window groups, top windows first, two at a time on the two lanes.  A group's fix-up runs on its lane; after
      // the LAST slice its bucket reduction and its share of the window combination follow on a high-priority stream
      // of their own, under the accumulation of the lower groups.
      CK(cudaEventRecord(cx.pl.ev_sorted, s));
      if (stagger) {
        for (int j = 0; j < NG; j++) CK(cudaStreamWaitEvent(cx.pl.stag_big[j], cx.pl.ev_sorted, 0));
      } else {
        for (int l = 0; l < nlanes; l++) CK(cudaStreamWaitEvent(cx.pl.big[l], cx.pl.ev_sorted, 0));
      }
      for (int g = NG - 1; g >= 0; g -= nlanes) {
        AffLanes ln{};
        ln.n = (g - 1 >= 0 && nlanes == 2) ? 2 : 1;
        ln.binv_stride = binv_stride;
        int gl[2] = {g, g - 1};
        for (int l = 0; l < ln.n; l++) {
          ln.seg0[l] = grp0[gl[l]];
          ln.segs[l] = grp0[gl[l] + 1] - grp0[gl[l]];
          if (stagger) {
            const int j = NG - 1 - g;   // lane 0 = top group = highest priority
            ln.big[l] = cx.pl.stag_big[j];
            ln.chain[l] = cx.pl.stag_chain[j];
            ln.ev_a[l] = cx.pl.stag_a[j];
            ln.ev_c[l] = cx.pl.stag_c[j];
            ln.binv_base0 = (size_t)j * binv_stride;
          } else {
            ln.big[l] = cx.pl.big[l];
            ln.chain[l] = cx.pl.chain[l];
            ln.ev_a[l] = cx.pl.ev_a[l];
            ln.ev_c[l] = cx.pl.ev_c[l];
          }
        }
        if constexpr (HasAffineTree<C>::value) {
          if (R > 0)
            g_launches += launch_affine_tree<C>(ln, keys[cur], vals[cur], pts_k, pstride, pk, R, NB, kb_, aw, chunk_rec, cps, heads, head_keys);
        }
        for (int l = 0; l < ln.n; l++) {
          const int s0 = ln.seg0[l], ns = ln.segs[l];
          cudaStream_t sl = ln.big[l];
          if (R == 0) {
            g_launches++;
            launch_accumulate<C>(sl, keys[cur] + (size_t)s0 * pk, vals[cur] + (size_t)s0 * pk, pts_k, pstride, pk, ns, chunk, cps, NB,
                                 kb_ + (size_t)s0 * NB, heads + (size_t)s0 * cps, head_keys + (size_t)s0 * cps);
          }
          CK(cudaGetLastError());
          {  // fold the chunk heads into the buckets: log_FAN(chunks) small levels
            uint32_t T = cps;
            Mem* hb[2] = {heads + (size_t)s0 * cps, heads2 + (size_t)s0 * cap2_per_seg};
            uint32_t* kb[2] = {head_keys + (size_t)s0 * cps, head_keys2 + (size_t)s0 * cap2_per_seg};
            int src = 0;
            for (;;) {
              uint32_t T_out = (T + FIXUP_FAN - 1) / FIXUP_FAN;
              int last = T_out == 1;
              g_launches++;
              launch_fixup_level<C>(sl, kb[src], hb[src], T, kb[src ^ 1], hb[src ^ 1], T_out, ns, NB, kb_ + (size_t)s0 * NB, last);
              CK(cudaGetLastError());
              if (last) break;
              T = T_out;
              src ^= 1;
            }
          }
          if (last_slice && split_tail) {
            // ---- this group's bucket reduction by levels + its share of the window combination ----
            cudaStream_t sp = cx.pl.post[gl[l]];
            CK(cudaEventRecord(cx.pl.ev_fix[gl[l]], sl));
            CK(cudaStreamWaitEvent(sp, cx.pl.ev_fix[gl[l]], 0));
            if (red2d && ns <= RED2D_MAX_WINDOWS) {
              const size_t rc_per = ((size_t)1 << (c - 1 - c / 2)) + ((size_t)1 << (c / 2));
              g_launches += launch_reduce_2d<C>(sp, buckets + (size_t)s0 * NB, K, slice_stride, ns, c, c * s0, red_rc + (size_t)s0 * rc_per,
                                                red_t + (size_t)s0 * c, partials + gl[l]);
              CK(cudaGetLastError());
            } else {
              const size_t ubase = (size_t)s0 << (c - 1 - log_mmin);
              const bool team = gl[l] == 0 && reduce_team;   // the bottom group is the one the caller ends up waiting for
              const int lm1 = team ? log_mt : log_m1;
              int logS = c - 1;
              size_t total_out = (size_t)ns << (logS - lm1);
              g_launches++;
              launch_reduce_first<C>(sp, buckets + (size_t)s0 * NB, K, slice_stride, total_out, lm1, Ub[0] + ubase, Vb[0] + ubase, team);
              CK(cudaGetLastError());
              logS -= lm1;
              int lv = 0;
              while (logS > 0) {
                int lm = logS > 3 ? 3 : logS;
                total_out = (size_t)ns << (logS - lm);
                g_launches++;
                launch_reduce_next<C>(sp, Ub[lv] + ubase, Vb[lv] + ubase, total_out, lm, Ub[lv ^ 1] + ubase, Vb[lv ^ 1] + ubase);
                CK(cudaGetLastError());
                logS -= lm;
                lv ^= 1;
              }
              g_launches++;
              launch_tail_group<C>(sp, Ub[lv] + ubase, ns, c, c * s0, partials + gl[l]);
              CK(cudaGetLastError());
            }
            CK(cudaEventRecord(cx.pl.ev_post[gl[l]], sp));
          }
        }
      }
      // the caller's stream resumes when the lanes are done (the next slice re-uses the pair arrays)
      if (stagger) {
        for (int j = 0; j < NG; j++) {
          CK(cudaEventRecord(cx.pl.stag_lane[j], cx.pl.stag_big[j]));
          CK(cudaStreamWaitEvent(s, cx.pl.stag_lane[j], 0));
        }
      } else {
        for (int l = 0; l < nlanes; l++) {
          if (cx.pl.big[l] == s) continue;
          CK(cudaEventRecord(cx.pl.ev_lane[l], cx.pl.big[l]));
          CK(cudaStreamWaitEvent(s, cx.pl.ev_lane[l], 0));
        }
      }
      CK(cudaEventRecord(ge[1], s));
      CK(cudaEventRecord(ge[2], s));
    }

    CK(cudaEventRecord(cx.ev[5], s));
    if (split_tail) {
      // ---- the groups' shares are summed and converted to the output representation ----
      for (int g = 0; g < NG; g++) CK(cudaStreamWaitEvent(s, cx.pl.ev_post[g], 0));
      CK(cudaEventRecord(cx.ev[6], s));
      g_launches++;
      launch_sum_points<C>(s, (const uint32_t*)partials, NG, OUT_XYZZ, out_mode, d_out);
      CK(cudaGetLastError());
    } else {
      // ---- bucket reduction by levels ----
      int logS = c - 1;
      size_t total_out = (size_t)nseg << (logS - log_m1);
      g_launches++;
      launch_reduce_first<C>(s, buckets, K, slice_stride, total_out, log_m1, Ub[0], Vb[0], 0);
      CK(cudaGetLastError());
      logS -= log_m1;
      int lv = 0;
      while (logS > 0) {
        int lm = logS > 3 ? 3 : logS;
        total_out = (size_t)nseg << (logS - lm);
        g_launches++;
        launch_reduce_next<C>(s, Ub[lv], Vb[lv], total_out, lm, Ub[lv ^ 1], Vb[lv ^ 1]);
        CK(cudaGetLastError());
        logS -= lm;
        lv ^= 1;
      }
      CK(cudaEventRecord(cx.ev[6], s));
      // ---- window combination (Horner) + output conversion ----
      g_launches++;
      launch_tail<C>(s, Ub[lv], nmsm, W, c, out_mode, d_out);
      CK(cudaGetLastError());
    }
    st.have_phases = true;
  }
  if (out_loc == ZKB200_DEVICE) {
    // result records stay on the device (multi-GPU combine without a host bounce): device-to-device, record by record
    for (int m = 0; m < nmsm; m++)
      CK(cudaMemcpyAsync((uint32_t*)out + (size_t)m * out_coords * L, d_out + (size_t)m * 4 * L, (size_t)out_coords * L * 4,
                         cudaMemcpyDeviceToDevice, s));
  } else {
    CK(cudaMemcpyAsync(h_out, d_out, (size_t)nmsm * 4 * L * 4, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaEventRecord(cx.ev[8], s));
  host_ms[1] = host_now();
  CK(cudaStreamSynchronize(s));
  CK(cudaStreamSynchronize(cx.s_copy));
  host_ms[2] = host_now();
  if (out_loc != ZKB200_DEVICE)
    for (int m = 0; m < nmsm; m++)
      memcpy((uint32_t*)out + (size_t)m * out_coords * L, h_out + (size_t)m * 4 * L, (size_t)out_coords * L * 4);
  if (st.have_phases) {
    // phase times summed over the slices (CUDA events on the launching stream)
    const int map[5][3] = {{3, 5, 1}, {5, 4, 2}, {4, 0, 3}, {0, 1, 4}, {1, 2, 5}};  // {from, to, stats slot}
    for (int k = 0; k < st.slices; k++) {
      if (!st.slice_ran[k]) continue;
      cudaEvent_t* ge = cx.gev + 6 * k;
      for (auto& mp : map) {
        float t;
        CK(cudaEventElapsedTime(&t, ge[mp[0]], ge[mp[1]]));
        st.ms[mp[2]] += t;
      }
    }
    CK(cudaEventElapsedTime(&st.ms[6], cx.ev[5], cx.ev[6]));
    CK(cudaEventElapsedTime(&st.ms[7], cx.ev[6], cx.ev[8]));
  }
  CK(cudaEventElapsedTime(&st.ms[8], cx.ev[0], cx.ev[8]));
  static const bool trace = getenv("ZKB200_TRACE") != nullptr;
  if (trace && st.have_phases) {
    // timeline of the window groups (ms since the start of the call): when a group's buckets were complete (its lane
    // finished the fix-up) and when its bucket reduction + share of the window combination were done
    float t_sorted = 0, t_lanes = 0, t_post = 0;
    CK(cudaEventElapsedTime(&t_sorted, cx.ev[0], cx.gev[4]));
    CK(cudaEventElapsedTime(&t_lanes, cx.ev[0], cx.ev[5]));
    CK(cudaEventElapsedTime(&t_post, cx.ev[0], cx.ev[6]));
    fprintf(stderr, "[zkmsm_b200 trace] n=%zu c=%d W=%d R=%d sorted %.3f lanes_done %.3f post_done %.3f total %.3f | host: arrays %.3f "
            "queued %.3f done %.3f |", n, c, W, st.aff_levels, t_sorted, t_lanes, t_post, st.ms[8], host_ms[0], host_ms[1], host_ms[2]);
    if (trace_groups > 1)
      for (int g = trace_groups - 1; g >= 0; g--) {
        float a = 0, b = 0;
        if (cudaEventElapsedTime(&a, cx.ev[0], cx.pl.ev_fix[g]) == cudaSuccess && cudaEventElapsedTime(&b, cx.ev[0], cx.pl.ev_post[g]) == cudaSuccess)
          fprintf(stderr, " g%d[%d,%d) buckets %.3f post %.3f |", g, trace_g0[g], trace_g0[g + 1], a, b);
        else (void)cudaGetLastError();
      }
    fprintf(stderr, "\n");
  }
  cx.stats = st;
}

// ---- devices used by the C-ABI entry points -----------------------------------------------------------
// $ZKB200_DEVICES = "all" | "0,1,2,3" makes every host-buffer call shard its work over several GPUs of the box
// from inside the library (one host thread per device), so that the unchanged single-process Haskell caller
// gets the multi-GPU path too.  Default: the single device of zkb200_set_device / $ZKB200_DEVICE.
std::mutex g_devlist_mu;
std::vector<int> g_devlist;
bool g_devlist_set = false;

std::vector<int> device_list() {
  std::lock_guard<std::mutex> lk(g_devlist_mu);
  if (!g_devlist_set) {
    g_devlist_set = true;
    const char* e = getenv("ZKB200_DEVICES");
    if (e && *e) {
      int count = 0;
      if (cudaGetDeviceCount(&count) != cudaSuccess) count = 0;
      if (!strcmp(e, "all")) {
        for (int i = 0; i < count && i < MAX_DEV; i++) g_devlist.push_back(i);
      } else {
        const char* p = e;
        while (*p) {
          char* end = nullptr;
          long v = strtol(p, &end, 10);
          if (end == p) break;
          if (v >= 0 && v < count && v < MAX_DEV) g_devlist.push_back((int)v);
          p = (*end == ',') ? end + 1 : end;
        }
      }
    }
  }
  return g_devlist;
}

template <class C>
void run_on(int dev, int nmsm, size_t n, const uint64_t* scalars, int sloc, const uint64_t* points, int ploc, int nl, int mont,
            int out_mode, int window, uint64_t* out, int out_loc = ZKB200_HOST) {
  DeviceCtx& cx = get_ctx(dev);
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  if (ploc == ZKB200_HOST && n > 0) {
    const size_t bytes = n * (size_t)(2 * C::Fp::L) * 4;
    uint64_t fp = 0;
    const uint32_t* resident = srs_lookup(cx, CurveId<C>::value, points, bytes, &fp);
    if (resident) {
      run_msm<C>(cx, nmsm, n, scalars, sloc, (const uint64_t*)resident, ZKB200_DEVICE, nl, mont, out_mode, window, out, out_loc);
      cx.stats.srs_hit = true;
      return;
    }
    run_msm<C>(cx, nmsm, n, scalars, sloc, points, ploc, nl, mont, out_mode, window, out, out_loc);
    if (fp) srs_insert(cx, CurveId<C>::value, points, bytes, fp, cx.buf[B_POINTS], cx.s_main);   // the upload is still there
    return;
  }
  run_msm<C>(cx, nmsm, n, scalars, sloc, points, ploc, nl, mont, out_mode, window, out, out_loc);
}

template <class C>
void run_sum(DeviceCtx& cx, int k, const uint64_t* in, int in_mode, int out_mode, uint64_t* out, int in_loc = ZKB200_HOST);

// Sharded execution over several devices (host buffers only).
//  * one MSM: contiguous slices of both vectors, one XYZZ partial per device, summed on the first device
//  * a batch over a shared point array: whole MSMs are dealt out to the devices
template <class C>
void run_multi(const std::vector<int>& devs, int nmsm, size_t n, const uint64_t* scalars, const uint64_t* points, int nl,
               int mont, int out_mode, int window, uint64_t* out) {
  constexpr int L = C::Fp::L;   // u64 words per affine point = 2 coordinates x L/2
  const int G = (int)devs.size();
  const int out_coords = out_mode == OUT_AFFINE ? 2 : (out_mode == OUT_XYZZ ? 4 : 3);
  std::vector<std::thread> th;
  if (nmsm == 1) {
    std::vector<uint64_t> parts((size_t)G * 4 * (L / 2));
    for (int g = 0; g < G; g++) {
      size_t lo = n * (size_t)g / G, hi = n * (size_t)(g + 1) / G;
      th.emplace_back([=, &parts] {
        run_on<C>(devs[g], 1, hi - lo, scalars + lo * nl, ZKB200_HOST, points + lo * L, ZKB200_HOST, nl, mont, OUT_XYZZ, window,
                  parts.data() + (size_t)g * 4 * (L / 2));
      });
    }
    for (auto& t : th) t.join();
    DeviceCtx& cx = get_ctx(devs[0]);
    DeviceGuard guard(cx.dev);
    std::lock_guard<std::mutex> lk(cx.mu);
    run_sum<C>(cx, G, parts.data(), OUT_XYZZ, out_mode, out);
  } else {
    for (int g = 0; g < G; g++) {
      int lo = (int)((long long)nmsm * g / G), hi = (int)((long long)nmsm * (g + 1) / G);
      if (hi <= lo) continue;
      th.emplace_back([=] {
        run_on<C>(devs[g], hi - lo, n, scalars + (size_t)lo * n * nl, ZKB200_HOST, points, ZKB200_HOST, nl, mont, out_mode, window,
                  out + (size_t)lo * out_coords * (L / 2));
      });
    }
    for (auto& t : th) t.join();
  }
}

void msm_entry(int curve, int nmsm, long n, const uint64_t* scalars, int sloc, const uint64_t* points, int ploc, int nl,
               int mont, int out_mode, int window, uint64_t* out, int out_loc = ZKB200_HOST) {
  if (n < 0) n = 0;
  if (curve < 0 || curve > ZKB200_BLS12_381_G2) { fprintf(stderr, "[zkmsm_b200] fatal: unknown curve id %d\n", curve); abort(); }
  std::vector<int> devs = device_list();
  if (n > 0) {
    // a device pointer must live on the GPU the call runs on (no peer access is set up): fail with a message, not a fault
    const int exec_dev = current_device_choice();
    const void* dp[3] = {sloc == ZKB200_DEVICE ? scalars : nullptr, ploc == ZKB200_DEVICE ? points : nullptr,
                         out_loc == ZKB200_DEVICE ? out : nullptr};
    for (const void* q : dp) {
      if (!q) continue;
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, q) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) || attr.device != exec_dev) {
        (void)cudaGetLastError();
        fprintf(stderr, "[zkmsm_b200] fatal: a pointer passed as device memory is not device memory of GPU %d (the device this "
                        "call runs on: zkb200_set_device / $ZKB200_DEVICE / a one-element device list)\n", exec_dev);
        abort();
      }
    }
  }
  const bool multi = devs.size() > 1 && sloc == ZKB200_HOST && ploc == ZKB200_HOST && out_loc == ZKB200_HOST && nmsm >= 1 &&
                     ((nmsm == 1 && (size_t)n >= ((size_t)1 << 16) * devs.size()) || (nmsm >= (int)devs.size()));
  if (multi) {
    switch (curve) {
      case ZKB200_BN128: run_multi<Bn254>(devs, nmsm, (size_t)n, scalars, points, nl, mont, out_mode, window, out); break;
      case ZKB200_BLS12_381: run_multi<Bls12381>(devs, nmsm, (size_t)n, scalars, points, nl, mont, out_mode, window, out); break;
      case ZKB200_BN128_G2: run_multi<Bn254G2>(devs, nmsm, (size_t)n, scalars, points, nl, mont, out_mode, window, out); break;
      default: run_multi<Bls12381G2>(devs, nmsm, (size_t)n, scalars, points, nl, mont, out_mode, window, out); break;
    }
    return;
  }
  const int dev = -1;   // current_device_choice()
  switch (curve) {
    case ZKB200_BN128: run_on<Bn254>(dev, nmsm, (size_t)n, scalars, sloc, points, ploc, nl, mont, out_mode, window, out, out_loc); break;
    case ZKB200_BLS12_381: run_on<Bls12381>(dev, nmsm, (size_t)n, scalars, sloc, points, ploc, nl, mont, out_mode, window, out, out_loc); break;
    case ZKB200_BN128_G2: run_on<Bn254G2>(dev, nmsm, (size_t)n, scalars, sloc, points, ploc, nl, mont, out_mode, window, out, out_loc); break;
    default: run_on<Bls12381G2>(dev, nmsm, (size_t)n, scalars, sloc, points, ploc, nl, mont, out_mode, window, out, out_loc); break;
  }
}

template <class C>
void run_sum(DeviceCtx& cx, int k, const uint64_t* in, int in_mode, int out_mode, uint64_t* out, int in_loc) {
  constexpr int L = C::Fp::L;
  const int in_coords = in_mode == OUT_XYZZ ? 4 : 3;
  const int out_coords = out_mode == OUT_AFFINE ? 2 : (out_mode == OUT_XYZZ ? 4 : 3);
  size_t in_bytes = (size_t)(k > 0 ? k : 1) * in_coords * L * 4;
  uint32_t* d_in = (uint32_t*)cx.ensure(B_SCALARS, in_bytes);
  uint32_t* d_out = (uint32_t*)cx.ensure(B_OUT, 4 * L * 4);
  uint32_t* h_out = cx.ensure_host(4 * L * 4);
  if (k > 0 && in_loc == ZKB200_DEVICE) d_in = (uint32_t*)in;   // partial points already on this device (NCCL all-gather output)
  else if (k > 0) CK(cudaMemcpyAsync(d_in, in, (size_t)k * in_coords * L * 4, cudaMemcpyHostToDevice, cx.s_main));
  g_launches++;
  launch_sum_points<C>(cx.s_main, d_in, k, in_mode, out_mode, d_out);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h_out, d_out, 4 * L * 4, cudaMemcpyDeviceToHost, cx.s_main));
  CK(cudaStreamSynchronize(cx.s_main));
  memcpy(out, h_out, (size_t)out_coords * L * 4);
}

// ---- IMAD throughput probe ------------------------------------------------------------------------
// Register-resident products with a DIFFERENT multiplier per repetition (bb[r], advanced every iteration): with one
// shared multiplier ptxas keeps a*b and turns the repeated mad.wide into 64-bit additions, which is what made the
// round-1 "mad.wide" figure look 1.7x faster than the carry chains (tools/imad_forms.cu documents the forms and the SASS).
template <int KIND>
__global__ void __launch_bounds__(256) k_imad_probe(uint32_t* out, int iters) {
  uint32_t a[8], E[8], O[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 2654435761u + i * 40503u + 1u; E[i] = a[i] ^ 0x9e3779b9u; O[i] = a[i] + i; }
  uint32_t b = blockIdx.x + 12345u;
  uint32_t bb[4] = {b, b * 3u + 1u, b * 5u + 2u, b * 7u + 3u};
  if (KIND == 0) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {  // 4 x (two independent 4-product carry chains) = 32 products
        cmad_row<8, false>(E, a, bb[r]);
        cmad_row<8, false>(O, a, bb[r] ^ 0x55u);
      }
      b += E[0];
#pragma unroll
      for (int q = 0; q < 4; q++) bb[q] += b;
    }
  } else if (KIND == 1) {
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = ((unsigned long long)E[i] << 32) | O[i];
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(bb[r]));
      b += (uint32_t)acc[0];
#pragma unroll
      for (int q = 0; q < 4; q++) bb[q] += b;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { E[i] = (uint32_t)acc[i]; O[i] = (uint32_t)(acc[i] >> 32); }
  } else if (KIND == 3) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++)   // product only, folded in by one 3-input logic op on the other pipe
          asm volatile("{ .reg .u64 t; .reg .u32 lo, hi; mul.wide.u32 t, %1, %2; mov.b64 {lo, hi}, t; lop3.b32 %0, %0, lo, hi, 0x96; }"
                       : "+r"(E[i]) : "r"(a[i]), "r"(bb[r]));
      b += E[0];
#pragma unroll
      for (int q = 0; q < 4; q++) bb[q] += b;
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(E[i]) : "r"(a[i]), "r"(bb[r]));
      b += E[0];
#pragma unroll
      for (int q = 0; q < 4; q++) bb[q] += b;
    }
  }
  uint32_t x = b;
#pragma unroll
  for (int i = 0; i < 8; i++) x ^= E[i] ^ O[i];
  if (x == 0x12345u) out[0] = x;  // keeps the chains alive without measurable traffic
}

template <class C>
void run_gen_chain(DeviceCtx& cx, unsigned long long start, long n, const uint64_t* p0, const uint64_t* d, uint64_t* out,
                          int out_loc) {
  constexpr int L = C::Fp::L;
  if (n <= 0) return;
  uint32_t* d_in = (uint32_t*)cx.ensure(B_SCALARS, 4 * L * 4);
  CK(cudaMemcpyAsync(d_in, p0, 2 * L * 4, cudaMemcpyHostToDevice, cx.s_main));
  CK(cudaMemcpyAsync(d_in + 2 * L, d, 2 * L * 4, cudaMemcpyHostToDevice, cx.s_main));
  size_t bytes = (size_t)n * 2 * L * 4;
  uint32_t* d_out = out_loc == ZKB200_DEVICE ? (uint32_t*)out : (uint32_t*)cx.ensure(B_POINTS, bytes);
  g_launches++;
  launch_gen_chain<C>(cx.s_main, d_in, start, (size_t)n, d_out);
  CK(cudaGetLastError());
  if (out_loc != ZKB200_DEVICE) CK(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, cx.s_main));
  CK(cudaStreamSynchronize(cx.s_main));
}

// batch conversions between the reference's representations (host buffers in, host buffers out)
template <class C>
void run_convert(int N, const uint64_t* src, uint64_t* tgt, int jac, int to_affine) {
  constexpr int L = C::Fp::L;
  if (N <= 0) return;
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  size_t n = (size_t)N;
  size_t in_bytes = n * (to_affine ? 3 : 2) * L * 4, out_bytes = n * (to_affine ? 2 : 3) * L * 4;
  uint32_t* d_in = (uint32_t*)cx.ensure(B_POINTS, in_bytes);
  uint32_t* d_out = (uint32_t*)cx.ensure(B_KEYS0, out_bytes);
  cudaStream_t s = cx.s_main;
  CK(cudaMemcpyAsync(d_in, src, in_bytes, cudaMemcpyHostToDevice, s));
  CK(cudaEventRecord(cx.ev[0], s));
  if (to_affine && n >= 4096) {
    // one inversion for the whole array: Z coordinates -> batch inversion tree -> scale every point by its own 1/Z
    uint32_t* d_z = (uint32_t*)cx.ensure(B_AFF_PRE, n * (size_t)L * 4);
    uint32_t* d_ws = (uint32_t*)cx.ensure(B_AFF_BINV, (binv_workspace_elems(n) + 64) * (size_t)L * 4);
    uint32_t* d_inv = nullptr;
    launch_convert_z<C>(s, d_in, n, d_z);
    CK(cudaGetLastError());
    g_launches += 2 + launch_batch_invert<C>(s, d_z, n, d_ws, &d_inv);
    CK(cudaGetLastError());
    launch_convert_apply<C>(s, d_in, d_inv, n, d_out, jac);
  } else {
    g_launches++;
    if (to_affine) launch_batch_to_affine<C>(s, d_in, n, d_out, jac);
    else launch_batch_from_affine<C>(s, d_in, n, d_out, jac);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(cx.ev[1], s));
  CK(cudaMemcpyAsync(tgt, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(&cx.last_op_ms, cx.ev[0], cx.ev[1]));
}

// Fr NTT (scope row 8f.2).  src / tgt may be host or device memory (device: the transform can feed
// zkb200_msm's device-resident scalar input without leaving the GPU); gen is always a host pointer.
template <class F>
void run_ntt(int m, const uint64_t* gen, const uint64_t* src, int src_loc, uint64_t* tgt, int tgt_loc, int inverse) {
  if (m < 0 || m > 30) { fprintf(stderr, "[zkmsm_b200] fatal: NTT size 2^%d unsupported\n", m); abort(); }
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  const size_t N = (size_t)1 << m, bytes = N * 32;
  uint32_t* d_gen = (uint32_t*)cx.ensure(B_OUT, 4096);
  uint32_t* d_tmp = (uint32_t*)cx.ensure(B_KEYS0, bytes);
  uint32_t* d_table = (uint32_t*)cx.ensure(B_NTT_TABLE, (N / 2 + 1) * 32);
  const uint32_t* d_src = (const uint32_t*)src;
  uint32_t* d_dst = (uint32_t*)tgt;
  if (tgt_loc != ZKB200_DEVICE || (const void*)tgt == (const void*)src) d_dst = (uint32_t*)cx.ensure(B_KEYS1, bytes);
  cudaStream_t s = cx.s_main;
  CK(cudaMemcpyAsync(d_gen, gen, 32, cudaMemcpyHostToDevice, s));
  if (src_loc != ZKB200_DEVICE) {
    uint32_t* p = (uint32_t*)cx.ensure(B_SCALARS, bytes);
    host_to_device(cx, p, src, bytes, s);
    d_src = p;
  }
  g_launches += 1 + (m + 8) / 9;
  CK(cudaEventRecord(cx.ev[0], s));
  ntt_device<F>(s, m, d_gen, d_src, d_tmp, d_dst, d_table, inverse);
  CK(cudaGetLastError());
  CK(cudaEventRecord(cx.ev[1], s));
  if (tgt_loc != ZKB200_DEVICE) CK(cudaMemcpyAsync(tgt, d_dst, bytes, cudaMemcpyDeviceToHost, s));
  else if ((void*)d_dst != (void*)tgt) CK(cudaMemcpyAsync(tgt, d_dst, bytes, cudaMemcpyDeviceToDevice, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(&cx.last_op_ms, cx.ev[0], cx.ev[1]));
}

// FFT of G1 group elements (scope row 8f.4): host buffers in, host buffers out
template <class C>
void run_gfft(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt, int inverse, int jac) {
  constexpr int L = C::Fp::L;
  if (m < 0 || m > 26) { fprintf(stderr, "[zkmsm_b200] fatal: group FFT size 2^%d unsupported\n", m); abort(); }
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  const size_t N = (size_t)1 << m, bytes = N * 3 * L * 4;
  uint32_t* d_src = (uint32_t*)cx.ensure(B_POINTS, bytes);
  uint32_t* d_dst = (uint32_t*)cx.ensure(B_KEYS0, bytes);
  void* d_work = cx.ensure(B_BUCKETS, N * sizeof(XyzzMem<typename C::Fp>));
  uint32_t* d_table = (uint32_t*)cx.ensure(B_NTT_TABLE, (N / 2 + 1) * 32);
  uint32_t* d_gen = (uint32_t*)cx.ensure(B_OUT, 4096);
  cudaStream_t s = cx.s_main;
  CK(cudaMemcpyAsync(d_gen, gen, 32, cudaMemcpyHostToDevice, s));
  host_to_device(cx, d_src, src, bytes, s);
  g_launches += 3 + m;
  CK(cudaEventRecord(cx.ev[0], s));
  gfft_device<C>(s, m, d_gen, d_src, d_work, d_table, d_dst, inverse, jac, glv_enabled() ? 1 : 0);
  CK(cudaGetLastError());
  CK(cudaEventRecord(cx.ev[1], s));
  CK(cudaMemcpyAsync(tgt, d_dst, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventElapsedTime(&cx.last_op_ms, cx.ev[0], cx.ev[1]));
}

}  // namespace

// ---- exported C ABI -----------------------------------------------------------------------------------
#pragma GCC visibility push(default)
extern "C" {

void zkb200_msm(int curve, int nmsm, long npoints, const uint64_t* scalars, int scalars_loc, const uint64_t* points,
                int points_loc, int expo_nlimbs, int mont_coeff, int out_mode, int window, uint64_t* out) {
  msm_entry(curve, nmsm, npoints, scalars, scalars_loc, points, points_loc, expo_nlimbs, mont_coeff, out_mode, window, out);
}

void zkb200_msm_ex(int curve, int nmsm, long npoints, const uint64_t* scalars, int scalars_loc, const uint64_t* points,
                   int points_loc, int expo_nlimbs, int mont_coeff, int out_mode, int window, uint64_t* out, int out_loc) {
  msm_entry(curve, nmsm, npoints, scalars, scalars_loc, points, points_loc, expo_nlimbs, mont_coeff, out_mode, window, out, out_loc);
}

void zkb200_sum_points_ex(int curve, int k, const uint64_t* in, int in_loc, int in_mode, int out_mode, uint64_t* out) {
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  if (curve == ZKB200_BN128) run_sum<Bn254>(cx, k, in, in_mode, out_mode, out, in_loc);
  else if (curve == ZKB200_BLS12_381) run_sum<Bls12381>(cx, k, in, in_mode, out_mode, out, in_loc);
  else if (curve == ZKB200_BN128_G2) run_sum<Bn254G2>(cx, k, in, in_mode, out_mode, out, in_loc);
  else if (curve == ZKB200_BLS12_381_G2) run_sum<Bls12381G2>(cx, k, in, in_mode, out_mode, out, in_loc);
  else { fprintf(stderr, "[zkmsm_b200] fatal: unknown curve id %d\n", curve); abort(); }
}

void zkb200_sum_points(int curve, int k, const uint64_t* in, int in_mode, int out_mode, uint64_t* out) {
  zkb200_sum_points_ex(curve, k, in, ZKB200_HOST, in_mode, out_mode, out);
}

void zkb200_set_device(int device) { g_device.store(device); }

void zkb200_set_devices(const int* devices, int count) {
  std::lock_guard<std::mutex> lk(g_devlist_mu);
  g_devlist.clear();
  for (int i = 0; i < count; i++) if (devices[i] >= 0 && devices[i] < MAX_DEV) g_devlist.push_back(devices[i]);
  g_devlist_set = true;
}

long long zkb200_launch_count(void) { return g_launches.load(); }

void* zkb200_device_upload(const void* host, size_t bytes) {
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  void* d = nullptr;
  CK(cudaMalloc(&d, bytes ? bytes : 1));
  host_to_device(cx, d, host, bytes, cx.s_copy);
  CK(cudaStreamSynchronize(cx.s_copy));
  return d;
}

void zkb200_device_free(void* device_ptr) {
  if (!device_ptr) return;
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  CK(cudaFree(device_ptr));
}

void zkb200_release_workspaces(void) {
  for (int d = 0; d < MAX_DEV; d++) {
    DeviceCtx& cx = g_ctx[d];
    if (!cx.ready) continue;
    DeviceGuard guard(cx.dev);
    std::lock_guard<std::mutex> lk(cx.mu);
    cx.release_workspaces();
  }
}

void zkb200_srs_cache_drop(void) {
  for (int d = 0; d < MAX_DEV; d++) {
    DeviceCtx& cx = g_ctx[d];
    if (!cx.ready) continue;
    DeviceGuard guard(cx.dev);
    std::lock_guard<std::mutex> lk(cx.mu);
    cx.srs_drop_all();
  }
}

void zkb200_set_glv(int on) { g_glv.store(on ? 1 : 0); }

float zkb200_last_op_ms(void) {
  DeviceCtx& cx = get_ctx();
  std::lock_guard<std::mutex> lk(cx.mu);
  return cx.last_op_ms;
}

int zkb200_last_srs_hit(void) {
  DeviceCtx& cx = get_ctx();
  std::lock_guard<std::mutex> lk(cx.mu);
  return cx.stats.srs_hit ? 1 : 0;
}

// Element-wise self-tests of the device primitives (selftest.cu): host arrays in, host array out, temporary device
// buffers (not the MSM workspaces).  field: 0 = bn128 Fp, 1 = bls12_381 Fp, 2 = bn128 Fr, 3 = bls12_381 Fr.
void zkb200_selftest_field(int field, int op, long n, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d,
                           uint64_t* out) {
  if (n <= 0) return;
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  const int L = (field == 1) ? 12 : 8;
  const size_t bytes = (size_t)n * L * 4;
  const uint64_t* in[4] = {a, b, c, d};
  uint32_t* din[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t* dout = nullptr;
  cudaStream_t s = cx.s_main;
  for (int k = 0; k < 4; k++)
    if (in[k]) { CK(cudaMalloc((void**)&din[k], bytes)); CK(cudaMemcpyAsync(din[k], in[k], bytes, cudaMemcpyHostToDevice, s)); }
  CK(cudaMalloc((void**)&dout, bytes));
  g_launches++;
  switch (field) {
    case 0: launch_selftest_field<Bn254Fp>(s, op, (size_t)n, din[0], din[1], din[2], din[3], dout); break;
    case 1: launch_selftest_field<Bls12381Fp>(s, op, (size_t)n, din[0], din[1], din[2], din[3], dout); break;
    case 2: launch_selftest_field<Bn254Fr>(s, op, (size_t)n, din[0], din[1], din[2], din[3], dout); break;
    case 3: launch_selftest_field<Bls12381Fr>(s, op, (size_t)n, din[0], din[1], din[2], din[3], dout); break;
    default: fprintf(stderr, "[zkmsm_b200] fatal: unknown field id %d\n", field); abort();
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int k = 0; k < 4; k++) if (din[k]) CK(cudaFree(din[k]));
  CK(cudaFree(dout));
}

void zkb200_selftest_group(int curve, int op, long n, const uint64_t* p1, const uint64_t* z1, const uint64_t* p2, const uint64_t* z2,
                           uint64_t* out_affine) {
  if (n <= 0) return;
  if (curve != ZKB200_BN128 && curve != ZKB200_BLS12_381) { fprintf(stderr, "[zkmsm_b200] fatal: selftest_group: G1 curves only\n"); abort(); }
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  const int L = curve == ZKB200_BLS12_381 ? 12 : 8;
  const size_t fb = (size_t)n * L * 4;
  uint32_t *dp1, *dz1, *dp2, *dz2, *dout;
  cudaStream_t s = cx.s_main;
  CK(cudaMalloc((void**)&dp1, 2 * fb)); CK(cudaMalloc((void**)&dp2, 2 * fb)); CK(cudaMalloc((void**)&dout, 2 * fb));
  CK(cudaMalloc((void**)&dz1, fb)); CK(cudaMalloc((void**)&dz2, fb));
  CK(cudaMemcpyAsync(dp1, p1, 2 * fb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dp2, p2, 2 * fb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dz1, z1, fb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dz2, z2, fb, cudaMemcpyHostToDevice, s));
  g_launches++;
  if (curve == ZKB200_BN128) launch_selftest_group<Bn254>(s, op, (size_t)n, dp1, dz1, dp2, dz2, dout);
  else launch_selftest_group<Bls12381>(s, op, (size_t)n, dp1, dz1, dp2, dz2, dout);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_affine, dout, 2 * fb, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaFree(dp1)); CK(cudaFree(dp2)); CK(cudaFree(dout)); CK(cudaFree(dz1)); CK(cudaFree(dz2));
}

void zkb200_ntt(int curve, int m, const uint64_t* gen, const uint64_t* src, int src_loc, uint64_t* tgt, int tgt_loc, int inverse) {
  if (curve == ZKB200_BN128) run_ntt<Bn254Fr>(m, gen, src, src_loc, tgt, tgt_loc, inverse);
  else if (curve == ZKB200_BLS12_381) run_ntt<Bls12381Fr>(m, gen, src, src_loc, tgt, tgt_loc, inverse);
  else { fprintf(stderr, "[zkmsm_b200] fatal: unknown curve id %d\n", curve); abort(); }
}

void zkb200_gen_chain(int curve, unsigned long long start, long n, const uint64_t* p0_affine, const uint64_t* d_affine,
                      uint64_t* out, int out_loc) {
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  if (curve == ZKB200_BN128) run_gen_chain<Bn254>(cx, start, n, p0_affine, d_affine, out, out_loc);
  else if (curve == ZKB200_BLS12_381) run_gen_chain<Bls12381>(cx, start, n, p0_affine, d_affine, out, out_loc);
  else if (curve == ZKB200_BN128_G2) run_gen_chain<Bn254G2>(cx, start, n, p0_affine, d_affine, out, out_loc);
  else if (curve == ZKB200_BLS12_381_G2) run_gen_chain<Bls12381G2>(cx, start, n, p0_affine, d_affine, out, out_loc);
  else { fprintf(stderr, "[zkmsm_b200] fatal: unknown curve id %d\n", curve); abort(); }
}

void zkb200_last_stats(float phase_ms[9], int* window_c, int* nwindows, long long* insertions) {
  DeviceCtx& cx = get_ctx();
  std::lock_guard<std::mutex> lk(cx.mu);
  if (phase_ms) memcpy(phase_ms, cx.stats.ms, sizeof(float) * 9);
  if (window_c) *window_c = cx.stats.c;
  if (nwindows) *nwindows = cx.stats.W;
  if (insertions) *insertions = cx.stats.insertions;
}

int zkb200_last_affine_levels(void) {
  DeviceCtx& cx = get_ctx();
  std::lock_guard<std::mutex> lk(cx.mu);
  return cx.stats.aff_levels;
}

double zkb200_imad_peak(int kind, int iters) {
  DeviceCtx& cx = get_ctx();
  DeviceGuard guard(cx.dev);
  std::lock_guard<std::mutex> lk(cx.mu);
  if (iters < 1) iters = 1;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cx.dev));
  int blocks = prop.multiProcessorCount * 8;
  uint32_t* d = (uint32_t*)cx.ensure(B_OUT, 256);
  cudaStream_t s = cx.s_main;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(cx.ev[0], s));
    if (kind == 0) k_imad_probe<0><<<blocks, 256, 0, s>>>(d, iters);
    else if (kind == 1) k_imad_probe<1><<<blocks, 256, 0, s>>>(d, iters);
    else if (kind == 3) k_imad_probe<3><<<blocks, 256, 0, s>>>(d, iters);
    else k_imad_probe<2><<<blocks, 256, 0, s>>>(d, iters);
    CK(cudaGetLastError());
    CK(cudaEventRecord(cx.ev[1], s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, cx.ev[0], cx.ev[1]));
    if (rep > 0 && ms < best) best = ms;
  }
  double products = (double)blocks * 256.0 * (double)iters * 32.0;
  return products / (best * 1e-3);
}

const char* zkb200_version(void) { return "zkmsm_b200 0.1 (sm_100a)"; }

#define ZK_REF_SYMBOLS(NAME, ID)                                                                                      \
  void NAME##_G1_proj_MSM_std_coeff_proj_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {      \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_PROJ, 0, t); }                                     \
  void NAME##_G1_proj_MSM_mont_coeff_proj_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {     \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 1, OUT_PROJ, 0, t); }                                     \
  void NAME##_G1_proj_MSM_std_coeff_affine_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {    \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_AFFINE, 0, t); }                                   \
  void NAME##_G1_proj_MSM_mont_coeff_affine_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {   \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 1, OUT_AFFINE, 0, t); }                                   \
  void NAME##_G1_jac_MSM_std_coeff_jac_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {        \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_JAC, 0, t); }                                      \
  void NAME##_G1_jac_MSM_mont_coeff_jac_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {       \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 1, OUT_JAC, 0, t); }                                      \
  void NAME##_G1_jac_MSM_std_coeff_affine_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {     \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_AFFINE, 0, t); }                                   \
  void NAME##_G1_jac_MSM_mont_coeff_affine_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {    \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 1, OUT_AFFINE, 0, t); }                                   \
  void NAME##_G1_proj_MSM_std_coeff_proj_out_variable(int n, const uint64_t* e, const uint64_t* g, uint64_t* t,       \
                                                      int nl, int ws) {                                               \
    if (ws < 1 || ws > 64) { fprintf(stderr, "[zkmsm_b200] fatal: window_size %d out of range\n", ws); abort(); }     \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_PROJ, ws, t); }                                    \
  void NAME##_G1_jac_MSM_std_coeff_jac_out_variable(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl, \
                                                    int ws) {                                                         \
    if (ws < 1 || ws > 64) { fprintf(stderr, "[zkmsm_b200] fatal: window_size %d out of range\n", ws); abort(); }     \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_JAC, ws, t); }

#define ZK_CONVERT_SYMBOLS(NAME, CURVE)                                                                               \
  void NAME##_G1_proj_batch_to_affine(int N, const uint64_t* src, uint64_t* tgt) { run_convert<CURVE>(N, src, tgt, 0, 1); }   \
  void NAME##_G1_jac_batch_to_affine(int N, const uint64_t* src, uint64_t* tgt) { run_convert<CURVE>(N, src, tgt, 1, 1); }    \
  void NAME##_G1_proj_batch_from_affine(int N, const uint64_t* src, uint64_t* tgt) { run_convert<CURVE>(N, src, tgt, 0, 0); } \
  void NAME##_G1_jac_batch_from_affine(int N, const uint64_t* src, uint64_t* tgt) { run_convert<CURVE>(N, src, tgt, 1, 0); }

#define ZK_NTT_SYMBOLS(NAME, CURVE)                                                                                   \
  void NAME##_poly_mont_ntt_forward(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                 \
    run_ntt<CURVE::Fr>(m, gen, src, ZKB200_HOST, tgt, ZKB200_HOST, 0); }                                                                        \
  void NAME##_poly_mont_ntt_inverse(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                 \
    run_ntt<CURVE::Fr>(m, gen, src, ZKB200_HOST, tgt, ZKB200_HOST, 1); }

#define ZK_GFFT_SYMBOLS(NAME, CURVE, CURVE2)                                                                          \
  void NAME##_G1_proj_fft_forward(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                   \
    run_gfft<CURVE>(m, gen, src, tgt, 0, 0); }                                                                        \
  void NAME##_G1_proj_fft_inverse(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                   \
    run_gfft<CURVE>(m, gen, src, tgt, 1, 0); }                                                                        \
  void NAME##_G1_jac_fft_forward(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                    \
    run_gfft<CURVE>(m, gen, src, tgt, 0, 1); }                                                                        \
  void NAME##_G1_jac_fft_inverse(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                    \
    run_gfft<CURVE>(m, gen, src, tgt, 1, 1); }                                                                        \
  void NAME##_G2_proj_fft_forward(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                   \
    run_gfft<CURVE2>(m, gen, src, tgt, 0, 0); }                                                                       \
  void NAME##_G2_proj_fft_inverse(int m, const uint64_t* gen, const uint64_t* src, uint64_t* tgt) {                   \
    run_gfft<CURVE2>(m, gen, src, tgt, 1, 0); }                                                                       \
  void NAME##_G2_proj_batch_to_affine(int N, const uint64_t* src, uint64_t* tgt) { run_convert<CURVE2>(N, src, tgt, 0, 1); }   \
  void NAME##_G2_proj_batch_from_affine(int N, const uint64_t* src, uint64_t* tgt) { run_convert<CURVE2>(N, src, tgt, 0, 0); }

ZK_GFFT_SYMBOLS(bn128, Bn254, Bn254G2)
ZK_GFFT_SYMBOLS(bls12_381, Bls12381, Bls12381G2)

ZK_NTT_SYMBOLS(bn128, Bn254)
ZK_NTT_SYMBOLS(bls12_381, Bls12381)

ZK_CONVERT_SYMBOLS(bn128, Bn254)
ZK_CONVERT_SYMBOLS(bls12_381, Bls12381)

// G2 (scope row 8f.3): proj + affine outputs only, like the reference (lib/cbits/curves/g2/proj/bn128_G2_proj.h:43-46)
#define ZK_REF_SYMBOLS_G2(NAME, ID)                                                                                   \
  void NAME##_G2_proj_MSM_std_coeff_proj_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {      \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_PROJ, 0, t); }                                     \
  void NAME##_G2_proj_MSM_mont_coeff_proj_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {     \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 1, OUT_PROJ, 0, t); }                                     \
  void NAME##_G2_proj_MSM_std_coeff_affine_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {    \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_AFFINE, 0, t); }                                   \
  void NAME##_G2_proj_MSM_mont_coeff_affine_out(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {   \
    msm_entry(ID, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 1, OUT_AFFINE, 0, t); }

// The reference also exports (and never calls) a slow sum-of-scalar-multiplications variant of the std-coefficient MSM
// (bn128_G1_proj.c:610-619, header name misspelt :47).  Same group element; served by the same pipeline.
#define ZK_SLOW_REF_SYMBOLS(NAME, ID1, ID2)                                                                                \
  void NAME##_G1_proj_MSM_std_coeff_proj_out_slow_reference(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) { \
    msm_entry(ID1, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_PROJ, 0, t); }                                          \
  void NAME##_G1_jac_MSM_std_coeff_jac_out_slow_reference(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) {   \
    msm_entry(ID1, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_JAC, 0, t); }                                           \
  void NAME##_G2_proj_MSM_std_coeff_proj_out_slow_reference(int n, const uint64_t* e, const uint64_t* g, uint64_t* t, int nl) { \
    msm_entry(ID2, 1, n, e, ZKB200_HOST, g, ZKB200_HOST, nl, 0, OUT_PROJ, 0, t); }

ZK_SLOW_REF_SYMBOLS(bn128, ZKB200_BN128, ZKB200_BN128_G2)
ZK_SLOW_REF_SYMBOLS(bls12_381, ZKB200_BLS12_381, ZKB200_BLS12_381_G2)

ZK_REF_SYMBOLS_G2(bn128, ZKB200_BN128_G2)
ZK_REF_SYMBOLS_G2(bls12_381, ZKB200_BLS12_381_G2)

ZK_REF_SYMBOLS(bn128, ZKB200_BN128)
ZK_REF_SYMBOLS(bls12_381, ZKB200_BLS12_381)

}  // extern "C"
#pragma GCC visibility pop
