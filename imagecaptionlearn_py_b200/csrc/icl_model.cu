// libicl_b200.so -- host orchestration + C-ABI (include/icl_b200.h) for the BiLSTM + mention-span-head path.
// Replaces the TF graph built by nn_utils/core.py:271-514,74-106 and executed by sess.run (core.py:625).
#include "../../include/icl_b200.h"
#include "icl_kernels.cuh"
#include "gemm_tcgen05.cuh"
#include "lstm_persistent.cuh"
#include "lstm_bptt.cuh"
#ifdef ICL_EXPERIMENTS
#include "lstm_bptt_ns.cuh"     // k_bptt_nsplit: measured slower than k_bptt_cluster (0.66 vs 0.36 ms), see DESIGN.md section 7
#endif
#include "lstm_fwd16.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <unordered_map>
#include <numeric>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <immintrin.h>
#include <vector>

using namespace icl;

static thread_local char g_err[1024] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
#define CK(x)                                                                                         \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
  } while (0)
#define CKI(x)                  \
  do {                          \
    int r_ = (x);               \
    if (r_ != 0) return r_;     \
  } while (0)

struct Param { std::string name; int rows, cols; int64_t off; };

// Inputs are double-buffered (device buffers + pinned staging): icl_upload packs batch i+1 and copies it on a separate copy
// stream while the compute stream still runs step i.  The "current" pointers below (Head::idx, icl_model::xraw, ...) are
// switched to the freshly uploaded set at the end of icl_upload.
// One set = ONE device allocation and its pinned mirror with the same layout: [ x | meta ], meta = the small integer arrays
// of the step layout followed by every head's index matrices / dense features / labels.  The sentence rows go up in chunks
// while they are being packed; everything else is ONE copy (a dozen small cudaMemcpyAsync calls on the copy stream stalled the
// concurrently running compute stream by ~80 us each -- measured 1.3 ms per step).
struct HeadIn {                                             // byte offsets into the set's blob
  size_t idx[ICL_N_INDEX] = {}, feats = 0, box = 0, bfeats = 0, labels = 0, boxrow = 0;
  // affinity layer-1 factorisation: group of every pair, representative pair of every group, CSR member lists (int32 each)
  size_t m_of = 0, b_of = 0, rep_m = 0, rep_boxrow = 0, m_start = 0, m_mem = 0, b_start = 0, b_mem = 0;
};
struct InSet {
  char *d_blob = nullptr, *h_blob = nullptr;
  size_t o_x = 0, o_lens = 0, o_rank = 0, o_tokstart = 0, o_tokseq = 0, o_tokrow = 0, o_off = 0, o_nact = 0, meta_off = 0, bytes = 0;
  template <typename T> T* dev(size_t o) const { return reinterpret_cast<T*>(d_blob + o); }
  template <typename T> T* host(size_t o) const { return reinterpret_cast<T*>(h_blob + o); }
  cudaEvent_t ev_copied = nullptr, ev_done = nullptr, ev_stats = nullptr;
  cudaEvent_t ev_c0 = nullptr, ev_s0 = nullptr;           // timing: first H2D copy issued / first kernel of the step
  bool copied_pending = false, done_pending = false, stats_pending = false;
};

struct Head {
  icl_head_config c;
  int D0 = 0, D0g = 0;               // input width; width of the prefix that carries gathered LSTM columns
  int ldbi = 0, lddbi = 0;           // row pitches of bi / dbi: D0 / D0g rounded up to 4 floats.  n_mention_feats is data-defined
                                     // (nn_utils/data.py:187), so D0 = 4H + F need not be a multiple of 4, and a TMA tensor map needs
                                     // 16-byte row pitches -- with a dense pitch layer 1 (the largest head GEMM) fell back to SIMT
  std::vector<int> dims;             // D0, w1..wL, C
  std::vector<int> pW, pB;           // param indices per layer (L hidden + softmax)
  SlotTable slots;                   // device pointers filled at create
  std::vector<int> slot_index_id;    // per slot: index-matrix id or -1
  float *bi = nullptr, *dbi = nullptr, *dA = nullptr, *dBuf = nullptr;
  std::vector<float*> act;           // y_k [B,w_k]
  std::vector<float*> dzb;           // gradient w.r.t. the pre-activation of hidden layer k [B,w_k] (one buffer per layer: the weight
                                     // gradients read it on the aux stream while the main stream already computes the next layer's)
  float *proba = nullptr, *dlogits = nullptr, *smb_part = nullptr, *row_loss = nullptr, *row_correct = nullptr, *scalars = nullptr;
  long long* pred = nullptr;
  int* idx[ICL_N_INDEX] = {};
  float *feats = nullptr, *box = nullptr, *bfeats = nullptr, *labels = nullptr;
  bool has_labels = false, active = true;
  float loss_w = 1.0f;               // d joint_loss / d loss of this head (icl_set_loss_weights)
  HeadIn in[2];
  float* h_out = nullptr; long long* h_pred = nullptr;   // pinned staging of the results
  // multi-head models (icl_multitask_lstm.py): the heads only share the encoder, so every head runs its forward / backward chain on
  // its OWN stream pair (hs: the dz chain, ha: its weight gradients), forked from and joined to the model's stream -- five heads of
  // latency-bound kernels side by side instead of one after another.  Single-head models use the model's stream / aux2.
  cudaStream_t hs = nullptr, ha = nullptr;
  cudaEvent_t ev_done = nullptr, ev_adone = nullptr;
  // Affinity layer 1, factorised (the concat it replaces: core.py:421-433,439): batch_input of a pair is
  // [mention columns (Dm) | box columns (Db)], so z1 = U[mention of the pair] + V[box of the pair] + b1 with
  // U = Xm W1[0:Dm] over the batch's DISTINCT mentions and V = Xb W1[Dm:D0] over its distinct boxes (k_pair_combine).
  bool fact_ok = false;              // an affinity head with a box block (and ICL_AFF_FACTOR != 0)
  bool fact = false;                 // the uploaded batch runs factorised (its groups are few enough, or ICL_AFF_FACTOR=2)
  bool box_compact = false;          // fact, host box rows: only the distinct rows were packed (row g of the box / b_feats blocks = group g)
  int Dm = 0, Db = 0, ldm = 0, ldb = 0, Mu = 0, Nu = 0;
  int64_t fact_batches = 0, fact_mu = 0, fact_nu = 0;     // statistics (icl_head_factor_stats)
  SlotTable slots_m, slots_b;        // the mention / box halves of `slots` (rows = groups)
  std::vector<int> slotm_src, slotb_src;
  float *Xm = nullptr, *Xb = nullptr, *U = nullptr, *V = nullptr, *dU = nullptr, *dV = nullptr, *dXm = nullptr;
  int *d_m_of = nullptr, *d_b_of = nullptr, *d_rep_m = nullptr, *d_rep_boxrow = nullptr, *d_m_start = nullptr, *d_m_mem = nullptr,
      *d_b_start = nullptr, *d_b_mem = nullptr, *d_boxrow = nullptr;
  cudaStream_t sb = nullptr;         // the box half of the forward pass: needs no LSTM state, runs beside the recurrence
  cudaEvent_t ev_bfork = nullptr, ev_box = nullptr, ev_du = nullptr;
};

enum { PH_PREP = 0, PH_PROJ, PH_REC_FWD, PH_HEADS_FWD, PH_HEADS_BWD, PH_REC_BWD, PH_WGRAD, PH_UPDATE, PH_N };

struct icl_model {
  icl_config cfg;
  int E, H, S_cap, T_cap;
  long Ntok_cap;
  std::vector<Param> params;
  std::map<std::string, int> pindex;
  int64_t n_params = 0;
  float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr, *Pr = nullptr;   // Pr: TF32-rounded copy of P (GEMM operands)
  bool pr_dirty = true;
  int round_ops = 1;
  int pK[2], pBias[2];
  int64_t step = 0;
  // optimizer slots: the reference's `alternate` multitask scheme builds one AdamOptimizer per task (icl_multitask_lstm.py:387-393),
  // i.e. separate m / v slots and beta-power accumulators for every variable each of them touches.  Slot 0 = M, V, step above.
  struct OptSlot { float *M = nullptr, *V = nullptr; int64_t step = 0; };
  std::vector<OptSlot> slots;
  int cur_slot = 0;
  // workspaces (step-major, see icl_kernels.cuh)
  // XH[d] = [rows, E+H+4]: columns [0,E) the prepared inputs (xd), [E,E+H) the TF32 h_{k-1} operand rows (Hp), column E+H = 1
  // on valid rows -- one matrix so that [dKernel; dbias] = XH^T dZ is a single GEMM (the ones column sums dZ into the bias
  // gradient, which follows the kernel in the flat gradient buffer); xd / Hp are views with row pitch ldx
  float *XH[2] = {};
  int ldx = 0;
  bool x_half = false;               // the current input set's sentence rows are fp16 (half-width wire format)
  float *xraw = nullptr, *xd[2] = {}, *Z[2] = {}, *Hx[2] = {}, *Hp[2] = {}, *Cc[2] = {}, *dHout[2] = {}, *dhrec[2] = {}, *dcc[2] = {},
        *R[2] = {};
  int *d_off = nullptr, *d_nact = nullptr, *d_rank = nullptr, *d_lens = nullptr, *d_tokseq = nullptr, *d_tokstart = nullptr;
  // device-resident token table (icl_set_token_table): batches then carry int32 row numbers instead of embedding rows
  float* tok_table = nullptr; int64_t tok_table_rows = 0;
  int* d_tokrow = nullptr; bool use_rows = false;
  // device-resident box-feature table (icl_set_box_table): affinity batches then carry one int32 row per pair
  float* box_table = nullptr; int64_t box_table_rows = 0;
  InSet in[2];
  int cur = -1;                                         // input set of the resident batch
  cudaStream_t copy = nullptr;
  float* h_stats = nullptr;                             // pinned [2 sets][ICL_MAX_HEADS][loss, accuracy]
  double* d_partial = nullptr; float* d_gnorm = nullptr;
  std::vector<Head> heads;
  // resident batch
  int S = 0, Tmax = 0; long Ntok = 0;
  int64_t seq_gid0 = 0, ex_gid0 = 0;
  std::vector<int> n_active, off;
  bool resident = false;
  cudaStream_t stream = nullptr, aux = nullptr, aux2 = nullptr, aux3 = nullptr, aux4 = nullptr;      // aux2: the heads' weight gradients
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_join3 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  bool G_external = false;             // the gradient buffer belongs to the caller (symmetric memory, icl_adopt_grad_buffer)
  cudaEvent_t ev_heads = nullptr;      // recorded when the heads' parameter gradients are complete (before the BPTT)
  cudaEvent_t ev_wg0 = nullptr;        // recorded when the forward direction's LSTM weight gradient is complete (before the backward direction's)
  cudaEvent_t ev_dz[ICL_MAX_LAYERS + 2] = {};   // heads' backward: "dz of this layer is ready" (main stream -> aux stream)
  bool heads_aux_pending = false;      // the aux stream holds weight-gradient work the main stream has not joined yet
  cudaEvent_t ev_ph[PH_N][2] = {};
  bool ph_used[PH_N] = {}, ph_on = false;      // per-phase timers: off unless asked for (icl_set_phase_timing / ICL_PHASE_EVENTS=1)
  // persistent recurrent kernels
  int rp_U = 0, rp_nsl = 0, rp_nkb = 0, rp_max_tiles = 0;   // rp_U == 0: not available for this H (per-step path)
  bool rp_on = true, wp_dirty = true;
  long rows_cap = 0, NtokP = 0;
  float* Wp[2] = {};
  unsigned* rp_flags = nullptr;
  long long* rp_trace = nullptr; int rp_trace_cta = 0;
#ifdef ICL_EXPERIMENTS
  RecFwdMaps rp_fmaps;
  RecBwdMaps rp_bmaps;
#endif
  unsigned* rp_bar = nullptr;
  bool rp_bwd_on = false;       // k_rec_bwd is correct but measured slower (0.84 ms) than the per-step path (0.56 ms): opt-in
  // fp16 input projection (K1): fp16 copies of the prepared inputs and of W_ih^T, kind::f16 GEMM
  bool k1_f16 = false, wih_dirty = true;
  int k1_Kp = 0;
  __half* X16[2] = {}; __half* Wih16[2] = {};
  CUtensorMap k1_ta[2], k1_tb[2], k1_tc[2];
  // second-generation forward recurrence (lstm_fwd16.cuh): fp16 recurrent operands, double-buffered x-projection boxes
  bool rf_on = false;
  struct RfVar {                 // one slicing of the fp16 forward recurrence (k_rec_fwd16<U>): U = 20 (H % 20 == 0) and / or U = 16 (any H)
    bool on = false, dirty = true;
    int U = 0, UP = 0, nsl = 0, KP = 0, nkb = 0, nk16 = 0, maxtpc = 0;
    __half* Hp16[2] = {}; __half* Wp16[2] = {};
    RecFwd16Maps maps;
  } rf20, rf16;
  RfVar* rf_last = nullptr;      // the slicing the last forward pass used: repacked eagerly after the update
  cudaEvent_t ev_packs = nullptr; bool packs_pending = false;
  cudaEvent_t ev_side = nullptr, ev_loss = nullptr; bool loss_pending = false;   // loss / accuracy sums run beside the backward pass
  cudaEvent_t ev_hfork = nullptr;      // multi-head models: the point of the model's stream every head stream starts from
  float last_keep = 1.0f; uint64_t last_seed = 0;      // of the last icl_run_resident (icl_get_batch_input of a factorised head)
  cudaEvent_t ev_zfork = nullptr, ev_zero = nullptr; bool zero_pending = false, dh_clean = false; int zero_early = 1, zero_blocks = 0; cudaEvent_t ev_dhzero = nullptr;   // ICL_ZERO_EARLY=0: A/B
  bool pdl = true;                     // ICL_PDL=0: the heads' GEMM chains without programmatic dependent launches (A/B)
  bool head_streams = true;            // ICL_HEAD_STREAMS=0: the heads of a multi-head model one after another on the model's stream (A/B)
  // fused BPTT step kernel (lstm_bptt.cuh): default backward recurrence in tensor-core mode
  bool bp_on = false, bp_cluster = true;     // bp_cluster: k_bptt_cluster (whole recurrence in one launch) when H <= 336
  int bp_cs = 4;
  BpttMaps bp_maps;
  // K3 second generation (lstm_bptt_ns.cuh): output units split over an 8-CTA cluster, fp16 operands with per-row scales
  bool ns_on = false, wb_dirty = true;
  int ns_UN = 0;
  __half* dZ16[2] = {}; float* S16[2] = {}; __half* Wb16[2] = {};
#ifdef ICL_EXPERIMENTS
  BpttNsMaps ns_maps;
#endif
  int n_sms = 148;
  int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
  float last_ms = 0.f; bool t_recorded = false;
  TmaCache tma;
};

// ----------------------------------------------------------------------------- small helpers
static double read_num(const void* p, int dtype, size_t i) {
  switch (dtype) {
    case ICL_F32: return ((const float*)p)[i];
    case ICL_F64: return ((const double*)p)[i];
    case ICL_I32: return ((const int32_t*)p)[i];
    default: return (double)((const int64_t*)p)[i];
  }
}
// Streaming row copies into the pinned input mirror: whole 64-byte lines of the destination are written with non-temporal stores
// (no read-for-ownership of a buffer the CPU never reads back, no cache pollution next to the thread that drives the GPU); the
// worker fences before the DMA is queued.  card2048 end to end: 1.80 -> 1.575 ms per step (profiles/r1i_pack_ab.txt).
// ICL_PACK_NT=0 restores memcpy.
static const bool g_pack_nt = [] { const char* e = getenv("ICL_PACK_NT"); return !e || atoi(e) != 0; }();
static void copy_f32_stream(float* dst, const float* src, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = src[i]; i++; }
  for (; i + 16 <= n; i += 16) {
    __m128 a = _mm_loadu_ps(src + i), b = _mm_loadu_ps(src + i + 4), c = _mm_loadu_ps(src + i + 8), d = _mm_loadu_ps(src + i + 12);
    _mm_stream_ps(dst + i, a); _mm_stream_ps(dst + i + 4, b); _mm_stream_ps(dst + i + 8, c); _mm_stream_ps(dst + i + 12, d);
  }
  for (; i < n; i++) dst[i] = src[i];
}
// float64 -> float32 (the reference feeds np.zeros() float64 arrays, nn_utils/data.py:375): cvtpd2ps rounds to nearest even like the cast
static void cvt_f64_stream(float* dst, const double* src, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = (float)src[i]; i++; }
  for (; i + 16 <= n; i += 16)
    for (int k = 0; k < 16; k += 4)
      _mm_stream_ps(dst + i + k, _mm_movelh_ps(_mm_cvtpd_ps(_mm_loadu_pd(src + i + k)), _mm_cvtpd_ps(_mm_loadu_pd(src + i + k + 2))));
  for (; i < n; i++) dst[i] = (float)src[i];
}
// Half-width wire format of the sentence rows (ICL_WIRE_FP16, default on where it applies: see icl_upload): the rows are rounded to
// fp16 (round to nearest even, F16C) while they are packed, so the pinned mirror and the H2D copy carry 2 bytes per element.  The
// device rounds the prepared inputs to 10 mantissa bits anyway (fp16 operand of the projection GEMM, TF32 operand of the weight
// gradient), so this moves that rounding in front of the input scaling instead of behind it.
__attribute__((target("avx,f16c"))) static void cvt_f32_h16_stream(uint16_t* dst, const float* src, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = _cvtss_sh(src[i], _MM_FROUND_TO_NEAREST_INT); i++; }
  for (; i + 32 <= n; i += 32)
    for (int k = 0; k < 32; k += 8)
      _mm_stream_si128((__m128i*)(dst + i + k), _mm256_cvtps_ph(_mm256_loadu_ps(src + i + k), _MM_FROUND_TO_NEAREST_INT));
  for (; i < n; i++) dst[i] = _cvtss_sh(src[i], _MM_FROUND_TO_NEAREST_INT);
}
__attribute__((target("avx,f16c"))) static void cvt_f64_h16_stream(uint16_t* dst, const double* src, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = _cvtss_sh((float)src[i], _MM_FROUND_TO_NEAREST_INT); i++; }
  for (; i + 32 <= n; i += 32)
    for (int k = 0; k < 32; k += 8) {
      const __m256 f = _mm256_set_m128(_mm256_cvtpd_ps(_mm256_loadu_pd(src + i + k + 4)), _mm256_cvtpd_ps(_mm256_loadu_pd(src + i + k)));
      _mm_stream_si128((__m128i*)(dst + i + k), _mm256_cvtps_ph(f, _MM_FROUND_TO_NEAREST_INT));
    }
  for (; i < n; i++) dst[i] = _cvtss_sh((float)src[i], _MM_FROUND_TO_NEAREST_INT);
}
static bool wire16_enabled() {            // read per upload, so a test can switch it
  static const bool cpu_ok = __builtin_cpu_supports("avx") && __builtin_cpu_supports("f16c");
  const char* e = getenv("ICL_WIRE_FP16");
  return cpu_ok && (!e || atoi(e) != 0);
}
static void to_f32(float* dst, const void* src, int dtype, size_t n) {
  if (dtype == ICL_F32) {
    if (g_pack_nt && n >= 64) copy_f32_stream(dst, (const float*)src, n); else memcpy(dst, src, n * sizeof(float));
    return;
  }
  if (dtype == ICL_F64) {
    const double* s = (const double*)src;
    if (g_pack_nt && n >= 64) cvt_f64_stream(dst, s, n); else for (size_t i = 0; i < n; i++) dst[i] = (float)s[i];
    return;
  }
  for (size_t i = 0; i < n; i++) dst[i] = (float)read_num(src, dtype, i);
}

// Host worker pool for the batch marshalling (pack / convert sentences into pinned memory).  Threads are created once
// per process: spawning them per call costs more than the 30 MB copy they parallelise.  How many of them pack is MEASURED, not
// assumed: the pool holds up to 12 threads (the host cores split between the ranks of the box, one left to the thread that drives
// the GPU) and the first large uploads of a process time the packing with 4 / 6 / 8 / 12 of them active and keep the fastest
// (ICL_HOST_THREADS fixes the number instead).  Why not simply "all": packing overlaps the previous step's kernels, and threads that
// saturate host DRAM slow the GPU's own command fetches over PCIe (round 1: 16 threads stretched a 1.77 ms step to 3.0 ms; round 2,
// 16-core single-GPU box, fp16 wire + non-temporal stores: e2e 1.97 / 1.47 / 1.29 / 1.29 / 1.29 / 1.45 ms per step with
// 2 / 4 / 6 / 8 / 12 / 16 threads), and on an 8-GPU box eight ranks share one memory system.
struct HostPool {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::function<void(int)> job;
  int n_items = 0, busy = 0;
  std::atomic<int> next{0};
  std::atomic<int> active{1};        // threads that take items, the caller included
  uint64_t gen = 0;
  bool stop = false;
  // tuner: candidate thread counts, best time seen for each, uploads measured so far
  std::vector<int> cand; std::vector<double> best; int trials = 0; bool tuned = false;
  static constexpr int TRIALS_PER = 4, WARM = 2;
  HostPool() {
    unsigned share = std::thread::hardware_concurrency();            // one process per GPU: split the host cores between the ranks
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) share /= (unsigned)std::max(1, atoi(e));
    if (share > 4) share -= 1;      // leave a core per rank to the thread that drives the GPU -- unless a rank has four cores or fewer
                                    // (8 ranks on a 32-core box: 4 packers measured 3.36 ms per e2e step, 3 packers 3.57)
    int cap = (int)std::max(1u, std::min(12u, share));
    if (const char* e = getenv("ICL_HOST_THREADS")) { cap = std::max(1, atoi(e)); tuned = true; }
    active = tuned ? cap : std::min(cap, 4);
    if (!tuned) {
      for (int c : {4, 6, 8, 12}) if (c <= cap) cand.push_back(c);
      if (cand.size() < 2) tuned = true;
      best.assign(cand.size(), 1e30);
    }
    for (int i = 1; i < cap; i++) th.emplace_back([this, i] { worker(i); });
  }
  // one large packing pass took `ms` with the current number of active threads: next candidate, or settle on the fastest
  void report(double ms) {
    if (tuned) return;
    const int t = trials++ - WARM;
    if (t < 0) return;
    const size_t c = (size_t)t / TRIALS_PER;
    if (c < cand.size()) best[c] = std::min(best[c], ms);
    const size_t nc = (size_t)(t + 1) / TRIALS_PER;
    if (nc < cand.size()) { active = cand[nc]; return; }
    size_t arg = 0;
    for (size_t i = 1; i < cand.size(); i++) if (best[i] < 0.97 * best[arg]) arg = i;      // more threads only for a clear gain
    active = cand[arg];
    tuned = true;
    if (getenv("ICL_HOST_THREADS_VERBOSE")) {
      fprintf(stderr, "icl_b200: host packing threads:");
      for (size_t i = 0; i < cand.size(); i++) fprintf(stderr, " %d -> %.3f ms", cand[i], best[i]);
      fprintf(stderr, "; using %d\n", active.load());
    }
  }
  ~HostPool() {
    { std::lock_guard<std::mutex> l(mu); stop = true; }
    cv.notify_all();
    for (auto& t : th) t.join();
  }
  void drain() { for (int i; (i = next.fetch_add(1)) < n_items;) job(i); }
  void worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return stop || gen != seen; });
        if (stop) return;
        seen = gen;
        if (id >= active.load()) continue;          // not among the threads that pack (see report())
        busy++;
      }
      drain();
      { std::lock_guard<std::mutex> l(mu); busy--; }
      cv_done.notify_one();
    }
  }
  // run job(i) for i in [0, n) on the pool + the calling thread; returns when all items are done
  void run(int n, std::function<void(int)> f) {
    if (n <= 0) return;
    {
      std::unique_lock<std::mutex> l(mu);
      cv_done.wait(l, [&] { return busy == 0; });          // a late waker of the previous run may still be draining
      job = std::move(f); n_items = n; next = 0; gen++;
    }
    cv.notify_all();
    drain();
    std::unique_lock<std::mutex> l(mu);
    cv_done.wait(l, [&] { return busy == 0 && next.load() >= n_items; });
  }
};
static HostPool& host_pool() { static HostPool* p = new HostPool(); return *p; }   // leaked on purpose (no exit-time join)


// ----------------------------------------------------------------------------- affinity layer-1 factorisation: host side
static int aff_mode() { const char* e = getenv("ICL_AFF_FACTOR"); return e ? atoi(e) : 1; }   // 0 off, 1 when it pays, 2 always
static uint64_t mix64(uint64_t h, uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); return h * 0xff51afd7ed558ccdull; }
static uint64_t hash_bytes(const void* p, size_t n, uint64_t h = 0x243f6a8885a308d3ull) {
  const unsigned char* c = (const unsigned char*)p;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t v; memcpy(&v, c + i, 8); h = mix64(h, v); }
  if (i < n) { uint64_t v = 0; memcpy(&v, c + i, n - i); h = mix64(h, v); }
  return h;
}
// 32 eight-byte samples spread over a long row: picks the candidate group, the full comparison decides
static uint64_t hash_sampled(const void* p, size_t n, uint64_t h = 0x13198a2e03707344ull) {
  if (n <= 512) return hash_bytes(p, n, h);
  const unsigned char* c = (const unsigned char*)p;
  const size_t stride = (n - 8) / 31;
  for (int i = 0; i < 32; i++) { uint64_t v; memcpy(&v, c + (size_t)i * stride, 8); h = mix64(h, v); }
  return h;
}
// groups of equal rows in order of first appearance: of[r] = group of row r, rep[g] = first row of group g.  `same` is the exact test.
template <typename Same>
static void group_rows(int B, const std::vector<uint64_t>& key, Same same, int* of, std::vector<int>& rep) {
  std::unordered_map<uint64_t, std::vector<int>> cand;
  cand.reserve((size_t)B * 2);
  rep.clear();
  for (int r = 0; r < B; r++) {
    std::vector<int>& c = cand[key[r]];
    int g = -1;
    for (int q : c) if (same(r, rep[q])) { g = q; break; }
    if (g < 0) { g = (int)rep.size(); rep.push_back(r); c.push_back(g); }
    of[r] = g;
  }
}
static void csr_lists(int B, int G, const int* of, int* start, int* mem) {
  for (int g = 0; g <= G; g++) start[g] = 0;
  for (int r = 0; r < B; r++) start[of[r] + 1]++;
  for (int g = 0; g < G; g++) start[g + 1] += start[g];
  std::vector<int> fill(start, start + G);
  for (int r = 0; r < B; r++) mem[fill[of[r]]++] = r;           // pair order inside a group: the order the segment sums add in
}
static size_t dtype_size(int dt) { return (dt == ICL_F64 || dt == ICL_I64) ? 8 : 4; }

// Distinct boxes of a batch: row r is (box-table row number | the xsz bytes of its host box row) + the bsz bytes of its b_feats row.
// Table rows are compared as they are grouped.  Host rows (16 KB each) are grouped by a SAMPLED hash alone, then every row is
// compared in full with its group's first row on the worker pool (8 MB of memcmp per 512 pairs, in parallel); only if two different
// rows shared a hash is the grouping redone with the exact comparison inside (tests/test_cabi_cpu.py builds such rows).
static void group_box_rows(int B, const char* xsrc, size_t xsz, const int32_t* table_rows, const char* bsrc, size_t bsz, int* b_of,
                           std::vector<int>& rep_b) {
  std::vector<uint64_t> key(B);
  const int per = 32;
  host_pool().run((B + per - 1) / per, [&](int it) {
    for (int r = it * per; r < std::min(B, (it + 1) * per); r++) {
      uint64_t k = table_rows ? mix64(0x3c6ef372fe94f82bull, (uint64_t)table_rows[r]) : hash_sampled(xsrc + r * xsz, xsz);
      if (bsz) k = hash_bytes(bsrc + r * bsz, bsz, k);
      key[r] = k;
    }
  });
  auto same_box = [&](int a, int c) {
    if (table_rows ? table_rows[a] != table_rows[c] : memcmp(xsrc + a * xsz, xsrc + c * xsz, xsz) != 0) return false;
    return bsz == 0 || memcmp(bsrc + a * bsz, bsrc + c * bsz, bsz) == 0;
  };
  group_rows(B, key, [&](int a, int c) { return table_rows ? same_box(a, c) : true; }, b_of, rep_b);
  if (!table_rows) {
    std::atomic<int> mismatch{0};
    host_pool().run((B + 15) / 16, [&](int it) {
      for (int r = it * 16; r < std::min(B, (it + 1) * 16); r++)
        if (rep_b[b_of[r]] != r && !same_box(r, rep_b[b_of[r]])) mismatch++;
    });
    if (mismatch.load()) group_rows(B, key, same_box, b_of, rep_b);
  }
}
// Host-only test hook (no device needed): the grouping icl_upload applies to the box rows of an affinity batch.
// group_of[r] = group of row r in order of first appearance; returns the number of groups through n_groups.
extern "C" int icl_group_rows(const void* rows, int64_t n_rows, int64_t row_bytes, int32_t* group_of, int32_t* n_groups) {
  if (!rows || !group_of || !n_groups || n_rows < 0 || row_bytes <= 0) return fail("icl_group_rows: bad argument");
  std::vector<int> rep;
  std::vector<int> of((size_t)n_rows);
  group_box_rows((int)n_rows, (const char*)rows, (size_t)row_bytes, nullptr, nullptr, 0, of.data(), rep);
  for (int64_t r = 0; r < n_rows; r++) group_of[r] = of[r];
  *n_groups = (int32_t)rep.size();
  return 0;
}

static void slot_plan(const icl_head_config& c, std::vector<int>& kinds /*index id or -1..-3*/) {
  // nn_utils/core.py:377-433.  -1 feats, -2 box, -3 bfeats
  kinds = {ICL_FIRST_I_BW, ICL_LAST_I_FW};
  if (c.encoding == ICL_ENC_FIRST_LAST_SENTENCE) { kinds.push_back(ICL_SENT_LAST_I_FW); kinds.push_back(ICL_SENT_FIRST_I_BW); }
  else { kinds.push_back(ICL_FIRST_I_FW); kinds.push_back(ICL_LAST_I_BW); }
  bool rel = c.task == ICL_TASK_REL_INTRA || c.task == ICL_TASK_REL_CROSS;
  if (rel) {
    kinds.push_back(ICL_FIRST_J_BW); kinds.push_back(ICL_LAST_J_FW); kinds.push_back(-1);
    if (c.encoding == ICL_ENC_FIRST_LAST_SENTENCE && c.task == ICL_TASK_REL_CROSS) { kinds.push_back(ICL_SENT_LAST_J_FW); kinds.push_back(ICL_SENT_FIRST_J_BW); }
    else if (c.encoding == ICL_ENC_FIRST_LAST_MENTION) { kinds.push_back(ICL_FIRST_J_FW); kinds.push_back(ICL_LAST_J_BW); }
  } else {
    kinds.push_back(-1);
    if (c.task == ICL_TASK_AFFINITY) { kinds.push_back(-2); if (c.n_box_feats > 0) kinds.push_back(-3); }
  }
}

static int add_param(icl_model* m, const std::string& name, int rows, int cols) {
  Param p{name, rows, cols, m->n_params};
  m->n_params += ((int64_t)rows * cols + 3) / 4 * 4;        // keep every tensor 16-byte aligned
  m->pindex[name] = (int)m->params.size();
  m->params.push_back(p);
  return (int)m->params.size() - 1;
}

template <typename T> static cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }

// ----------------------------------------------------------------------------- zero fills on the SMs
// cudaMemsetAsync may be executed by a copy engine: with the next batch's H2D transfer in flight on the copy stream the
// compute stream's memsets queued behind 4 MB chunks (measured: a 1.77 ms step stretched to 2.9 ms).  These run as kernels.
__global__ void k_zero_words(uint32_t* __restrict__ p, size_t n_words) {
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    uint4* p4 = reinterpret_cast<uint4*>(p);
    const size_t n4 = n_words >> 2;
    for (size_t i = i0; i < n4; i += stride) p4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (n4 << 2) + i0; i < n_words; i += stride) p[i] = 0u;
  } else {
    for (size_t i = i0; i < n_words; i += stride) p[i] = 0u;
  }
}
__global__ void k_zero_2d(float* __restrict__ p0, float* __restrict__ p1, long pitch, int width, long rows) {   // floats, % 4 == 0
  float* p = blockIdx.y ? p1 : p0;
  const long w4 = width >> 2, n = rows * w4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    *reinterpret_cast<float4*>(p + (i / w4) * pitch + (i % w4) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}
// several zero fills in ONE launch (blockIdx.y = range): the backward pass needs five cleared buffers before its first kernel, and
// a launch costs more than clearing 9 MB
struct ZeroRanges { uint32_t* p[6]; size_t n[6]; };
__global__ void k_zero_multi(const ZeroRanges r) {
  uint32_t* p = r.p[blockIdx.y];
  const size_t n_words = r.n[blockIdx.y];
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    uint4* p4 = reinterpret_cast<uint4*>(p);
    const size_t n4 = n_words >> 2;
    for (size_t i = i0; i < n4; i += stride) __stcs(p4 + i, make_uint4(0u, 0u, 0u, 0u));      // streaming: the zeros are read once, much later
    for (size_t i = (n4 << 2) + i0; i < n_words; i += stride) p[i] = 0u;
  } else {
    for (size_t i = i0; i < n_words; i += stride) p[i] = 0u;
  }
}
// max_blocks > 0: a deliberately NARROW launch (that many CTAs in all) for fills that run beside a latency-bound persistent kernel:
// they then live on the SMs that kernel leaves idle instead of taking issue slots and L2 from it
static cudaError_t zero_multi_async(std::initializer_list<std::pair<void*, size_t>> ranges, cudaStream_t st, int max_blocks = 0) {
  ZeroRanges r; int n = 0; size_t mx = 0;
  for (auto& pr : ranges) if (pr.second) { r.p[n] = reinterpret_cast<uint32_t*>(pr.first); r.n[n] = pr.second / 4; mx = std::max(mx, r.n[n]); n++; }
  if (n == 0) return cudaSuccess;
  unsigned blocks = (unsigned)std::min<size_t>(296, (mx / 4 + 255) / 256 + 1);
  if (max_blocks > 0) blocks = std::max(1u, std::min(blocks, (unsigned)max_blocks / (unsigned)n));
  k_zero_multi<<<dim3(blocks, n), max_blocks > 0 ? 512 : 256, 0, st>>>(r);
  return cudaGetLastError();
}
__global__ void k_fill_u16(uint16_t* __restrict__ p, size_t n, uint16_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
static cudaError_t zero_async(void* p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  const size_t words = bytes / 4;
  const unsigned blocks = (unsigned)std::min<size_t>(1184, (words / 4 + 255) / 256 + 1);
  k_zero_words<<<blocks, 256, 0, st>>>(reinterpret_cast<uint32_t*>(p), words);
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------- GEMM dispatch
// splits > 1: split-K over gridDim.z with a red.global epilogue (plain epilogue only; C is zeroed here)
// prezeroed: C already holds zeros (split-K accumulates into it); pdl: programmatic dependent launch
static int gemm(icl_model* m, cudaStream_t st, bool a_mn, bool b_mn, const GemmArgs& g, int force_mode = -1, int splits = 1,
                bool prezeroed = false, bool pdl = false) {
  int mode = force_mode >= 0 ? force_mode : m->cfg.gemm_mode;
  if (g.M <= 0 || g.N <= 0) return 0;
  if (splits == 0) {     // auto: few output tiles + a long contraction -> split K so that about one wave of CTAs shares it
    const int BN = g.N >= 512 ? 256 : 128, tiles = ((g.M + 127) / 128) * ((g.N + BN - 1) / BN), kb = (g.K + 31) / 32;
    splits = (tiles >= 100 || kb < 16 || g.epi.mode != EPI_PLAIN || g.epi.bias || g.ldc != g.N) ? 1
             : std::max(1, std::min(std::min(16, 148 / tiles), kb / 8));
  }
  if (mode == ICL_GEMM_TCGEN05_TF32 && tcgen05_gemm_supported(g, a_mn, b_mn)) {
    if (splits > 1) {
      if (g.ldc != g.N) return fail("split-K gemm needs a dense C");
      if (!prezeroed) CK(zero_async(g.C, (size_t)g.M * g.N * 4, st));
    }
    int r = tcgen05_gemm_launch(m->tma, st, a_mn, b_mn, g, splits, pdl);
    if (r != 0) return fail("tcgen05 gemm launch failed (%d): %s", r, cudaGetErrorString(cudaGetLastError()));
    m->launches++;
    return 0;
  }
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64);
  if (!a_mn && !b_mn) k_gemm_simt<true, true><<<grid, 256, 0, st>>>(g);
  else if (!a_mn && b_mn) k_gemm_simt<true, false><<<grid, 256, 0, st>>>(g);
  else if (a_mn && b_mn) k_gemm_simt<false, false><<<grid, 256, 0, st>>>(g);
  else k_gemm_simt<false, true><<<grid, 256, 0, st>>>(g);
  m->launches++;
  CK(cudaGetLastError());
  return 0;
}

static GemmArgs mk_gemm(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, int K) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.epi.drop.keep = 1.0f;
  return g;
}

// ----------------------------------------------------------------------------- persistent recurrent kernels: host side
static int box_map(icl_model* m, const float* ptr, uint64_t cols, uint64_t rows, uint32_t box_c, uint32_t box_r, int swizzle,
                   CUtensorMap* out, uint64_t ld = 0) {
  int r = m->tma.get(ptr, cols, rows, ld ? ld : cols, box_c, box_r, swizzle, out);
  return r ? fail("cuTensorMapEncodeTiled failed (%d) for a {%u,%u} box over [%llu,%llu]", r, box_c, box_r,
                  (unsigned long long)rows, (unsigned long long)cols) : 0;
}
#ifdef ICL_EXPERIMENTS
template <int U> static int rec_set_attr(int nkb) {
  cudaError_t e = cudaFuncSetAttribute(k_rec_fwd<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, rec_fwd_smem<U>(nkb));
  return e == cudaSuccess ? 0 : fail("cudaFuncSetAttribute(k_rec_fwd): %s", cudaGetErrorString(e));
}
static int rec_init(icl_model* m) {
  const int H = m->H, U = m->rp_U;
  const uint64_t RC = (uint64_t)m->rows_cap;
  const int NONE = (int)CU_TENSOR_MAP_SWIZZLE_NONE, SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  RecFwdMaps& f = m->rp_fmaps;
  for (int d = 0; d < 2; d++) {
    CKI(box_map(m, m->Hp[d], H, RC, 32, 128, SW128, &f.a[d], m->ldx));
    CKI(box_map(m, m->Wp[d], (uint64_t)m->rp_nkb * 32, (uint64_t)m->rp_nsl * 4 * U, 32, 4 * U, SW128, &f.w[d]));
    CKI(box_map(m, m->Z[d], 4 * H, RC, U, 32, NONE, &f.z[d]));
    CKI(box_map(m, m->Cc[d], H, RC, U, 32, NONE, &f.cc[d]));
    CKI(box_map(m, m->Hx[d], H, RC, U, 32, NONE, &f.hx[d]));
    CKI(box_map(m, m->Hp[d], H, RC, U, 32, NONE, &f.hp[d], m->ldx));
  }
  for (int d = 0; d < 2; d++) {
    CKI(box_map(m, m->Z[d], 4 * H, RC, 32, 128, SW128, &m->rp_bmaps.za[d]));
    const float* Whh = m->Pr + m->params[m->pK[d]].off + (size_t)m->E * 4 * H;
    CKI(box_map(m, Whh, 4 * H, H, 32, 128, SW128, &m->rp_bmaps.wb[d]));
  }
  if (cudaFuncSetAttribute(k_rec_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_rec_bwd) failed");
  return U == 20 ? rec_set_attr<20>(m->rp_nkb) : rec_set_attr<16>(m->rp_nkb);
}

#else
static int rec_init(icl_model*) { return 0; }
#endif

static int k1_init(icl_model* m) {
  m->k1_f16 = m->cfg.gemm_mode == ICL_GEMM_TCGEN05_TF32;
  if (const char* e = getenv("ICL_K1_FP16")) m->k1_f16 = m->k1_f16 && atoi(e) != 0;
  if (!m->k1_f16) return 0;
  const int E = m->E, H = m->H, SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  m->k1_Kp = (E + 63) / 64 * 64;
  const uint64_t RC = (uint64_t)m->rows_cap;
  for (int d = 0; d < 2; d++) {
    CK(cudaMalloc((void**)&m->X16[d], RC * m->k1_Kp * 2)); CK(cudaMemset(m->X16[d], 0, RC * m->k1_Kp * 2));
    CK(cudaMalloc((void**)&m->Wih16[d], (size_t)4 * H * m->k1_Kp * 2));
    int r = TmaCache::get16(m->X16[d], (uint64_t)m->k1_Kp, RC, (uint64_t)m->k1_Kp, 64, 128, SW128, &m->k1_ta[d]);
    if (!r) r = TmaCache::get16(m->Wih16[d], (uint64_t)m->k1_Kp, (uint64_t)4 * H, (uint64_t)m->k1_Kp, 64, 256, SW128, &m->k1_tb[d]);
    if (r) return fail("cuTensorMapEncodeTiled failed (%d) for the fp16 input-projection maps", r);
    CKI(box_map(m, m->Z[d], 4 * H, RC, 32, 32, SW128, &m->k1_tc[d]));
  }
  return 0;
}

template <int U> static int rec16_set_attr(int nkb) {
  cudaError_t e = cudaFuncSetAttribute(k_rec_fwd16<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, rec_fwd16_smem<U>(nkb));
  return e == cudaSuccess ? 0 : fail("cudaFuncSetAttribute(k_rec_fwd16): %s", cudaGetErrorString(e));
}
static int rec16_setup(icl_model* m, icl_model::RfVar& v, int U) {
  const int H = m->H;
  v.U = U; v.UP = U == 20 ? RF<20>::UP : RF<16>::UP; v.maxtpc = U == 20 ? RF<20>::MAXTPC : RF<16>::MAXTPC;
  v.nsl = (H + U - 1) / U;
  v.KP = (v.nsl * v.UP + 63) / 64 * 64;
  v.nkb = v.KP / 64;
  v.nk16 = (v.nsl * v.UP + 15) / 16;
  const size_t smem = U == 20 ? rec_fwd16_smem<20>(v.nkb) : rec_fwd16_smem<16>(v.nkb);
  if (smem > 227 * 1024 || v.nkb > 10) return 0;                     // stays off
  const uint64_t RC = (uint64_t)m->rows_cap;
  const int NONE = (int)CU_TENSOR_MAP_SWIZZLE_NONE, SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B, SW64 = (int)CU_TENSOR_MAP_SWIZZLE_64B;
  for (int d = 0; d < 2; d++) {
    CK(cudaMalloc((void**)&v.Hp16[d], RC * v.KP * 2)); CK(cudaMemset(v.Hp16[d], 0, RC * v.KP * 2));
    const size_t wn = (size_t)v.nsl * 4 * U * v.KP;
    CK(cudaMalloc((void**)&v.Wp16[d], wn * 2)); CK(cudaMemset(v.Wp16[d], 0, wn * 2));
    int r = TmaCache::get16(v.Hp16[d], (uint64_t)v.KP, RC, (uint64_t)v.KP, 64, 128, SW128, &v.maps.a[d]);
    if (!r) r = TmaCache::get16(v.Wp16[d], (uint64_t)v.KP, (uint64_t)v.nsl * 4 * U, (uint64_t)v.KP, 64, 4 * U, SW128, &v.maps.w[d]);
    if (!r) r = TmaCache::get16(v.Hp16[d], (uint64_t)v.KP, RC, (uint64_t)v.KP, v.UP, 32, NONE, &v.maps.hp16[d]);
    if (r) return fail("cuTensorMapEncodeTiled failed (%d) for the fp16 recurrence maps", r);
    if (U == 16) {      // [rows][4 gates][H] as a 3-D tensor / [rows][H]: boxes of the last slice are clipped at H; 64-byte rows, swizzled
      r = TmaCache::get3(m->Z[d], (uint64_t)H, 4, RC, (uint64_t)H * 4, (uint64_t)4 * H * 4, 16, 1, 32, SW64, &v.maps.z[d]);
      if (r) return fail("cuTensorMapEncodeTiled failed (%d) for the 3-D gate map", r);
      CKI(box_map(m, m->Cc[d], H, RC, 16, 32, SW64, &v.maps.cc[d]));
    } else {
      CKI(box_map(m, m->Z[d], 4 * H, RC, U, 32, NONE, &v.maps.z[d]));
      CKI(box_map(m, m->Cc[d], H, RC, U, 32, NONE, &v.maps.cc[d]));
    }
  }
  CKI(U == 20 ? rec16_set_attr<20>(v.nkb) : rec16_set_attr<16>(v.nkb));
  v.on = true;
  return 0;
}
static int rec16_init(icl_model* m) {
  m->rf_on = m->rp_U != 0;
  if (const char* e = getenv("ICL_REC_FP16")) m->rf_on = m->rf_on && atoi(e) != 0;
  if (!m->rf_on) return 0;
  if (m->H % 20 == 0) CKI(rec16_setup(m, m->rf20, 20));
  if (m->H % 4 == 0) CKI(rec16_setup(m, m->rf16, 16));
  m->rf_on = m->rf20.on || m->rf16.on;
  return 0;
}
// the slicing for this batch: U = 20 where H allows it, else U = 16 (any H; the whole h tile fits its ring).  On card2048 the two
// measure the same (0.314 - 0.330 vs 0.319 - 0.322 ms, profiles/r2j_*): ICL_RF_U=16|20 forces one for A/B runs
static icl_model::RfVar* rf_pick(icl_model* m) {
  if (!m->rf_on || !m->rp_on || m->Tmax > RP_MAXT) return nullptr;
  const int tiles = (m->n_active[0] + 127) / 128;
  auto fits = [&](icl_model::RfVar& v) {
    if (!v.on) return false;
    const int P = std::max(1, std::min(148 / (2 * v.nsl), tiles));
    return (tiles + P - 1) / P <= v.maxtpc;
  };
  static const int force = getenv("ICL_RF_U") ? atoi(getenv("ICL_RF_U")) : 0;      // A/B: 16 or 20
  if (force == 20 && fits(m->rf20)) return &m->rf20;
  if (force == 16 && fits(m->rf16)) return &m->rf16;
  if (fits(m->rf20)) return &m->rf20;
  if (fits(m->rf16)) return &m->rf16;
  return nullptr;
}

// make input set s the "current" one: the device pointers every launch site reads
static void use_input_set(icl_model* m, int s) {
  InSet& I = m->in[s];
  m->xraw = I.dev<float>(I.o_x); m->d_off = I.dev<int>(I.o_off); m->d_nact = I.dev<int>(I.o_nact); m->d_rank = I.dev<int>(I.o_rank);
  m->d_lens = I.dev<int>(I.o_lens); m->d_tokseq = I.dev<int>(I.o_tokseq); m->d_tokstart = I.dev<int>(I.o_tokstart);
  m->d_tokrow = I.dev<int>(I.o_tokrow);
  for (Head& h : m->heads) {
    HeadIn& hin = h.in[s];
    for (int i = 0; i < ICL_N_INDEX; i++) h.idx[i] = I.dev<int>(hin.idx[i]);
    h.feats = I.dev<float>(hin.feats); h.box = I.dev<float>(hin.box); h.bfeats = I.dev<float>(hin.bfeats);
    h.labels = I.dev<float>(hin.labels);
    for (int i = 0; i < h.slots.n_slots; i++) {
      int id = h.slot_index_id[i];
      h.slots.rowidx[i] = nullptr;
      if (id >= 0) h.slots.idx[i] = h.idx[id];
      else h.slots.dense[i] = id == -1 ? h.feats : id == -2 ? h.box : h.bfeats;
    }
    h.d_boxrow = I.dev<int>(hin.boxrow);
    if (h.fact_ok) {
      h.d_m_of = I.dev<int>(hin.m_of); h.d_b_of = I.dev<int>(hin.b_of); h.d_rep_m = I.dev<int>(hin.rep_m);
      h.d_rep_boxrow = I.dev<int>(hin.rep_boxrow); h.d_m_start = I.dev<int>(hin.m_start); h.d_m_mem = I.dev<int>(hin.m_mem);
      h.d_b_start = I.dev<int>(hin.b_start); h.d_b_mem = I.dev<int>(hin.b_mem);
    }
  }
}
// the mention / box halves of a factorised head's slot table for the resident batch (after icl_upload has settled the box source)
static void fact_slots(Head& h, const float* box_table) {
  for (size_t j = 0; j < h.slotm_src.size(); j++) {
    const int i = h.slotm_src[j];
    h.slots_m.idx[j] = h.slots.idx[i]; h.slots_m.dense[j] = h.slots.dense[i]; h.slots_m.rowidx[j] = nullptr;
  }
  h.slots_m.rowmap = h.Mu == h.c.batch_size ? nullptr : h.d_rep_m;     // row g = the group's first pair: its index rows, its m_feats row
  for (size_t j = 0; j < h.slotb_src.size(); j++) {
    const int i = h.slotb_src[j];
    const bool table = h.slot_index_id[i] == -2 && !h.box_compact;
    h.slots_b.dense[j] = table ? box_table : h.slots.dense[i];           // compact blocks: row g = group g
    h.slots_b.rowidx[j] = table ? h.d_rep_boxrow : nullptr;
  }
  h.slots_b.rowmap = nullptr;
}

static int bptt_init(icl_model* m) {
  m->bp_on = m->cfg.gemm_mode == ICL_GEMM_TCGEN05_TF32;
  if (const char* e = getenv("ICL_BPTT_FUSED")) m->bp_on = m->bp_on && atoi(e) != 0;
  if (const char* e = getenv("ICL_BPTT_CS")) { int c = atoi(e); if (c == 1 || c == 2 || c == 4) m->bp_cs = c; }
  if (const char* e = getenv("ICL_BPTT_MODE")) m->bp_cluster = strcmp(e, "step") != 0;
  if (m->H > BC_NACC * BP_BN) m->bp_cluster = false;
  if (!m->bp_on) return 0;
  const int H = m->H, SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  for (int d = 0; d < 2; d++) {
    CKI(box_map(m, m->Z[d], 4 * H, (uint64_t)m->rows_cap, 32, 128, SW128, &m->bp_maps.za[d]));
    const float* Whh = m->Pr + m->params[m->pK[d]].off + (size_t)m->E * 4 * H;
    CKI(box_map(m, Whh, 4 * H, H, 32, BP_BN, SW128, &m->bp_maps.wb[d]));
  }
  if (cudaFuncSetAttribute(k_bptt_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BP_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_bptt_step<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BP_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_bptt_step<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BP_SMEM) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_bptt_step) failed");
  if (cudaFuncSetAttribute(k_bptt_cluster<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_bptt_cluster<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_bptt_cluster<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_bptt_cluster) failed");
  return 0;
}


#ifdef ICL_EXPERIMENTS
template <int UN> static int ns_setup(icl_model* m) {
  using C = BNS<UN>;
  const int SW128 = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  const uint64_t RC = (uint64_t)m->rows_cap;
  for (int d = 0; d < 2; d++) {
    CK(cudaMalloc((void**)&m->dZ16[d], RC * C::KTOT * 2)); CK(cudaMemset(m->dZ16[d], 0, RC * C::KTOT * 2));
    CK(cudaMalloc((void**)&m->S16[d], RC * BN_CS * 4)); CK(cudaMemset(m->S16[d], 0, RC * BN_CS * 4));
    CK(cudaMalloc((void**)&m->Wb16[d], (size_t)BN_CS * C::NP * C::KTOT * 2));
    int r = TmaCache::get16(m->dZ16[d], (uint64_t)C::KTOT, RC, (uint64_t)C::KTOT, 64, 128, SW128, &m->ns_maps.a[d]);
    if (!r) r = TmaCache::get16(m->Wb16[d], (uint64_t)C::KTOT, (uint64_t)BN_CS * C::NP, (uint64_t)C::KTOT, 64, C::NP, SW128, &m->ns_maps.w[d]);
    if (r) return fail("cuTensorMapEncodeTiled failed (k_bptt_nsplit)");
  }
  if (cudaFuncSetAttribute(k_bptt_nsplit<UN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_bptt_nsplit<UN>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
    return fail("cudaFuncSetAttribute(k_bptt_nsplit) failed");
  return 0;
}
static int ns_init(icl_model* m) {
  const int H = m->H;
  m->ns_on = m->bp_on && H % 4 == 0 && H <= 320 && getenv("ICL_BPTT_NSPLIT") && atoi(getenv("ICL_BPTT_NSPLIT")) != 0;
  if (!m->ns_on) return 0;
  m->ns_UN = H <= 64 ? 8 : H <= 128 ? 16 : H <= 224 ? 28 : 40;
  switch (m->ns_UN) {
    case 8: return ns_setup<8>(m);
    case 16: return ns_setup<16>(m);
    case 28: return ns_setup<28>(m);
    default: return ns_setup<40>(m);
  }
}

#else
static int ns_init(icl_model*) { return 0; }
#endif

// ----------------------------------------------------------------------------- API
extern "C" const char* icl_last_error(void) { return g_err; }
extern "C" int icl_version(void) { return 2; }

extern "C" void icl_destroy(icl_model* m) {
  if (!m) return;
  cudaDeviceSynchronize();
  auto F = [](void* p) { if (p) cudaFree(p); };
  F(m->P); if (!m->G_external) F(m->G); F(m->Pr); F(m->tok_table); F(m->box_table);
  if (m->slots.empty()) { F(m->M); F(m->V); }
  else {                       // m->M / m->V alias the live slot
    m->slots[m->cur_slot].M = m->M; m->slots[m->cur_slot].V = m->V;
    for (auto& o : m->slots) { F(o.M); F(o.V); }
  }
  for (InSet& I : m->in) {
    F(I.d_blob);
    if (I.h_blob) cudaFreeHost(I.h_blob);
    for (cudaEvent_t e : {I.ev_copied, I.ev_done, I.ev_stats, I.ev_c0, I.ev_s0}) if (e) cudaEventDestroy(e);
  }
  if (m->h_stats) cudaFreeHost(m->h_stats);
  if (m->copy) cudaStreamDestroy(m->copy);
  for (int d = 0; d < 2; d++) {
    F(m->XH[d]); F(m->Z[d]); F(m->Hx[d]); F(m->Cc[d]); F(m->dHout[d]); F(m->dhrec[d]); F(m->dcc[d]); F(m->R[d]);
  }
  F(m->Wp[0]); F(m->Wp[1]); F(m->rp_flags); F(m->rp_trace); F(m->rp_bar);
  for (auto* v : {&m->rf20, &m->rf16}) { F(v->Hp16[0]); F(v->Hp16[1]); F(v->Wp16[0]); F(v->Wp16[1]); } F(m->dZ16[0]); F(m->dZ16[1]); F(m->S16[0]); F(m->S16[1]); F(m->Wb16[0]); F(m->Wb16[1]); F(m->X16[0]); F(m->X16[1]); F(m->Wih16[0]); F(m->Wih16[1]);
  F(m->d_partial); F(m->d_gnorm);
  for (auto& h : m->heads) {
    F(h.bi); F(h.dbi); F(h.dA); F(h.dBuf); for (auto a : h.act) F(a); for (auto a : h.dzb) F(a);
    F(h.proba); F(h.dlogits); F(h.smb_part); F(h.row_loss); F(h.row_correct); F(h.scalars); F(h.pred);
    if (h.h_out) cudaFreeHost(h.h_out);
    if (h.h_pred) cudaFreeHost(h.h_pred);
    F(h.Xm); F(h.Xb); F(h.U); F(h.V); F(h.dU); F(h.dV); F(h.dXm);
    if (h.sb) cudaStreamDestroy(h.sb);
    for (cudaEvent_t e : {h.ev_bfork, h.ev_box, h.ev_du}) if (e) cudaEventDestroy(e);
    if (h.hs) cudaStreamDestroy(h.hs);
    if (h.ha) cudaStreamDestroy(h.ha);
    if (h.ev_done) cudaEventDestroy(h.ev_done);
    if (h.ev_adone) cudaEventDestroy(h.ev_adone);
  }
  if (m->aux) cudaStreamDestroy(m->aux);
  if (m->aux2) cudaStreamDestroy(m->aux2);
  if (m->aux3) cudaStreamDestroy(m->aux3);
  if (m->aux4) cudaStreamDestroy(m->aux4);
  for (cudaEvent_t e : {m->ev_fork, m->ev_join, m->ev_join2, m->ev_join3, m->ev_t0, m->ev_t1, m->ev_heads, m->ev_wg0, m->ev_packs, m->ev_side, m->ev_loss, m->ev_hfork, m->ev_zfork, m->ev_zero, m->ev_dhzero}) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev_dz) if (e) cudaEventDestroy(e);
  for (int i = 0; i < PH_N; i++) for (int j = 0; j < 2; j++) if (m->ev_ph[i][j]) cudaEventDestroy(m->ev_ph[i][j]);
  delete m;
}

extern "C" int icl_create(const icl_config* cfg, icl_model** out) {
  if (!cfg || !out) return fail("icl_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("icl_create: no CUDA device -- this library has no CPU fallback");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail("icl_create: device sm_%d%d is not Blackwell sm_100 (kernels are sm_100a-only)", prop.major, prop.minor);
  if (cfg->embed_width % 4 || cfg->lstm_hidden % 4) return fail("icl_create: embed_width and lstm_hidden must be multiples of 4");
  if (cfg->n_heads < 1 || cfg->n_heads > ICL_MAX_HEADS) return fail("icl_create: n_heads out of range");
  if (cfg->max_seqs < 1 || cfg->max_seq_len < 1) return fail("icl_create: max_seqs / max_seq_len must be positive");
  icl_model* m = new icl_model();
  m->cfg = *cfg;
  if (m->cfg.beta1 <= 0) m->cfg.beta1 = 0.9f;
  if (m->cfg.beta2 <= 0) m->cfg.beta2 = 0.999f;
  m->round_ops = cfg->gemm_mode == ICL_GEMM_TCGEN05_TF32;
  int E = m->E = cfg->embed_width, H = m->H = cfg->lstm_hidden;
  m->S_cap = cfg->max_seqs; m->T_cap = cfg->max_seq_len;
  m->Ntok_cap = (long)m->S_cap * m->T_cap;
  m->rows_cap = (long)m->T_cap * ((m->S_cap + 127) / 128 * 128);      // every step block is padded to 128 rows
  // parameters, named like the TF variables
  const char* dn[2] = {"fw", "bw"};
  for (int d = 0; d < 2; d++) {
    std::string base = std::string("bidirectional_lstm/bidirectional_rnn/") + dn[d] + "/basic_lstm_cell/";
    m->pK[d] = add_param(m, base + "kernel", E + H, 4 * H);
    m->pBias[d] = add_param(m, base + "bias", 1, 4 * H);
  }
  m->heads.resize(cfg->n_heads);
  for (int hi = 0; hi < cfg->n_heads; hi++) {
    Head& h = m->heads[hi];
    h.c = cfg->heads[hi];
    if (h.c.n_hidden < 1 || h.c.n_hidden > ICL_MAX_LAYERS) { icl_destroy(m); return fail("head %d: n_hidden out of range", hi); }
    if (h.c.n_classes < 2 || h.c.n_classes > 32) { icl_destroy(m); return fail("head %d: n_classes must be in [2,32]", hi); }
    std::vector<int> plan;
    slot_plan(h.c, plan);
    memset(&h.slots, 0, sizeof(h.slots));
    int col = 0;
    for (size_t i = 0; i < plan.size(); i++) {
      int w = plan[i] >= 0 ? H : plan[i] == -1 ? h.c.n_feats : plan[i] == -2 ? h.c.box_width : h.c.n_box_feats;
      h.slots.kind[i] = plan[i] >= 0 ? 0 : 1;
      h.slots.col[i] = col;
      h.slots.width[i] = w;
      col += w;
      if (plan[i] >= 0) h.D0g = col;
    }
    h.slots.n_slots = (int)plan.size();
    h.slot_index_id = plan;
    h.D0 = col;
    h.ldbi = (h.D0 + 3) / 4 * 4; h.lddbi = (std::max(h.D0g, 1) + 3) / 4 * 4;
    {   // mention / box halves of the column plan
      int aff_mode = 1;
      if (const char* e = getenv("ICL_AFF_FACTOR")) aff_mode = atoi(e);
      h.fact_ok = aff_mode != 0 && h.c.task == ICL_TASK_AFFINITY && h.c.box_width > 0;
      memset(&h.slots_m, 0, sizeof(h.slots_m)); memset(&h.slots_b, 0, sizeof(h.slots_b));
      if (h.fact_ok) {
        for (size_t i = 0; i < plan.size(); i++) {
          const bool boxside = plan[i] == -2 || plan[i] == -3;
          SlotTable& t = boxside ? h.slots_b : h.slots_m;
          if (boxside && h.slots_b.n_slots == 0) h.Dm = h.slots.col[i];
          const int j = t.n_slots++;
          t.kind[j] = h.slots.kind[i]; t.width[j] = h.slots.width[i]; t.col[j] = h.slots.col[i] - (boxside ? h.Dm : 0);
          (boxside ? h.slotb_src : h.slotm_src).push_back((int)i);
        }
        h.Db = h.D0 - h.Dm;
        h.ldm = (h.Dm + 3) / 4 * 4; h.ldb = (h.Db + 3) / 4 * 4;
        // box columns must be the tail of the plan (they are: core.py:421-433) and the W1 row split 16-byte aligned
        if (h.Dm <= 0 || h.Db <= 0 || ((int64_t)h.Dm * h.c.widths[0]) % 4 != 0) h.fact_ok = false;
        for (size_t i = 0; i + 1 < plan.size(); i++) if ((plan[i] == -2 || plan[i] == -3) && plan[i + 1] >= -1) h.fact_ok = false;
      }
    }
    h.dims.push_back(h.D0);
    for (int k = 0; k < h.c.n_hidden; k++) h.dims.push_back(h.c.widths[k]);
    h.dims.push_back(h.c.n_classes);
    // multitask heads: setup_ffw / the softmax layer open tf.variable_scope(scope_name + "hdn_k") with scope_name = "<task>/" while
    // ALREADY inside `with tf.variable_scope(task)` (core.py:166-172,497 under icl_multitask_lstm.py:62), and TF nests a string
    // scope under the current one: the variables of a checkpoint are "<task>/<task>/hdn_k/Variable"
    std::string pre = h.c.scope[0] ? std::string(h.c.scope) + "/" + std::string(h.c.scope) + "/" : "";
    for (int k = 0; k <= h.c.n_hidden; k++) {
      std::string sc = k < h.c.n_hidden ? pre + "hdn_" + std::to_string(k + 1) : pre + "softmax";
      h.pW.push_back(add_param(m, sc + "/Variable", h.dims[k], h.dims[k + 1]));
      h.pB.push_back(add_param(m, sc + "/Variable_1", 1, h.dims[k + 1]));
    }
  }
  *out = m;   // from here on failures leave a destroyable handle
#define CKD(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail("icl_create: %s -> %s", #x, cudaGetErrorString(e_)); icl_destroy(m); *out = nullptr; return -1; } } while (0)
  size_t np = (size_t)m->n_params;
  CKD(dmalloc(&m->P, np)); CKD(dmalloc(&m->G, np)); CKD(dmalloc(&m->M, np)); CKD(dmalloc(&m->V, np)); CKD(dmalloc(&m->Pr, np));
  CKD(cudaMemset(m->P, 0, np * 4)); CKD(cudaMemset(m->G, 0, np * 4)); CKD(cudaMemset(m->M, 0, np * 4)); CKD(cudaMemset(m->V, 0, np * 4));
  CKD(cudaMemset(m->Pr, 0, np * 4));
  const size_t RC = (size_t)m->rows_cap, SP = (size_t)(m->S_cap + 127) / 128 * 128;
  // zero-filled once: rows past nact[k] of a step block only ever hold finite don't-care values
#define ZALLOC(p, n) do { CKD(dmalloc(&(p), (n))); CKD(cudaMemset((p), 0, (n) * 4)); } while (0)
  for (int d = 0; d < 2; d++) {
    m->ldx = E + H + 4;
    ZALLOC(m->XH[d], RC * m->ldx); m->xd[d] = m->XH[d]; m->Hp[d] = m->XH[d] + E;
    ZALLOC(m->Z[d], RC * 4 * H); ZALLOC(m->Hx[d], RC * H);
    ZALLOC(m->Cc[d], RC * H); ZALLOC(m->dHout[d], RC * H); ZALLOC(m->dhrec[d], SP * H); ZALLOC(m->dcc[d], SP * H);
    ZALLOC(m->R[d], SP * 4 * H);
  }
  m->rp_U = (H % 20 == 0) ? 20 : (H % 16 == 0) ? 16 : 0;
  m->rp_nkb = (H + 31) / 32;
  if (m->rp_U && (m->rp_nkb > 10 || m->T_cap > RP_MAXT)) m->rp_U = 0;
  if (const char* e = getenv("ICL_PERSISTENT")) m->rp_on = atoi(e) != 0;
  if (const char* e = getenv("ICL_PERSISTENT_BWD")) m->rp_bwd_on = atoi(e) != 0;
  m->n_sms = prop.multiProcessorCount;
  if (m->cfg.gemm_mode != ICL_GEMM_TCGEN05_TF32) m->rp_U = 0;      // the fp32 validation mode keeps the per-step SIMT path
  if (m->rp_U) {
    m->rp_nsl = H / m->rp_U;
    m->rp_max_tiles = (int)(SP / 128);
    for (int d = 0; d < 2; d++) ZALLOC(m->Wp[d], (size_t)m->rp_nsl * 4 * m->rp_U * m->rp_nkb * 32);
    CKD(dmalloc(&m->rp_flags, (size_t)2 * m->rp_max_tiles));
    CKD(dmalloc(&m->rp_bar, 4));
  }
#undef ZALLOC
  for (InSet& I : m->in) {
    size_t o = 0;
    auto take = [&o](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    I.o_x = take((size_t)m->Ntok_cap * E * 4);
    I.meta_off = o;
    I.o_lens = take((size_t)m->S_cap * 4); I.o_rank = take((size_t)m->S_cap * 4); I.o_tokstart = take((size_t)m->S_cap * 4);
    I.o_off = take((size_t)(m->T_cap + 1) * 4); I.o_nact = take((size_t)(m->T_cap + 1) * 4);
    const int si = (int)(&I - m->in);
    for (Head& h : m->heads) {
      HeadIn& hin = h.in[si];
      const size_t B = h.c.batch_size;
      for (int i = 0; i < ICL_N_INDEX; i++) hin.idx[i] = take(B * 3 * 4);
      hin.feats = take(B * h.c.n_feats * 4); hin.box = take(B * h.c.box_width * 4); hin.bfeats = take(B * h.c.n_box_feats * 4);
      hin.labels = take(B * h.c.n_classes * 4);
      hin.boxrow = take(B * 4);
      if (h.fact_ok) {
        hin.m_of = take(B * 4); hin.b_of = take(B * 4); hin.rep_m = take(B * 4); hin.rep_boxrow = take(B * 4);
        hin.m_start = take((B + 1) * 4); hin.m_mem = take(B * 4); hin.b_start = take((B + 1) * 4); hin.b_mem = take(B * 4);
      }
    }
    I.o_tokrow = take((size_t)m->Ntok_cap * 4);
    I.o_tokseq = take((size_t)m->Ntok_cap * 4);              // last: only its used prefix is copied
    I.bytes = o;
    CKD(cudaMalloc((void**)&I.d_blob, I.bytes));
    CKD(cudaMallocHost((void**)&I.h_blob, I.bytes));
    CKD(cudaEventCreate(&I.ev_copied)); CKD(cudaEventCreate(&I.ev_done)); CKD(cudaEventCreate(&I.ev_c0)); CKD(cudaEventCreate(&I.ev_s0));
    CKD(cudaEventCreateWithFlags(&I.ev_stats, cudaEventDisableTiming));
  }
  CKD(cudaMallocHost((void**)&m->h_stats, (size_t)2 * ICL_MAX_HEADS * 2 * 4));
  CKD(cudaStreamCreateWithFlags(&m->copy, cudaStreamNonBlocking));
  CKD(dmalloc(&m->d_partial, 1024)); CKD(dmalloc(&m->d_gnorm, 4));
  if (const char* e = getenv("ICL_HEAD_STREAMS")) m->head_streams = atoi(e) != 0;
  if (const char* e = getenv("ICL_PDL")) m->pdl = atoi(e) != 0;
  if (const char* e = getenv("ICL_PHASE_EVENTS")) m->ph_on = atoi(e) != 0;
  if (const char* e = getenv("ICL_ZERO_EARLY")) m->zero_early = atoi(e);
  if (const char* e = getenv("ICL_ZERO_BLOCKS")) m->zero_blocks = atoi(e);
  for (auto& h : m->heads) {
    int B = h.c.batch_size, C = h.c.n_classes;
    int maxw = 0;
    for (int k = 1; k <= h.c.n_hidden; k++) maxw = std::max(maxw, h.dims[k]);
    CKD(dmalloc(&h.bi, (size_t)B * h.ldbi)); CKD(dmalloc(&h.dbi, (size_t)B * h.lddbi));
    CKD(cudaMemset(h.bi, 0, (size_t)B * h.ldbi * 4)); CKD(cudaMemset(h.dbi, 0, (size_t)B * h.lddbi * 4));
    CKD(dmalloc(&h.dA, (size_t)B * maxw)); CKD(dmalloc(&h.dBuf, (size_t)B * maxw));
    for (int k = 1; k <= h.c.n_hidden; k++) { float* a; CKD(dmalloc(&a, (size_t)B * h.dims[k])); h.act.push_back(a); }
    for (int k = 1; k <= h.c.n_hidden; k++) { float* a; CKD(dmalloc(&a, (size_t)B * h.dims[k])); h.dzb.push_back(a); }
    CKD(dmalloc(&h.proba, (size_t)B * C)); CKD(dmalloc(&h.dlogits, (size_t)B * C));
    CKD(dmalloc(&h.smb_part, (size_t)((B + SMB_ROWS - 1) / SMB_ROWS) * ((size_t)h.dims[h.c.n_hidden] * C + C)));
    CKD(dmalloc(&h.row_loss, B)); CKD(dmalloc(&h.row_correct, B)); CKD(dmalloc(&h.scalars, 4)); CKD(dmalloc(&h.pred, B));
    CKD(cudaMallocHost((void**)&h.h_out, ((size_t)B * C + 4) * 4));
    CKD(cudaMallocHost((void**)&h.h_pred, (size_t)B * 8));
    if (h.fact_ok) {
      const size_t w1 = (size_t)h.dims[1];
      CKD(dmalloc(&h.Xm, (size_t)B * h.ldm)); CKD(cudaMemset(h.Xm, 0, (size_t)B * h.ldm * 4));
      CKD(dmalloc(&h.Xb, (size_t)B * h.ldb)); CKD(cudaMemset(h.Xb, 0, (size_t)B * h.ldb * 4));
      CKD(dmalloc(&h.U, B * w1)); CKD(dmalloc(&h.V, B * w1)); CKD(dmalloc(&h.dU, B * w1)); CKD(dmalloc(&h.dV, B * w1));
      CKD(dmalloc(&h.dXm, (size_t)B * h.lddbi)); CKD(cudaMemset(h.dXm, 0, (size_t)B * h.lddbi * 4));
      CKD(cudaStreamCreateWithFlags(&h.sb, cudaStreamNonBlocking));
      CKD(cudaEventCreateWithFlags(&h.ev_bfork, cudaEventDisableTiming)); CKD(cudaEventCreateWithFlags(&h.ev_box, cudaEventDisableTiming));
      CKD(cudaEventCreateWithFlags(&h.ev_du, cudaEventDisableTiming));
    }
    if (m->heads.size() > 1 && m->head_streams) {
      CKD(cudaStreamCreateWithFlags(&h.hs, cudaStreamNonBlocking)); CKD(cudaStreamCreateWithFlags(&h.ha, cudaStreamNonBlocking));
      CKD(cudaEventCreateWithFlags(&h.ev_done, cudaEventDisableTiming)); CKD(cudaEventCreateWithFlags(&h.ev_adone, cudaEventDisableTiming));
    }
  }
  use_input_set(m, 0);
  CKD(cudaStreamCreateWithFlags(&m->aux, cudaStreamNonBlocking));
  CKD(cudaStreamCreateWithFlags(&m->aux2, cudaStreamNonBlocking));
  CKD(cudaStreamCreateWithFlags(&m->aux3, cudaStreamNonBlocking));
  CKD(cudaStreamCreateWithFlags(&m->aux4, cudaStreamNonBlocking));
  CKD(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_join2, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_join3, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_heads, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_wg0, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_packs, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_side, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_loss, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_hfork, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_zfork, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_zero, cudaEventDisableTiming));
  CKD(cudaEventCreateWithFlags(&m->ev_dhzero, cudaEventDisableTiming));
  for (auto& e : m->ev_dz) CKD(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CKD(cudaEventCreate(&m->ev_t0)); CKD(cudaEventCreate(&m->ev_t1));
  for (int i = 0; i < PH_N; i++) for (int j = 0; j < 2; j++) CKD(cudaEventCreate(&m->ev_ph[i][j]));
  if (tcgen05_gemm_init() != 0) { fail("icl_create: cannot resolve cuTensorMapEncodeTiled"); icl_destroy(m); *out = nullptr; return -1; }
  if (m->rp_U && rec_init(m) != 0) { icl_destroy(m); *out = nullptr; return -1; }
  if (m->rp_U && rec16_init(m) != 0) { icl_destroy(m); *out = nullptr; return -1; }
  if (k1_init(m) != 0) { icl_destroy(m); *out = nullptr; return -1; }
  if (bptt_init(m) != 0) { icl_destroy(m); *out = nullptr; return -1; }
  if (ns_init(m) != 0) { icl_destroy(m); *out = nullptr; return -1; }
#undef CKD
  return 0;
}

// Device-resident box-feature table [n_rows, box_width] (fp32): every box of the corpus once (the reference re-reads and re-parses
// the per-image feature files for every batch, nn_utils/data.py:506-524); affinity batches then reference rows (icl_head_batch.box_rows).
extern "C" int icl_set_box_table(icl_model* m, const float* table, int64_t n_rows, int32_t width) {
  CK(cudaStreamSynchronize(m->stream));
  if (m->box_table) { CK(cudaFree(m->box_table)); m->box_table = nullptr; m->box_table_rows = 0; }
  if (!table || n_rows <= 0) return 0;
  for (auto& h : m->heads) if (h.c.box_width > 0 && h.c.box_width != width) return fail("icl_set_box_table: width %d != the head's box_width %d", width, h.c.box_width);
  CK(dmalloc(&m->box_table, (size_t)n_rows * width));
  CK(cudaMemcpy(m->box_table, table, (size_t)n_rows * width * 4, cudaMemcpyHostToDevice));
  m->box_table_rows = n_rows;
  return 0;
}

// Device-resident token table [n_rows, E] (fp32): e.g. every caption matrix of the corpus concatenated (nn_utils/data.py keeps
// them per caption in data_dict['sentences']), uploaded ONCE; batches then reference rows (icl_batch.token_rows).
extern "C" int icl_set_token_table(icl_model* m, const float* table, int64_t n_rows) {
  CK(cudaStreamSynchronize(m->stream));
  if (m->tok_table) { CK(cudaFree(m->tok_table)); m->tok_table = nullptr; m->tok_table_rows = 0; }
  if (!table || n_rows <= 0) return 0;
  CK(dmalloc(&m->tok_table, (size_t)n_rows * m->E));
  CK(cudaMemcpy(m->tok_table, table, (size_t)n_rows * m->E * 4, cudaMemcpyHostToDevice));
  m->tok_table_rows = n_rows;
  return 0;
}

// CRC-32C (Castagnoli), host only: the tensor checksums of TensorFlow Saver-V2 checkpoints (tf_checkpoint.py).  Slice-by-8 tables.
extern "C" uint32_t icl_crc32c(const void* data, uint64_t n, uint32_t crc) {
  static uint32_t T[8][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      T[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
      for (int t = 1; t < 8; t++) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFF];
    init = true;
  }
  const uint8_t* p = (const uint8_t*)data;
  crc = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = T[7][lo & 0xFF] ^ T[6][(lo >> 8) & 0xFF] ^ T[5][(lo >> 16) & 0xFF] ^ T[4][lo >> 24] ^ T[3][hi & 0xFF] ^ T[2][(hi >> 8) & 0xFF] ^
          T[1][(hi >> 16) & 0xFF] ^ T[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) crc = T[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

extern "C" int icl_set_stream(icl_model* m, void* s) { m->stream = (cudaStream_t)s; return 0; }
extern "C" int icl_sync(icl_model* m) { CK(cudaStreamSynchronize(m->stream)); CK(cudaStreamSynchronize(m->aux)); CK(cudaStreamSynchronize(m->aux2)); return 0; }
extern "C" int icl_param_count(icl_model* m) { return (int)m->params.size(); }
extern "C" int icl_param_info(icl_model* m, int i, const char** name, int32_t* rows, int32_t* cols, int64_t* off) {
  if (i < 0 || i >= (int)m->params.size()) return fail("param index out of range");
  const Param& p = m->params[i];
  if (name) *name = p.name.c_str();
  if (rows) *rows = p.rows;
  if (cols) *cols = p.cols;
  if (off) *off = p.off;
  return 0;
}
static float* kind_buf(icl_model* m, int kind) { return kind == 0 ? m->P : kind == 1 ? m->G : kind == 2 ? m->M : kind == 3 ? m->V : nullptr; }
extern "C" int icl_get_tensor(icl_model* m, int kind, const char* name, float* host) {
  auto it = m->pindex.find(name);
  if (it == m->pindex.end() || !kind_buf(m, kind)) return fail("unknown tensor '%s' (kind %d)", name, kind);
  const Param& p = m->params[it->second];
  CK(cudaStreamSynchronize(m->stream));
  CK(cudaMemcpy(host, kind_buf(m, kind) + p.off, (size_t)p.rows * p.cols * 4, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int icl_set_tensor(icl_model* m, int kind, const char* name, const float* host) {
  auto it = m->pindex.find(name);
  if (it == m->pindex.end() || !kind_buf(m, kind)) return fail("unknown tensor '%s' (kind %d)", name, kind);
  const Param& p = m->params[it->second];
  CK(cudaStreamSynchronize(m->stream));
  CK(cudaMemcpy(kind_buf(m, kind) + p.off, host, (size_t)p.rows * p.cols * 4, cudaMemcpyHostToDevice));
  if (kind == 0) m->pr_dirty = m->wp_dirty = m->wih_dirty = m->wb_dirty = true;
  return 0;
}
extern "C" int icl_get_step(icl_model* m, int64_t* t) { *t = m->step; return 0; }
extern "C" int icl_set_step(icl_model* m, int64_t t) { m->step = t; return 0; }
extern "C" int icl_grad_buffer(icl_model* m, void** p, int64_t* n) { *p = m->G; *n = m->n_params; return 0; }
// Data parallel over NVSwitch: the caller hands the library a gradient buffer it allocated in SYMMETRIC memory (same size on every
// rank, mapped into a multicast object: torch.distributed._symmetric_memory) -- every kernel then writes its gradients there -- and
// icl_nvls_allreduce sums it over the ranks IN the switch: rank r reads slice r of all ranks' buffers with ONE multimem.ld_reduce per
// 16 bytes and writes the sum back to slice r of all of them with one multimem.st.  No NCCL kernel, no staging; the caller brackets
// the call with cross-rank barriers (all gradients complete before / all slices written after).
extern "C" int icl_adopt_grad_buffer(icl_model* m, void* buf, int64_t n_floats) {
  if (n_floats < m->n_params) return fail("icl_adopt_grad_buffer: %lld floats < %lld parameters", (long long)n_floats, (long long)m->n_params);
  if (((uintptr_t)buf & 15) != 0) return fail("icl_adopt_grad_buffer: the buffer must be 16-byte aligned");
  CK(cudaStreamSynchronize(m->stream));
  if (m->G && !m->G_external) CK(cudaFree(m->G));
  m->G = reinterpret_cast<float*>(buf);
  m->G_external = true;
  CK(cudaMemsetAsync(m->G, 0, (size_t)m->n_params * 4, m->stream));
  return 0;
}
__device__ __forceinline__ float4 mm_ld_reduce(float* a) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* a, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__global__ void k_nvls_allreduce(float* __restrict__ mc, long n4, int rank, int world) {
  // slice of this rank in float4 units; every element of the buffer is read-reduced and written by exactly one rank.  A reduction
  // through the switch is a ~3 us round trip: four are kept in flight per thread (one per thread measured 34 us for 4.7 MB)
  const long per = (n4 + world - 1) / world, lo = rank * per, hi = lo + per < n4 ? lo + per : n4;
  const long T = (long)gridDim.x * blockDim.x;
  long i = lo + (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * T < hi; i += 4 * T) {
    const float4 v0 = mm_ld_reduce(mc + i * 4), v1 = mm_ld_reduce(mc + (i + T) * 4), v2 = mm_ld_reduce(mc + (i + 2 * T) * 4),
                 v3 = mm_ld_reduce(mc + (i + 3 * T) * 4);
    mm_st(mc + i * 4, v0); mm_st(mc + (i + T) * 4, v1); mm_st(mc + (i + 2 * T) * 4, v2); mm_st(mc + (i + 3 * T) * 4, v3);
  }
  for (; i < hi; i += T) mm_st(mc + i * 4, mm_ld_reduce(mc + i * 4));
}
extern "C" int icl_nvls_allreduce(icl_model* m, void* multicast_ptr, int32_t rank, int32_t world) {
  if (!m->G_external) return fail("icl_nvls_allreduce: no symmetric gradient buffer adopted");
  if (world < 1 || rank < 0 || rank >= world) return fail("icl_nvls_allreduce: rank %d of %d", rank, world);
  const long n4 = (m->n_params + 3) / 4;             // the symmetric buffer is allocated in whole float4s
  const long per = (n4 + world - 1) / world;
  const int blocks = (int)std::max<long>(1, std::min<long>(592, (per + 255) / 256));
  k_nvls_allreduce<<<blocks, 256, 0, m->stream>>>(reinterpret_cast<float*>(multicast_ptr), n4, rank, world);
  m->launches++;
  return cudaGetLastError() == cudaSuccess ? 0 : fail("k_nvls_allreduce launch failed");
}
// Overlapping the gradient all-reduce with the backward pass: the flat gradient buffer is [LSTM | heads]; the heads' part is
// final once the heads' backward has run, long before the BPTT and the weight-gradient GEMMs finish.  `first_head_float` = where
// the heads' gradients start; icl_wait_head_grads makes `cuda_stream` wait for them (cudaStreamWaitEvent), so a collective
// enqueued on that stream afterwards runs concurrently with the rest of the step on the compute stream.
extern "C" int icl_grad_split(icl_model* m, int64_t* first_head_float) {
  *first_head_float = m->heads.empty() ? m->n_params : m->params[m->heads[0].pW[0]].off;
  return 0;
}
extern "C" int icl_wait_head_grads(icl_model* m, void* cuda_stream) {
  CK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, m->ev_heads, 0));
  return 0;
}
// The LSTM's gradient floats are [0, first_head_float): the forward direction's kernel + bias first, then the backward direction's
// from *first_bw_float on.  The weight-gradient GEMM of the forward direction runs first; icl_wait_fw_lstm_grads makes the given
// stream wait for it, so a collective on floats [0, first_bw_float) overlaps the backward direction's GEMM.
// Work the library has queued on its side streams behind the last call (today: the fp16 repack of the LSTM weights after an update,
// which normally overlaps the next step's input preparation) is joined into the main stream: a caller that brackets ONE step with
// events calls this before the closing event so that the step's time includes all of its work.
// Test hook: fills the fp16 h_{k-1} operand rows of the forward recurrence (every slicing) with 65504.0 on the main stream.  A run
// publishes every row before another CTA reads it, so the poison must never reach a result; a read of an unpublished row would.
// (A large finite value rather than NaN: the K-padding columns are multiplied by zero weight rows.)
extern "C" int icl_debug_poison_recurrence(icl_model* m) {
  for (auto* v : {&m->rf20, &m->rf16})
    for (int d = 0; d < 2; d++)
      if (v->on && v->Hp16[d]) {
        const size_t n = (size_t)m->rows_cap * v->KP;
        k_fill_u16<<<1184, 256, 0, m->stream>>>(reinterpret_cast<uint16_t*>(v->Hp16[d]), n, (uint16_t)0x7BFF);
        CK(cudaGetLastError());
      }
  return 0;
}
extern "C" int icl_join_side_work(icl_model* m) {
  if (m->packs_pending) { CK(cudaStreamWaitEvent(m->stream, m->ev_packs, 0)); m->packs_pending = false; }
  return 0;
}
extern "C" int icl_grad_split_lstm(icl_model* m, int64_t* first_bw_float) {
  *first_bw_float = m->params[m->pK[1]].off;
  return 0;
}
extern "C" int icl_wait_fw_lstm_grads(icl_model* m, void* cuda_stream) {
  CK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, m->ev_wg0, 0));
  return 0;
}
extern "C" int icl_param_buffer(icl_model* m, void** p, int64_t* n) { *p = m->P; *n = m->n_params; m->pr_dirty = m->wp_dirty = m->wih_dirty = m->wb_dirty = true; return 0; }
extern "C" int icl_kernel_launches(icl_model* m, int64_t* n) { *n = m->launches; return 0; }
extern "C" int icl_last_step_ms(icl_model* m, float* ms) { *ms = m->last_ms; return 0; }
extern "C" int icl_copy_bytes(icl_model* m, int64_t* h2d, int64_t* d2h) { *h2d = m->h2d_bytes; *d2h = m->d2h_bytes; return 0; }
extern "C" int icl_batch_stats(icl_model* m, int64_t* n_seqs, int64_t* n_tokens, int32_t* t_max) {
  *n_seqs = m->S; *n_tokens = m->Ntok; *t_max = m->Tmax; return 0;
}
// device time of each phase of the last icl_run_resident (ms; 0 where the phase did not run).  Synchronises the stream.
extern "C" int icl_set_phase_timing(icl_model* m, int on) { m->ph_on = on != 0; return 0; }
extern "C" int icl_phase_ms(icl_model* m, float* ms) {
  CK(cudaStreamSynchronize(m->stream));
  for (int i = 0; i < PH_N; i++) {
    ms[i] = 0.f;
    if (m->ph_used[i]) CK(cudaEventElapsedTime(&ms[i], m->ev_ph[i][0], m->ev_ph[i][1]));
  }
  return 0;
}

// ----------------------------------------------------------------------------- upload (H2D of one batch_tensors dict)
#define H2D(dst, src, bytes, st) do { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st)); m->h2d_bytes += (int64_t)(bytes); } while (0)
extern "C" int icl_upload(icl_model* m, const icl_batch* b) {
  if (!m || !b) return fail("icl_upload: null argument");
  int S = b->n_seqs, E = m->E;
  if (S < 1 || S > m->S_cap) return fail("icl_upload: n_seqs=%d exceeds capacity %d", S, m->S_cap);
  if (b->n_heads != m->cfg.n_heads) return fail("icl_upload: batch has %d heads, model has %d", b->n_heads, m->cfg.n_heads);
  const bool by_rows = b->token_rows != nullptr;
  if (by_rows && !m->tok_table) return fail("icl_upload: token_rows given but no token table was set (icl_set_token_table)");
  // per-head sentence tensors (multi-head calls): sequence s lives in the tensor of the head whose range contains it
  const bool per_head = !by_rows && !b->sentences;
  std::vector<const char*> seq_src;               // per sequence: first byte of its padded row block (per_head only)
  std::vector<int> seq_T;
  if (per_head) {
    seq_src.assign(S, nullptr); seq_T.assign(S, 0);
    const size_t es = b->sent_dtype == ICL_F64 ? 8 : 4;
    int tmaxp = 0;
    for (int hi = 0; hi < b->n_heads; hi++) {
      const icl_head_batch& hb = b->heads[hi];
      if (hb.inactive || !hb.sentences) continue;
      if (hb.n_seqs < 0 || hb.sent_offset < 0 || hb.sent_offset + hb.n_seqs > S) return fail("icl_upload: head %d sentence range [%d,%d) outside [0,%d)", hi, hb.sent_offset, hb.sent_offset + hb.n_seqs, S);
      if (hb.padded_T < 1 || hb.padded_T > m->T_cap) return fail("icl_upload: head %d padded_T=%d exceeds capacity %d", hi, hb.padded_T, m->T_cap);
      for (int s = 0; s < hb.n_seqs; s++) {
        seq_src[hb.sent_offset + s] = (const char*)hb.sentences + (size_t)s * hb.padded_T * E * es;
        seq_T[hb.sent_offset + s] = hb.padded_T;
      }
      tmaxp = std::max(tmaxp, hb.padded_T);
    }
    for (int s = 0; s < S; s++) if (!seq_src[s]) return fail("icl_upload: neither sentences nor token_rows given (sequence %d has no source)", s);
    if (b->sent_packed) return fail("icl_upload: per-head sentence tensors must be padded, not packed");
  }
  if (!by_rows && !per_head && !b->sent_packed && (b->padded_T < 1 || b->padded_T > m->T_cap)) return fail("icl_upload: padded_T=%d exceeds capacity %d", b->padded_T, m->T_cap);
  int T = (by_rows || b->sent_packed) ? m->T_cap : per_head ? *std::max_element(seq_T.begin(), seq_T.end()) : b->padded_T;
  // next input set: its pinned staging is free once the copies issued from it two uploads ago have completed, its device
  // buffers once the step that consumed them has (the copy stream waits for that; the host does not)
  const int set = (m->cur + 1) & 1;
  InSet& I = m->in[set];
  if (I.copied_pending) { CK(cudaEventSynchronize(I.ev_copied)); I.copied_pending = false; }
  if (I.done_pending) { CK(cudaStreamWaitEvent(m->copy, I.ev_done, 0)); I.done_pending = false; }
  m->resident = false;
  m->cur = set;
  use_input_set(m, set);
  m->h2d_bytes = 0;
  int *lens = I.host<int>(I.o_lens), *rank = I.host<int>(I.o_rank), *tokstart = I.host<int>(I.o_tokstart),
      *tokseq = I.host<int>(I.o_tokseq), *offs = I.host<int>(I.o_off), *nact = I.host<int>(I.o_nact);
  float* const h_x = I.host<float>(I.o_x);
  long ntok = 0; int tmax = 0;
  for (int s = 0; s < S; s++) {
    double l = read_num(b->seq_lengths, b->len_dtype, s);
    int li = (int)l;
    if (li != l || li < 0 || li > (per_head ? seq_T[s] : T)) return fail("icl_upload: seq_lengths[%d]=%g outside [0,%d]", s, l, per_head ? seq_T[s] : T);
    lens[s] = li; tokstart[s] = (int)ntok;
    ntok += li; tmax = std::max(tmax, li);
  }
  // rank by length (descending, stable): step k runs ranks [0, nact[k])
  std::vector<int> order(S);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return lens[a] > lens[c]; });
  for (int i = 0; i < S; i++) rank[order[i]] = i;
  m->n_active.assign(tmax + 1, 0);
  for (int s = 0; s < S; s++) for (int k = 0; k < lens[s]; k++) m->n_active[k]++;
  m->off.assign(tmax + 2, 0);
  for (int k = 0; k <= tmax; k++) m->off[k + 1] = m->off[k] + (m->n_active[k] + 127) / 128 * 128;   // blocks padded to 128 rows
  for (int k = 0; k <= tmax; k++) { offs[k] = m->off[k]; nact[k] = m->n_active[k]; }
  // pack valid tokens (caption-major) into pinned memory as fp32
  size_t esz = b->sent_dtype == ICL_F64 ? 8 : 4;
  if (!by_rows && b->sent_dtype != ICL_F32 && b->sent_dtype != ICL_F64) return fail("icl_upload: sentences must be float32/float64");
  // pack + convert on several host threads, in chunks, so the H2D copy of chunk i overlaps the packing of chunk i+1
  cudaStream_t st = m->copy;
  if (m->ph_on) CK(cudaEventRecord(I.ev_c0, st));
  m->use_rows = by_rows;
  // half-width wire (see cvt_f32_h16_stream): only where the device would round these rows to 10 mantissa bits next anyway (product
  // mode with the fp16 projection operands) and no l2-normalisation sits in between (its sum of squares keeps the fp32 inputs)
  const bool wire16 = !by_rows && wire16_enabled() && m->round_ops && m->k1_f16 && !m->cfg.data_norm;
  m->x_half = wire16;
  if (by_rows) {                      // packed caption-major row numbers into the resident token table: 4 bytes per token
    int* tokrow = I.host<int>(I.o_tokrow);
    for (int s = 0; s < S; s++)
      for (int t = 0; t < lens[s]; t++) {
        const int32_t r = b->token_rows[tokstart[s] + t];
        if (r < 0 || r >= m->tok_table_rows) return fail("icl_upload: token_rows[%d]=%d outside the token table [0,%lld)", tokstart[s] + t, r, (long long)m->tok_table_rows);
        tokrow[tokstart[s] + t] = r;
        tokseq[tokstart[s] + t] = s;
      }
  } else {
    const int n_chunks = ntok * (long)E * 4 > (4 << 20) ? 8 : 1;
    const auto t_pack0 = std::chrono::steady_clock::now();
    int s0 = 0;
    for (int c = 0; c < n_chunks; c++) {
      const long tok_end = ntok * (c + 1) / n_chunks;
      int s1 = s0;
      while (s1 < S && (c == n_chunks - 1 || tokstart[s1] + lens[s1] <= tok_end)) s1++;
      if (s1 == s0) continue;
      const int per = 16, items = (s1 - s0 + per - 1) / per;           // 16 sequences per work item
      host_pool().run(items, [&, s0, s1](int it) {
        for (int s = s0 + it * per; s < std::min(s1, s0 + (it + 1) * per); s++) {
          const char* src = per_head ? seq_src[s] : (const char*)b->sentences + (b->sent_packed ? (size_t)tokstart[s] * E : (size_t)s * T * E) * esz;
          if (wire16) {
            uint16_t* d16 = reinterpret_cast<uint16_t*>(h_x) + (size_t)tokstart[s] * E;
            if (b->sent_dtype == ICL_F32) cvt_f32_h16_stream(d16, (const float*)src, (size_t)lens[s] * E);
            else cvt_f64_h16_stream(d16, (const double*)src, (size_t)lens[s] * E);
          } else to_f32(h_x + (size_t)tokstart[s] * E, src, b->sent_dtype, (size_t)lens[s] * E);
          for (int t = 0; t < lens[s]; t++) tokseq[tokstart[s] + t] = s;
        }
        if (g_pack_nt || wire16) _mm_sfence();
      });
      const long t0 = tokstart[s0], t1 = (long)tokstart[s1 - 1] + lens[s1 - 1];
      if (t1 > t0) {
        if (wire16) H2D(reinterpret_cast<uint16_t*>(m->xraw) + t0 * E, reinterpret_cast<uint16_t*>(h_x) + t0 * E, (size_t)(t1 - t0) * E * 2, st);
        else H2D(m->xraw + t0 * E, h_x + t0 * E, (size_t)(t1 - t0) * E * 4, st);
      }
      s0 = s1;
    }
    if (n_chunks > 1)       // a large packing pass: one sample for the pool's thread-count tuner
      host_pool().report(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pack0).count());
  }
  std::vector<std::pair<size_t, size_t>> holes;          // byte ranges of the meta region nothing on the device will read
  for (int hi = 0; hi < b->n_heads; hi++) {
    Head& h = m->heads[hi];
    HeadIn& hin = h.in[set];
    const icl_head_batch& hb = b->heads[hi];
    int B = h.c.batch_size, C = h.c.n_classes;
    h.active = !hb.inactive;
    if (!h.active) { h.has_labels = false; continue; }
    for (int sl = 0; sl < h.slots.n_slots; sl++) {
      int id = h.slot_index_id[sl];
      if (id < 0) continue;
      if (!hb.idx[id]) return fail("icl_upload: head %d is missing index matrix %d", hi, id);
      int* dst = I.host<int>(hin.idx[id]);
      for (int r = 0; r < B; r++) {
        double d = read_num(hb.idx[id], hb.idx_dtype, r * 3), sq = read_num(hb.idx[id], hb.idx_dtype, r * 3 + 1) + hb.sent_offset,
               w = read_num(hb.idx[id], hb.idx_dtype, r * 3 + 2);
        // TF-CPU gather_nd raises on an out-of-range index (core.py:350); so do we
        if (!(d == 0 || d == 1) || sq < 0 || sq >= S || w < 0 || w >= T || sq != (int)sq || w != (int)w)
          return fail("icl_upload: head %d index matrix %d row %d = [%g,%g,%g] out of range (S=%d,T=%d)", hi, id, r, d, sq, w, S, T);
        dst[r * 3] = (int)d; dst[r * 3 + 1] = (int)sq; dst[r * 3 + 2] = (int)w;
      }
    }
    auto up = [&](const void* src, int dt, size_t off, size_t n, const char* what) -> int {
      if (n == 0) return 0;
      if (!src) return fail("icl_upload: head %d is missing %s", hi, what);
      float* dst = I.host<float>(off);
      if (n < (64u << 10)) { to_f32(dst, src, dt, n); if (g_pack_nt) _mm_sfence(); return 0; }
      const size_t per = 16u << 10, esz_ = dtype_size(dt);          // large blocks (m_feats of 2048 examples: 4 MB of float64): on the pool
      host_pool().run((int)((n + per - 1) / per), [&](int it) {
        const size_t a = (size_t)it * per, c = std::min(n, a + per);
        to_f32(dst + a, (const char*)src + a * esz_, dt, c - a);
        if (g_pack_nt) _mm_sfence();
      });
      return 0;
    };
    CKI(up(hb.feats, hb.feats_dtype, hin.feats, (size_t)B * h.c.n_feats, "m_feats/ij_feats"));
    const bool box_table_rows = hb.box_rows && h.c.box_width > 0;
    if (box_table_rows) {                 // rows of the resident box table instead of [B, 4096] floats
      if (!m->box_table) return fail("icl_upload: head %d gives box_rows but no box table was set (icl_set_box_table)", hi);
      int* br = I.host<int>(hin.boxrow);
      for (int r = 0; r < B; r++) {
        if (hb.box_rows[r] < 0 || hb.box_rows[r] >= m->box_table_rows)
          return fail("icl_upload: head %d box_rows[%d]=%d outside the box table [0,%lld)", hi, r, hb.box_rows[r], (long long)m->box_table_rows);
        br[r] = hb.box_rows[r];
      }
      for (int sl = 0; sl < h.slots.n_slots; sl++)
        if (h.slot_index_id[sl] == -2) { h.slots.dense[sl] = m->box_table; h.slots.rowidx[sl] = I.dev<int>(hin.boxrow); }
    } else if (h.c.box_width > 0 && !hb.box) return fail("icl_upload: head %d is missing box_embeddings", hi);
    if (h.c.n_box_feats > 0 && !hb.bfeats) return fail("icl_upload: head %d is missing b_feats", hi);
    // Affinity layer-1 factorisation: the distinct mentions (index rows + m_feats row) and distinct boxes (table row, or the
    // bytes of the host row, + b_feats row) of the batch; a hash picks the candidates, an exact comparison decides
    h.fact = false; h.box_compact = false;
    const int fmode = h.fact_ok ? aff_mode() : 0;
    if (fmode != 0) {
      int *m_of = I.host<int>(hin.m_of), *b_of = I.host<int>(hin.b_of);
      std::vector<uint64_t> key(B);
      std::vector<int> rep_m, rep_b;
      const size_t fsz = (size_t)h.c.n_feats * dtype_size(hb.feats_dtype), xsz = (size_t)h.c.box_width * dtype_size(hb.box_dtype),
                   bsz = (size_t)h.c.n_box_feats * dtype_size(hb.bfeats_dtype);
      const char *fsrc = (const char*)hb.feats, *xsrc = (const char*)hb.box, *bsrc = (const char*)hb.bfeats;
      std::vector<const int*> ixs;
      for (int i : h.slotm_src) if (h.slot_index_id[i] >= 0) ixs.push_back(I.host<int>(hin.idx[h.slot_index_id[i]]));
      for (int r = 0; r < B; r++) {
        uint64_t k = 0x452821e638d01377ull;
        for (const int* ix : ixs) k = hash_bytes(ix + r * 3, 12, k);      // the m_feats rows are compared, not hashed
        key[r] = k;
      }
      group_rows(B, key, [&](int a, int c) {
        for (const int* ix : ixs) if (memcmp(ix + a * 3, ix + c * 3, 12) != 0) return false;
        return fsz == 0 || memcmp(fsrc + a * fsz, fsrc + c * fsz, fsz) == 0;
      }, m_of, rep_m);
      group_box_rows(B, box_table_rows ? nullptr : xsrc, xsz, box_table_rows ? hb.box_rows : nullptr, bsrc, bsz, b_of, rep_b);
      h.Mu = (int)rep_m.size(); h.Nu = (int)rep_b.size();
      // reference-shaped TRAINING batches carry one copy of the caption per pair (nn_utils/data.py:397-403: every copy draws its own
      // dropout masks), so their mentions do not repeat and only the box half (4096 of the 5552 columns) collapses; prediction
      // batches built with load_batch(dedup=True) share captions and collapse on both sides
      h.fact = fmode >= 2 || h.Nu * 2 <= B;
      if (h.fact) {
        h.box_compact = !box_table_rows;
        memcpy(I.host<int>(hin.rep_m), rep_m.data(), (size_t)h.Mu * 4);
        int* rbr = I.host<int>(hin.rep_boxrow);
        for (int g = 0; g < h.Nu; g++) rbr[g] = box_table_rows ? hb.box_rows[rep_b[g]] : rep_b[g];
        csr_lists(B, h.Mu, m_of, I.host<int>(hin.m_start), I.host<int>(hin.m_mem));
        csr_lists(B, h.Nu, b_of, I.host<int>(hin.b_start), I.host<int>(hin.b_mem));
        // only the distinct box / b_feats rows are packed: row g of the block = group g
        host_pool().run(h.Nu, [&](int g) {
          if (!box_table_rows) to_f32(I.host<float>(hin.box) + (size_t)g * h.c.box_width, xsrc + rep_b[g] * xsz, hb.box_dtype, h.c.box_width);
          if (bsz) to_f32(I.host<float>(hin.bfeats) + (size_t)g * h.c.n_box_feats, bsrc + rep_b[g] * bsz, hb.bfeats_dtype, h.c.n_box_feats);
          if (g_pack_nt) _mm_sfence();
        });
        fact_slots(h, m->box_table);
        h.fact_batches++; h.fact_mu += h.Mu; h.fact_nu += h.Nu;
        // the unused tails of the two blocks need not cross PCIe
        if (!box_table_rows) holes.push_back({hin.box + (size_t)h.Nu * h.c.box_width * 4, hin.box + (size_t)B * h.c.box_width * 4});
        if (bsz) holes.push_back({hin.bfeats + (size_t)h.Nu * h.c.n_box_feats * 4, hin.bfeats + (size_t)B * h.c.n_box_feats * 4});
      }
    }
    if (!h.fact) {
      if (!box_table_rows) CKI(up(hb.box, hb.box_dtype, hin.box, (size_t)B * h.c.box_width, "box_embeddings"));
      CKI(up(hb.bfeats, hb.bfeats_dtype, hin.bfeats, (size_t)B * h.c.n_box_feats, "b_feats"));
    }
    h.has_labels = hb.labels != nullptr;
    if (h.has_labels) CKI(up(hb.labels, hb.labels_dtype, hin.labels, (size_t)B * C, "labels"));
  }
  // everything but the sentence rows: ONE copy of the meta region (the token->sequence map is last: only its used prefix)
  // (a factorised affinity head packs only its distinct box rows: the copy skips the rest of that block when it is worth a second call)
  {
    size_t pos = I.meta_off;
    const size_t end = I.o_tokseq + (size_t)ntok * 4;
    std::sort(holes.begin(), holes.end());
    for (const auto& hole : holes) {
      const size_t h0 = (hole.first + 255) & ~(size_t)255, h1 = hole.second & ~(size_t)255;
      if (h1 <= h0 || h1 - h0 < (512u << 10) || h0 < pos) continue;
      H2D(I.d_blob + pos, I.h_blob + pos, h0 - pos, st);
      pos = h1;
    }
    H2D(I.d_blob + pos, I.h_blob + pos, end - pos, st);
  }
  CK(cudaEventRecord(I.ev_copied, st));
  I.copied_pending = true;
  m->S = S; m->Tmax = tmax; m->Ntok = ntok; m->NtokP = m->off[tmax];
  m->seq_gid0 = b->seq_gid_offset; m->ex_gid0 = b->ex_gid_offset;
  m->resident = true;
  return 0;
}

// ----------------------------------------------------------------------------- forward / backward on resident data
#define LAUNCHED(m) do { (m)->launches++; CK(cudaGetLastError()); } while (0)
// per-phase CUDA-event timers (icl_phase_ms): sixteen timing-event records per step.  A timing event is a serialisation point of
// the stream -- measured: card2048 1.295 -> 1.245 ms per step, nonvis512 0.845 -> 0.802 without them -- so they are only recorded
// when a caller asks for phase times (icl_set_phase_timing; bench.py takes them in a separate pass from the headline number)
#define PH_BEGIN(m, ph) do { if ((m)->ph_on) CK(cudaEventRecord((m)->ev_ph[ph][0], (m)->stream)); } while (0)
#define PH_END(m, ph) do { if ((m)->ph_on) { CK(cudaEventRecord((m)->ev_ph[ph][1], (m)->stream)); (m)->ph_used[ph] = true; } } while (0)

static StepLayout mk_layout(icl_model* m) { StepLayout L; L.off = m->d_off; L.nact = m->d_nact; L.rank = m->d_rank; L.lens = m->d_lens; return L; }
static const float* wbase(icl_model* m) { return m->round_ops ? m->Pr : m->P; }   // GEMM-operand view of the parameters

static int refresh_rounded_params(icl_model* m) {
  if (m->round_ops && m->pr_dirty) {
    k_round_copy<<<296, 256, 0, m->stream>>>(m->P, m->Pr, (long)m->n_params); LAUNCHED(m);
  }
  m->pr_dirty = false;
  return 0;
}

// rows [nact[k], padded end) of every step block -> 0 (operands of the time-batched weight-gradient GEMMs)
__global__ void k_zero_pad_rows(float* __restrict__ buf0, float* __restrict__ buf1, StepLayout L, int W) {
  const int k = blockIdx.x;
  float* buf = blockIdx.y ? buf1 : buf0;
  const long r0 = (long)L.off[k] + L.nact[k], r1 = L.off[k + 1];
  float4* p = reinterpret_cast<float4*>(buf + r0 * W);
  const long n4 = (r1 - r0) * W / 4;
  for (long i = threadIdx.x; i < n4; i += blockDim.x) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

static bool rec_usable(icl_model* m) {
  if (m->rf_on) return rf_pick(m) != nullptr;
#ifndef ICL_EXPERIMENTS
  return false;
#endif
  if (m->rp_U == 0 || !m->rp_on || m->Tmax > RP_MAXT) return false;
  int tiles = (m->n_active[0] + 127) / 128, P = std::max(1, std::min(148 / (2 * m->rp_nsl), tiles));
  return (tiles + P - 1) / P <= RP_MAXTPC;          // the kernel carries the cell state of <= RP_MAXTPC tiles per CTA in registers
}

#ifdef ICL_EXPERIMENTS
static RecArgs rec_args(icl_model* m, int training) {
  RecArgs a;
  a.off = m->d_off; a.nact = m->d_nact; a.Tmax = m->Tmax; a.H = m->H; a.nsl = m->rp_nsl;
  int tiles = (m->n_active[0] + 127) / 128;
  a.P = std::max(1, std::min(148 / (2 * m->rp_nsl), tiles));
  a.nkb = m->rp_nkb; a.nk8 = (m->H + 7) / 8; a.max_tiles = m->rp_max_tiles; a.training = training; a.flags = m->rp_flags;
  a.trace = m->rp_trace; a.trace_cta = m->rp_trace_cta;
  return a;
}

#endif

static void pack_whh16(icl_model* m, icl_model::RfVar& v, cudaStream_t st) {
  const int E = m->E, H = m->H;
  const float* W0 = m->P + m->params[m->pK[0]].off + (size_t)E * 4 * H;
  const float* W1 = m->P + m->params[m->pK[1]].off + (size_t)E * 4 * H;
  k_pack_whh_fwd16<<<dim3((4 * H + 31) / 32, (H + 31) / 32, 2), dim3(32, 8), 0, st>>>(W0, W1, v.Wp16[0], v.Wp16[1], H, v.U, v.UP, v.nsl, v.KP);
  m->launches++;
  v.dirty = false;
}
static void pack_wih16(icl_model* m, cudaStream_t st) {
  const int E = m->E, H = m->H;
  k_pack_wih16<<<dim3((4 * H + 31) / 32, m->k1_Kp / 32, 2), dim3(32, 8), 0, st>>>(m->P + m->params[m->pK[0]].off, m->P + m->params[m->pK[1]].off,
                                                                                 m->Wih16[0], m->Wih16[1], E, 4 * H, m->k1_Kp);
  m->launches++;
  m->wih_dirty = false;
}
// the fp16 operand copies of the LSTM weights follow every update: packed on the aux stream right behind the Adam kernel, so they
// overlap the host work between two steps and the next step's input preparation instead of standing in front of K1 / K2
static int repack_after_update(icl_model* m) {
  if (!m->rf_on || !m->rf_last || !m->k1_f16) return 0;
  CK(cudaEventRecord(m->ev_fork, m->stream));
  CK(cudaStreamWaitEvent(m->aux, m->ev_fork, 0));
  pack_wih16(m, m->aux);
  pack_whh16(m, *m->rf_last, m->aux);
  CK(cudaEventRecord(m->ev_packs, m->aux));
  m->packs_pending = true;
  return 0;
}

static int rec_forward_fp16(icl_model* m, int training) {
  cudaStream_t st = m->stream;
  icl_model::RfVar* vp = rf_pick(m);
  if (!vp) return fail("k_rec_fwd16: no slicing fits this batch");
  icl_model::RfVar& v = *vp;
  const int H = m->H, U = v.U;
  if (m->wp_dirty) { m->rf20.dirty = m->rf16.dirty = true; m->wp_dirty = false; }
  if (v.dirty) pack_whh16(m, v, st);
  m->rf_last = vp;
  CK(zero_async(m->rp_flags, (size_t)2 * m->rp_max_tiles * 4, st));
  RecFwd16Args a;
  a.off = m->d_off; a.nact = m->d_nact; a.Tmax = m->Tmax; a.H = H; a.nsl = v.nsl;
  const int tiles = (m->n_active[0] + 127) / 128;
  a.P = std::max(1, std::min(148 / (2 * v.nsl), tiles));
  a.nkb = v.nkb; a.nk16 = v.nk16; a.nst = U == 20 ? rec_fwd16_stages<20>(a.nkb) : rec_fwd16_stages<16>(a.nkb);
  a.max_tiles = m->rp_max_tiles; a.training = training; a.ldx = m->ldx; a.flags = m->rp_flags;
  for (int d = 0; d < 2; d++) { a.Z[d] = m->Z[d]; a.Cc[d] = m->Cc[d]; a.Hx[d] = m->Hx[d]; a.Hp[d] = m->Hp[d]; }
  a.trace = m->rp_trace; a.trace_cta = m->rp_trace_cta;
  void* args[] = {(void*)&v.maps, (void*)&a};
  dim3 grid(2 * a.P * a.nsl), block(RF_THREADS);
  cudaError_t e;
  if (U == 20) e = cudaLaunchCooperativeKernel((void*)k_rec_fwd16<20>, grid, block, args, rec_fwd16_smem<20>(a.nkb), st);
  else e = cudaLaunchCooperativeKernel((void*)k_rec_fwd16<16>, grid, block, args, rec_fwd16_smem<16>(a.nkb), st);
  if (e != cudaSuccess) return fail("k_rec_fwd16<%d> launch failed: %s", U, cudaGetErrorString(e));
  m->launches++;
  return 0;
}

static int rec_forward_persistent(icl_model* m, int training) {
  if (m->rf_on) return rec_forward_fp16(m, training);
#ifndef ICL_EXPERIMENTS
  return fail("the TF32 first-generation recurrence (ICL_REC_FP16=0) is only in -DICL_EXPERIMENTS builds");
#else
  cudaStream_t st = m->stream;
  const int E = m->E, H = m->H, U = m->rp_U;
  if (m->wp_dirty) {
    for (int d = 0; d < 2; d++) {
      const float* Whh = m->P + m->params[m->pK[d]].off + (size_t)E * 4 * H;
      k_pack_whh_fwd<<<148, 256, 0, st>>>(Whh, m->Wp[d], H, U, m->rp_nsl, m->rp_nkb * 32); LAUNCHED(m);
    }
    m->wp_dirty = false;
  }
  CK(zero_async(m->rp_flags, (size_t)2 * m->rp_max_tiles * 4, st));
  RecArgs a = rec_args(m, training);
  void* args[] = {(void*)&m->rp_fmaps, (void*)&a};
  dim3 grid(2 * a.P * a.nsl), block(RP_FWD_THREADS);
  cudaError_t e;
  if (U == 20) e = cudaLaunchCooperativeKernel((void*)k_rec_fwd<20>, grid, block, args, rec_fwd_smem<20>(a.nkb), st);
  else e = cudaLaunchCooperativeKernel((void*)k_rec_fwd<16>, grid, block, args, rec_fwd_smem<16>(a.nkb), st);
  if (e != cudaSuccess) return fail("k_rec_fwd launch failed: %s", cudaGetErrorString(e));
  m->launches++;
  return 0;
#endif
}

static int heads_prefetch(icl_model* m);
static int zero_backward_early(icl_model* m);
static int lstm_forward(icl_model* m, float keep_in, uint64_t seed, int training) {
  int E = m->E, H = m->H;
  long Ntok = m->Ntok, NP = m->NtokP;
  cudaStream_t st = m->stream;
  if (Ntok == 0) return heads_prefetch(m);      // a batch of empty sequences: no recurrence, but the heads still run (on zero rows)
  PH_BEGIN(m, PH_PREP);
  // the pad rows of XH and the zero state in front of step 0 are cleared on the aux stream beside k_prep_x (disjoint rows)
  CK(cudaEventRecord(m->ev_side, st));
  CK(cudaStreamWaitEvent(m->aux, m->ev_side, 0));
  k_zero_pad_rows<<<dim3(m->Tmax, 2), 256, 0, m->aux>>>(m->XH[0], m->XH[1], mk_layout(m), m->ldx); LAUNCHED(m);
  k_zero_2d<<<dim3(74, 2), 256, 0, m->aux>>>(m->Hp[0], m->Hp[1], m->ldx, H, m->off[1]); LAUNCHED(m);   // h_prev of step 0 (A rows of dW_hh)
  CK(cudaEventRecord(m->ev_join, m->aux));
  k_prep_x<<<(unsigned)((Ntok * 32 + 255) / 256), 256, 0, st>>>(m->xraw, m->d_tokseq, m->d_tokstart, mk_layout(m), (int)Ntok, E, m->T_cap,
                                                               m->cfg.data_norm, keep_in, seed, m->seq_gid0, m->round_ops, m->xd[0], m->xd[1], m->ldx, E + H,
                                                               m->use_rows ? m->d_tokrow : nullptr, m->tok_table, m->k1_f16 ? m->X16[0] : nullptr,
                                                               m->k1_f16 ? m->X16[1] : nullptr, m->k1_Kp, m->x_half && !m->use_rows);
  LAUNCHED(m);
  CK(cudaStreamWaitEvent(st, m->ev_join, 0));
  PH_END(m, PH_PREP);
  // K1: time-batched input projection  Z = Xd * W_ih + b   (W_ih = kernel rows [0,E)), both directions
  PH_BEGIN(m, PH_PROJ);
  if (m->packs_pending) { CK(cudaStreamWaitEvent(st, m->ev_packs, 0)); m->packs_pending = false; }     // the eager repack of the last update
  if (m->k1_f16 && m->wih_dirty) pack_wih16(m, st);
  for (int d = 0; d < 2; d++) {
    const float* K = wbase(m) + m->params[m->pK[d]].off;
    GemmArgs g = mk_gemm(m->xd[d], m->ldx, K, 4 * H, m->Z[d], 4 * H, (int)NP, 4 * H, m->k1_f16 ? m->k1_Kp : E);
    g.epi.bias = m->P + m->params[m->pBias[d]].off;
    if (m->k1_f16) {        // fp16 operands (prepared inputs and W_ih^T), kind::f16, fp32 accumulate: half the operand traffic of TF32
      if (tcgen05_gemm_f16_launch(st, m->k1_ta[d], m->k1_tb[d], m->k1_tc[d], g) != 0) return fail("fp16 input-projection GEMM launch failed");
      m->launches++;
    } else CKI(gemm(m, st, false, true, g));
  }
  PH_END(m, PH_PROJ);
  CKI(heads_prefetch(m));      // box halves of factorised affinity heads: beside the recurrence (which leaves SMs idle), not beside K1
  if (training && (m->zero_early == 1 || m->zero_early == 3)) CKI(zero_backward_early(m));
  // K2: the recurrence
  PH_BEGIN(m, PH_REC_FWD);
  if (rec_usable(m)) {
    CKI(rec_forward_persistent(m, training));
    PH_END(m, PH_REC_FWD);
    return 0;
  }
  CK(cudaEventRecord(m->ev_fork, st));
  CK(cudaStreamWaitEvent(m->aux, m->ev_fork, 0));
  for (int k = 0; k < m->Tmax; k++) {
    for (int d = 0; d < 2; d++) {        // directions interleaved on two streams so the host feeds both concurrently
      cudaStream_t sd = d ? m->aux : st;
      const float* Whh = wbase(m) + m->params[m->pK[d]].off + (size_t)E * 4 * H;
      int n = m->n_active[k], n_next = m->n_active[k + 1];
      long o = m->off[k];
      if (k > 0) {   // R = h_{k-1} * W_hh for the running rows
        GemmArgs r = mk_gemm(m->Hp[d] + o * m->ldx, m->ldx, Whh, 4 * H, m->R[d], 4 * H, n, 4 * H, H);
        CKI(gemm(m, sd, false, true, r));
      }
      long nthr = (long)n * (H / 4);
      k_lstm_cell_fwd<<<(unsigned)((nthr + 127) / 128), 128, 0, sd>>>(
          m->Z[d] + o * 4 * H, k > 0 ? m->R[d] : nullptr, k > 0 ? m->Cc[d] + (long)m->off[k - 1] * H : nullptr, m->Cc[d] + o * H,
          m->Hx[d] + o * H, m->Hp[d] + (long)m->off[k + 1] * m->ldx, m->ldx, n, n_next, H, m->round_ops);
      LAUNCHED(m);
    }
  }
  CK(cudaEventRecord(m->ev_join, m->aux));
  CK(cudaStreamWaitEvent(st, m->ev_join, 0));
  PH_END(m, PH_REC_FWD);
  return 0;
}

static Drop mk_drop(uint64_t seed, uint32_t stream, float keep, int64_t gid0) { Drop d; d.seed = seed; d.stream = stream; d.keep = keep; d.row_gid0 = gid0; return d; }

// The box half of a factorised affinity head's layer 1 (V = Xb W1[Dm:D0], one row per distinct box) needs no LSTM state: it is
// issued at the start of the step on the head's box stream and runs beside the recurrence; k_pair_combine waits for it.
// No split-K here (the forward pass stays free of float atomics: bit-reproducible predictions).
static int heads_prefetch(icl_model* m) {
  for (size_t hi = 0; hi < m->heads.size(); hi++) {
    Head& h = m->heads[hi];
    if (!h.active || !h.fact) continue;
    const int w1 = h.dims[1];
    CK(cudaEventRecord(h.ev_bfork, m->stream));
    CK(cudaStreamWaitEvent(h.sb, h.ev_bfork, 0));
    k_gather_concat<<<dim3(h.Nu, h.slots_b.n_slots), 96, 0, h.sb>>>(h.slots_b, m->Hx[0], m->Hx[1], mk_layout(m), m->H, m->T_cap, h.ldb,
                                                                     mk_drop(0, 0, 1.0f, 0), m->round_ops, h.Xb);
    LAUNCHED(m);
    const Param& pw = m->params[h.pW[0]];
    GemmArgs g = mk_gemm(h.Xb, h.ldb, wbase(m) + pw.off + (int64_t)h.Dm * w1, w1, h.V, w1, h.Nu, w1, h.Db);
    CKI(gemm(m, h.sb, false, true, g));
    CK(cudaEventRecord(h.ev_box, h.sb));
  }
  return 0;
}

static int heads_forward(icl_model* m, float keep, uint64_t seed) {
  const cudaStream_t st0 = m->stream;
  int H = m->H;
  PH_BEGIN(m, PH_HEADS_FWD);
  const bool fork = m->heads.size() > 1 && m->heads[0].hs != nullptr;
  if (fork) CK(cudaEventRecord(m->ev_hfork, st0));
  for (size_t hi = 0; hi < m->heads.size(); hi++) {
    Head& h = m->heads[hi];
    if (!h.active) continue;
    const cudaStream_t st = fork ? h.hs : st0;
    if (fork) CK(cudaStreamWaitEvent(st, m->ev_hfork, 0));
    int B = h.c.batch_size, C = h.c.n_classes, L = h.c.n_hidden;
    const float* in = h.bi;
    int k0 = 0;
    if (h.fact) {          // layer 1 = U[mention] + V[box] + b1, activation, dropout (see Head::fact)
      const int w1 = h.dims[1];
      k_gather_concat<<<dim3(h.Mu, h.slots_m.n_slots), 96, 0, st>>>(h.slots_m, m->Hx[0], m->Hx[1], mk_layout(m), H, m->T_cap, h.ldm,
                                                                    mk_drop(seed, 0, keep, m->seq_gid0), m->round_ops, h.Xm);
      LAUNCHED(m);
      GemmArgs gu = mk_gemm(h.Xm, h.ldm, wbase(m) + m->params[h.pW[0]].off, w1, h.U, w1, h.Mu, w1, h.Dm);
      CKI(gemm(m, st, false, true, gu));
      CK(cudaStreamWaitEvent(st, h.ev_box, 0));
      Epilogue e; memset(&e, 0, sizeof(e));
      e.mode = EPI_BIAS_ACT_DROP; e.bias = m->P + m->params[h.pB[0]].off; e.act = h.c.activation;
      e.drop = mk_drop(seed, STREAM_HEAD + (uint32_t)hi * 8 + 0, keep, m->ex_gid0);
      k_pair_combine<<<B, 128, 0, st>>>(h.U, h.V, h.d_m_of, h.d_b_of, w1, e, h.act[0]);
      LAUNCHED(m);
      in = h.act[0];
      k0 = 1;
    } else {
      k_gather_concat<<<dim3(B, h.slots.n_slots), 96, 0, st>>>(h.slots, m->Hx[0], m->Hx[1], mk_layout(m), H, m->T_cap, h.ldbi,
                                                               mk_drop(seed, 0, keep, m->seq_gid0), m->round_ops, h.bi);
      LAUNCHED(m);
    }
    for (int k = k0; k < L; k++) {
      const Param& pw = m->params[h.pW[k]];
      GemmArgs g = mk_gemm(in, k == 0 ? h.ldbi : h.dims[k], wbase(m) + pw.off, h.dims[k + 1], h.act[k], h.dims[k + 1], B, h.dims[k + 1], h.dims[k]);
      g.epi.mode = EPI_BIAS_ACT_DROP; g.epi.bias = m->P + m->params[h.pB[k]].off; g.epi.act = h.c.activation;
      g.epi.drop = mk_drop(seed, STREAM_HEAD + (uint32_t)hi * 8 + k, keep, m->ex_gid0);
      CKI(gemm(m, st, false, true, g, -1, 1, false, m->pdl));      // programmatic dependent launch: the prologue overlaps the previous layer's tail
      in = h.act[k];
    }
    float scale = h.c.weighted_classes ? 1.0f / B : 1.0f;   // "weighted" as executed == mean CE (core.py:244-267)
    k_softmax_ce<<<(B * 32 + 127) / 128, 128, 0, st>>>(in, h.dims[L], m->P + m->params[h.pW[L]].off, m->P + m->params[h.pB[L]].off, C,
                                                      h.has_labels ? h.labels : nullptr, B, scale, h.loss_w, h.proba, h.pred, h.row_loss,
                                                      h.row_correct, h.dlogits);
    LAUNCHED(m);
    if (h.has_labels) {      // loss sum, accuracy mean: nothing on the device waits for them -- beside the backward pass, joined at the end of the call
      CK(cudaEventRecord(m->ev_side, st));
      CK(cudaStreamWaitEvent(m->aux2, m->ev_side, 0));
      k_reduce_sum2<<<2, 256, 0, m->aux2>>>(h.row_loss, h.row_correct, B, h.scalars, (float)B); LAUNCHED(m);
      CK(cudaEventRecord(m->ev_loss, m->aux2));
      m->loss_pending = true;
    }
    if (fork) { CK(cudaEventRecord(h.ev_done, st)); CK(cudaStreamWaitEvent(st0, h.ev_done, 0)); }
  }
  PH_END(m, PH_HEADS_FWD);
  return 0;
}

static int colsum(icl_model* m, cudaStream_t st, const float* X, long rows, int N, long ld, float* out) {   // out pre-zeroed
  int rpb = 64;
  dim3 grid((N + 127) / 128, (unsigned)((rows + rpb - 1) / rpb));
  k_colsum_atomic<<<grid, 128, 0, st>>>(X, rows, N, ld, out, rpb);
  LAUNCHED(m);
  return 0;
}

// which: 1 = the flat gradient buffer, 2 = dHout of both directions + the dc carry, 3 = both
static int zero_backward_buffers(icl_model* m, cudaStream_t st, int max_blocks = 0, int which = 3) {
  const int H = m->H;
  const size_t dh = (which & 2) ? (size_t)m->NtokP * H * 4 : 0, dc = (which & 2) ? (size_t)m->S * H * 4 : 0, g = (which & 1) ? (size_t)m->n_params * 4 : 0;
  if (dh + dc + g == 0) return 0;
  // narrow side-stream fills (max_blocks < 0 = sized here): a 512-thread CTA clears ~14 GB/s; 20 CTAs (the SMs a persistent recurrence
  // leaves idle) or as many as clear the lot in ~200 us, i.e. inside the recurrence they run beside.  Measured beside K2 against the
  // full-width launch (the default, ICL_ZERO_BLOCKS=0): nonvis512 -1.2 %, card2048 -0.4 %, but rel_cross512 +4.5 %, multitask512 +2.7 %
  // (their larger fills outlast the recurrence when narrow): not the default.
  if (max_blocks < 0) max_blocks = (int)std::min<size_t>(296, std::max<size_t>(20, (2 * dh + 2 * dc + g) / (3u << 20)));
  CK(zero_multi_async({{m->dHout[0], dh}, {m->dHout[1], dh}, {m->G, g}, {m->dcc[0], dc}, {m->dcc[1], dc}}, st, max_blocks));
  LAUNCHED(m);
  return 0;
}
// Nothing touches the backward pass's accumulators before heads_backward: a training step clears them on a side stream while the
// forward pass runs (ICL_ZERO_EARLY: 1 = beside the recurrence, 2 = beside the heads' forward chain, 0 = in front of heads_backward)
static int zero_backward_early(icl_model* m) {
  CK(cudaEventRecord(m->ev_zfork, m->stream));
  CK(cudaStreamWaitEvent(m->aux3, m->ev_zfork, 0));
  CKI(zero_backward_buffers(m, m->aux3, m->zero_blocks, m->dh_clean ? 1 : 3));
  CK(cudaEventRecord(m->ev_zero, m->aux3));
  m->zero_pending = true;
  return 0;
}
static int heads_backward(icl_model* m, float keep, uint64_t seed) {
  const cudaStream_t st0 = m->stream;
  int H = m->H;
  PH_BEGIN(m, PH_HEADS_BWD);
  const bool fork = m->heads.size() > 1 && m->heads[0].hs != nullptr;
  for (size_t hi = 0; hi < m->heads.size(); hi++)
    if (m->heads[hi].active && !m->heads[hi].has_labels) return fail("backward needs labels for head %zu", hi);
  // ONE launch clears everything the backward pass accumulates into: dHout of both directions (span scatter), the WHOLE flat
  // gradient buffer (split-K weight gradients and bias column sums of every head and of the LSTM add into it; heads that are not
  // fed keep a zero gradient) and the dc carry of the BPTT
  // -- issued at the START of a training step on a side stream (zero_backward_buffers): 70 MB of fills beside the forward pass
  if (m->zero_pending) { CK(cudaStreamWaitEvent(st0, m->ev_zero, 0)); m->zero_pending = false; }
  else CKI(zero_backward_buffers(m, st0, 0, m->dh_clean ? 1 : 3));
  m->dh_clean = false;        // the span scatters below write dHout; lstm_backward clears it again behind the recurrence
  if (fork) CK(cudaEventRecord(m->ev_hfork, st0));
  for (size_t hi = 0; hi < m->heads.size(); hi++) {
    Head& h = m->heads[hi];
    if (!h.active) continue;                  // not fed in this call: its parameters keep the zero gradient
    // st: this head's dz chain; sa: its weight / bias gradients (span scatters of concurrent heads meet in dHout as float atomics)
    const cudaStream_t st = fork ? h.hs : st0, sa = fork ? h.ha : m->aux2;
    if (fork) CK(cudaStreamWaitEvent(st, m->ev_hfork, 0));
    int B = h.c.batch_size, L = h.c.n_hidden;
    // Layer by layer from the softmax down: the chain dz_k -> dz_{k-1} = (dz_k W_k^T) * act' stays on the main stream; the weight
    // and bias gradients of layer k only need dz_k, so they run on the aux stream concurrently with the rest of the chain
    // (joined before the weight-gradient phase of the LSTM / the update).
    // The softmax layer (N = n_classes, 2..12 columns) never goes near a GEMM: ONE fused kernel computes its weight and bias
    // gradients and dz of the last hidden layer (k_softmax_bwd).
    {
      const Param &pw = m->params[h.pW[L]], &pb = m->params[h.pB[L]];
      const int Kl = h.dims[L], C = h.dims[L + 1], nb = (B + SMB_ROWS - 1) / SMB_ROWS;
      Epilogue e; memset(&e, 0, sizeof(e));
      e.mode = EPI_DACT; e.act = h.c.activation; e.aux = h.act[L - 1]; e.ldaux = Kl;
      e.drop = mk_drop(seed, STREAM_HEAD + (uint32_t)hi * 8 + (L - 1), keep, m->ex_gid0); e.round_out = m->round_ops;
      k_softmax_bwd<<<nb, SMB_THREADS, (size_t)(SMB_ROWS * C + Kl * (C + 1)) * 4, st>>>(h.act[L - 1], h.dlogits, m->P + pw.off, B, Kl, C, e,
                                                                                         h.dzb[L - 1], h.smb_part);
      LAUNCHED(m);
      // its partial sums only feed the softmax layer's own weight / bias gradient: reduced on the aux stream like the other layers' dW
      CK(cudaEventRecord(m->ev_side, st));
      CK(cudaStreamWaitEvent(sa, m->ev_side, 0));
      k_softmax_bwd_reduce<<<(Kl * C + C + 31) / 32, dim3(32, 16), 0, sa>>>(h.smb_part, nb, Kl * C, C, m->G + pw.off, m->G + pb.off);
      LAUNCHED(m);
      m->heads_aux_pending = true;
    }
    const float* dz = h.dzb[L - 1];    // gradient w.r.t. the pre-activation of layer k+1
    for (int k = L - 1; k >= 0; k--) {
      const float* in = k == 0 ? h.bi : h.act[k - 1];
      int din = h.dims[k], dout = h.dims[k + 1];
      const Param& pw = m->params[h.pW[k]];
      CK(cudaEventRecord(m->ev_dz[k], st));
      CK(cudaStreamWaitEvent(sa, m->ev_dz[k], 0));
      if (k == 0 && h.fact) {
        // factorised layer 1 in reverse: dU[g] / dV[g] = sums of dz1 over the pairs of a mention / a box (ordered, no atomics), then
        //   dW1[0:Dm] = Xm^T dU,  dW1[Dm:D0] = Xb^T dV  (contractions over a few dozen groups instead of the batch),
        //   d(mention columns) = dU W1[0:D0g]^T -> span scatter, once per distinct mention
        CKI(colsum(m, sa, dz, B, dout, dout, m->G + m->params[h.pB[0]].off));
        const float* dU = dz;                      // no mention repeats (Mu == B): the groups are the pairs
        if (h.Mu != B) {
          k_segment_sum<<<h.Mu, 128, 0, st>>>(dz, dout, h.d_m_start, h.d_m_mem, m->round_ops, h.dU); LAUNCHED(m);
          CK(cudaEventRecord(h.ev_du, st));
          CK(cudaStreamWaitEvent(sa, h.ev_du, 0));
          dU = h.dU;
        }
        GemmArgs gm = mk_gemm(h.Xm, h.ldm, dU, dout, m->G + pw.off, dout, h.Dm, dout, h.Mu);
        CKI(gemm(m, sa, true, true, gm, -1, 0, true));
        k_segment_sum<<<h.Nu, 128, 0, sa>>>(dz, dout, h.d_b_start, h.d_b_mem, m->round_ops, h.dV); LAUNCHED(m);
        GemmArgs gb = mk_gemm(h.Xb, h.ldb, h.dV, dout, m->G + pw.off + (int64_t)h.Dm * dout, dout, h.Db, dout, h.Nu);
        CKI(gemm(m, sa, true, true, gb, -1, 0, true));
        m->heads_aux_pending = true;
        if (h.D0g > 0) {
          GemmArgs gx = mk_gemm(dU, dout, wbase(m) + pw.off, dout, h.dXm, h.lddbi, h.Mu, h.D0g, dout);
          CKI(gemm(m, st, false, false, gx));
          k_scatter_spans<<<dim3(h.Mu, h.slots_m.n_slots), 96, 0, st>>>(h.slots_m, h.dXm, mk_layout(m), H, m->T_cap, h.lddbi,
                                                                        mk_drop(seed, 0, keep, m->seq_gid0), m->dHout[0], m->dHout[1]);
          LAUNCHED(m);
        }
        continue;
      }
      // dW = in^T * dz   (contraction over the batch: both operands MN-major)
      GemmArgs gw = mk_gemm(in, k == 0 ? h.ldbi : din, dz, dout, m->G + pw.off, dout, din, dout, B);
      CKI(gemm(m, sa, true, true, gw, -1, 0, true));
      CKI(colsum(m, sa, dz, B, dout, dout, m->G + m->params[h.pB[k]].off));
      m->heads_aux_pending = true;
      // d(in) = dz * W^T  (W [din,dout] row-major is K-major as the B operand)
      if (k > 0) {
        GemmArgs gx = mk_gemm(dz, dout, wbase(m) + pw.off, dout, h.dzb[k - 1], din, B, din, dout);
        gx.epi.mode = EPI_DACT; gx.epi.act = h.c.activation; gx.epi.aux = h.act[k - 1]; gx.epi.ldaux = din;
        gx.epi.drop = mk_drop(seed, STREAM_HEAD + (uint32_t)hi * 8 + (k - 1), keep, m->ex_gid0);
        gx.epi.round_out = m->round_ops;
        CKI(gemm(m, st, false, false, gx, -1, 1, false, m->pdl));
        dz = h.dzb[k - 1];
      } else if (h.D0g > 0) {
        GemmArgs gx = mk_gemm(dz, dout, wbase(m) + pw.off, dout, h.dbi, h.lddbi, B, h.D0g, dout);
        CKI(gemm(m, st, false, false, gx, -1, 1, false, m->pdl));
        k_scatter_spans<<<dim3(B, h.slots.n_slots), 96, 0, st>>>(h.slots, h.dbi, mk_layout(m), H, m->T_cap, h.lddbi, mk_drop(seed, 0, keep, m->seq_gid0),
                                          m->dHout[0], m->dHout[1]);
        LAUNCHED(m);
      }
    }
    if (fork) {      // chain -> the model's stream; weight gradients -> aux2, where ev_heads marks "every head gradient complete"
      CK(cudaEventRecord(h.ev_done, st)); CK(cudaStreamWaitEvent(st0, h.ev_done, 0));
      CK(cudaEventRecord(h.ev_adone, sa)); CK(cudaStreamWaitEvent(m->aux2, h.ev_adone, 0));
    }
  }
  PH_END(m, PH_HEADS_BWD);
  CK(cudaEventRecord(m->ev_heads, m->heads_aux_pending ? m->aux2 : st0));   // the head gradients can be all-reduced while the BPTT runs
  return 0;
}
// the heads' weight gradients (aux stream) must be complete before anything reads the gradient buffer on the main stream
static int join_heads_aux(icl_model* m) {
  if (m->heads_aux_pending) { CK(cudaStreamWaitEvent(m->stream, m->ev_heads, 0)); m->heads_aux_pending = false; }
  return 0;
}

#ifdef ICL_EXPERIMENTS
// K3 (opt-in): every BPTT step of both directions in one cooperative launch (lstm_persistent.cuh)
static int rec_backward_persistent(icl_model* m) {
  const int H = m->H, S = m->S;
  cudaStream_t st = m->stream;
  for (int d = 0; d < 2; d++) {
    CK(zero_async(m->dhrec[d], (size_t)S * H * 4, st));
    CK(zero_async(m->dcc[d], (size_t)S * H * 4, st));
  }
  CK(zero_async(m->rp_bar, 4, st));
  RecBwdArgs a;
  a.off = m->d_off; a.nact = m->d_nact; a.Tmax = m->Tmax; a.H = H; a.round_ops = m->round_ops; a.bar = m->rp_bar;
  a.trace = m->rp_trace; a.trace_cta = m->rp_trace_cta;
  for (int d = 0; d < 2; d++) { a.Z[d] = m->Z[d]; a.Cc[d] = m->Cc[d]; a.dHout[d] = m->dHout[d]; a.dhrec[d] = m->dhrec[d]; a.dcc[d] = m->dcc[d]; }
  void* args[] = {(void*)&m->rp_bmaps, (void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_rec_bwd, dim3(m->n_sms), dim3(RB_THREADS), args, RB_SMEM, st);
  if (e != cudaSuccess) return fail("k_rec_bwd launch failed: %s", cudaGetErrorString(e));
  m->launches++;
  return 0;
}

#endif

// K3 (default in tensor-core mode): one fused launch per step for both directions -- dh_rec = dZ_{k+1} W_hh^T with the
// contraction split over a thread-block cluster, reduced through distributed shared memory, cell backward in the epilogue
static int rec_backward_fused(icl_model* m) {
  const int H = m->H, S = m->S, cs = m->bp_cs;
  cudaStream_t st = m->stream;
  for (int d = 0; d < 2; d++) CK(zero_async(m->dcc[d], (size_t)S * H * 4, st));
  const int Nt = (H + BP_BN - 1) / BP_BN;
  for (int k = m->Tmax - 1; k >= 0; k--) {
    BpttArgs a;
    for (int d = 0; d < 2; d++) { a.Z[d] = m->Z[d]; a.Cc[d] = m->Cc[d]; a.dHout[d] = m->dHout[d]; a.dcc[d] = m->dcc[d]; }
    a.H = H; a.k = k; a.o_k = m->off[k]; a.o_kp1 = m->off[k + 1]; a.o_km1 = k > 0 ? m->off[k - 1] : 0;
    a.n_k = m->n_active[k]; a.n_kp1 = k + 1 < m->Tmax ? m->n_active[k + 1] : 0;
    a.n_km1 = k > 0 ? m->n_active[k - 1] : 0; a.o_km2 = k > 1 ? m->off[k - 2] : 0;
    a.round_ops = m->round_ops; a.cs = cs;
    { static const char* e = getenv("ICL_BPTT_DBG"); a.dbg = e ? atoi(e) : 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(Nt * cs), (unsigned)((a.n_k + 127) / 128), 2);
    cfg.blockDim = dim3(BP_THREADS); cfg.dynamicSmemBytes = BP_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    cudaError_t e = cs == 4 ? cudaLaunchKernelEx(&cfg, k_bptt_step<4>, m->bp_maps, a)
                    : cs == 2 ? cudaLaunchKernelEx(&cfg, k_bptt_step<2>, m->bp_maps, a)
                              : cudaLaunchKernelEx(&cfg, k_bptt_step<1>, m->bp_maps, a);
    if (e != cudaSuccess) return fail("k_bptt_step launch failed (step %d): %s", k, cudaGetErrorString(e));
    m->launches++;
  }
  return 0;
}

// K3 (default in tensor-core mode, H <= 336): the whole backward recurrence in ONE launch, a cluster of 4 CTAs per
// (direction, 128-row tile) walking all of the tile's time steps with cluster barriers only (lstm_bptt.cuh)
static int rec_backward_cluster(icl_model* m) {
  const int H = m->H;
  cudaStream_t st = m->stream;                  // (the dc carry was cleared with the other backward buffers, heads_backward)
  BpttClusterArgs a;
  for (int d = 0; d < 2; d++) { a.Z[d] = m->Z[d]; a.Cc[d] = m->Cc[d]; a.dHout[d] = m->dHout[d]; a.dcc[d] = m->dcc[d]; }
  a.off = m->d_off; a.nact = m->d_nact; a.H = H; a.Tmax = m->Tmax; a.round_ops = m->round_ops;
  a.trace = m->rp_trace; a.trace_cta = m->rp_trace_cta;
  // The dependent chain of a tile's time steps bounds the phase, and a step is the faster the more CTAs share the tile (measured
  // 12 / 20 / 35 us per step with 8 / 4 / 2 CTAs), but everything must be co-resident on the 148 SMs to run concurrently.  Tiles
  // are ordered by chain length (the longest sequences come first), so: the first n8 tiles get 8-CTA clusters, the last n2 tiles
  // (a handful of steps each) 2-CTA clusters, the rest 4 -- the split that minimises the longest chain time under the SM budget.
  // Three concurrent launches (clusters are independent of each other).
  // Tile shapes.  The first s 128-row blocks (the longest chains) may be cut into 64-row tiles on 8-CTA clusters: twice the
  // independent chains, half the park / pull / cell work per step (12 -> 9 us), for 16 CTAs per block and direction instead of 8.
  // Batches of <= 512 sequences cut everything (2 x 8 x 8 = 128 CTAs); larger batches have no CTAs to spare.  The search below minimises the longest chain time over (s, n8, n4, n2) under the SM budget.
  const int blocks = (m->n_active[0] + 127) / 128;
  auto chain = [&](int row) { int k = 0; while (k < m->Tmax && (m->n_active[k] + 127) / 128 * 128 > row) k++; return k; };   // steps of the tile starting at `row`
  int ns = 0, n8 = 0, n2 = 0;
  {
    const double t64 = 8.7, t8 = 12.0, t4 = 20.0, t2 = 35.0;       // measured us per step
    double best = 1e30;
    int best_ctas = 1 << 30;
    const int budget = m->n_sms - 8;                 // a few SMs of slack: exactly 148 CTAs measured 0.41 ms, 140 CTAs 0.36 ms
    const int force_tm = getenv("ICL_BPTT_TM") ? atoi(getenv("ICL_BPTT_TM")) : 0;      // 128: never cut, 64: cut every block
    const bool mixed = getenv("ICL_BPTT_MIXED") && atoi(getenv("ICL_BPTT_MIXED")) != 0;
    for (int s = 0; s <= blocks; s++) {
      if ((force_tm == 128 && s != 0) || (force_tm == 64 && s != blocks)) continue;
      // all or nothing unless asked for (ICL_BPTT_MIXED=1): the model prefers cutting only the first block of a 1024-sequence batch
      // (rel_cross512: 204 vs 240 us), measured it is slower (rec_bwd 0.214 -> 0.259 ms with 136 instead of 104 CTAs)
      if (s != 0 && s != blocks && !mixed) continue;
      const int rest = blocks - s;
      for (int a8 = 0; a8 <= rest; a8++)
        for (int a2 = 0; a8 + a2 <= rest; a2++) {
          const int a4 = rest - a8 - a2, ctas = 2 * (16 * s + 8 * a8 + 4 * a4 + 2 * a2);
          if (ctas > budget && !(s == 0 && a8 == 0 && a2 == rest)) continue;
          const double cost = std::max({s ? chain(0) * t64 : 0.0, a8 ? chain(128 * s) * t8 : 0.0, a4 ? chain(128 * (s + a8)) * t4 : 0.0,
                                        a2 ? chain(128 * (s + a8 + a4)) * t2 : 0.0});
          if (cost < best - 1e-9 || (cost < best + 1e-9 && ctas < best_ctas)) { best = cost; best_ctas = ctas; ns = s; n8 = a8; n2 = a2; }
        }
    }
  }
  if (const char* e = getenv("ICL_BPTT_CLUSTER_CS")) { ns = 0; n2 = 0; n8 = atoi(e) == 8 ? blocks : atoi(e) == 4 ? 0 : n8; }
  if (const char* e = getenv("ICL_BPTT_N8")) { n8 = std::max(0, std::min(blocks - ns, atoi(e))); n2 = std::min(n2, blocks - ns - n8); }
  if (const char* e = getenv("ICL_BPTT_N2")) n2 = std::max(0, std::min(blocks - ns - n8, atoi(e)));
  const int n4 = blocks - ns - n8 - n2;
  static const int trace_cs = getenv("ICL_TRACE_CS") ? atoi(getenv("ICL_TRACE_CS")) : 8;     // bring-up trace: one of the launches
  auto launch = [&](int cs, int tm, int row0, int ntiles, cudaStream_t s) -> int {
    a.row0 = row0; a.tm = tm;
    a.trace = (cs == trace_cs && tm == 128) ? m->rp_trace : nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs, (unsigned)(2 * ntiles), 1);
    cfg.blockDim = dim3(BC_THREADS); cfg.dynamicSmemBytes = BC_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cs == 8   ? cudaLaunchKernelEx(&cfg, k_bptt_cluster<8>, m->bp_maps, a)
                    : cs == 4 ? cudaLaunchKernelEx(&cfg, k_bptt_cluster<4>, m->bp_maps, a)
                              : cudaLaunchKernelEx(&cfg, k_bptt_cluster<2>, m->bp_maps, a);
    if (e != cudaSuccess) return fail("k_bptt_cluster<%d> launch failed: %s", cs, cudaGetErrorString(e));
    m->launches++;
    return 0;
  };
  // the group with the longest chains stays on the main stream; the others fork to the side streams and join
  struct Grp { int cs, tm, row0, n; } grp[4] = {{8, 64, 0, 2 * ns}, {8, 128, 128 * ns, n8}, {4, 128, 128 * (ns + n8), n4}, {2, 128, 128 * (ns + n8 + n4), n2}};
  cudaStream_t side[3] = {m->aux, m->aux3, m->aux4};
  cudaEvent_t join[3] = {m->ev_join, m->ev_join2, m->ev_join3};
  int used = 0, first = -1;
  for (int i = 0; i < 4; i++) if (grp[i].n > 0 && first < 0) first = i;
  CK(cudaEventRecord(m->ev_fork, st));
  for (int i = 0; i < 4; i++) {
    if (grp[i].n == 0 || i == first) continue;
    CK(cudaStreamWaitEvent(side[used], m->ev_fork, 0));
    CKI(launch(grp[i].cs, grp[i].tm, grp[i].row0, grp[i].n, side[used]));
    CK(cudaEventRecord(join[used], side[used]));
    used++;
  }
  CKI(launch(grp[first].cs, grp[first].tm, grp[first].row0, grp[first].n, st));
  for (int i = 0; i < used; i++) CK(cudaStreamWaitEvent(st, join[i], 0));
  return 0;
}


#ifdef ICL_EXPERIMENTS
// K3 experiment (ICL_BPTT_NSPLIT=1 in an -DICL_EXPERIMENTS build): k_bptt_nsplit (lstm_bptt_ns.cuh), ONE launch: an 8-CTA cluster per (direction,
// 128-row tile), the clusters in chain-length order (the longest sequences first; a batch with more than 18 clusters runs in waves)
template <int UN> static int ns_launch(icl_model* m, const BpttNsArgs& a, int tiles, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(BN_CS, (unsigned)(2 * tiles), 1);
  cfg.blockDim = dim3(BN_THREADS); cfg.dynamicSmemBytes = BNS<UN>::SMEM; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = BN_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_bptt_nsplit<UN>, m->ns_maps, a);
  if (e != cudaSuccess) return fail("k_bptt_nsplit<%d> launch failed: %s", UN, cudaGetErrorString(e));
  m->launches++;
  return 0;
}
static int rec_backward_nsplit(icl_model* m) {
  const int E = m->E, H = m->H, UN = m->ns_UN;
  cudaStream_t st = m->stream;
  if (m->wb_dirty) {
    const float* W0 = m->P + m->params[m->pK[0]].off + (size_t)E * 4 * H;
    const float* W1 = m->P + m->params[m->pK[1]].off + (size_t)E * 4 * H;
    const int NP = (UN + 15) / 16 * 16;
    k_pack_whh_bwd16<<<dim3(148, 2), 256, 0, st>>>(W0, W1, m->Wb16[0], m->Wb16[1], H, UN, NP, BN_CS * 4 * UN);
    LAUNCHED(m);
    m->wb_dirty = false;
  }
  BpttNsArgs a;
  for (int d = 0; d < 2; d++) { a.Z[d] = m->Z[d]; a.Cc[d] = m->Cc[d]; a.dHout[d] = m->dHout[d]; a.dZ16[d] = m->dZ16[d]; a.S16[d] = m->S16[d]; }
  a.off = m->d_off; a.nact = m->d_nact; a.H = H; a.Tmax = m->Tmax; a.round_ops = m->round_ops; a.tile0 = 0;
  a.trace = m->rp_trace; a.trace_cta = m->rp_trace_cta;
  const int tiles = (m->n_active[0] + 127) / 128;
  switch (UN) {
    case 8: return ns_launch<8>(m, a, tiles, st);
    case 16: return ns_launch<16>(m, a, tiles, st);
    case 28: return ns_launch<28>(m, a, tiles, st);
    default: return ns_launch<40>(m, a, tiles, st);
  }
}

#else
static int rec_backward_nsplit(icl_model*) { return fail("k_bptt_nsplit is not in this build"); }
#endif

// K3 (fp32 validation mode): one cell kernel + one split-K GEMM per step and direction, directions interleaved on two streams
static int rec_backward_steps(icl_model* m) {
  const int E = m->E, H = m->H, S = m->S;
  cudaStream_t st = m->stream;
  CK(cudaEventRecord(m->ev_fork, st));
  CK(cudaStreamWaitEvent(m->aux, m->ev_fork, 0));
  for (int d = 0; d < 2; d++) {
    cudaStream_t sd = d ? m->aux : st;
    CK(zero_async(m->dhrec[d], (size_t)S * H * 4, sd));
    CK(zero_async(m->dcc[d], (size_t)S * H * 4, sd));
  }
  for (int k = m->Tmax - 1; k >= 0; k--) {
    for (int d = 0; d < 2; d++) {        // directions interleaved on two streams so the host feeds both concurrently
      cudaStream_t sd = d ? m->aux : st;
      const float* Whh = wbase(m) + m->params[m->pK[d]].off + (size_t)E * 4 * H;
      int n = m->n_active[k];
      long o = m->off[k];
      long nthr = (long)n * (H / 4);
      {   // cell backward of step k; launched with PDL so it is resident before the GEMM of step k+1 drains
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((nthr + 127) / 128)); cfg.blockDim = dim3(128); cfg.stream = sd;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        float* zk = m->Z[d] + o * 4 * H; const float* ck = m->Cc[d] + o * H;
        const float* cprev = k > 0 ? m->Cc[d] + (long)m->off[k - 1] * H : nullptr;
        const float* dhk = m->dHout[d] + o * H;
        CK(cudaLaunchKernelEx(&cfg, k_lstm_cell_bwd, zk, ck, cprev, dhk, m->dhrec[d], m->dcc[d], n, H, m->round_ops));
        LAUNCHED(m);
      }
      if (k > 0) {   // dh_{k-1} = dZ_k * W_hh^T for the running rows
        GemmArgs r = mk_gemm(m->Z[d] + o * 4 * H, 4 * H, Whh, 4 * H, m->dhrec[d], H, n, H, 4 * H);
        int tiles = ((n + 127) / 128) * ((H + 127) / 128);
        // split-K (about one wave per direction) accumulating into dhrec, which the cell kernel left zeroed
        CKI(gemm(m, sd, false, false, r, -1, std::max(1, std::min(4, 148 / tiles)), true, m->cfg.gemm_mode == ICL_GEMM_TCGEN05_TF32));
      }
    }
  }
  CK(cudaEventRecord(m->ev_join, m->aux));
  CK(cudaStreamWaitEvent(st, m->ev_join, 0));
  return 0;
}

static int lstm_backward(icl_model* m) {
  const int E = m->E, H = m->H;
  const long Ntok = m->NtokP;            // time-batched GEMMs run over the padded row space (pad rows are zero)
  cudaStream_t st = m->stream;
  if (m->Ntok == 0) return 0;
  PH_BEGIN(m, PH_REC_BWD);
  if (m->ns_on && m->Tmax <= RP_MAXT) CKI(rec_backward_nsplit(m));                       // all three write the pad rows of dZ as zeros
  else if (m->bp_on && m->bp_cluster && m->Tmax <= RP_MAXT) CKI(rec_backward_cluster(m));
  else if (m->bp_on) CKI(rec_backward_fused(m));
  else {
#ifdef ICL_EXPERIMENTS
    if (m->rp_U != 0 && m->rp_on && m->rp_bwd_on && m->Tmax <= RP_MAXT) CKI(rec_backward_persistent(m));
    else
#endif
      CKI(rec_backward_steps(m));
    // pad rows of dZ must be exactly zero for the time-batched weight-gradient GEMM (they still hold gates / Zx)
    k_zero_pad_rows<<<dim3(m->Tmax, 2), 256, 0, st>>>(m->Z[0], m->Z[1], mk_layout(m), 4 * H); LAUNCHED(m);
  }
  PH_END(m, PH_REC_BWD);
  if (m->zero_early == 3) {
    // ICL_ZERO_EARLY=3 (measured, not the default): dHout and the dc carry are dead once the recurrence has run, so they can be
    // cleared HERE, beside the tensor-bound weight-gradient GEMMs, for the next step (all rows are zero between steps; whatever
    // fails before this point leaves dh_clean false and the next backward pass clears them itself), leaving only the flat gradient
    // buffer to clear beside the next step's recurrence.  card2048 1.240 -> 1.235 ms, but multitask512 (100 MB of fills beside two
    // short GEMMs) 1.34 -> 1.37-1.40 ms, whatever the width of the fill (20 / 40 / auto / 1480 CTAs).
    CK(cudaEventRecord(m->ev_zfork, st));
    CK(cudaStreamWaitEvent(m->aux3, m->ev_zfork, 0));
    CKI(zero_backward_buffers(m, m->aux3, m->zero_blocks, 2));
    CK(cudaEventRecord(m->ev_dhzero, m->aux3));
  }
  // time-batched weight gradients: ONE split-K GEMM per direction (contraction over all tokens)
  CKI(join_heads_aux(m));
  PH_BEGIN(m, PH_WGRAD);
  // 128x256 tiles: ceil((E+H+1)/128) x ceil(4H/256) of them; split-K so that ~one wave of 148 CTAs covers the contraction
  int tiles = ((E + H + 1 + 127) / 128) * ((4 * H + 255) / 256);
  int splits = (int)std::max<long>(1, std::min<long>(std::min<long>(16, 148 / std::max(1, tiles)), Ntok / 1024));
  for (int d = 0; d < 2; d++) {
    float* dK = m->G + m->params[m->pK[d]].off;
    // [dKernel; dbias] = [Xd | Hprev | 1]^T dZ  (the bias follows the kernel in the flat buffers)
    if (m->params[m->pBias[d]].off != m->params[m->pK[d]].off + (int64_t)(E + H) * 4 * H) return fail("kernel/bias not contiguous");
    GemmArgs gk = mk_gemm(m->XH[d], m->ldx, m->Z[d], 4 * H, dK, 4 * H, E + H + 1, 4 * H, (int)Ntok);
    CKI(gemm(m, st, true, true, gk, -1, splits, /*prezeroed=*/true));
    if (d == 0) CK(cudaEventRecord(m->ev_wg0, st));
  }
  if (m->zero_early == 3) { CK(cudaStreamWaitEvent(st, m->ev_dhzero, 0)); m->dh_clean = true; }      // the fills are joined behind the weight gradients
  PH_END(m, PH_WGRAD);
  return 0;
}

// Select the Adam state (m, v, step) the next updates use; slots are created zeroed on first use.  icl_get/set_tensor kinds 2, 3
// and icl_get/set_step address the selected slot.
extern "C" int icl_set_optimizer_slot(icl_model* m, int slot) {
  if (slot < 0 || slot >= 64) return fail("optimizer slot out of range");
  if (m->slots.empty()) m->slots.resize(1);
  m->slots[m->cur_slot].M = m->M; m->slots[m->cur_slot].V = m->V; m->slots[m->cur_slot].step = m->step;     // park the live one
  while ((int)m->slots.size() <= slot) {
    icl_model::OptSlot o;
    CK(dmalloc(&o.M, (size_t)m->n_params)); CK(dmalloc(&o.V, (size_t)m->n_params));
    CK(cudaMemset(o.M, 0, (size_t)m->n_params * 4)); CK(cudaMemset(o.V, 0, (size_t)m->n_params * 4));
    m->slots.push_back(o);
  }
  m->cur_slot = slot;
  m->M = m->slots[slot].M; m->V = m->slots[slot].V; m->step = m->slots[slot].step;
  return 0;
}
extern "C" int icl_optimizer_slots(icl_model* m) { return std::max<int>(1, (int)m->slots.size()); }

extern "C" int icl_set_loss_weights(icl_model* m, const float* w) {
  for (size_t hi = 0; hi < m->heads.size(); hi++) m->heads[hi].loss_w = w ? w[hi] : 1.0f;
  return 0;
}

static int apply_update(icl_model* m, double extra_sumsq);
extern "C" int icl_apply_update(icl_model* m) { return apply_update(m, 0.0); }
// extra_sumsq: squared norm of gradients the caller keeps on the host (they take part in clip_by_global_norm);
// gnorm_out (may be NULL): the global norm used, read back synchronously
extern "C" int icl_apply_update_ex(icl_model* m, double extra_sumsq, float* gnorm_out) {
  CKI(apply_update(m, extra_sumsq));
  if (gnorm_out) {
    *gnorm_out = 0.f;
    if (m->cfg.clip_norm > 0) {
      CK(cudaMemcpyAsync(gnorm_out, m->d_gnorm, 4, cudaMemcpyDeviceToHost, m->stream));
      CK(cudaStreamSynchronize(m->stream));
    }
  }
  return 0;
}
static int apply_update(icl_model* m, double extra_sumsq) {
  cudaStream_t st = m->stream;
  long n = (long)m->n_params;
  PH_BEGIN(m, PH_UPDATE);
  if (m->cfg.clip_norm > 0) {       // heads that were not fed have an all-zero gradient: the norm over everything is the norm TF takes
    k_sumsq_partial<<<592, 256, 0, st>>>(m->G, n, m->d_partial); LAUNCHED(m);
    k_sumsq_final<<<1, 256, 0, st>>>(m->d_partial, 592, m->d_gnorm, extra_sumsq); LAUNCHED(m);
  }
  m->step++;
  double b1 = m->cfg.beta1, b2 = m->cfg.beta2;
  float lr_t = (float)(m->cfg.learn_rate * std::sqrt(1.0 - std::pow(b2, (double)m->step)) / (1.0 - std::pow(b1, (double)m->step)));
  // variables without a gradient (the heads that were not fed in this step) are not touched by the optimizer -- neither the
  // weights nor their m / v (tf.train.Optimizer skips None gradients): update the contiguous runs of live parameters only
  std::vector<std::pair<int64_t, int64_t>> runs;      // [begin, end) in floats
  auto add_run = [&](int64_t b0, int64_t e0) { if (!runs.empty() && runs.back().second == b0) runs.back().second = e0; else runs.push_back({b0, e0}); };
  add_run(0, m->heads.empty() ? m->n_params : m->params[m->heads[0].pW[0]].off);
  for (size_t hi = 0; hi < m->heads.size(); hi++) {
    const int64_t b0 = m->params[m->heads[hi].pW[0]].off;
    const int64_t e0 = hi + 1 < m->heads.size() ? m->params[m->heads[hi + 1].pW[0]].off : m->n_params;
    if (m->heads[hi].active) add_run(b0, e0);
  }
  for (auto& r : runs) {
    const long len = (long)(r.second - r.first);
    const int blocks = (int)std::max<long>(1, std::min<long>(592, (len / 4 + 255) / 256));
    k_adam<<<blocks, 256, 0, st>>>(m->P + r.first, m->G + r.first, m->M + r.first, m->V + r.first, m->round_ops ? m->Pr + r.first : nullptr,
                                  len, m->d_gnorm, m->cfg.clip_norm, lr_t, (float)b1, (float)b2, m->cfg.adam_epsilon);
    LAUNCHED(m);
  }
  m->wih_dirty = true;
  if (!getenv("ICL_DEBUG_STALE_WP")) m->wp_dirty = true;     // the packed recurrent weights of the persistent forward kernel follow the
  m->wb_dirty = true;                                        // update (the env knob exists so that a test can prove it catches staleness)
  PH_END(m, PH_UPDATE);
  if (!getenv("ICL_DEBUG_STALE_WP")) {
    m->rf20.dirty = m->rf16.dirty = true; m->wp_dirty = false;
    CKI(repack_after_update(m));
  }
  return 0;
}

extern "C" int icl_run_resident(icl_model* m, int op, float keep_in, float keep, uint64_t seed) {
  if (!m->resident) return fail("icl_run_resident: no batch uploaded");
  if (!(keep_in > 0 && keep_in <= 1 && keep > 0 && keep <= 1)) return fail("keep probabilities must be in (0,1]");
  for (int i = 0; i < PH_N; i++) m->ph_used[i] = false;
  InSet& I = m->in[m->cur];
  CK(cudaStreamWaitEvent(m->stream, I.ev_copied, 0));          // the batch's H2D copies (copy stream) have landed
  if (m->ph_on) {                                              // timing events (icl_last_step_ms, icl_debug_timeline): see PH_BEGIN
    CK(cudaEventRecord(I.ev_s0, m->stream));
    CK(cudaEventRecord(m->ev_t0, m->stream));
  }
  m->t_recorded = m->ph_on;
  CKI(refresh_rounded_params(m));
  m->last_keep = keep; m->last_seed = seed;
  CKI(lstm_forward(m, keep_in, seed, op >= ICL_OP_GRADS));
  if (op >= ICL_OP_GRADS && m->zero_early == 2) CKI(zero_backward_early(m));
  CKI(heads_forward(m, keep, seed));
  if (op >= ICL_OP_GRADS) {
    CKI(heads_backward(m, keep, seed));
    CKI(lstm_backward(m));
    CKI(join_heads_aux(m));
  }
  if (op == ICL_OP_TRAIN) CKI(icl_apply_update(m));
  if (m->loss_pending) { CK(cudaStreamWaitEvent(m->stream, m->ev_loss, 0)); m->loss_pending = false; }
  if (m->ph_on) CK(cudaEventRecord(m->ev_t1, m->stream));
  CK(cudaEventRecord(I.ev_done, m->stream));                   // this set's device buffers may be overwritten after this point
  I.done_pending = true;
  return 0;
}

// Asynchronous read-back of the step's scalars: enqueues a D2H copy of (loss, accuracy) of every head of the step just issued
// and hands out the scalars of the PREVIOUS call (NaN when there is none) -- the host never waits for the step in flight.
extern "C" int icl_poll_stats(icl_model* m, icl_head_out* prev) {
  if (m->cur < 0) return fail("icl_poll_stats: no batch uploaded");
  InSet& I = m->in[m->cur];
  float* hs = m->h_stats + (size_t)m->cur * ICL_MAX_HEADS * 2;
  m->d2h_bytes = 0;
  for (size_t hi = 0; hi < m->heads.size(); hi++) {
    CK(cudaMemcpyAsync(hs + hi * 2, m->heads[hi].scalars, 8, cudaMemcpyDeviceToHost, m->stream));
    m->d2h_bytes += 8;
  }
  CK(cudaEventRecord(I.ev_stats, m->stream));
  I.stats_pending = true;
  InSet& Pv = m->in[m->cur ^ 1];
  const float* ps = m->h_stats + (size_t)(m->cur ^ 1) * ICL_MAX_HEADS * 2;
  const bool have = Pv.stats_pending;
  if (have) { CK(cudaEventSynchronize(Pv.ev_stats)); Pv.stats_pending = false; }
  for (size_t hi = 0; hi < m->heads.size() && prev; hi++) {
    prev[hi].loss = have ? ps[hi * 2] : NAN;
    prev[hi].accuracy = have ? ps[hi * 2 + 1] : NAN;
  }
  return 0;
}

// bring-up: device timeline of the two most recent pipelined steps, ms relative to the older step's first copy:
// out[0..3] = older {copy start, copy end, compute start, compute end}, out[4..7] = newer.  Synchronises.
extern "C" int icl_debug_timeline(icl_model* m, float* out) {
  CK(cudaDeviceSynchronize());
  if (m->cur < 0) return fail("no batch");
  InSet &A = m->in[m->cur ^ 1], &B = m->in[m->cur];
  cudaEvent_t ev[8] = {A.ev_c0, A.ev_copied, A.ev_s0, A.ev_done, B.ev_c0, B.ev_copied, B.ev_s0, B.ev_done};
  for (int i = 0; i < 8; i++) if (cudaEventElapsedTime(&out[i], A.ev_c0, ev[i]) != cudaSuccess) { out[i] = NAN; cudaGetLastError(); }
  return 0;
}

// One pipelined train step: upload (copy stream, double-buffered) + forward/BPTT/update on the compute stream + icl_poll_stats.
// Returns as soon as the work is enqueued; errors of the device work surface at the next synchronising call.
extern "C" int icl_train_async(icl_model* m, const icl_batch* b, float keep_in, float keep, uint64_t seed, icl_head_out* prev) {
  CKI(icl_upload(m, b));
  CKI(icl_run_resident(m, ICL_OP_TRAIN, keep_in, keep, seed));
  return icl_poll_stats(m, prev);
}

extern "C" int icl_fetch(icl_model* m, icl_head_out* out) {
  cudaStream_t st = m->stream;
  m->d2h_bytes = 0;
  for (size_t hi = 0; hi < m->heads.size(); hi++) {
    Head& h = m->heads[hi];
    int B = h.c.batch_size, C = h.c.n_classes;
    CK(cudaMemcpyAsync(h.h_out, h.proba, (size_t)B * C * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h.h_out + (size_t)B * C, h.scalars, 2 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h.h_pred, h.pred, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    m->d2h_bytes += (int64_t)B * C * 4 + 8 + (int64_t)B * 8;
  }
  CK(cudaStreamSynchronize(st));
  if (m->t_recorded) cudaEventElapsedTime(&m->last_ms, m->ev_t0, m->ev_t1); else m->last_ms = NAN;
  for (size_t hi = 0; hi < m->heads.size() && out; hi++) {
    Head& h = m->heads[hi];
    int B = h.c.batch_size, C = h.c.n_classes;
    if (out[hi].proba) memcpy(out[hi].proba, h.h_out, (size_t)B * C * 4);
    if (out[hi].pred) memcpy(out[hi].pred, h.h_pred, (size_t)B * 8);
    if (!h.active) {
      if (out[hi].proba) for (size_t i = 0; i < (size_t)B * C; i++) out[hi].proba[i] = NAN;
      out[hi].loss = out[hi].accuracy = NAN;
      continue;
    }
    out[hi].loss = h.has_labels ? h.h_out[(size_t)B * C] : NAN;
    out[hi].accuracy = h.has_labels ? h.h_out[(size_t)B * C + 1] : NAN;
  }
  return 0;
}

extern "C" int icl_run(icl_model* m, int op, const icl_batch* b, float keep_in, float keep, uint64_t seed, icl_head_out* out) {
  CKI(icl_upload(m, b));
  CKI(icl_run_resident(m, op, keep_in, keep, seed));
  return icl_fetch(m, out);
}

// ----------------------------------------------------------------------------- hooks
extern "C" int icl_get_lstm_outputs(icl_model* m, int dir, float* host) {
  if (!m->resident) return fail("no batch resident");
  int T = m->T_cap;
  float* tmp;
  size_t n = (size_t)m->S * T * m->H;
  CK(dmalloc(&tmp, n));
  k_unpack_outputs<<<m->S * T, 128, 0, m->stream>>>(m->Hx[dir], mk_layout(m), m->S, T, m->H, dir, tmp);
  cudaError_t e = cudaMemcpyAsync(host, tmp, n * 4, cudaMemcpyDeviceToHost, m->stream);
  cudaStreamSynchronize(m->stream);
  cudaFree(tmp);
  CK(e);
  return 0;
}
extern "C" int icl_get_batch_input(icl_model* m, int head, float* host) {
  if (head < 0 || head >= (int)m->heads.size()) return fail("head out of range");
  Head& h = m->heads[head];
  CK(cudaStreamSynchronize(m->stream));
  if (h.fact) {      // a factorised affinity head never builds the concatenated rows: build them here, for the caller (tests)
    SlotTable t = h.slots;
    for (int i = 0; i < t.n_slots; i++)
      if ((h.slot_index_id[i] == -2 && h.box_compact) || h.slot_index_id[i] == -3) t.rowidx[i] = h.d_b_of;     // compact blocks: row = group
    k_gather_concat<<<dim3(h.c.batch_size, t.n_slots), 96, 0, m->stream>>>(t, m->Hx[0], m->Hx[1], mk_layout(m), m->H, m->T_cap, h.ldbi,
                                                                           mk_drop(m->last_seed, 0, m->last_keep, m->seq_gid0), m->round_ops, h.bi);
    LAUNCHED(m);
    CK(cudaStreamSynchronize(m->stream));
  }
  CK(cudaMemcpy2D(host, (size_t)h.D0 * 4, h.bi, (size_t)h.ldbi * 4, (size_t)h.D0 * 4, h.c.batch_size, cudaMemcpyDeviceToHost));
  return 0;
}
// Affinity layer-1 factorisation of head `head`: whether the resident batch runs factorised, its distinct mentions / boxes, and the
// totals over all uploads so far (batches that ran factorised, their distinct mentions and boxes).
extern "C" int icl_head_factor_stats(icl_model* m, int head, int32_t* factorised, int32_t* n_mentions, int32_t* n_boxes, int64_t* totals) {
  if (head < 0 || head >= (int)m->heads.size()) return fail("head out of range");
  const Head& h = m->heads[head];
  if (factorised) *factorised = h.fact ? 1 : 0;
  if (n_mentions) *n_mentions = h.fact ? h.Mu : 0;
  if (n_boxes) *n_boxes = h.fact ? h.Nu : 0;
  if (totals) { totals[0] = h.fact_batches; totals[1] = h.fact_mu; totals[2] = h.fact_nu; }
  return 0;
}
extern "C" int icl_get_activation(icl_model* m, int head, int layer, float* host) {
  if (head < 0 || head >= (int)m->heads.size()) return fail("head out of range");
  Head& h = m->heads[head];
  if (layer < 0 || layer >= h.c.n_hidden) return fail("layer out of range");
  CK(cudaStreamSynchronize(m->stream));
  CK(cudaMemcpy(host, h.act[layer], (size_t)h.c.batch_size * h.dims[layer + 1] * 4, cudaMemcpyDeviceToHost));
  return 0;
}
// bring-up: trace the roles of one CTA of the persistent kernels during the next runs; host gets [4][RP_TRACE_EV][4] int64
extern "C" int icl_rec_trace(icl_model* m, int cta, long long* host) {
  if (!m->rp_trace) { CK(cudaMalloc((void**)&m->rp_trace, (size_t)4 * RP_TRACE_EV * 4 * 8)); }
  if (host) {
    CK(cudaStreamSynchronize(m->stream));
    CK(cudaMemcpy(host, m->rp_trace, (size_t)4 * RP_TRACE_EV * 4 * 8, cudaMemcpyDeviceToHost));
  }
  CK(cudaMemset(m->rp_trace, 0xff, (size_t)4 * RP_TRACE_EV * 4 * 8));
  m->rp_trace_cta = cta;
  return 0;
}
extern "C" int icl_pack_rows(const void* src, int32_t src_dtype, int64_t n, void* dst, int32_t half) {
  if (!src || !dst || n < 0) return fail("icl_pack_rows: null buffer or negative count");
  if (src_dtype != ICL_F32 && src_dtype != ICL_F64) return fail("icl_pack_rows: sentences must be float32/float64");
  if (!half) { to_f32((float*)dst, src, src_dtype, (size_t)n); _mm_sfence(); return 0; }
  if (!(__builtin_cpu_supports("avx") && __builtin_cpu_supports("f16c"))) return fail("icl_pack_rows: this CPU has no F16C");
  if (src_dtype == ICL_F32) cvt_f32_h16_stream((uint16_t*)dst, (const float*)src, (size_t)n);
  else cvt_f64_h16_stream((uint16_t*)dst, (const double*)src, (size_t)n);
  _mm_sfence();
  return 0;
}

extern "C" int icl_debug_mask(icl_model* m, uint64_t seed, uint32_t stream, int64_t first, int64_t n, float keep, float* host) {
  float* tmp;
  CK(dmalloc(&tmp, (size_t)n));
  k_debug_mask<<<(unsigned)((n + 255) / 256), 256, 0, m->stream>>>(seed, stream, first, n, keep, tmp);
  cudaError_t e = cudaMemcpyAsync(host, tmp, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream);
  cudaStreamSynchronize(m->stream);
  cudaFree(tmp);
  CK(e);
  return 0;
}
extern "C" int icl_gemm(icl_model* m, int mode, int a_mn, int b_mn, int M, int N, int K, const float* A, const float* B, float* C,
                        int splits) {
  float *dA, *dB, *dC;
  CK(dmalloc(&dA, (size_t)M * K)); CK(dmalloc(&dB, (size_t)K * N)); CK(dmalloc(&dC, (size_t)M * N));
  CK(cudaMemcpy(dA, A, (size_t)M * K * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B, (size_t)K * N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
  GemmArgs g = mk_gemm(dA, a_mn ? M : K, dB, b_mn ? N : K, dC, N, M, N, K);
  int r = gemm(m, m->stream, a_mn, b_mn, g, mode, splits < 1 ? 1 : splits);
  cudaError_t e = cudaStreamSynchronize(m->stream);
  if (r == 0 && e == cudaSuccess) e = cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  if (r != 0) return r;
  CK(e);
  return 0;
}
