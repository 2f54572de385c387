// Device kernels for the BiLSTM + mention-span-head path (everything except the tcgen05 GEMM).
// All arithmetic is fp32; dropout masks are a pure function of (seed, stream, element index) so forward and
// backward regenerate identical masks and nothing is stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace icl {

// ----------------------------------------------------------------------------- counter-based RNG (Philox4x32-10)
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x, hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// tf.nn.dropout's binary tensor floor(keep + U[0,1)) for 4 consecutive elements idx4*4 .. idx4*4+3
__host__ __device__ __forceinline__ void drop4(uint64_t seed, uint32_t stream, uint64_t idx4, float keep, float out[4]) {
  uint4 r = philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), stream, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float s = 1.0f / 16777216.0f;
  out[0] = (keep + (float)(r.x >> 8) * s) >= 1.0f ? 1.0f : 0.0f;
  out[1] = (keep + (float)(r.y >> 8) * s) >= 1.0f ? 1.0f : 0.0f;
  out[2] = (keep + (float)(r.z >> 8) * s) >= 1.0f ? 1.0f : 0.0f;
  out[3] = (keep + (float)(r.w >> 8) * s) >= 1.0f ? 1.0f : 0.0f;
}

__host__ __device__ __forceinline__ float drop1(uint64_t seed, uint32_t stream, uint64_t idx, float keep) {
  float m[4];
  drop4(seed, stream, idx >> 2, keep, m);
  return m[idx & 3];
}

enum { STREAM_IN_FW = 0, STREAM_IN_BW = 1, STREAM_OUT_FW = 2, STREAM_OUT_BW = 3, STREAM_HEAD = 16 };

struct Drop {          // dropout descriptor for an epilogue / elementwise kernel
  uint64_t seed;
  uint32_t stream;
  float keep;          // >= 1: off
  int64_t row_gid0;    // global id of row 0
};

// ----------------------------------------------------------------------------- activations
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float act_fwd(float z, int act) {
  switch (act) {
    case 1: return sigmoidf_(z);
    case 2: return tanhf(z);
    case 3: return fmaxf(z, 0.0f);
    case 4: return fmaxf(z, 0.01f * z);
    default: return z;
  }
}
// derivative expressed through the activation OUTPUT a (pre-dropout)
__device__ __forceinline__ float act_bwd_from_out(float a, int act) {
  switch (act) {
    case 1: return a * (1.0f - a);
    case 2: return 1.0f - a * a;
    case 3: return a > 0.0f ? 1.0f : 0.0f;
    case 4: return a > 0.0f ? 1.0f : 0.01f;
    default: return 1.0f;
  }
}

// ----------------------------------------------------------------------------- GEMM epilogue (shared by SIMT and tcgen05)
enum { EPI_PLAIN = 0, EPI_BIAS_ACT_DROP = 1, EPI_DACT = 2 };
struct Epilogue {
  int mode;
  const float* bias;   // [N] (mode 1, or mode 0 if non-null)
  int act;
  Drop drop;           // mode 1: applied to output; mode 2: mask of the layer whose output is `aux`
  const float* aux;    // mode 2: post-dropout output y of the producing layer, [M, ldaux]
  long ldaux;
  float beta;          // 0: overwrite, 1: accumulate into C
};

__device__ __forceinline__ float epilogue_apply(const Epilogue& e, float v, long m, long n, int N, float cold) {
  if (e.bias) v += e.bias[n];
  if (e.mode == EPI_BIAS_ACT_DROP) {
    v = act_fwd(v, e.act);
    if (e.drop.keep < 1.0f) {
      float mk = drop1(e.drop.seed, e.drop.stream, (uint64_t)((e.drop.row_gid0 + m) * N + n), e.drop.keep);
      v = v / e.drop.keep * mk;
    }
  } else if (e.mode == EPI_DACT) {
    float y = e.aux[m * e.ldaux + n];
    float mk = 1.0f, a = y;
    if (e.drop.keep < 1.0f) {
      mk = drop1(e.drop.seed, e.drop.stream, (uint64_t)((e.drop.row_gid0 + m) * N + n), e.drop.keep);
      a = y * e.drop.keep;         // where mk==1; irrelevant where mk==0
      v = v / e.drop.keep * mk;
    }
    v *= act_bwd_from_out(a, e.act);
  }
  if (e.beta != 0.0f) v += e.beta * cold;
  return v;
}

// ----------------------------------------------------------------------------- row maps for the recurrent steps
// Packed, padded layout: sequence s owns rows start[s] .. start[s]+len[s] (len+1 rows).  For direction d and
// token q: data row = start+q+d (x-projection, gates, dZ, h_prev, c_prev); state-out row = start+q+1-d.
struct SeqMap {
  const int* order;   // sequences sorted by length, descending
  const int* start;   // [S]
  const int* lens;    // [S]
  int k;              // recurrence step
  int dir;            // 0 fw, 1 bw
  int mode;           // 0: identity rows; 1: data row of sorted position m at step k; 2: seq id (per-seq buffers)
};
__device__ __forceinline__ long seqmap_row(const SeqMap& s, int m) {
  if (s.mode == 0) return m;
  int sq = s.order[m];
  if (s.mode == 2) return sq;
  int q = s.dir ? (s.lens[sq] - 1 - s.k) : s.k;
  return (long)s.start[sq] + q + s.dir;
}

// ----------------------------------------------------------------------------- SIMT fp32 GEMM (validation path and tiny-N layers)
// C[M,N] = epi(A*B).  A(m,k) = A_KMAJOR ? A[row(m)*lda+k] : A[k*lda+m];  B(k,n) = B_KMAJOR ? B[n*ldb+k] : B[k*ldb+n].
struct GemmArgs {
  const float* A; long lda;
  const float* B; long ldb;
  float* C; long ldc;
  int M, N, K;
  SeqMap amap, cmap;     // optional row gathers (A must be K-major when amap.mode != 0)
  Epilogue epi;
};

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256) k_gemm_simt(const GemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * 256;
      int mm, kk;
      if (A_KMAJOR) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      float v = 0.0f;
      if (m0 + mm < g.M && k0 + kk < g.K) {
        if (A_KMAJOR) v = g.A[seqmap_row(g.amap, m0 + mm) * g.lda + k0 + kk];
        else v = g.A[(long)(k0 + kk) * g.lda + m0 + mm];
      }
      As[kk][mm] = v;
      int nn;
      if (B_KMAJOR) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      v = 0.0f;
      if (n0 + nn < g.N && k0 + kk < g.K) {
        if (B_KMAJOR) v = g.B[(long)(n0 + nn) * g.ldb + k0 + kk];
        else v = g.B[(long)(k0 + kk) * g.ldb + n0 + nn];
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    long crow = seqmap_row(g.cmap, m);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx + 16 * j;
      if (n >= g.N) continue;
      float* p = g.C + crow * g.ldc + n;
      float cold = g.epi.beta != 0.0f ? *p : 0.0f;
      *p = epilogue_apply(g.epi, acc[i][j], m, n, g.N, cold);
    }
  }
}

// ----------------------------------------------------------------------------- input preparation
// One warp per valid token: optional l2-normalise (core.py:289-290), then the two directions' input dropout
// (core.py:309-312) written to the padded layouts Xd_fw[start+t], Xd_bw[start+t+1].
__global__ void k_prep_x(const float* __restrict__ xraw, const int* __restrict__ tok_seq, const int* __restrict__ tokstart,
                         const int* __restrict__ start, int ntok, int E, int Tcap, int data_norm, float keep_in,
                         uint64_t seed, int64_t seq_gid0, float* __restrict__ xfw, float* __restrict__ xbw) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ntok) return;
  int s = tok_seq[warp];
  int t = warp - tokstart[s];
  const float* src = xraw + (long)warp * E;
  float scale = 1.0f;
  if (data_norm) {
    float ss = 0.0f;
    for (int e = lane; e < E; e += 32) { float v = src[e]; ss += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    scale = rsqrtf(fmaxf(ss, 1e-12f));
  }
  long rfw = (long)start[s] + t, rbw = rfw + 1;
  uint64_t base = (uint64_t)((seq_gid0 + s) * Tcap + t) * (uint64_t)E;
  for (int e4 = lane * 4; e4 < E; e4 += 128) {
    float mf[4] = {1, 1, 1, 1}, mb[4] = {1, 1, 1, 1};
    if (keep_in < 1.0f) {      // E % 4 == 0 is enforced at create
      drop4(seed, STREAM_IN_FW, (base + e4) >> 2, keep_in, mf);
      drop4(seed, STREAM_IN_BW, (base + e4) >> 2, keep_in, mb);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float v = src[e4 + j] * scale;
      float vf = v, vb = v;
      if (keep_in < 1.0f) { vf = v / keep_in * mf[j]; vb = v / keep_in * mb[j]; }
      xfw[rfw * E + e4 + j] = vf;
      xbw[rbw * E + e4 + j] = vb;
    }
  }
}

// zero the per-sequence pad rows: which=0: fw-style (row start+len), which=1: bw-style (row start)
__global__ void k_zero_rows(float* buf, const int* __restrict__ start, const int* __restrict__ lens, int S, int W, int at_end) {
  int s = blockIdx.x;
  if (s >= S) return;
  long r = (long)start[s] + (at_end ? lens[s] : 0);
  for (int e = threadIdx.x; e < W; e += blockDim.x) buf[r * W + e] = 0.0f;
}

// ----------------------------------------------------------------------------- LSTM cell, forward (BasicLSTMCell, i,j,f,o, forget_bias 1)
// z (full pre-activation, x-projection + bias + recurrent term) is in Z[data row]; overwritten with the
// activated gates (sig i, tanh j, sig f, sig o) for the backward pass.
__global__ void k_lstm_cell_fwd(float* __restrict__ Z, float* __restrict__ HP, float* __restrict__ CP, SeqMap sm, int n_active, int H) {
  int pos = blockIdx.x;
  if (pos >= n_active) return;
  int sq = sm.order[pos];
  int q = sm.dir ? (sm.lens[sq] - 1 - sm.k) : sm.k;
  long row = (long)sm.start[sq] + q + sm.dir, hrow = (long)sm.start[sq] + q + 1 - sm.dir;
  float* z = Z + row * 4 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) {
    float si = sigmoidf_(z[u]), tj = tanhf(z[H + u]), sf = sigmoidf_(z[2 * H + u] + 1.0f), so = sigmoidf_(z[3 * H + u]);
    float c = CP[row * H + u] * sf + si * tj;
    float h = tanhf(c) * so;
    z[u] = si; z[H + u] = tj; z[2 * H + u] = sf; z[3 * H + u] = so;
    CP[hrow * H + u] = c;
    HP[hrow * H + u] = h;
  }
}

// backward of the cell at one step: gates in G[data row], writes dZ[data row]; carries dh (recurrent) and dc per sequence
__global__ void k_lstm_cell_bwd(const float* G, const float* __restrict__ CP, const float* __restrict__ dHout,
                                float* dZ /* may alias G: in-place gates -> dZ */, float* __restrict__ dhrec, float* __restrict__ dccarry,
                                SeqMap sm, int n_active, int H) {
  int pos = blockIdx.x;
  if (pos >= n_active) return;
  int sq = sm.order[pos];
  int q = sm.dir ? (sm.lens[sq] - 1 - sm.k) : sm.k;
  long row = (long)sm.start[sq] + q + sm.dir, hrow = (long)sm.start[sq] + q + 1 - sm.dir;
  const float* g = G + row * 4 * H;
  float* dz = dZ + row * 4 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) {
    float si = g[u], tj = g[H + u], sf = g[2 * H + u], so = g[3 * H + u];
    float c = CP[hrow * H + u], cprev = CP[row * H + u];
    float tc = tanhf(c);
    float dh = dHout[hrow * H + u] + dhrec[(long)sq * H + u];
    float dc = dccarry[(long)sq * H + u] + dh * so * (1.0f - tc * tc);
    dz[u] = dc * tj * si * (1.0f - si);
    dz[H + u] = dc * si * (1.0f - tj * tj);
    dz[2 * H + u] = dc * cprev * sf * (1.0f - sf);
    dz[3 * H + u] = dh * tc * so * (1.0f - so);
    dccarry[(long)sq * H + u] = dc * sf;
  }
}

// ----------------------------------------------------------------------------- span gather + concat (core.py:335-440)
struct SlotTable {
  int n_slots;
  int kind[16];        // 0: gathered LSTM rows (width H), 1: dense block
  int col[16];         // first column in batch_input
  int width[16];
  const int* idx[16];  // kind 0: [B,3] int32 on device
  const float* dense[16];  // kind 1: [B,width]
};

// one block per example; output dropout of the LSTM (core.py:312) is applied here, on the gathered rows
__global__ void k_gather_concat(SlotTable st, const float* __restrict__ hp_fw, const float* __restrict__ hp_bw,
                                const int* __restrict__ start, const int* __restrict__ lens, int H, int Tcap, int D0,
                                Drop drop, float* __restrict__ out) {
  int b = blockIdx.x;
  float* o = out + (long)b * D0;
  for (int sl = 0; sl < st.n_slots; sl++) {
    if (st.kind[sl] == 1) {
      const float* src = st.dense[sl] + (long)b * st.width[sl];
      for (int e = threadIdx.x; e < st.width[sl]; e += blockDim.x) o[st.col[sl] + e] = src[e];
    } else {
      const int* ix = st.idx[sl] + b * 3;
      int d = ix[0], s = ix[1], w = ix[2];
      bool valid = w < lens[s];            // dynamic_rnn emits zeros past the sequence length
      long row = (long)start[s] + w + 1 - d;
      const float* src = (d ? hp_bw : hp_fw) + row * H;
      uint64_t base = (uint64_t)((drop.row_gid0 + s) * Tcap + w) * (uint64_t)H;
      for (int u = threadIdx.x; u < H; u += blockDim.x) {
        float v = valid ? src[u] : 0.0f;
        if (drop.keep < 1.0f) v = v / drop.keep * drop1(drop.seed, STREAM_OUT_FW + d, base + u, drop.keep);
        o[st.col[sl] + u] = v;
      }
    }
  }
}

// backward: scatter-add d(batch_input) into dHout_fw/bw (pre-dropout gradient, dropout scaling applied here)
__global__ void k_scatter_spans(SlotTable st, const float* __restrict__ dbi, const int* __restrict__ start,
                                const int* __restrict__ lens, int H, int Tcap, int D0, Drop drop,
                                float* __restrict__ dh_fw, float* __restrict__ dh_bw) {
  int b = blockIdx.x;
  const float* g = dbi + (long)b * D0;
  for (int sl = 0; sl < st.n_slots; sl++) {
    if (st.kind[sl] == 1) continue;
    const int* ix = st.idx[sl] + b * 3;
    int d = ix[0], s = ix[1], w = ix[2];
    if (w >= lens[s]) continue;
    long row = (long)start[s] + w + 1 - d;
    float* dst = (d ? dh_bw : dh_fw) + row * H;
    uint64_t base = (uint64_t)((drop.row_gid0 + s) * Tcap + w) * (uint64_t)H;
    for (int u = threadIdx.x; u < H; u += blockDim.x) {
      float v = g[st.col[sl] + u];
      if (drop.keep < 1.0f) v = v / drop.keep * drop1(drop.seed, STREAM_OUT_FW + d, base + u, drop.keep);
      atomicAdd(dst + u, v);
    }
  }
}

// ----------------------------------------------------------------------------- softmax layer + CE + metrics (core.py:202-268,509-511)
// one warp per example: logits = a*W + b, softmax, argmax, per-row CE, dlogits = (p - y) * scale
__global__ void k_softmax_ce(const float* __restrict__ a, int K, const float* __restrict__ W, const float* __restrict__ bias,
                             int C, const float* __restrict__ y, int B, float scale, float* __restrict__ proba,
                             long long* __restrict__ pred, float* __restrict__ row_loss, float* __restrict__ row_correct,
                             float* __restrict__ dlogits) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float logit[32];
  float mx = -INFINITY;
  for (int c = 0; c < C; c++) {
    float p = 0.0f;
    for (int k = lane; k < K; k += 32) p = fmaf(a[(long)b * K + k], W[(long)k * C + c], p);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    logit[c] = p + bias[c];
    mx = fmaxf(mx, logit[c]);
  }
  if (lane != 0) return;
  float sum = 0.0f;
  for (int c = 0; c < C; c++) sum += expf(logit[c] - mx);
  float lse = logf(sum);
  int am = 0, ay = 0;
  float best = -INFINITY, besty = -INFINITY, loss = 0.0f;
  for (int c = 0; c < C; c++) {
    float p = expf(logit[c] - mx) / sum;
    proba[(long)b * C + c] = p;
    if (p > best) { best = p; am = c; }
    if (y) {
      float yc = y[(long)b * C + c];
      if (yc > besty) { besty = yc; ay = c; }
      loss -= yc * (logit[c] - mx - lse);
      dlogits[(long)b * C + c] = (p - yc) * scale;
    }
  }
  pred[b] = am;
  if (y) { row_loss[b] = loss * scale; row_correct[b] = (am == ay) ? 1.0f : 0.0f; }
}

// deterministic single-block sum of n floats into out[0] (and mean into out[1])
__global__ void k_reduce_sum(const float* __restrict__ v, int n, float* out, float mean_div) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(mean_div > 0 ? sh[0] / mean_div : sh[0]);
}

// column sums of X[rows, N] -> out[N]  (bias gradients); grid.x covers N in 32-column strips, 256 threads = 8 row lanes
__global__ void k_colsum(const float* __restrict__ X, long rows, int N, long ld, float* __restrict__ out) {
  __shared__ float sh[8][33];
  int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  float s = 0.0f;
  if (c < N)
    for (long r = rl; r < rows; r += 8) s += X[r * ld + c];
  sh[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += sh[i][threadIdx.x & 31];
    out[c] = t;
  }
}
// wide variant: many row-blocks accumulate with atomics (for the [Np,4H] LSTM dZ)
__global__ void k_colsum_atomic(const float* __restrict__ X, long rows, int N, long ld, float* __restrict__ out, int rows_per_block) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  long r0 = (long)blockIdx.y * rows_per_block, r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float s = 0.0f;
  for (long r = r0; r < r1; r++) s += X[r * ld + c];
  atomicAdd(out + c, s);
}

// ----------------------------------------------------------------------------- clip_by_global_norm + Adam (core.py:94-103)
__global__ void k_sumsq_partial(const float* __restrict__ g, long n, double* __restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double v = g[i];
    s += v * v;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void k_sumsq_final(const double* __restrict__ partial, int n, float* __restrict__ gnorm) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gnorm[0] = (float)sqrt(sh[0]);
}
// TF-1.x Adam: theta -= lr_t * m / (sqrt(v) + eps), lr_t computed on the host from the step count
__global__ void k_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                       const float* __restrict__ gnorm, float clip, float lr_t, float b1, float b2, float eps) {
  float scale = 1.0f;
  if (clip > 0.0f) scale = clip / fmaxf(gnorm[0], clip);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gi = g[i] * scale;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// ----------------------------------------------------------------------------- test hooks
__global__ void k_unpack_outputs(const float* __restrict__ hp, const int* __restrict__ start, const int* __restrict__ lens,
                                 int S, int T, int H, int dir, float* __restrict__ out) {
  int s = blockIdx.x / T, t = blockIdx.x % T;
  float* o = out + ((long)s * T + t) * H;
  bool valid = t < lens[s];
  const float* src = hp + ((long)start[s] + t + 1 - dir) * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) o[u] = valid ? src[u] : 0.0f;
}
__global__ void k_debug_mask(uint64_t seed, uint32_t stream, int64_t first, int64_t n, float keep, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = drop1(seed, stream, (uint64_t)(first + i), keep);
}

}  // namespace icl
