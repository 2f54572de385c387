// Device kernels for the BiLSTM + mention-span-head path (everything except the tcgen05 GEMMs and the persistent
// recurrent kernels).  All arithmetic is fp32; dropout masks are a pure function of (seed, stream, element index)
// so forward and backward regenerate identical masks and nothing is stored.
//
// HBM layout of the recurrent tensors ("step-major"): sequences are ranked by length (descending, stable); step k
// of a direction touches the first nact[k] ranks, and all per-token tensors of that direction (input rows, gates,
// h, c, dH) keep step k's rows contiguously at [off[k], off[k]+nact[k]).  Forward step k consumes token k, backward
// step k consumes token len-1-k (reverse_sequence, nn_utils/core.py:324-329), so token (s,t) of direction d lives
// at row off[d ? len[s]-1-t : t] + rank[s].  Every per-step operand is therefore one dense row block that a single
// TMA box can fetch -- no gathers inside the recurrence.
#pragma once
#include <cuda_fp16.h>
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

// ----------------------------------------------------------------------------- numerics helpers
// The tensor cores read fp32 containers as TF32 by TRUNCATING the low 13 mantissa bits (a systematic ~ -1e-3
// relative bias per product).  Every tensor that is only ever an MMA operand is therefore stored pre-rounded
// with round-to-nearest (cvt.rna), which makes the truncation a no-op and the error unbiased and ~4x smaller.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float maybe_round(float x, int on) { return on ? tf32_rna(x) : x; }

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
  int round_out;       // store the result RNA-rounded to TF32 (it is only an MMA operand downstream)
};

__device__ __forceinline__ float epilogue_apply(const Epilogue& e, float v, long m, long n, int N) {
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
  return e.round_out ? tf32_rna(v) : v;
}

// four consecutive columns n..n+3 of row m at once: one Philox call yields the four dropout bits (N % 4 == 0, n % 4 == 0)
__device__ __forceinline__ float4 epilogue_apply4(const Epilogue& e, float4 v, long m, int n, int N) {
  float x[4] = {v.x, v.y, v.z, v.w};
  if (e.bias) {
    const float4 b = *reinterpret_cast<const float4*>(e.bias + n);
    x[0] += b.x; x[1] += b.y; x[2] += b.z; x[3] += b.w;
  }
  float mk[4] = {1.f, 1.f, 1.f, 1.f};
  const bool drop = e.mode != EPI_PLAIN && e.drop.keep < 1.0f;
  if (drop) drop4(e.drop.seed, e.drop.stream, (uint64_t)((e.drop.row_gid0 + m) * N + n) >> 2, e.drop.keep, mk);
  if (e.mode == EPI_BIAS_ACT_DROP) {
    const float inv = drop ? 1.0f / e.drop.keep : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) { x[j] = act_fwd(x[j], e.act); if (drop) x[j] = x[j] / e.drop.keep * mk[j]; }
    (void)inv;
  } else if (e.mode == EPI_DACT) {
    const float4 y4 = *reinterpret_cast<const float4*>(e.aux + m * e.ldaux + n);
    const float y[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float a = y[j];
      if (drop) { a = y[j] * e.drop.keep; x[j] = x[j] / e.drop.keep * mk[j]; }
      x[j] *= act_bwd_from_out(a, e.act);
    }
  }
  if (e.round_out) {
#pragma unroll
    for (int j = 0; j < 4; j++) x[j] = tf32_rna(x[j]);
  }
  return make_float4(x[0], x[1], x[2], x[3]);
}

// ----------------------------------------------------------------------------- SIMT fp32 GEMM (validation mode and tiny / unaligned shapes)
// C[M,N] = epi(A*B).  A(m,k) = A_KMAJOR ? A[m*lda+k] : A[k*lda+m];  B(k,n) = B_KMAJOR ? B[n*ldb+k] : B[k*ldb+n].
struct GemmArgs {
  const float* A; long lda;
  const float* B; long ldb;
  float* C; long ldc;
  int M, N, K;
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
        if (A_KMAJOR) v = g.A[(long)(m0 + mm) * g.lda + k0 + kk];
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
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx + 16 * j;
      if (n >= g.N) continue;
      g.C[(long)m * g.ldc + n] = epilogue_apply(g.epi, acc[i][j], m, n, g.N);
    }
  }
}

// ----------------------------------------------------------------------------- step-major layout
struct StepLayout {
  const int* off;    // [Tmax+1] first row of step k
  const int* nact;   // [Tmax]   sequences still running at step k
  const int* rank;   // [S]      rank of sequence s in the length-sorted order
  const int* lens;   // [S]
};
__device__ __forceinline__ long token_row(const StepLayout& L, int dir, int s, int t) {
  int k = dir ? (L.lens[s] - 1 - t) : t;
  return (long)L.off[k] + L.rank[s];
}

// ----------------------------------------------------------------------------- input preparation
// One warp per valid token: optional l2-normalise (core.py:289-290), then the two directions' input dropout
// (core.py:309-312), written to each direction's step-major row; TF32-rounded when the rows feed the tensor cores.
__global__ void k_prep_x(const float* __restrict__ xraw, const int* __restrict__ tok_seq, const int* __restrict__ tokstart,
                         StepLayout L, int ntok, int E, int Tcap, int data_norm, float keep_in, uint64_t seed,
                         int64_t seq_gid0, int round_ops, float* __restrict__ xfw, float* __restrict__ xbw, int ldx, int ones_col,
                         const int* __restrict__ tok_row, const float* __restrict__ table, __half* __restrict__ x16fw,
                         __half* __restrict__ x16bw, int ld16, int raw_half) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ntok) return;
  int s = tok_seq[warp];
  int t = warp - tokstart[s];
  // the token's embedding row: from this batch's uploaded rows, or gathered from the device-resident token table
  // raw_half: the uploaded rows crossed PCIe as fp16 (icl_upload's half-width wire format; never together with the token table)
  const float* src = tok_row ? table + (long)tok_row[warp] * E : xraw + (long)warp * E;
  const __half* src16 = reinterpret_cast<const __half*>(xraw) + (long)warp * E;
  float scale = 1.0f;
  if (data_norm) {
    float ss = 0.0f;
    for (int e = lane; e < E; e += 32) { float v = raw_half ? __half2float(src16[e]) : src[e]; ss += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    scale = rsqrtf(fmaxf(ss, 1e-12f));
  }
  long rfw = token_row(L, 0, s, t), rbw = token_row(L, 1, s, t);
  if (lane == 0) { xfw[rfw * ldx + ones_col] = 1.0f; xbw[rbw * ldx + ones_col] = 1.0f; }   // sums dZ into the bias gradient
  uint64_t base = (uint64_t)((seq_gid0 + s) * Tcap + t) * (uint64_t)E;
  for (int e4 = lane * 4; e4 < E; e4 += 128) {       // E % 4 == 0 is enforced at create
    float mf[4] = {1, 1, 1, 1}, mb[4] = {1, 1, 1, 1};
    if (keep_in < 1.0f) {
      drop4(seed, STREAM_IN_FW, (base + e4) >> 2, keep_in, mf);
      drop4(seed, STREAM_IN_BW, (base + e4) >> 2, keep_in, mb);
    }
    float4 x;
    if (raw_half) {
      const uint2 u = *reinterpret_cast<const uint2*>(src16 + e4);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      x = make_float4(a.x, a.y, b.x, b.y);
    } else x = *reinterpret_cast<const float4*>(src + e4);
    float xv[4] = {x.x * scale, x.y * scale, x.z * scale, x.w * scale}, vf[4], vb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      vf[j] = xv[j]; vb[j] = xv[j];
      if (keep_in < 1.0f) { vf[j] = xv[j] / keep_in * mf[j]; vb[j] = xv[j] / keep_in * mb[j]; }
      vf[j] = maybe_round(vf[j], round_ops); vb[j] = maybe_round(vb[j], round_ops);
    }
    *reinterpret_cast<float4*>(xfw + rfw * ldx + e4) = make_float4(vf[0], vf[1], vf[2], vf[3]);
    *reinterpret_cast<float4*>(xbw + rbw * ldx + e4) = make_float4(vb[0], vb[1], vb[2], vb[3]);
    if (x16fw) {               // fp16 copies for the input-projection GEMM (the fp32 rows above feed the weight-gradient GEMM)
      __half2 f01 = __floats2half2_rn(vf[0], vf[1]), f23 = __floats2half2_rn(vf[2], vf[3]);
      __half2 b01 = __floats2half2_rn(vb[0], vb[1]), b23 = __floats2half2_rn(vb[2], vb[3]);
      *reinterpret_cast<uint2*>(x16fw + rfw * ld16 + e4) = make_uint2(*reinterpret_cast<uint32_t*>(&f01), *reinterpret_cast<uint32_t*>(&f23));
      *reinterpret_cast<uint2*>(x16bw + rbw * ld16 + e4) = make_uint2(*reinterpret_cast<uint32_t*>(&b01), *reinterpret_cast<uint32_t*>(&b23));
    }
  }
}

// ----------------------------------------------------------------------------- LSTM cell (BasicLSTMCell: i,j,f,o, forget_bias 1.0)
// shared by the per-step kernels here and the fused epilogues of the persistent kernels
struct CellOut { float si, tj, sf, so, c, h; };
__device__ __forceinline__ CellOut lstm_cell(float zi, float zj, float zf, float zo, float cprev) {
  CellOut o;
  o.si = sigmoidf_(zi); o.tj = tanhf(zj); o.sf = sigmoidf_(zf + 1.0f); o.so = sigmoidf_(zo);
  o.c = cprev * o.sf + o.si * o.tj;
  o.h = tanhf(o.c) * o.so;
  return o;
}
struct CellGrad { float di, dj, df, dg_o, dc_prev; };
__device__ __forceinline__ CellGrad lstm_cell_bwd(float si, float tj, float sf, float so, float c, float cprev, float dh,
                                                  float dc_in) {
  CellGrad g;
  float tc = tanhf(c);
  float dc = dc_in + dh * so * (1.0f - tc * tc);
  g.di = dc * tj * si * (1.0f - si);
  g.dj = dc * si * (1.0f - tj * tj);
  g.df = dc * cprev * sf * (1.0f - sf);
  g.dg_o = dh * tc * so * (1.0f - so);
  g.dc_prev = dc * sf;
  return g;
}

// Per-step forward cell (the non-persistent path): z = Zx (+ R, the recurrent product of this step) for rows
// [0,n) of step k.  Z rows are overwritten with the activated gates for the backward pass.  Each thread = 4 units.
__global__ void k_lstm_cell_fwd(float* __restrict__ Zk, const float* __restrict__ R, const float* __restrict__ Cprev,
                                float* __restrict__ Ck, float* __restrict__ Hk, float* __restrict__ Hp_next, int ldhp, int n,
                                int n_next, int H, int round_ops) {
  int q = H >> 2;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n * q) return;
  int m = (int)(i / q), u = (int)(i % q) * 4;
  float* z = Zk + (long)m * 4 * H + u;
  float4 g[4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    g[a] = *reinterpret_cast<float4*>(z + a * H);
    if (R) {
      float4 r = *reinterpret_cast<const float4*>(R + (long)m * 4 * H + a * H + u);
      g[a].x += r.x; g[a].y += r.y; g[a].z += r.z; g[a].w += r.w;
    }
  }
  float4 cp = Cprev ? *reinterpret_cast<const float4*>(Cprev + (long)m * H + u) : make_float4(0, 0, 0, 0);
  const float* gi = &g[0].x; const float* gj = &g[1].x; const float* gf = &g[2].x; const float* go = &g[3].x;
  const float* cpv = &cp.x;
  float si[4], tj[4], sf[4], so[4], c[4], h[4], hr[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    CellOut o = lstm_cell(gi[j], gj[j], gf[j], go[j], cpv[j]);
    si[j] = o.si; tj[j] = o.tj; sf[j] = o.sf; so[j] = o.so; c[j] = o.c; h[j] = o.h; hr[j] = maybe_round(o.h, round_ops);
  }
  *reinterpret_cast<float4*>(z) = make_float4(si[0], si[1], si[2], si[3]);
  *reinterpret_cast<float4*>(z + H) = make_float4(tj[0], tj[1], tj[2], tj[3]);
  *reinterpret_cast<float4*>(z + 2 * H) = make_float4(sf[0], sf[1], sf[2], sf[3]);
  *reinterpret_cast<float4*>(z + 3 * H) = make_float4(so[0], so[1], so[2], so[3]);
  *reinterpret_cast<float4*>(Ck + (long)m * H + u) = make_float4(c[0], c[1], c[2], c[3]);
  *reinterpret_cast<float4*>(Hk + (long)m * H + u) = make_float4(h[0], h[1], h[2], h[3]);
  if (m < n_next) *reinterpret_cast<float4*>(Hp_next + (long)m * ldhp + u) = make_float4(hr[0], hr[1], hr[2], hr[3]);
}

// Per-step backward cell: gates in Zk rows -> dZ (in place); dh = dHout + dhrec (rank-indexed carry), dc carry.
__global__ void k_lstm_cell_bwd(float* __restrict__ Zk, const float* __restrict__ Ck, const float* __restrict__ Cprev,
                                const float* __restrict__ dHk, float* __restrict__ dhrec, float* __restrict__ dcc, int n, int H,
                                int round_ops) {
  asm volatile("griddepcontrol.wait;" ::: "memory");               // PDL: the recurrent GEMM before us has fully completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int q = H >> 2;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n * q) return;
  int m = (int)(i / q), u = (int)(i % q) * 4;
  float* z = Zk + (long)m * 4 * H + u;
  float4 g4[4];
#pragma unroll
  for (int a = 0; a < 4; a++) g4[a] = *reinterpret_cast<float4*>(z + a * H);
  float4 c4 = *reinterpret_cast<const float4*>(Ck + (long)m * H + u);
  float4 cp4 = Cprev ? *reinterpret_cast<const float4*>(Cprev + (long)m * H + u) : make_float4(0, 0, 0, 0);
  float4 dh4 = *reinterpret_cast<const float4*>(dHk + (long)m * H + u);
  float4 dr4 = *reinterpret_cast<const float4*>(dhrec + (long)m * H + u);
  *reinterpret_cast<float4*>(dhrec + (long)m * H + u) = make_float4(0.f, 0.f, 0.f, 0.f);   // ready for the next split-K accumulation
  float4 dc4 = *reinterpret_cast<const float4*>(dcc + (long)m * H + u);
  const float *si = &g4[0].x, *tj = &g4[1].x, *sf = &g4[2].x, *so = &g4[3].x, *c = &c4.x, *cp = &cp4.x, *dh = &dh4.x, *dr = &dr4.x,
              *dc = &dc4.x;
  float di[4], dj[4], df[4], dgo[4], dcp[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    CellGrad g = lstm_cell_bwd(si[j], tj[j], sf[j], so[j], c[j], cp[j], dh[j] + dr[j], dc[j]);
    di[j] = maybe_round(g.di, round_ops); dj[j] = maybe_round(g.dj, round_ops); df[j] = maybe_round(g.df, round_ops);
    dgo[j] = maybe_round(g.dg_o, round_ops); dcp[j] = g.dc_prev;
  }
  *reinterpret_cast<float4*>(z) = make_float4(di[0], di[1], di[2], di[3]);
  *reinterpret_cast<float4*>(z + H) = make_float4(dj[0], dj[1], dj[2], dj[3]);
  *reinterpret_cast<float4*>(z + 2 * H) = make_float4(df[0], df[1], df[2], df[3]);
  *reinterpret_cast<float4*>(z + 3 * H) = make_float4(dgo[0], dgo[1], dgo[2], dgo[3]);
  *reinterpret_cast<float4*>(dcc + (long)m * H + u) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
}

// ----------------------------------------------------------------------------- span gather + concat (core.py:335-440)
struct SlotTable {
  int n_slots;
  int kind[16];        // 0: gathered LSTM rows (width H), 1: dense block
  int col[16];         // first column in batch_input
  int width[16];
  const int* idx[16];  // kind 0: [B,3] int32 on device
  const float* dense[16];  // kind 1: [B,width], or (rowidx set) a resident table [n_rows,width]
  const int* rowidx[16];   // kind 1, optional: [B] row of example b in the resident table (icl_set_box_table)
  const int* rowmap;       // optional: output row b is built from example rowmap[b] (index rows, dense rows without rowidx) -- the
                           // distinct mentions of an affinity batch (layer-1 factorisation) are gathered once each
};

// one block per (example, slot) -- every slot's index chain (index row -> length / rank / step offset -> state row) resolves
// concurrently instead of one after the other; the LSTM's output dropout (core.py:312) is applied here, on the gathered rows.
// batch_input is only ever a GEMM operand (layer-1 forward, layer-1 weight gradient): stored TF32-rounded.
__global__ void k_gather_concat(SlotTable st, const float* __restrict__ h_fw, const float* __restrict__ h_bw, StepLayout L,
                                int H, int Tcap, int D0, Drop drop, int round_ops, float* __restrict__ out) {
  const int bo = blockIdx.x, b = st.rowmap ? st.rowmap[bo] : bo;
  float* o = out + (long)bo * D0;
  {
    const int sl = blockIdx.y;
    if (st.kind[sl] == 1) {
      const float* src = st.dense[sl] + (long)(st.rowidx[sl] ? st.rowidx[sl][bo] : b) * st.width[sl];
      if ((st.width[sl] & 3) == 0 && (st.col[sl] & 3) == 0 && (D0 & 3) == 0) {
        for (int e = threadIdx.x * 4; e < st.width[sl]; e += blockDim.x * 4) {
          float4 v = *reinterpret_cast<const float4*>(src + e);
          *reinterpret_cast<float4*>(o + st.col[sl] + e) =
              make_float4(maybe_round(v.x, round_ops), maybe_round(v.y, round_ops), maybe_round(v.z, round_ops), maybe_round(v.w, round_ops));
        }
      } else {
        for (int e = threadIdx.x; e < st.width[sl]; e += blockDim.x) o[st.col[sl] + e] = maybe_round(src[e], round_ops);
      }
    } else {
      const int* ix = st.idx[sl] + b * 3;
      int d = ix[0], s = ix[1], w = ix[2];
      bool valid = w < L.lens[s];            // dynamic_rnn emits zeros past the sequence length
      const float* src = (d ? h_bw : h_fw) + (valid ? token_row(L, d, s, w) : 0) * H;
      uint64_t base = (uint64_t)((drop.row_gid0 + s) * Tcap + w) * (uint64_t)H;
      for (int u = threadIdx.x * 4; u < H; u += blockDim.x * 4) {        // H % 4 == 0; one Philox call per 4 units
        float4 v4 = valid ? *reinterpret_cast<const float4*>(src + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        float x[4] = {v4.x, v4.y, v4.z, v4.w};
        if (drop.keep < 1.0f) {
          float mk[4];
          drop4(drop.seed, STREAM_OUT_FW + d, (base + u) >> 2, drop.keep, mk);
#pragma unroll
          for (int j = 0; j < 4; j++) x[j] = x[j] / drop.keep * mk[j];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) o[st.col[sl] + u + j] = maybe_round(x[j], round_ops);
      }
    }
  }
}

// backward: scatter-add d(batch_input) into dH_fw/bw (pre-dropout gradient, dropout scaling applied here)
__global__ void k_scatter_spans(SlotTable st, const float* __restrict__ dbi, StepLayout L, int H, int Tcap, int D0, Drop drop,
                                float* __restrict__ dh_fw, float* __restrict__ dh_bw) {
  const int bo = blockIdx.x, b = st.rowmap ? st.rowmap[bo] : bo;
  const float* g = dbi + (long)bo * D0;
  {                                                  // one block per (example, slot), like k_gather_concat
    const int sl = blockIdx.y;
    if (st.kind[sl] == 1) return;
    const int* ix = st.idx[sl] + b * 3;
    int d = ix[0], s = ix[1], w = ix[2];
    if (w >= L.lens[s]) return;
    float* dst = (d ? dh_bw : dh_fw) + token_row(L, d, s, w) * H;
    uint64_t base = (uint64_t)((drop.row_gid0 + s) * Tcap + w) * (uint64_t)H;
    for (int u = threadIdx.x * 4; u < H; u += blockDim.x * 4) {
      float mk[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop.keep < 1.0f) drop4(drop.seed, STREAM_OUT_FW + d, (base + u) >> 2, drop.keep, mk);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        v[j] = g[st.col[sl] + u + j];
        if (drop.keep < 1.0f) v[j] = v[j] / drop.keep * mk[j];
      }
      if (v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f || v[3] != 0.0f)
        atomicAdd(reinterpret_cast<float4*>(dst + u), make_float4(v[0], v[1], v[2], v[3]));     // one 16-byte red (sm_90+)
    }
  }
}

// ----------------------------------------------------------------------------- affinity layer 1, factorised
// batch_input of a mention-box pair is [mention encoding | m_feats | box row (| b_feats)] (core.py:421-433), so
//   z1[p] = [enc, m_feats](mention of p) . W1[0:Dm]  +  [box (, b_feats)](box of p) . W1[Dm:D0]  +  b1  =  U[m_of[p]] + V[b_of[p]] + b1
// with U computed once per DISTINCT mention and V once per distinct box of the batch (a batch of 512 pairs of ~2 images holds ~30
// mentions and ~40 boxes).  This kernel is the layer's epilogue: the pair's two rows, bias, activation, dropout -- bit for bit
// the epilogue of the concatenated GEMM (same mask stream / element numbering).  One block per pair.
__global__ void k_pair_combine(const float* __restrict__ U, const float* __restrict__ V, const int* __restrict__ m_of,
                               const int* __restrict__ b_of, int N, Epilogue e, float* __restrict__ out) {
  const int p = blockIdx.x;
  const float* u = U + (long)m_of[p] * N;
  const float* v = V + (long)b_of[p] * N;
  float* o = out + (long)p * N;
  if ((N & 3) == 0) {
    for (int n = threadIdx.x * 4; n < N; n += blockDim.x * 4) {
      const float4 a = *reinterpret_cast<const float4*>(u + n), c = *reinterpret_cast<const float4*>(v + n);
      *reinterpret_cast<float4*>(o + n) = epilogue_apply4(e, make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w), p, n, N);
    }
  } else {
    for (int n = threadIdx.x; n < N; n += blockDim.x) o[n] = epilogue_apply(e, u[n] + v[n], p, n, N);
  }
}
// backward of the row fan-out: dU[g] = sum of dz1[p] over the pairs p of group g, in list order (CSR lists built by the host in
// pair order: no atomics, bit-reproducible); stored TF32-rounded when it only feeds tensor-core contractions.  One block per group.
__global__ void k_segment_sum(const float* __restrict__ dz, int N, const int* __restrict__ start, const int* __restrict__ members,
                              int round_ops, float* __restrict__ out) {
  const int g = blockIdx.x, p0 = start[g], p1 = start[g + 1];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.0f;
    for (int i = p0; i < p1; i++) acc += dz[(long)members[i] * N + n];
    out[(long)g * N + n] = maybe_round(acc, round_ops);
  }
}

// ----------------------------------------------------------------------------- softmax layer + CE + metrics (core.py:202-268,509-511)
// one warp per example: logits = a*W + b, softmax, argmax, per-row CE * scale, dlogits = (p - y) * scale * gscale
// (gscale = d joint_loss / d loss of this head: 1, or the row sum of the trainable loss-mixing matrix of `weighted_joint`)
__global__ void k_softmax_ce(const float* __restrict__ a, int K, const float* __restrict__ W, const float* __restrict__ bias,
                             int C, const float* __restrict__ y, int B, float scale, float gscale, float* __restrict__ proba,
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
      dlogits[(long)b * C + c] = (p - yc) * (scale * gscale);
    }
  }
  pred[b] = am;
  if (y) { row_loss[b] = loss * scale; row_correct[b] = (am == ay) ? 1.0f : 0.0f; }
}

// Softmax layer backward (core.py:202-232 in reverse), ONE kernel: with dl = dlogits [B,C] (C = 2..32 classes),
//   dz_{L-1}[b,k] = (sum_c dl[b,c] W[k,c]) * dropout-mask/keep * act'(...)   (the EPI_DACT epilogue of the hidden layers' GEMMs)
//   dW[k,c] = sum_b a[b,k] dl[b,c],   db[c] = sum_b dl[b,c]
// A C-column product is bandwidth work, not a tensor-core tile (and C = 2 cannot even be a TMA row pitch): every block takes
// SMB_ROWS examples, thread k owns hidden unit k (coalesced along k) and writes its partial sums of dW / db to the block's row of
// `part` [blocks][K*C + C]; k_softmax_bwd_reduce adds the rows in block order (no atomics: the gradient is bit-reproducible).
constexpr int SMB_ROWS = 8, SMB_THREADS = 128;
__global__ void __launch_bounds__(SMB_THREADS) k_softmax_bwd(const float* __restrict__ a, const float* __restrict__ dl,
                                                             const float* __restrict__ W, int B, int K, int C, Epilogue e,
                                                             float* __restrict__ dz, float* __restrict__ part) {
  extern __shared__ float sm_dyn[];
  float* const dW = part + (long)blockIdx.x * (K * C + C);
  float* const db = dW + K * C;
  float* s_dl = sm_dyn;                         // [SMB_ROWS][C]
  float* s_W = sm_dyn + SMB_ROWS * C;           // [K][C + 1]  (odd pitch: no bank conflicts between neighbouring k)
  const int r0 = blockIdx.x * SMB_ROWS, nr = min(SMB_ROWS, B - r0), CP = C + 1;
  // the epilogue's parameters in registers (kernel parameters live in the constant bank: a dependent load per use otherwise)
  const int act = e.act, round_out = e.round_out;
  const float keep = e.drop.keep, inv_keep = 1.0f / e.drop.keep;
  const bool drop = keep < 1.0f;
  const uint64_t seed = e.drop.seed;
  const uint32_t stream = e.drop.stream;
  const long gid0 = e.drop.row_gid0;
  for (int i = threadIdx.x; i < nr * C; i += blockDim.x) s_dl[i] = dl[(long)r0 * C + i];
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) s_W[(i / C) * CP + i % C] = W[i];
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float av[SMB_ROWS];                         // a = the last hidden layer's post-dropout output: operand of dW AND `aux` of the epilogue
#pragma unroll
    for (int r = 0; r < SMB_ROWS; r++) av[r] = r < nr ? a[(long)(r0 + r) * K + k] : 0.0f;
#pragma unroll
    for (int r = 0; r < SMB_ROWS; r++) {
      if (r < nr) {
        float v = 0.0f;
        for (int c = 0; c < C; c++) v = fmaf(s_dl[r * C + c], s_W[k * CP + c], v);
        float y = av[r];
        if (drop) {                              // EPI_DACT (see epilogue_apply): mask of the layer whose output is a
          const float mk = drop1(seed, stream, (uint64_t)((gid0 + r0 + r) * K + k), keep);
          y *= keep;
          v = v * inv_keep * mk;
        }
        v *= act_bwd_from_out(y, act);
        dz[(long)(r0 + r) * K + k] = round_out ? tf32_rna(v) : v;
      }
    }
    for (int c = 0; c < C; c++) {
      float acc = 0.0f;
#pragma unroll
      for (int r = 0; r < SMB_ROWS; r++) acc = fmaf(av[r], r < nr ? s_dl[r * C + c] : 0.0f, acc);
      dW[k * C + c] = acc;
    }
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.0f;
    for (int r = 0; r < nr; r++) acc += s_dl[r * C + c];
    db[c] = acc;
  }
}
// dW [K*C] and db [C] of the softmax layer = the sum of the nb partial rows of k_softmax_bwd in a FIXED association order (bit-
// reproducible): block (32 columns, 16 row groups); group y adds rows y, y+16, ... in order, then the 16 group sums are added in order.
__global__ void __launch_bounds__(512) k_softmax_bwd_reduce(const float* __restrict__ part, int nb, int KC, int C, float* __restrict__ dW,
                                                            float* __restrict__ db) {
  __shared__ float sh[16][33];
  const int i = blockIdx.x * 32 + threadIdx.x, n = KC + C;
  float s = 0.0f;
  if (i < n)
    for (int b = threadIdx.y; b < nb; b += 16) s += part[(long)b * n + i];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = 0.0f;
#pragma unroll
    for (int y = 0; y < 16; y++) t += sh[y][threadIdx.x];
    if (i < KC) dW[i] = t; else db[i - KC] = t;
  }
}

// deterministic single-block sums: block 0 sums v0 into out[0], block 1 sums v1 into out[1] divided by mean_div
__global__ void k_reduce_sum2(const float* __restrict__ v0, const float* __restrict__ v1, int n, float* out, float mean_div1) {
  __shared__ double sh[256];
  const float* v = blockIdx.x ? v1 : v0;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(blockIdx.x && mean_div1 > 0 ? sh[0] / mean_div1 : sh[0]);
}
// deterministic single-block sum of n floats into out[0] (divided by mean_div when > 0)
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

// wide variant: many row-blocks accumulate with atomics (for the [Ntok,4H] LSTM dZ)
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
  long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 v = g4[i];
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long i = n4 << 2; i < n; i++) s += (double)g[i] * g[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void k_sumsq_final(const double* __restrict__ partial, int n, float* __restrict__ gnorm, double extra) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gnorm[0] = (float)sqrt(sh[0] + extra);      // extra: squared norm of gradients kept on the host
}
// TF-1.x Adam: theta -= lr_t * m / (sqrt(v) + eps), lr_t computed on the host from the step count.  Also writes the
// TF32-rounded copy of the new parameters that the next step's GEMMs read (pr; may be null).
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       float* __restrict__ pr, long n, const float* __restrict__ gnorm, float clip, float lr_t, float b1, float b2,
                       float eps) {
  float scale = 1.0f;
  if (clip > 0.0f) scale = clip / fmaxf(gnorm[0], clip);
  long n4 = n >> 2;     // the flat buffers are padded to a multiple of 4 floats
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 g4 = reinterpret_cast<const float4*>(g)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i],
           p4 = reinterpret_cast<float4*>(p)[i];
    float *gp = &g4.x, *mp = &m4.x, *vp = &v4.x, *pp = &p4.x, r[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float gi = gp[j] * scale;
      mp[j] = b1 * mp[j] + (1.0f - b1) * gi;
      vp[j] = b2 * vp[j] + (1.0f - b2) * gi * gi;
      pp[j] -= lr_t * mp[j] / (sqrtf(vp[j]) + eps);
      r[j] = tf32_rna(pp[j]);
    }
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
    if (pr) reinterpret_cast<float4*>(pr)[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
}
// W_ih (rows [0,E) of the TF kernel [E+H, 4H]) as fp16, transposed to K-major [4H][Kp] for the fp16 input-projection GEMM.
// 32x32 shared-memory tile transpose: reads coalesced along the 4H columns, writes coalesced along k.  grid (4H/32, Kp/32, 2 dirs),
// block (32, 8).  Columns k >= E are written as zeros.
__global__ void k_pack_wih16(const float* __restrict__ K0, const float* __restrict__ K1, __half* __restrict__ out0,
                             __half* __restrict__ out1, int E, int N4H, int Kp) {
  __shared__ float tile[32][33];
  const float* K = blockIdx.z ? K1 : K0;
  __half* out = blockIdx.z ? out1 : out0;
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < E && n < N4H) ? K[(long)k * N4H + n] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N4H && k < Kp) out[(long)n * Kp + k] = __float2half_rn(tile[threadIdx.x][r]);
  }
}
__global__ void k_round_copy(const float* __restrict__ src, float* __restrict__ dst, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = tf32_rna(src[i]);
}

// ----------------------------------------------------------------------------- test hooks
__global__ void k_unpack_outputs(const float* __restrict__ h, StepLayout L, int S, int T, int H, int dir, float* __restrict__ out) {
  int s = blockIdx.x / T, t = blockIdx.x % T;
  float* o = out + ((long)s * T + t) * H;
  bool valid = t < L.lens[s];
  const float* src = h + (valid ? token_row(L, dir, s, t) : 0) * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) o[u] = valid ? src[u] : 0.0f;
}
__global__ void k_debug_mask(uint64_t seed, uint32_t stream, int64_t first, int64_t n, float keep, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = drop1(seed, stream, (uint64_t)(first + i), keep);
}

}  // namespace icl
