// Persistent recurrent kernels of the BiLSTM (sm_100a): one launch runs EVERY time step of BOTH directions.
//
// Forward (k_rec_fwd, "K2"): direction d's hidden units are cut into nsl slices of U units; CTA (d, part p, slice j) keeps
// its slice of W_hh -- all four gates of its U units, [4U x H] TF32, K-major, 128B-swizzled -- resident in shared memory
// for the whole launch.  Rows (length-ranked sequences) are cut into 128-row tiles; tile t belongs to part t % P for
// every step, so a tile's recurrent state is produced and consumed by the same nsl CTAs and the only cross-CTA
// dependency is "all nsl slices have published h_{k-1} of tile t" -- a per-tile counter in global memory, not a grid
// barrier.  Per (step, tile):
//     producer warp : wait counter -> TMA h_{k-1} tile (K chunks of 32) -> smem ring
//     MMA warp      : tcgen05.mma kind::tf32 (M=128, N=4U) into a TMEM accumulator ring (TMEM = 512 columns)
//     loader warp   : TMA the x-projection boxes {U x 128} of the four gates + c_{k-1} -> smem
//     4 epilogue warps: TMEM + Zx -> gates -> c_k, h_k (fused BasicLSTMCell, forget_bias 1) -> smem boxes -> TMA stores of
//                     gates (for BPTT), c_k, h_k (exact, for the span heads) and TF32(h_k) into step k+1's operand block,
//                     then release the tile's counter.
// The roles are decoupled by mbarriers, so step k+1 of tile 0 starts while tiles 1.. of step k are still in flight.
// All global traffic of the recurrence is TMA boxes (coalesced by the copy engine); no LSU gathers.
//
// Step-major blocks are padded to 128 rows (icl_model.cu), so a full-tile box never touches another step's rows and
// rows past nact[k] only ever hold finite don't-care values.
#pragma once
#include "gemm_tcgen05.cuh"

namespace icl {

#ifndef RP_ASTAGES_OVERRIDE
#define RP_ASTAGES_OVERRIDE 3
#endif
constexpr int RP_ROWS = 128, RP_ASTAGES = RP_ASTAGES_OVERRIDE, RP_THREADS = 224, RP_MAXT = 128, RP_MAXACC = 8;

struct RecMaps {          // [direction]
  CUtensorMap a[2];       // A operand: box {32 k, 128 rows}, SWIZZLE_128B, over Hp (fwd) / unused (bwd)
  CUtensorMap hp[2];      // box {U, 128} over Hp  (TF32-rounded h, operand rows of the next step)
  CUtensorMap hx[2];      // box {U, 128} over Hx  (exact h)
  CUtensorMap cc[2];      // box {U, 128} over Cc
  CUtensorMap z[2];       // box {U, 128} over Z   [rows, 4H]
  CUtensorMap w[2];       // resident weight slice boxes over the packed copy
  CUtensorMap dh[2];      // bwd: box {U,128} over dHout (load) ; reduce target uses dhr
  CUtensorMap dhr[2];     // bwd: box {32, 128} over dHout for cp.reduce.async.bulk (add)
  CUtensorMap dc[2];      // bwd: box {U, 128} over the dc carry [S_pad, H]
};

struct RecArgs {
  const int* off;         // [Tmax+1] padded step offsets
  const int* nact;        // [Tmax+1]
  int Tmax, H, nsl, P, nkb, nk8, max_tiles, training;
  unsigned* flags;        // [2][max_tiles] per-tile publication counters (zeroed before the launch)
  long long* trace;       // optional bring-up trace of CTA trace_cta: [4 roles][RP_TRACE_EV][4] = (event, step, tile, clock64)
  int trace_cta;
};
constexpr int RP_TRACE_EV = 2048;
struct Tracer {
  long long* p; int n;
  __device__ __forceinline__ void init(long long* trace, int cta, int role) { p = (trace && (int)blockIdx.x == cta) ? trace + (long)role * RP_TRACE_EV * 4 : nullptr; n = 0; }
  __device__ __forceinline__ void init(const RecArgs& g, int role) { init(g.trace, g.trace_cta, role); }
  __device__ __forceinline__ void ev(int e, int k, int t) {
    if (p && n < RP_TRACE_EV) { p[n * 4] = e; p[n * 4 + 1] = k; p[n * 4 + 2] = t; p[n * 4 + 3] = clock64(); n++; }
  }
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// Publication of a tile.  The data was written by a TMA bulk store whose completion the SAME thread has waited for
// (cp.async.bulk.wait_group); the counter then moves BEHIND A PROXY FENCE: the store was performed by the async proxy, the counter
// is a generic-proxy atomic, and a `red` right after the wait let a consumer on another SM see the counter and still read the
// rows' previous contents (round 2: one tile of a FRESH session -- zero-filled rows -- off by 2-3 % in
// ~2 of 3 runs of the multitask test once the consumer's ring held a whole tile and its loads followed the counter at once; a
// session that repeats a batch re-reads identical stale values, which is why the 50-run reproducibility test never saw it).
// tests/test_gpu_configs.py::test_recurrence_never_reads_unpublished_rows poisons the fp16 rows before every run to catch any
// such read deterministically.  (The publishing lane has no other bulk copy in flight: the gate / c boxes belong to other lanes.)
__device__ __forceinline__ void flag_release_add(unsigned* p) {
#ifndef ICL_NO_PUBLISH_FENCE      // (defining it reproduces the bug: the poison test then fails)
  asm volatile("fence.proxy.async;" ::: "memory");
#endif
#ifdef ICL_FLAG_RELEASE
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");        // measured: +0.04 ms on K2, no further effect
#else
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
#endif
}
// bounded spin on a publication counter: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void flag_wait(const unsigned* p, unsigned target) {
  const long long t0 = clock64();
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= target) return;
    if (clock64() - t0 > 4000000000LL) __trap();
    __nanosleep(32);
  }
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Packed W_hh for the forward kernel: row (j*4U + n) = gate column of slice j, n = c*16 + gate*4 + i <-> unit j*U + 4c + i
// (so every 16-column accumulator chunk holds i,j,f,o of four consecutive units); K padded with zeros to Kp.
__global__ void k_pack_whh_fwd(const float* __restrict__ Whh, float* __restrict__ Wp, int H, int U, int nsl, int Kp) {
  long total = (long)nsl * 4 * U * Kp;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int k = (int)(idx % Kp);
    int row = (int)(idx / Kp);
    int j = row / (4 * U), n = row % (4 * U);
    int c = n / 16, gate = (n % 16) / 4, i = n % 4;
    int u = j * U + c * 4 + i;
    Wp[idx] = (k < H && u < H) ? tf32_rna(Whh[(long)k * 4 * H + gate * H + u]) : 0.0f;
  }
}

// fast gate nonlinearities for the fused epilogues: ex2.approx-based, abs. error ~2e-7 (vs 1e-3 parity tolerance)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// The four gate nonlinearities of one LSTM cell with ONE reciprocal (Montgomery's batch inversion): sigma(x) = 1/(1+e^-x) and
// tanh(x) = 1 - 2/(1+e^2x) share r = 1/((1+a)(1+b)(1+c)(1+d)), so a cell costs 4 ex2 + 1 rcp on the MUFU pipe (16 lanes per
// clock per SM, the scarce unit of the fused epilogue) instead of 4 ex2 + 4 rcp; the extra products run on the FMA pipe.  Inputs are
// clamped to +-20 (sigma) / +-10 (tanh) so that the product of the four denominators stays below 6e34: the clamped functions differ
// from the exact ones by < 5e-9.  xf already includes the forget bias.
__device__ __forceinline__ void lstm_gates_fast(float xi, float xj, float xf, float xo, float& si, float& tj, float& sf, float& so) {
  constexpr float L2E = 1.4426950408889634f;
  xi = fminf(fmaxf(xi, -20.0f), 20.0f); xf = fminf(fmaxf(xf, -20.0f), 20.0f); xo = fminf(fmaxf(xo, -20.0f), 20.0f);
  xj = fminf(fmaxf(xj, -10.0f), 10.0f);
  const float A = 1.0f + ex2_approx(-L2E * xi), B = 1.0f + ex2_approx(2.0f * L2E * xj);
  const float Cf = 1.0f + ex2_approx(-L2E * xf), D = 1.0f + ex2_approx(-L2E * xo);
  const float AB = A * B, CD = Cf * D;
  const float r = rcp_approx(AB * CD);
  const float rAB = r * CD, rCD = r * AB;            // 1 / (A B), 1 / (C D)
  si = rAB * B;
  tj = fmaf(-2.0f * rAB, A, 1.0f);
  sf = rCD * D;
  so = rCD * Cf;
}

constexpr int RP_EW = 8;          // epilogue warps: two per TMEM lane quarter, splitting the slice's units
constexpr int RP_MAXTPC = 4;      // tiles per CTA whose cell state c is carried in registers across the steps
constexpr int RP_FWD_THREADS = 64 + 32 * RP_EW;

template <int U> struct RecSplit {           // units of a slice handled by epilogue half 0 / half 1 (multiples of 4)
  static constexpr int U0 = (U / 4 + 1) / 2 * 4, U1 = U - U0;
};
struct RecFwdMaps {       // [direction]; z/cc/hx/hp: boxes {U units, 32 rows} of one TMEM lane quarter
  CUtensorMap a[2], w[2];
  CUtensorMap z[2], cc[2], hx[2], hp[2];
};

#ifdef ICL_EXPERIMENTS      // the first-generation kernels (k_rec_fwd: TF32 operands, everything through narrow TMA boxes; k_rec_bwd:
                            // cooperative grid-barrier BPTT): superseded by lstm_fwd16.cuh / lstm_bptt.cuh, kept for A/B builds only
template <int U>
constexpr int rec_fwd_smem(int nkb) { return nkb * 4 * U * 128 + RP_ASTAGES * 16384 + 7 * RP_ROWS * U * 4 + 512 + 1024; }

template <int U>
__global__ void __launch_bounds__(RP_FWD_THREADS, 1) k_rec_fwd(const __grid_constant__ RecFwdMaps maps, const RecArgs g) {
  constexpr int N = 4 * U;
  constexpr int U0 = RecSplit<U>::U0, U1 = RecSplit<U>::U1;
  constexpr int NACC = (512 / N) < RP_MAXACC ? (512 / N) : RP_MAXACC;
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_off[RP_MAXT + 2], s_n[RP_MAXT + 2];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sW = base;
  const uint32_t sA = sW + (uint32_t)g.nkb * N * 128;
  const uint32_t sE = sA + RP_ASTAGES * 16384;                           // 4 quarters x 7 boxes [32 rows x U]
  const uint32_t bars = sE + 7 * RP_ROWS * U * 4;
  const uint32_t full0 = bars, empty0 = bars + 8 * RP_ASTAGES, wfull = bars + 16 * RP_ASTAGES, efull0 = wfull + 8,
                 tfull0 = efull0 + 8 * RP_EW, tempty0 = tfull0 + 8 * RP_MAXACC, tmem_slot = tempty0 + 8 * RP_MAXACC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_dir = g.P * g.nsl;
  const int d = blockIdx.x / per_dir, p = (blockIdx.x % per_dir) / g.nsl, j = blockIdx.x % g.nsl;
  const int H = g.H, Tmax = g.Tmax;
  unsigned* flags = g.flags + (size_t)d * g.max_tiles;

  for (int i = threadIdx.x; i <= Tmax; i += blockDim.x) { s_off[i] = g.off[i]; s_n[i] = g.nact[i]; }
  if (threadIdx.x == 0) {
    for (int s = 0; s < RP_ASTAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(wfull, 1);
    for (int w = 0; w < 4; w++) mbar_init(efull0 + 8 * w, 1);
    for (int a = 0; a < NACC; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, RP_EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const unsigned per_step = (unsigned)(g.nsl * 4);                      // counter increments per (step, tile): slices x quarters

  if (warp == 0) {
    if (lane == 0) {                                                   // ---- A producer (h_{k-1} tiles) + resident W slice
      mbar_expect_tx(wfull, (uint32_t)g.nkb * N * 128);
      for (int kb = 0; kb < g.nkb; kb++) tma_load_2d(sW + kb * N * 128, &maps.w[d], kb * 32, j * N, wfull);
      uint32_t it = 0;
      Tracer tr; tr.init(g, 0);
      for (int k = 1; k < Tmax; k++) {
        for (int t = p; t * RP_ROWS < s_n[k]; t += g.P) {
          tr.ev(0, k, t);
          flag_wait(flags + t, per_step * k);
          fence_async_all();
          tr.ev(1, k, t);
          for (int kb = 0; kb < g.nkb; kb++, it++) {
            const uint32_t s = it % RP_ASTAGES;
            mbar_wait(empty0 + 8 * s, ((it / RP_ASTAGES) & 1) ^ 1);
            mbar_expect_tx(full0 + 8 * s, 16384);
            tma_load_2d(sA + s * 16384, &maps.a[d], kb * 32, s_off[k] + t * RP_ROWS, full0 + 8 * s);
          }
          tr.ev(2, k, t);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                                   // ---- MMA issuer
      constexpr uint32_t idesc = make_idesc(2, false, false, RP_ROWS, N);
      mbar_wait(wfull, 0);
      uint32_t it = 0, acc = 0;
      Tracer tr; tr.init(g, 1);
      for (int k = 1; k < Tmax; k++) {
        for (int t = p; t * RP_ROWS < s_n[k]; t += g.P, acc++) {
          const uint32_t slot = acc % NACC;
          mbar_wait(tempty0 + 8 * slot, ((acc / NACC) & 1) ^ 1);
          tc_fence_after();
          tr.ev(0, k, t);
          for (int kb = 0; kb < g.nkb; kb++, it++) {
            const uint32_t s = it % RP_ASTAGES;
            mbar_wait(full0 + 8 * s, (it / RP_ASTAGES) & 1);
            tc_fence_after();
            const int nk = min(4, g.nk8 - kb * 4);
            for (int kk = 0; kk < nk; kk++) {
              const uint64_t ad = make_smem_desc(sA + s * 16384 + kk * 32, 16, 1024);
              const uint64_t bd = make_smem_desc(sW + kb * N * 128 + kk * 32, 16, 1024);
              tc_mma_tf32(tmem + slot * N, ad, bd, idesc, (kb | kk) != 0);
            }
            tc_commit(empty0 + 8 * s);
          }
          tc_commit(tfull0 + 8 * slot);
          tr.ev(1, k, t);
        }
      }
    }
  } else {
    // ---- 8 epilogue warps = 4 TMEM lane quarters x 2 unit halves.  A quarter owns rows [32q, 32q+32) x the slice's U units
    // of EVERY tile of this CTA: it TMA-loads its own Zx boxes {U x 32}, its two warps split the units (hs = 0/1), keep
    // their cell state c in registers across the steps, and the quarter TMA-stores its outputs and publishes them on the
    // tile's counter.  Only a 64-thread named barrier couples the two warps; no CTA-wide barrier in the step loop.
    const int ew = warp - 2, q = warp & 3, hs = ew >> 2;
    const int UH = hs ? U1 : U0, ubase = hs ? U0 : 0;                      // my units of the slice: [ubase, ubase + UH)
    constexpr int BOXB = 32 * U * 4;
    const uint32_t sMine = sE + (uint32_t)q * (7 * BOXB);
    float* zb = reinterpret_cast<float*>(gbase + (sMine - base)) + ubase;   // [4 gates][32][U], then c, h, hr [32][U]
    float* cb = zb + 4 * 32 * U;
    float* hb = cb + 32 * U;
    float* hrb = hb + 32 * U;
    const CUtensorMap *mz = &maps.z[d], *mc = &maps.cc[d], *mh = &maps.hx[d], *mp = &maps.hp[d];
    const uint32_t efull = efull0 + 8 * q;
    const int ucol = j * U;                                              // first hidden unit of the slice
    const bool issuer = hs == 0 && lane == 0;                            // the quarter's TMA thread
    float cst[RP_MAXTPC][U0];                                            // carried cell state (U0 >= U1)
#pragma unroll
    for (int i = 0; i < RP_MAXTPC; i++)
#pragma unroll
      for (int u = 0; u < U0; u++) cst[i][u] = 0.0f;
    Tracer tr; tr.init(g, 3);
    if (ew != 0 || lane != 0) tr.p = nullptr;
    uint32_t e = 0, acc = 0;
    if (issuer && p * RP_ROWS < s_n[0]) {                                // prefetch the first tile's Zx boxes
      mbar_expect_tx(efull, 4 * BOXB);
      for (int gate = 0; gate < 4; gate++) tma_load_2d(sMine + gate * BOXB, mz, gate * H + ucol, s_off[0] + p * RP_ROWS + 32 * q, efull);
    }
    for (int k = 0; k < Tmax; k++) {
#pragma unroll
      for (int i = 0; i < RP_MAXTPC; i++) {
        const int t = p + i * g.P;
        if (t * RP_ROWS >= s_n[k]) break;
        tr.ev(0, k, t);
        mbar_wait(efull, e & 1);
        e++;
        tr.ev(1, k, t);
        uint32_t slot = 0;
        if (k > 0) {
          slot = acc % NACC;
          mbar_wait(tfull0 + 8 * slot, (acc / NACC) & 1);
          tc_fence_after();
        }
        tr.ev(2, k, t);
#pragma unroll
        for (int c = 0; c < U0 / 4; c++) {
          if (c * 4 < UH) {
            uint32_t a[16];
            if (k > 0) tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + slot * N + (ubase / 4 + c) * 16, a);
            else {
#pragma unroll
              for (int x = 0; x < 16; x++) a[x] = 0u;
            }
            float4 z4[4];
#pragma unroll
            for (int gate = 0; gate < 4; gate++) z4[gate] = *reinterpret_cast<float4*>(zb + gate * 32 * U + lane * U + c * 4);
            const float *zi = &z4[0].x, *zj = &z4[1].x, *zf = &z4[2].x, *zo = &z4[3].x;
            float si[4], tj[4], sf[4], so[4], hn[4], hr[4];
#pragma unroll
            for (int x = 0; x < 4; x++) {
              si[x] = sigmoid_fast(zi[x] + __uint_as_float(a[x]));
              tj[x] = tanh_fast(zj[x] + __uint_as_float(a[4 + x]));
              sf[x] = sigmoid_fast(zf[x] + __uint_as_float(a[8 + x]) + 1.0f);
              so[x] = sigmoid_fast(zo[x] + __uint_as_float(a[12 + x]));
              const float cn = cst[i][c * 4 + x] * sf[x] + si[x] * tj[x];
              cst[i][c * 4 + x] = cn;
              hn[x] = tanh_fast(cn) * so[x];
              hr[x] = tf32_rna(hn[x]);
            }
            if (g.training) {
              *reinterpret_cast<float4*>(zb + 0 * 32 * U + lane * U + c * 4) = make_float4(si[0], si[1], si[2], si[3]);
              *reinterpret_cast<float4*>(zb + 1 * 32 * U + lane * U + c * 4) = make_float4(tj[0], tj[1], tj[2], tj[3]);
              *reinterpret_cast<float4*>(zb + 2 * 32 * U + lane * U + c * 4) = make_float4(sf[0], sf[1], sf[2], sf[3]);
              *reinterpret_cast<float4*>(zb + 3 * 32 * U + lane * U + c * 4) = make_float4(so[0], so[1], so[2], so[3]);
              *reinterpret_cast<float4*>(cb + lane * U + c * 4) =
                  make_float4(cst[i][c * 4], cst[i][c * 4 + 1], cst[i][c * 4 + 2], cst[i][c * 4 + 3]);
            }
            *reinterpret_cast<float4*>(hb + lane * U + c * 4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            *reinterpret_cast<float4*>(hrb + lane * U + c * 4) = make_float4(hr[0], hr[1], hr[2], hr[3]);
          }
        }
        if (k > 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * slot);
          acc++;
        }
        fence_async_smem();                                             // generic-proxy smem writes -> visible to the TMA engine
        asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");       // both unit halves of the quarter are in smem
        tr.ev(3, k, t);
        if (issuer) {
          const int row = s_off[k] + t * RP_ROWS + 32 * q;
          if (t * RP_ROWS < s_n[k + 1]) tma_store_2d(mp, sMine + 6 * BOXB, ucol, s_off[k + 1] + t * RP_ROWS + 32 * q);
          bulk_commit();                                               // group A: what step k+1 of the other slices waits for
          tma_store_2d(mh, sMine + 5 * BOXB, ucol, row);
          if (g.training) {
            tma_store_2d(mc, sMine + 4 * BOXB, ucol, row);
            for (int gate = 0; gate < 4; gate++) tma_store_2d(mz, sMine + gate * BOXB, gate * H + ucol, row);
          }
          bulk_commit();                                               // group B
          bulk_wait_read<0>();                                         // boxes may be refilled
          tr.ev(4, k, t);
          int k2 = k, t2 = t + g.P;                                    // next tile of this CTA's schedule
          if (i + 1 >= RP_MAXTPC || t2 * RP_ROWS >= s_n[k]) { k2 = k + 1; t2 = p; }
          if (k2 < Tmax && t2 * RP_ROWS < s_n[k2]) {
            mbar_expect_tx(efull, 4 * BOXB);
            for (int gate = 0; gate < 4; gate++)
              tma_load_2d(sMine + gate * BOXB, mz, gate * H + ucol, s_off[k2] + t2 * RP_ROWS + 32 * q, efull);
          }
          bulk_wait<1>();                                               // group A is complete in global memory
          tr.ev(5, k, t);
        }
        if (hs == 0) {
          __syncwarp();
          if (lane == 1) flag_release_add(flags + t);                   // a lane with no bulk copies in flight publishes the tile
          if (lane == 0) tr.ev(6, k, t);
        }
      }
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// =====================================================================================================================
// Backward recurrence (k_rec_bwd, "K3"): ONE cooperative launch runs every BPTT step of both directions.
//
// Per step k (descending) the whole grid alternates two phases separated by a software grid barrier:
//   E  all threads: backward of the LSTM cell for every running row of both directions (float4, coalesced along the hidden
//      units): dh = dHout_k + dh_rec, gates -> dZ_k in place (TF32-rounded), dc carry; dh_rec is left zeroed;
//   G  dh_rec[0:n_k) = dZ_k W_hh^T as 128x128 tcgen05 tiles (TMA -> 6-stage ring -> UMMA kind::tf32 -> TMEM -> red.global),
//      tasks (direction, row tile, unit tile, K split) dealt round-robin to the CTAs; W_hh^T streams from L2 (a [128 x 4H]
//      slice is 614 KB, it cannot stay resident next to the operand ring).
// dZ_k is written by generic stores and read by TMA in the next phase: writers fence.proxy.async before the barrier.
// =====================================================================================================================
constexpr int RB_THREADS = 256, RB_STAGES = 6;
constexpr int RB_SMEM = RB_STAGES * 2 * 16384 + 4 * 32 * 33 * 4 + 256 + 1024;

struct RecBwdMaps { CUtensorMap za[2], wb[2]; };   // A: Z[d] {4H, rows}; B: W_hh {4H, H} (rows = hidden units, K-major); boxes {32, 128}
struct RecBwdArgs {
  const int* off; const int* nact;
  int Tmax, H, round_ops;
  float* Z[2]; const float* Cc[2]; const float* dHout[2]; float* dhrec[2]; float* dcc[2];
  unsigned* bar;                                    // grid-barrier counter, zeroed before the launch
  long long* trace; int trace_cta;                  // optional bring-up trace (see Tracer)
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    for (;;) {                                        // tight poll: the barrier is on the critical path of every step
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 4000000000LL) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(RB_THREADS, 1) k_rec_bwd(const __grid_constant__ RecBwdMaps maps, const RecBwdArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_off[RP_MAXT + 2], s_n[RP_MAXT + 2];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = base + RB_STAGES * 16384, sStg = sB + RB_STAGES * 16384;
  const uint32_t bars = sStg + 4 * 32 * 33 * 4;
  const uint32_t full0 = bars, empty0 = bars + 8 * RB_STAGES, tfull = bars + 16 * RB_STAGES, tempty = tfull + 8, tmem_slot = tempty + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = g.H, Tmax = g.Tmax, q4 = H >> 2;
  const int total_kb = (4 * H + 31) / 32, Nt = (H + 127) / 128;

  for (int i = threadIdx.x; i <= Tmax; i += blockDim.x) { s_off[i] = g.off[i]; s_n[i] = g.nact[i]; }
  if (threadIdx.x == 0) {
    for (int s = 0; s < RB_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  uint32_t it = 0, acc = 0;                            // running k-block / accumulator counters (barrier parities)
  unsigned bar_target = 0;
  Tracer tr; tr.init(g.trace, g.trace_cta, warp == 0 ? 0 : warp == 1 ? 1 : warp == 2 ? 2 : 3);
  if (lane != 0 || warp > 2) tr.p = nullptr;
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (long)gridDim.x * blockDim.x;
  for (int k = Tmax - 1; k >= 0; k--) {
    const int n = s_n[k];
    const long o = s_off[k], op = k > 0 ? s_off[k - 1] : 0;
    // ---------------------------------------------------------------- E: cell backward, both directions
    const long per_dir = (long)n * q4;
    tr.ev(0, k, 0);
    for (long idx = gtid; idx < 2 * per_dir; idx += gthreads) {
      const int d = idx >= per_dir;
      const long r = idx - (d ? per_dir : 0);
      const int m = (int)(r / q4), u = (int)(r % q4) * 4;
      float* z = g.Z[d] + (o + m) * 4 * H + u;
      float4 g4[4];
#pragma unroll
      for (int a = 0; a < 4; a++) g4[a] = *reinterpret_cast<float4*>(z + a * H);
      const float4 c4 = *reinterpret_cast<const float4*>(g.Cc[d] + (o + m) * H + u);
      const float4 cp4 = k > 0 ? *reinterpret_cast<const float4*>(g.Cc[d] + (op + m) * H + u) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 dh4 = *reinterpret_cast<const float4*>(g.dHout[d] + (o + m) * H + u);
      float* drp = g.dhrec[d] + (long)m * H + u;
      const float4 dr4 = *reinterpret_cast<float4*>(drp);
      *reinterpret_cast<float4*>(drp) = make_float4(0.f, 0.f, 0.f, 0.f);
      float* dcp = g.dcc[d] + (long)m * H + u;
      const float4 dc4 = *reinterpret_cast<float4*>(dcp);
      const float *si = &g4[0].x, *tj = &g4[1].x, *sf = &g4[2].x, *so = &g4[3].x, *c = &c4.x, *cp = &cp4.x, *dh = &dh4.x, *dr = &dr4.x,
                  *dc = &dc4.x;
      float di[4], dj[4], df[4], dgo[4], dcn[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const CellGrad cg = lstm_cell_bwd(si[j], tj[j], sf[j], so[j], c[j], cp[j], dh[j] + dr[j], dc[j]);
        di[j] = maybe_round(cg.di, g.round_ops); dj[j] = maybe_round(cg.dj, g.round_ops); df[j] = maybe_round(cg.df, g.round_ops);
        dgo[j] = maybe_round(cg.dg_o, g.round_ops); dcn[j] = cg.dc_prev;
      }
      *reinterpret_cast<float4*>(z) = make_float4(di[0], di[1], di[2], di[3]);
      *reinterpret_cast<float4*>(z + H) = make_float4(dj[0], dj[1], dj[2], dj[3]);
      *reinterpret_cast<float4*>(z + 2 * H) = make_float4(df[0], df[1], df[2], df[3]);
      *reinterpret_cast<float4*>(z + 3 * H) = make_float4(dgo[0], dgo[1], dgo[2], dgo[3]);
      *reinterpret_cast<float4*>(dcp) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
    }
    if (k == 0) break;
    tr.ev(1, k, 0);
    fence_async_all();                                  // dZ_k (generic stores) -> visible to the TMA reads of phase G
    tr.ev(2, k, 0);
    bar_target += gridDim.x;
    grid_barrier(g.bar, bar_target);
    tr.ev(3, k, 0);
    // ---------------------------------------------------------------- G: dh_rec = dZ_k W_hh^T
    const int Mt = (n + 127) / 128;
    int sp = (int)gridDim.x / max(1, 2 * Mt * Nt);
    sp = max(1, min(min(sp, 4), total_kb / 4));
    const int kb_per = (total_kb + sp - 1) / sp;
    const int tasks = 2 * Mt * Nt * sp;
    if (warp == 0) {
      if (lane == 0) {
        fence_async_all();
        for (int t = blockIdx.x; t < tasks; t += gridDim.x) {
          const int s = t % sp, nt = (t / sp) % Nt, mt = (t / (sp * Nt)) % Mt, d = t / (sp * Nt * Mt);
          const int kb0 = s * kb_per, nkb = min(kb_per, total_kb - kb0);
          for (int kb = 0; kb < nkb; kb++, it++) {
            const uint32_t st = it % RB_STAGES;
            mbar_wait(empty0 + 8 * st, ((it / RB_STAGES) & 1) ^ 1);
            mbar_expect_tx(full0 + 8 * st, 32768);
            tma_load_2d(sA + st * 16384, &maps.za[d], (kb0 + kb) * 32, (int)o + mt * 128, full0 + 8 * st);
            tma_load_2d(sB + st * 16384, &maps.wb[d], (kb0 + kb) * 32, nt * 128, full0 + 8 * st);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(2, false, false, 128, 128);
        for (int t = blockIdx.x; t < tasks; t += gridDim.x, acc++) {
          const int s = t % sp;
          const int kb0 = s * kb_per, nkb = min(kb_per, total_kb - kb0);
          mbar_wait(tempty, (acc & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < nkb; kb++, it++) {
            const uint32_t st = it % RB_STAGES;
            mbar_wait(full0 + 8 * st, (it / RB_STAGES) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
              tc_mma_tf32(tmem, make_smem_desc(sA + st * 16384 + kk * 32, 16, 1024), make_smem_desc(sB + st * 16384 + kk * 32, 16, 1024),
                          idesc, (kb | kk) != 0);
            tc_commit(empty0 + 8 * st);
          }
          tc_commit(tfull);
        }
      }
    } else if (warp < 6) {
      const int q = warp & 3;
      float* stg = reinterpret_cast<float*>(gbase + (sStg - base)) + q * (32 * 33);
      for (int t = blockIdx.x; t < tasks; t += gridDim.x, acc++) {
        const int nt = (t / sp) % Nt, mt = (t / (sp * Nt)) % Mt, d = t / (sp * Nt * Mt);
        mbar_wait(tfull, acc & 1);
        tc_fence_after();
        const int mrow0 = mt * 128 + q * 32, rows = min(32, n - mrow0);
        float* out = g.dhrec[d];
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          const int ncol = nt * 128 + c + lane;
          if (nt * 128 + c >= H) break;
          uint32_t r[32];
          tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, r);
#pragma unroll
          for (int j = 0; j < 32; j++) stg[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          if (ncol < H) {
            float* cp = out + (long)mrow0 * H + ncol;
            if (sp > 1) {
#pragma unroll 8
              for (int rr = 0; rr < rows; rr++) atomicAdd(cp + (long)rr * H, stg[rr * 33 + lane]);
            } else {
#pragma unroll 8
              for (int rr = 0; rr < rows; rr++) cp[(long)rr * H] = stg[rr * 33 + lane];
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
      }
    }
    tr.ev(4, k, 0);
    bar_target += gridDim.x;
    grid_barrier(g.bar, bar_target);
    tr.ev(5, k, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
  }
}

#endif  // ICL_EXPERIMENTS

}  // namespace icl
