// Persistent forward recurrence, second generation ("K2", k_rec_fwd16): same decomposition as k_rec_fwd (lstm_persistent.cuh) --
// CTA (direction, part, slice) keeps its slice of W_hh resident in shared memory for the whole launch, tiles are published on
// per-tile counters, no grid barrier -- with three changes that the clock64 traces of the first generation asked for:
//
//  1. fp16 operands for the recurrent product (tcgen05.mma kind::f16, fp32 accumulation in TMEM).  h is in (-1, 1) and W_hh is
//     O(0.1): fp16 carries the same 10 explicit mantissa bits as TF32 for these magnitudes (values below 6e-5 lose bits, an
//     absolute error < 3e-8), so the products are as exact as the TF32 ones; the resident W slice (61 KB instead of 102 KB) and
//     the streamed h tile (96 KB instead of 153 KB) halve.  The fp32 TF32-rounded copy of h that the weight-gradient GEMM
//     contracts over (XH) is still written.
//  2. The freed shared memory double-buffers the x-projection boxes: the boxes of tile i+2 are requested when tile i has been
//     consumed, so their latency (the dominant term of the 4 us per-tile epilogue chain) is off the critical path.
//  3. Only what the NEXT STEP needs goes through TMA + the publication protocol: the fp16 h rows (one 48-byte-row box per TMEM
//     lane quarter).  Gates, c, exact h and the TF32 h are read by later kernels only and are stored straight from registers
//     (7 narrow TMA boxes per tile and quarter -> 1).
// Measured (card2048): 0.399 -> 0.378 ms.  The clock64 trace (tools/trace_fwd16.py) shows what bounds both generations: 4.2 us
// per tile between "accumulator ready" and "tile handed over" -- every CTA writes a 20-unit-wide column slab of the gate / c / h
// matrices, i.e. 80-byte row segments (2.5 sectors), through the LSU (here) or as narrow TMA boxes (first generation).  A variant
// with a dedicated publisher warp (store completion + counter off the epilogue loop) was not faster (0.399 ms) and was dropped.
//
// fp16 h layout: Hp16 [rows][KP], slice j's 20 units at columns [24 j, 24 j + 20) + 4 zero columns, so that a quarter's box row
// is 48 bytes (TMA needs multiples of 16); the packed W slice has zero rows at the pad positions.
#pragma once
#include <cuda_fp16.h>

#include "lstm_persistent.cuh"

namespace icl {

constexpr int RF_MAXST = 6, RF_EW = 8, RF_SW = 4, RF_THREADS = 64 + 32 * (RF_EW + RF_SW) + 32 /* counter-polling warp */, RF_MAXACC = 8;
template <int U> struct RF {
  static constexpr int UP = (U * 2) % 16 == 0 ? U : (U + 7) / 8 * 8;         // units of a slice as stored in Hp16 (padded to 16 bytes)
  static constexpr int N = 4 * U;
  static constexpr int U0 = RecSplit<U>::U0, U1 = RecSplit<U>::U1;
  static constexpr bool SWZ = U == 16;        // 64-byte box rows: SWIZZLE_64B boxes (a lane-per-row float4 read of plain 64-byte rows is a
                                              // 4-way bank conflict; 80-byte rows are conflict-free as they are) + the ragged-H tensor maps
  static constexpr int MAXTPC = U == 16 ? 6 : 4;                              // tiles per CTA whose cell state is carried in registers
  static constexpr int ZBOX = 32 * U * 4;                                     // one gate box of a quarter
  static constexpr int HBOX = 32 * UP * 2;                                    // the fp16 h box of a quarter
};
// The h_{k-1} ring is as deep as shared memory allows, up to a whole tile (H = 300: 3 of 6 k-blocks, H = 200: all 4)
template <int U> constexpr int rec_fwd16_fixed(int nkb) { return nkb * RF<U>::N * 128 + 4 * (2 * 5 * RF<U>::ZBOX + 2 * RF<U>::HBOX) + 512 + 1024; }
template <int U> constexpr int rec_fwd16_stages(int nkb) {
  int st = (227 * 1024 - rec_fwd16_fixed<U>(nkb)) / 16384;
  return st > RF_MAXST ? RF_MAXST : st > nkb ? nkb : st;
}
template <int U> constexpr int rec_fwd16_smem(int nkb) { return rec_fwd16_fixed<U>(nkb) + rec_fwd16_stages<U>(nkb) * 16384; }

struct RecFwd16Maps { CUtensorMap a[2], w[2], z[2], cc[2], hp16[2]; };   // a: Hp16 box {64 halves, 128 rows} SW128; w: packed W box {64, 4U} SW128;
                                                                  // z: Z box {U, 32}; hp16: Hp16 box {UP, 32}
struct RecFwd16Args {
  const int* off; const int* nact;
  int Tmax, H, nsl, P, nkb, nk16, nst, max_tiles, training, ldx;
  unsigned* flags;
  float* Z[2]; float* Cc[2]; float* Hx[2]; float* Hp[2];           // generic-store targets (Hp = TF32 h rows inside XH, pitch ldx)
  long long* trace; int trace_cta;
};

// Packed fp16 W_hh: row (j*4U + n), n = c*16 + gate*4 + i <-> unit j*U + 4c + i (as k_pack_whh_fwd); column k' = (k / U) * UP + k % U.
// 32x32 tile transpose: reads coalesced along the 4H gate columns of W_hh [H, 4H], writes coalesced along k'.  The pad columns
// (k' % UP >= U, k' >= nsl*UP) are never written: the buffer is zeroed once at allocation.  grid (4H/32, H/32, 2 dirs), block (32, 8).
__global__ void k_pack_whh_fwd16(const float* __restrict__ Whh0, const float* __restrict__ Whh1, __half* __restrict__ Wp0,
                                 __half* __restrict__ Wp1, int H, int U, int UP, int nsl, int Kp) {
  __shared__ float tile[32][33];
  const float* Whh = blockIdx.z ? Whh1 : Whh0;
  __half* Wp = blockIdx.z ? Wp1 : Wp0;
  const int col0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int k = k0 + r, col = col0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < H && col < 4 * H) ? Whh[(long)k * 4 * H + col] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int col = col0 + r, k = k0 + threadIdx.x;
    if (col < 4 * H && k < H) {
      const int gate = col / H, u = col % H, j = u / U, uu = u % U;
      const int row = j * 4 * U + (uu / 4) * 16 + gate * 4 + (uu % 4);
      Wp[(long)row * Kp + (k / U) * UP + k % U] = __float2half_rn(tile[threadIdx.x][r]);
    }
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// U = 20: H must be a multiple of 20; the x-projection / gate map `z` is 2-D over [rows][4H].
// U = 16: any H (multiple of 4): `z` is 3-D over [rows][4 gates][H] and `cc` 2-D over [rows][H], so the boxes of the LAST slice are
// clipped at H (loads read zeros, stores skip) instead of running into the next gate's columns; the slice's pad units compute
// finite values that are never stored (their W_hh rows and fp16 h columns are zero).
template <int U>
__global__ void __launch_bounds__(RF_THREADS, 1) k_rec_fwd16(const __grid_constant__ RecFwd16Maps maps, const RecFwd16Args g) {
  constexpr int N = RF<U>::N, UP = RF<U>::UP, U0 = RF<U>::U0, U1 = RF<U>::U1, ZBOX = RF<U>::ZBOX, HBOX = RF<U>::HBOX;
  constexpr int RF_MAXTPC = RF<U>::MAXTPC;
  constexpr bool SWZ = RF<U>::SWZ;
  constexpr int NACC = (512 / N) < RF_MAXACC ? (512 / N) : RF_MAXACC;
  constexpr int QBYTES = 2 * 5 * ZBOX + 2 * HBOX;                            // per quarter: two sets of (4 gate boxes + c box) + two fp16 h boxes
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_off[RP_MAXT + 2], s_n[RP_MAXT + 2];
  __shared__ int s_flagged;                                                 // tiles whose h_{k-1} rows are known to be published
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sW = base;
  const uint32_t sA = sW + (uint32_t)g.nkb * N * 128;
  const int NST = g.nst;                                                    // stages of the h_{k-1} ring
  const uint32_t sE = sA + (uint32_t)NST * 16384;
  const uint32_t bars = sE + 4 * QBYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * RF_MAXST, wfull = bars + 16 * RF_MAXST, efull0 = wfull + 8,   // efull: [quarter][set]
                 sready0 = efull0 + 8 * 8,                                                    // sready: [quarter][set], as efull
                 tfull0 = sready0 + 8 * 8, tempty0 = tfull0 + 8 * RF_MAXACC, tmem_slot = tempty0 + 8 * RF_MAXACC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_dir = g.P * g.nsl;
  const int d = blockIdx.x / per_dir, p = (blockIdx.x % per_dir) / g.nsl, j = blockIdx.x % g.nsl;
  const int H = g.H, Tmax = g.Tmax;
  unsigned* flags = g.flags + (size_t)d * g.max_tiles;

  for (int i = threadIdx.x; i <= Tmax; i += blockDim.x) { s_off[i] = g.off[i]; s_n[i] = g.nact[i]; }
  if (threadIdx.x == 0) {
    s_flagged = 0;
    for (int s = 0; s < RF_MAXST; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(wfull, 1);
    for (int w = 0; w < 8; w++) { mbar_init(efull0 + 8 * w, 1); mbar_init(sready0 + 8 * w, 2); }
    for (int a = 0; a < NACC; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, RF_EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the pad columns of the fp16 h boxes stay zero for the whole launch
  for (int i = threadIdx.x; i < 4 * HBOX; i += blockDim.x) {              // 2 boxes of HBOX / 2 halves per quarter
    const int q = i / HBOX, e = i % HBOX;
    reinterpret_cast<__half*>(gbase + (sE - base) + q * QBYTES + 2 * 5 * ZBOX)[e] = __float2half_rn(0.0f);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const unsigned per_step = (unsigned)(g.nsl * 4);                      // counter increments per (step, tile): slices x quarters

  if (warp == 0) {
    // ---- A producer (fp16 h_{k-1} tiles) + resident W slice.  The tile rate of the kernel is set HERE.  What the clock64 traces
    // and tools/probe_tma.cu (profiles/r2_probe_tma*.txt) say: ONE thread gets a TMA load accepted every 300 - 430 ns whatever its
    // size or the ring depth, several lanes issue concurrently (100 ns per load from six lanes); an ld.acquire.gpu of a publication
    // counter is an L2 round trip of 1.2 - 1.6 us even when the tile was published long ago.  So: the counters are polled AHEAD by
    // the last warp, which reports in shared memory how far the producer may go, and when the whole tile fits the ring (U = 16:
    // NST == nkb) its k-blocks are requested by nkb lanes at once.  With a shallower ring (U = 20, H = 300: 3 stages for 6
    // k-blocks) a stage's turn-around (~1.2 us) bounds the tile and one lane issues in order (parallel lanes measured slower there).
    if (lane == 0) {
      mbar_expect_tx(wfull, (uint32_t)g.nkb * N * 128);
      for (int kb = 0; kb < g.nkb; kb++) tma_load_2d(sW + kb * N * 128, &maps.w[d], kb * 64, j * N, wfull);
    }
    auto wait_published = [&](int seq) {
      const long long t0 = clock64();
      for (;;) {
        int v;
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&s_flagged)) : "memory");
        if (v > seq) break;
        if (clock64() - t0 > 4000000000LL) __trap();
      }
      fence_async_all();
    };
    Tracer tr; tr.init(g.trace, g.trace_cta, 0);
    if (lane != 0) tr.p = nullptr;
    if (NST >= g.nkb) {
      if (lane < g.nkb) {                                              // lane = k-block = ring stage
        int seq = 0;
        for (int k = 1; k < Tmax; k++) {
          for (int t = p; t * RP_ROWS < s_n[k]; t += g.P, seq++) {
            tr.ev(0, k, t);
            wait_published(seq);
            tr.ev(1, k, t);
            mbar_wait(empty0 + 8 * lane, (seq & 1) ^ 1);
            mbar_expect_tx(full0 + 8 * lane, 16384);
            tma_load_2d(sA + lane * 16384, &maps.a[d], lane * 64, s_off[k] + t * RP_ROWS, full0 + 8 * lane);
            tr.ev(2, k, t);
          }
        }
      }
    } else if (lane == 0) {
      uint32_t it = 0;
      int seq = 0;
      for (int k = 1; k < Tmax; k++) {
        for (int t = p; t * RP_ROWS < s_n[k]; t += g.P, seq++) {
          tr.ev(0, k, t);
          wait_published(seq);
          tr.ev(1, k, t);
          for (int kb = 0; kb < g.nkb; kb++, it++) {
            const uint32_t s = it % NST;
            mbar_wait(empty0 + 8 * s, ((it / NST) & 1) ^ 1);
            mbar_expect_tx(full0 + 8 * s, 16384);
            tma_load_2d(sA + s * 16384, &maps.a[d], kb * 64, s_off[k] + t * RP_ROWS, full0 + 8 * s);
          }
          tr.ev(2, k, t);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                                   // ---- MMA issuer
      constexpr uint32_t idesc = make_idesc(0, false, false, RP_ROWS, N);   // kind::f16, fp16 x fp16 -> fp32
      mbar_wait(wfull, 0);
      uint32_t it = 0, acc = 0;
      const bool par = NST >= g.nkb;                                   // lane-parallel producer: stage = k-block, phase = tile
      Tracer tr; tr.init(g.trace, g.trace_cta, 1);
      for (int k = 1; k < Tmax; k++) {
        for (int t = p; t * RP_ROWS < s_n[k]; t += g.P, acc++) {
          const uint32_t slot = acc % NACC;
          mbar_wait(tempty0 + 8 * slot, ((acc / NACC) & 1) ^ 1);
          tc_fence_after();
          tr.ev(0, k, t);
          for (int kb = 0; kb < g.nkb; kb++, it++) {
            const uint32_t s = par ? (uint32_t)kb : it % NST;
            mbar_wait(full0 + 8 * s, par ? (acc & 1) : ((it / NST) & 1));
            tc_fence_after();
            if (kb == 0) tr.ev(1, k, t);
            const int nk = min(4, g.nk16 - kb * 4);                    // UMMA_K = 16 halves = 32 bytes
            for (int kk = 0; kk < nk; kk++) {
              const uint64_t ad = make_smem_desc(sA + s * 16384 + kk * 32, 16, 1024);
              const uint64_t bd = make_smem_desc(sW + kb * N * 128 + kk * 32, 16, 1024);
              tc_mma_f16(tmem + slot * N, ad, bd, idesc, (kb | kk) != 0);
            }
            tc_commit(empty0 + 8 * s);
          }
          tc_commit(tfull0 + 8 * slot);
          tr.ev(2, k, t);
        }
      }
    }
  } else if (warp < 2 + RF_EW) {
    // ---- 8 cell warps = 4 TMEM lane quarters x 2 unit halves (as k_rec_fwd).  They only compute: a tile's boxes (gates, c, fp16 h)
    // are handed to the quarter's store warp through the `sready` mbarrier, so the next tile's cells start at once
    const int ew = warp - 2, q = warp & 3, hs = ew >> 2;
    const int UH = hs ? U1 : U0, ubase = hs ? U0 : 0;
    const uint32_t sMine = sE + (uint32_t)q * QBYTES;
    float* zset[2] = {reinterpret_cast<float*>(gbase + (sMine - base)), reinterpret_cast<float*>(gbase + (sMine - base) + 5 * ZBOX)};
    __half* const hbox0 = reinterpret_cast<__half*>(gbase + (sMine - base) + 2 * 5 * ZBOX);   // two boxes, alternating per tile
    const int ucol = j * U;
    float* const Hd = g.Hx[d]; float* const Hpd = g.Hp[d];
    float cst[RF_MAXTPC][U0];
#pragma unroll
    for (int i = 0; i < RF_MAXTPC; i++)
#pragma unroll
      for (int u = 0; u < U0; u++) cst[i][u] = 0.0f;
    uint32_t n_tile = 0, acc = 0;                                        // tiles processed (set = n_tile & 1), accumulators consumed
    Tracer tr; tr.init(g.trace, g.trace_cta, 3);
    if (ew != 0 || lane != 0) tr.p = nullptr;
    for (int k = 0; k < Tmax; k++) {
      const int n_k = s_n[k], n_k1 = s_n[k + 1];
      const long o_k = s_off[k], o_k1 = s_off[k + 1];
#pragma unroll
      for (int i = 0; i < RF_MAXTPC; i++) {
        const int t = p + i * g.P;
        if (t * RP_ROWS >= n_k) break;
        const int set = n_tile & 1;
        tr.ev(0, k, t);
        mbar_wait(efull0 + 8 * (q * 2 + set), (n_tile >> 1) & 1);
        tr.ev(1, k, t);
        uint32_t slot = 0;
        if (k > 0) {
          slot = acc % NACC;
          mbar_wait(tfull0 + 8 * slot, (acc / NACC) & 1);
          tc_fence_after();
        }
        tr.ev(2, k, t);
        float* zb = zset[set];
        __half* hbox = hbox0 + set * (HBOX / 2);
        // position of unit chunk ch (4 floats) of this lane's row inside a gate box: plain, or XOR-ed as SWIZZLE_64B stores it
        auto zpos = [&](int gate, int ch) { return gate * 32 * U + lane * U + (SWZ ? ((ch ^ ((lane >> 1) & 3)) << 2) : (ch << 2)); };
        const long row = o_k + t * RP_ROWS + 32 * q + lane;
        const bool has_next = t * RP_ROWS < n_k1;
        const long row_n = o_k1 + t * RP_ROWS + 32 * q + lane;
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
            for (int gate = 0; gate < 4; gate++) z4[gate] = *reinterpret_cast<const float4*>(zb + zpos(gate, ubase / 4 + c));
            const float *zi = &z4[0].x, *zj = &z4[1].x, *zf = &z4[2].x, *zo = &z4[3].x;
            float si[4], tj[4], sf[4], so[4], hn[4], hr[4];
#pragma unroll
            for (int x = 0; x < 4; x++) {
              lstm_gates_fast(zi[x] + __uint_as_float(a[x]), zj[x] + __uint_as_float(a[4 + x]), zf[x] + __uint_as_float(a[8 + x]) + 1.0f,
                              zo[x] + __uint_as_float(a[12 + x]), si[x], tj[x], sf[x], so[x]);
              const float cn = cst[i][c * 4 + x] * sf[x] + si[x] * tj[x];
              cst[i][c * 4 + x] = cn;
              hn[x] = tanh_fast(cn) * so[x];
              hr[x] = tf32_rna(hn[x]);
            }
            const int u = ucol + ubase + c * 4;
            if (g.training) {                                          // read by the backward pass only
              // the gates go back into the x-projection boxes and out as four TMA boxes (the LSU path alone was the bottleneck:
              // every lane-per-row float4 store is 32 wavefronts); c / h / TF32 h stay on the LSU, the two engines overlap (moving
              // c to the LSU as well to free a 4th stage of the h_{k-1} ring measured slower: 0.337 vs 0.321 ms)
              *reinterpret_cast<float4*>(zb + zpos(0, ubase / 4 + c)) = make_float4(si[0], si[1], si[2], si[3]);
              *reinterpret_cast<float4*>(zb + zpos(1, ubase / 4 + c)) = make_float4(tj[0], tj[1], tj[2], tj[3]);
              *reinterpret_cast<float4*>(zb + zpos(2, ubase / 4 + c)) = make_float4(sf[0], sf[1], sf[2], sf[3]);
              *reinterpret_cast<float4*>(zb + zpos(3, ubase / 4 + c)) = make_float4(so[0], so[1], so[2], so[3]);
              *reinterpret_cast<float4*>(zb + zpos(4, ubase / 4 + c)) = make_float4(cst[i][c * 4], cst[i][c * 4 + 1], cst[i][c * 4 + 2], cst[i][c * 4 + 3]);
              if (has_next && u < H) *reinterpret_cast<float4*>(Hpd + row_n * g.ldx + u) = make_float4(hr[0], hr[1], hr[2], hr[3]);
            }
            if (u < H) *reinterpret_cast<float4*>(Hd + row * H + u) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            __half2 h01 = __floats2half2_rn(hn[0], hn[1]), h23 = __floats2half2_rn(hn[2], hn[3]);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
            *reinterpret_cast<uint2*>(hbox + lane * UP + ubase + c * 4) = pk;       // next step's operand rows (fp16)
          }
        }
        tr.ev(4, k, t);
        if (k > 0) tc_fence_before();
        fence_async_smem();                                             // the boxes (generic smem writes) -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if (k > 0) mbar_arrive(tempty0 + 8 * slot);
          mbar_arrive(sready0 + 8 * (q * 2 + set));                     // this half of the quarter's boxes is ready to leave
        }
        if (k > 0) acc++;
        tr.ev(3, k, t);
        n_tile++;
        // hazards: set `set` (gate boxes + fp16 box) is next written for tile n_tile + 2, after its `efull` barrier has fired; the
        // store warp requests that refill only when the TMA stores of THIS tile have read the boxes
      }
    }
  } else if (warp == 2 + RF_EW + RF_SW) {
    if (lane == 0) {                                                   // ---- publication counters, polled ahead of the producer
      int seq = 0;
      for (int k = 1; k < Tmax; k++) {
        for (int t = p; t * RP_ROWS < s_n[k]; t += g.P) {
          flag_wait(flags + t, per_step * k);
          seq++;
          asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(&s_flagged)), "r"(seq) : "memory");
        }
      }
    }
  } else {
    // ---- 4 store warps, one per TMEM lane quarter: x-projection boxes in, gate / c / fp16-h boxes out, the tile's publication
    // counter.  Everything here used to sit at the end of the cell warps' tile loop (1.6 - 2 us of the 3.4 - 4 us per tile,
    // profiles/r2b_rec_fwd16_trace.txt).  The TMA operations of a tile are spread over SIX LANES (a thread's bulk copies are
    // processed about two at a time): lane g < 4 stores gate box g and requests the x-projection box that replaces it, lane 4
    // stores the c box, lane 5 stores the fp16 h box and publishes the tile.
    const int q = warp & 3;
    const uint32_t sMine = sE + (uint32_t)q * QBYTES;
    const int ucol = j * U;
    auto tile_rows = [&](int k, int i) { return s_off[k] + (p + i * g.P) * RP_ROWS + 32 * q; };
    auto load_z = [&](int k, int i, int set) {                          // x-projection boxes of tile (k, i) of this quarter: lanes 0-3
      const uint32_t bar = efull0 + 8 * (q * 2 + set);
      const int r0 = tile_rows(k, i);
      if (lane == 0) mbar_expect_tx(bar, 4 * ZBOX);
      if (lane < 4) {
        if constexpr (SWZ) tma_load_3d(sMine + set * 5 * ZBOX + lane * ZBOX, &maps.z[d], ucol, lane, r0, bar);
        else tma_load_2d(sMine + set * 5 * ZBOX + lane * ZBOX, &maps.z[d], lane * H + ucol, r0, bar);
      }
    };
    // (k2, i2) runs two tiles ahead of the tile being stored: the boxes it names are requested into the set that has just left
    int k2 = 0, i2 = 0;
    auto ahead_valid = [&]() { return k2 < Tmax; };
    auto ahead_next = [&]() {
      i2++;
      if (i2 >= RF_MAXTPC || (p + i2 * g.P) * RP_ROWS >= s_n[k2]) {
        i2 = 0; k2++;
        if (k2 < Tmax && p * RP_ROWS >= s_n[k2]) k2 = Tmax;             // running rows only shrink: nothing left for this part
      }
    };
    if (p * RP_ROWS >= s_n[0]) k2 = Tmax;
    if (ahead_valid()) { load_z(k2, i2, 0); ahead_next(); }             // prime both sets
    if (ahead_valid()) { load_z(k2, i2, 1); ahead_next(); }
    uint32_t n_tile = 0;
    Tracer tr; tr.init(g.trace, g.trace_cta, 2);
    if (q != 2 || lane != 5) tr.p = nullptr;                              // the publishing lane of the quarter of cell warp 0 (warp 2)
    for (int k = 0; k < Tmax; k++) {
      const int n_k = s_n[k], n_k1 = s_n[k + 1];
#pragma unroll 1
      for (int i = 0; i < RF_MAXTPC; i++) {
        const int t = p + i * g.P;
        if (t * RP_ROWS >= n_k) break;
        const int set = n_tile & 1;
        const bool has_next = t * RP_ROWS < n_k1;
        mbar_wait(sready0 + 8 * (q * 2 + set), (n_tile >> 1) & 1);
        tr.ev(0, k, t);
        const int r0 = tile_rows(k, i);
        if (lane == 5) {
          if (has_next) tma_store_2d(&maps.hp16[d], sMine + 2 * 5 * ZBOX + set * HBOX, j * UP, s_off[k + 1] + (p + i * g.P) * RP_ROWS + 32 * q);
          bulk_commit();
          bulk_wait<0>();                                              // the fp16 rows are complete in global memory
          tr.ev(1, k, t);
          flag_release_add(flags + t);                                  // what the next step of the other slices waits for
        } else if (g.training && lane < 5) {
          if (lane < 4) {
            if constexpr (SWZ) tma_store_3d(&maps.z[d], sMine + set * 5 * ZBOX + lane * ZBOX, ucol, lane, r0);
            else tma_store_2d(&maps.z[d], sMine + set * 5 * ZBOX + lane * ZBOX, lane * H + ucol, r0);
          }
          else tma_store_2d(&maps.cc[d], sMine + set * 5 * ZBOX + 4 * ZBOX, ucol, r0);
          bulk_commit();
          bulk_wait_read<0>();                                         // the box has left shared memory
        }
        __syncwarp();                                                   // every box of this set has been read: it may be refilled
        tr.ev(2, k, t);
        if (ahead_valid()) { load_z(k2, i2, set); ahead_next(); }
        tr.ev(3, k, t);
        n_tile++;
      }
    }
    if (lane < 5) bulk_wait<0>();                                        // the last gate boxes have left shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace icl
