// Fused BPTT step of the BiLSTM (sm_100a), "K3": ONE launch per time step runs, for both directions,
//
//     dh_rec = dZ_{k+1} . W_hh^T                      (tcgen05.mma kind::tf32, accumulator in TMEM)
//     dZ_k   = cell_backward(gates_k, c_k, c_{k-1}, dHout_k + dh_rec, dc)      (fused epilogue, in place over the gates)
//
// replacing the per-step pair {k_lstm_cell_bwd, split-K k_gemm_tcgen05 + red.global + memset} of the first version.
//
// Grid = (Nt * CS, Mt, 2 directions), thread-block clusters of CS CTAs along x.  A cluster owns one 128-row x 112-unit tile
// of dh_rec; its CTAs split the contraction (K = 4H) into CS ranges of 32-wide k-blocks (TMA -> 3-stage ring -> UMMA
// M=128, N=112).  The partial accumulators are then reduced through DISTRIBUTED SHARED MEMORY instead of red.global: every
// CTA parks its partial tile in its own (by then idle) operand ring, and after a cluster barrier the rows of the tile are
// dealt to the CTAs (128/CS rows each): a CTA pulls the CS partials of its rows with coalesced ld.shared::cluster, sums them
// and runs the LSTM cell backward on them, row-wise so that every global access is a coalesced run along the hidden units.
// (A first version pushed the partials with st.shared::cluster, one row per lane: 4.4 us per step; the pull is row-coalesced.)
// Deterministic (no atomics), no dh_rec round trip through HBM, pad rows of dZ are written as zeros (the time-batched
// weight-gradient GEMM contracts over them).
//
// Programmatic dependent launch: the prologue (barriers, TMEM allocation) overlaps the tail of step k+1's launch, and every
// launch bulk-prefetches into L2 the gates / cell states / dH rows that the NEXT launch's epilogue will read (they were
// produced by the forward pass and have long left L2), so the HBM reads of the cell backward overlap the contraction.
#pragma once
#include "gemm_tcgen05.cuh"
#include "lstm_persistent.cuh"

namespace icl {

constexpr int BP_BN = 112, BP_STAGES = 3, BP_THREADS = 320;          // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int BP_A_BYTES = 128 * 128, BP_B_BYTES = BP_BN * 128, BP_STAGE_BYTES = BP_A_BYTES + BP_B_BYTES;
constexpr int BP_SLOT_LD = 116;                                        // floats per exchanged row (conflict-free float4 rows)
constexpr int BP_SMEM = 1024 + BP_STAGES * BP_STAGE_BYTES + 128;
static_assert(128 * BP_SLOT_LD * 4 <= BP_STAGES * BP_STAGE_BYTES, "exchange slots reuse the operand ring");

struct BpttMaps { CUtensorMap za[2], wb[2]; };     // A: Z[d] box {32 k, 128 rows}; B: W_hh [H, 4H] box {32 k, 112 units}
struct BpttArgs {
  float* Z[2]; const float* Cc[2]; const float* dHout[2]; float* dcc[2];
  int H, k;
  int o_k, o_kp1, o_km1;      // first row of step k / k+1 / k-1 blocks
  int n_k, n_kp1;             // running rows of step k and of step k+1 (0 when k is the last step: no recurrent term)
  int round_ops, cs;          // cs = cluster size (1, 2 or 4)
  int n_km1, o_km2;           // rows of step k-1 and first row of step k-2 (L2 prefetch for the next launch)
  int dbg;                    // bring-up timing experiments: 1 = skip the final phase, 2 = skip the contraction, 4 = skip the exchange
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// bulk L2 prefetch of [p, p + bytes): bytes a multiple of 16, p 16-byte aligned
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// grid-strided bulk prefetch of a contiguous region in 8 KB pieces, one piece per calling thread `who` of `nwho`
__device__ __forceinline__ void prefetch_region(const float* p, long bytes, long who, long nwho) {
  constexpr long CH = 8192;
  for (long o = who * CH; o < bytes; o += nwho * CH) prefetch_l2_bulk(reinterpret_cast<const char*>(p) + o, (uint32_t)min(CH, bytes - o));
}

template <int CS>
__global__ void __launch_bounds__(BP_THREADS, 2) k_bptt_step(const __grid_constant__ BpttMaps maps, const BpttArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + BP_STAGES * BP_STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * BP_STAGES, tmem_full = bars + 16 * BP_STAGES, tmem_slot = tmem_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int cs = CS;
  const int H = g.H;
  const int rank = (int)cluster_ctarank();
  const int nt = blockIdx.x / cs, m0 = blockIdx.y * 128, d = blockIdx.z;
  const int total_kb = (4 * H + 31) / 32, kb_per = (total_kb + cs - 1) / cs;
  const int kb0 = rank * kb_per, num_kb = (g.dbg & 2) ? 0 : max(0, min(kb_per, total_kb - kb0));
  const bool has_mma = m0 < g.n_kp1;                 // uniform over the cluster: some row of the tile has a recurrent term
  constexpr int rows_own = 128 / cs;                     // rows [rank*rows_own, +rows_own) of the tile are finished by this CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < BP_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.za[d]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wb[d]) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2 && lane == 0 && g.k > 0 && !(g.dbg & 8)) {
    // next launch's epilogue operands -> L2 (block k-1 of the gates and dH, block k-2 of the cell states; block k-1 of the cell
    // states was fetched as this launch's c_{k-1}); spread over the CTAs of this direction
    const long who = (long)blockIdx.y * gridDim.x + blockIdx.x, nwho = (long)gridDim.x * gridDim.y;
    prefetch_region(g.Z[d] + (long)g.o_km1 * 4 * H, (long)g.n_km1 * 4 * H * 4, who, nwho);
    prefetch_region(g.dHout[d] + (long)g.o_km1 * H, (long)g.n_km1 * H * 4, who, nwho);
    if (g.k > 1) prefetch_region(g.Cc[d] + (long)g.o_km2 * H, (long)g.n_km1 * H * 4, who, nwho);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  asm volatile("griddepcontrol.wait;" ::: "memory");               // step k+1's launch has completed: dZ_{k+1} and dc are visible
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (has_mma) {
    if (warp == 0) {
      if (lane == 0) {                                                 // ---- TMA producer
        for (int kb = 0; kb < num_kb; kb++) {
          const int s = kb % BP_STAGES;
          mbar_wait(empty0 + 8 * s, ((kb / BP_STAGES) & 1) ^ 1);
          const uint32_t fb = full0 + 8 * s;
          mbar_expect_tx(fb, BP_STAGE_BYTES);
          tma_load_2d(base + s * BP_STAGE_BYTES, &maps.za[d], (kb0 + kb) * 32, g.o_kp1 + m0, fb);
          tma_load_2d(base + s * BP_STAGE_BYTES + BP_A_BYTES, &maps.wb[d], (kb0 + kb) * 32, nt * BP_BN, fb);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {                                                 // ---- MMA issuer
        constexpr uint32_t idesc = make_idesc(2, false, false, 128, BP_BN);
        for (int kb = 0; kb < num_kb; kb++) {
          const int s = kb % BP_STAGES;
          mbar_wait(full0 + 8 * s, (kb / BP_STAGES) & 1);
          tc_fence_after();
          const uint32_t a = base + s * BP_STAGE_BYTES, b = a + BP_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; kk++)
            tc_mma_tf32(tmem_acc, make_smem_desc(a + kk * 32, 16, 1024), make_smem_desc(b + kk * 32, 16, 1024), idesc, (kb | kk) != 0);
          tc_commit(empty0 + 8 * s);
        }
        tc_commit(tmem_full);
      }
      __syncwarp();
    } else if (!(g.dbg & 4)) {
      // park my partial tile [128 rows][BP_SLOT_LD] in my own operand ring (idle once my MMAs have retired)
      const int q = warp & 3, half = (warp - 2) >> 2;
      const uint32_t dst = base + (uint32_t)((q * 32 + lane) * BP_SLOT_LD) * 4;
      if (num_kb > 0) { mbar_wait(tmem_full, 0); tc_fence_after(); }
#pragma unroll 1
      for (int c = half * 32; c < BP_BN; c += 64) {
        uint32_t r[32];
        if (num_kb > 0) tc_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + c, r);
        else {
#pragma unroll
          for (int j = 0; j < 32; j++) r[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (c + j < BP_BN)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)(c + j) * 4), "r"(r[j]), "r"(r[j + 1]), "r"(r[j + 2]),
                         "r"(r[j + 3]) : "memory");
      }
      tc_fence_before();
    }
    cluster_sync_all();                                                // every partial of the tile is parked (release / acquire)
  }

  // ---- final phase, all threads: sum the CS partials of my rows (DSMEM pull), LSTM cell backward, write dZ_k and dc.
  // items = (row of my row range, group of 4 hidden units); a thread has at most BP_IPT of them
  constexpr int CG = BP_BN / 4, n_items = rows_own * CG, IPT = (n_items + BP_THREADS - 1) / BP_THREADS;
  float* const Zb = g.Z[d] + (long)g.o_k * 4 * H;
  const float* const Ck = g.Cc[d] + (long)g.o_k * H;
  const float* const Cp = g.k > 0 ? g.Cc[d] + (long)g.o_km1 * H : nullptr;
  const float* const dHk = g.dHout[d] + (long)g.o_k * H;
  float* const dcc = g.dcc[d];
  const bool pull = has_mma && !(g.dbg & 4), fin = !(g.dbg & 1);
#pragma unroll
  for (int b0 = 0; b0 < IPT; b0 += 2) {
    float4 g4[2][4], c4[2], cp4[2], dh4[2], dc4[2];
    int kind[2];                                                       // 0: nothing, 1: pad row (dZ = 0), 2: cell backward
    long ro[2]; int uu[2];
    // (1) issue the global loads of both items (L2 hits: prefetched by the previous launch)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int it = threadIdx.x + (b0 + b) * BP_THREADS;
      const int row = it / CG, u = nt * BP_BN + (it % CG) * 4, grow = m0 + rank * rows_own + row;
      kind[b] = (b0 + b >= IPT || it >= n_items || u >= H) ? 0 : grow >= g.n_k ? 1 : 2;
      ro[b] = grow; uu[b] = u;
      if (kind[b] == 2 && fin) {
        const float* z = Zb + (long)grow * 4 * H + u;
#pragma unroll
        for (int a = 0; a < 4; a++) g4[b][a] = *reinterpret_cast<const float4*>(z + a * H);
        c4[b] = *reinterpret_cast<const float4*>(Ck + (long)grow * H + u);
        cp4[b] = Cp ? *reinterpret_cast<const float4*>(Cp + (long)grow * H + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        dh4[b] = *reinterpret_cast<const float4*>(dHk + (long)grow * H + u);
        dc4[b] = *reinterpret_cast<const float4*>(dcc + (long)grow * H + u);
      }
    }
    // (2) pull and sum the CS partials of both items while those loads are in flight
    float4 dhr[2];
#pragma unroll
    for (int b = 0; b < 2; b++) {
      dhr[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int it = threadIdx.x + (b0 + b) * BP_THREADS;
      if (pull && b0 + b < IPT && it < n_items) {
        const uint32_t la = base + (uint32_t)(((rank * rows_own + it / CG) * BP_SLOT_LD) + (it % CG) * 4) * 4;
        float4 p[CS];
#pragma unroll
        for (int s = 0; s < CS; s++) p[s] = ld_cluster_f4(mapa_shared(la, (uint32_t)s));
#pragma unroll
        for (int s = 0; s < CS; s++) { dhr[b].x += p[s].x; dhr[b].y += p[s].y; dhr[b].z += p[s].z; dhr[b].w += p[s].w; }
      }
    }
    if (b0 + 2 >= IPT && has_mma) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // my pulls are done
    // (3) cell backward + stores
#pragma unroll
    for (int b = 0; b < 2; b++) {
      if (kind[b] == 0 || !fin) continue;
      float* z = Zb + ro[b] * 4 * H + uu[b];
      if (kind[b] == 1) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(z) = z4; *reinterpret_cast<float4*>(z + H) = z4;
        *reinterpret_cast<float4*>(z + 2 * H) = z4; *reinterpret_cast<float4*>(z + 3 * H) = z4;
        continue;
      }
      if (ro[b] < g.n_kp1) { dh4[b].x += dhr[b].x; dh4[b].y += dhr[b].y; dh4[b].z += dhr[b].z; dh4[b].w += dhr[b].w; }
      const float *si = &g4[b][0].x, *tj = &g4[b][1].x, *sf = &g4[b][2].x, *so = &g4[b][3].x, *c = &c4[b].x, *cp = &cp4[b].x,
                  *dh = &dh4[b].x, *dc = &dc4[b].x;
      float di[4], dj[4], df[4], dgo[4], dcp[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const CellGrad cg_ = lstm_cell_bwd(si[j], tj[j], sf[j], so[j], c[j], cp[j], dh[j], dc[j]);
        di[j] = maybe_round(cg_.di, g.round_ops); dj[j] = maybe_round(cg_.dj, g.round_ops); df[j] = maybe_round(cg_.df, g.round_ops);
        dgo[j] = maybe_round(cg_.dg_o, g.round_ops); dcp[j] = cg_.dc_prev;
      }
      *reinterpret_cast<float4*>(z) = make_float4(di[0], di[1], di[2], di[3]);
      *reinterpret_cast<float4*>(z + H) = make_float4(dj[0], dj[1], dj[2], dj[3]);
      *reinterpret_cast<float4*>(z + 2 * H) = make_float4(df[0], df[1], df[2], df[3]);
      *reinterpret_cast<float4*>(z + 3 * H) = make_float4(dgo[0], dgo[1], dgo[2], dgo[3]);
      *reinterpret_cast<float4*>(dcc + ro[b] * H + uu[b]) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_acc) : "memory");
  }
  if (has_mma) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // nobody still reads my parked partial
}


// =====================================================================================================================
// k_bptt_cluster: the WHOLE backward recurrence in one launch, one thread-block cluster per (direction, 128-row tile).
//
// dZ_k of a row tile depends only on dZ_{k+1} of the SAME rows, so a cluster that owns a row tile can walk all of its time
// steps with nothing but cluster barriers -- no grid barrier, no flags in global memory, no launch per step.  Per step:
//   K-loop   : the cluster's CS (4, or 8 when the batch has few row tiles) CTAs split the contraction (K = 4H) into k-block ranges; per k-block one TMA box of dZ_{k+1}
//              [128 x 32] and three boxes of W_hh [112 units x 32] feed UMMAs into THREE TMEM accumulators (all H <= 336 units
//              of the tile: dZ is read once, not once per 112-unit column tile as in k_bptt_step);
//   park     : TMEM -> the CTA's own (idle) operand ring as a [128 x 340] fp32 tile;             cluster barrier S1
//   pull     : every CTA finishes 128/CS rows: it sums the CS partials of its rows with coalesced ld.shared::cluster
//              (barrier S2 = "my pulls are done", waited for before the ring is refilled);
//   cell bwd : gates/c/dH rows (L2-prefetched one step ahead) -> dZ_k in place, dc carry;  fence.proxy.async + barrier S3
//              make dZ_k visible to the TMA loads of step k-1 issued by the other CTAs of the cluster.
// Clusters are independent: the launch needs no co-residency (a batch with more than 18 row tiles per direction simply
// runs in waves).
// =====================================================================================================================
constexpr int BC_NACC = 3, BC_STAGES = 3, BC_THREADS = 320;     // cluster size CS = 4 or 8 (template parameter)
constexpr int BC_STAGE_BYTES = BP_A_BYTES + BC_NACC * BP_B_BYTES;          // 16 KB + 3 x 14 KB
constexpr int BC_PARK_LD = BC_NACC * BP_BN + 4;                            // 340 floats per parked row
constexpr int BC_SMEM = 1024 + BC_STAGES * BC_STAGE_BYTES + 128;
static_assert(128 * BC_PARK_LD * 4 <= BC_STAGES * BC_STAGE_BYTES, "the parked tile reuses the operand ring");

struct BpttClusterArgs {
  float* Z[2]; const float* Cc[2]; const float* dHout[2]; float* dcc[2];
  const int* off; const int* nact;       // [Tmax+1] step offsets / running rows
  int H, Tmax, round_ops;
  int row0;                             // first row of this launch (it covers gridDim.y / 2 tiles of tm rows from there)
  int tm;                               // rows per tile: 128 or 64
  long long* trace; int trace_cta;      // optional bring-up trace (see Tracer): CTA index = y * gridDim.x + x
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int BC_CS>
__global__ void __launch_bounds__(BC_THREADS, 1) k_bptt_cluster(const __grid_constant__ BpttMaps maps, const BpttClusterArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_off[RP_MAXT + 2], s_n[RP_MAXT + 2];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + BC_STAGES * BC_STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * BC_STAGES, tmem_full = bars + 16 * BC_STAGES, tmem_slot = tmem_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = g.H, Tmax = g.Tmax, HQ = H >> 2;
  const int rank = (int)cluster_ctarank();
  // grid.y = 2 * tiles with the direction in the low bit: the long chains (low tiles) of BOTH directions are scheduled first
  // g.tm = rows of a tile: 128, or 64 for small batches (twice as many independent chains, half the park / pull / cell work per
  // step; the A box still carries 128 rows and the upper 64 accumulator rows are ignored -- the tensor pipe is idle anyway)
  const int TM = g.tm;
  const int m0 = g.row0 + (blockIdx.y >> 1) * TM, d = blockIdx.y & 1;
  const int total_kb = (4 * H + 31) / 32, kb_per = (total_kb + BC_CS - 1) / BC_CS;
  const int kb0 = rank * kb_per, num_kb = max(0, min(kb_per, total_kb - kb0));
  const int nacc = (H + BP_BN - 1) / BP_BN;                                   // accumulators in use (<= BC_NACC)
  const int ROWS_OWN = TM / BC_CS;

  for (int i = threadIdx.x; i <= Tmax; i += blockDim.x) { s_off[i] = g.off[i]; s_n[i] = g.nact[i]; }
  if (threadIdx.x == 0) {
    for (int s = 0; s < BC_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.za[d]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wb[d]) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));
  asm volatile("griddepcontrol.wait;" ::: "memory");

  float* const Zd = g.Z[d];
  const float* const Cd = g.Cc[d];
  const float* const dHd = g.dHout[d];
  float* const dcc = g.dcc[d];
  const int n_items = ROWS_OWN * HQ;                    // (row of my 32 rows, group of 4 hidden units)
  uint32_t it = 0, nfull = 0;                           // running k-block / accumulator-phase counters (mbarrier parities)

  Tracer tr;
  {
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    tr.p = (g.trace && cta == g.trace_cta && lane == 0 && warp <= 2) ? g.trace + (long)warp * RP_TRACE_EV * 4 : nullptr;
    tr.n = 0;
  }
  // first step in which this tile has running rows (n_k grows as k decreases)
  int k_first = -1;
  // (step blocks are padded to 128 rows and the pad rows of dZ must be written as zeros: with 64-row tiles the upper half of a
  // block's last 128 rows may hold no running row at all and still has to be visited)
  for (int k = Tmax - 1; k >= 0; k--) if (m0 < ((s_n[k] + 127) & ~127)) { k_first = k; break; }

  // L2 prefetch of the cell-backward operands of step k for my rows.  The rows of a CTA are contiguous in the step-major layout, so
  // each operand is ONE region: four bulk prefetches per step (two halves of the gate rows, dH, c) from four lanes of warp 2.  (One
  // prefetch per row -- 96 per step -- kept the CTA's TMA unit busy for 3.7 us in front of the K-loop's operand loads,
  // profiles/r1g_bptt_cluster_trace.txt: "S23 -> step start".)
  auto prefetch_step = [&](int k) {
    if (warp != 2 || k < 0 || lane > 3) return;
    const int r0 = m0 + rank * ROWS_OWN, nr = min(ROWS_OWN, s_n[k] - r0);
    if (nr <= 0) return;
    const long row0 = (long)s_off[k] + r0;
    if (lane < 2) {
      const int h0 = lane ? nr / 2 : 0, h1 = lane ? nr : nr / 2;
      if (h1 > h0) prefetch_l2_bulk(Zd + (row0 + h0) * 4 * H, (uint32_t)((h1 - h0) * 4 * H * 4));
    } else if (lane == 2) {
      prefetch_l2_bulk(dHd + row0 * H, (uint32_t)(nr * H * 4));
    } else {
      if (k > 0) prefetch_l2_bulk(Cd + ((long)s_off[k - 1] + r0) * H, (uint32_t)(nr * H * 4));
      if (k == k_first) prefetch_l2_bulk(Cd + row0 * H, (uint32_t)(nr * H * 4));
    }
  };
  prefetch_step(k_first);

  for (int k = k_first; k >= 0; k--) {
    const int n_k = s_n[k], n_kp1 = k + 1 < Tmax ? s_n[k + 1] : 0;
    const long o_k = s_off[k];
    const bool has_mma = m0 < n_kp1;                  // uniform over the cluster
    prefetch_step(k - 1);
    tr.ev(0, k, 0);
    if (has_mma) {
      if (warp == 0) {
        // ---- TMA producer: the 1 + nacc boxes of a k-block are requested by 1 + nacc lanes, lane 0 arms the stage's barrier for all of
        // them.  (In tools/probe_tma.cu several lanes issue TMA loads concurrently; in this kernel -- as in the GEMM and in
        // k_rec_fwd16 -- it measured no different from one issuing lane: the step is a chain of round trips, not issue-bound.)
        if (lane <= nacc) {
          for (int kb = 0; kb < num_kb; kb++, it++) {
            const int s = it % BC_STAGES;
            mbar_wait(empty0 + 8 * s, ((it / BC_STAGES) & 1) ^ 1);
            const uint32_t fb = full0 + 8 * s, st = base + s * BC_STAGE_BYTES;
            if (lane == 0) {
              mbar_expect_tx(fb, BP_A_BYTES + nacc * BP_B_BYTES);
              tma_load_2d(st, &maps.za[d], (kb0 + kb) * 32, s_off[k + 1] + m0, fb);
            } else {
              tma_load_2d(st + BP_A_BYTES + (lane - 1) * BP_B_BYTES, &maps.wb[d], (kb0 + kb) * 32, (lane - 1) * BP_BN, fb);
            }
          }
        }
        __syncwarp();
      } else if (warp == 1) {
        if (lane == 0) {                                               // ---- MMA issuer
          constexpr uint32_t idesc = make_idesc(2, false, false, 128, BP_BN);
          for (int kb = 0; kb < num_kb; kb++, it++) {
            const int s = it % BC_STAGES;
            mbar_wait(full0 + 8 * s, (it / BC_STAGES) & 1);
            tc_fence_after();
            const uint32_t a = base + s * BC_STAGE_BYTES;
            for (int ac = 0; ac < nacc; ac++) {
              const uint32_t b = a + BP_A_BYTES + ac * BP_B_BYTES;
#pragma unroll
              for (int kk = 0; kk < 4; kk++)
                tc_mma_tf32(tmem_acc + ac * 128, make_smem_desc(a + kk * 32, 16, 1024), make_smem_desc(b + kk * 32, 16, 1024), idesc,
                            (kb | kk) != 0);
            }
            tc_commit(empty0 + 8 * s);
          }
          if (num_kb > 0) tc_commit(tmem_full);
        }
        __syncwarp();
      } else {
        // park my partial [128 rows][BC_PARK_LD] in my own operand ring (idle once my MMAs have retired)
        const int q = warp & 3, half = (warp - 2) >> 2;
        const uint32_t dst = base + (uint32_t)((q * 32 + lane) * BC_PARK_LD) * 4;
        if (num_kb > 0) { mbar_wait(tmem_full, nfull & 1); tc_fence_after(); }
        tr.ev(1, k, 0);
        for (int ac = 0; ac < nacc && q * 32 < TM; ac++) {           // rows >= TM of the accumulator belong to nobody
#pragma unroll 1
          for (int c = half * 32; c < BP_BN; c += 64) {
            uint32_t r[32];
            if (num_kb > 0) tc_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + ac * 128 + c, r);
            else {
#pragma unroll
              for (int j = 0; j < 32; j++) r[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (c + j < BP_BN)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)(ac * BP_BN + c + j) * 4), "r"(r[j]),
                             "r"(r[j + 1]), "r"(r[j + 2]), "r"(r[j + 3]) : "memory");
          }
        }
        tc_fence_before();
        tr.ev(2, k, 0);
      }
      nfull++;
      cluster_arrive(); cluster_wait();                                // S1: every partial of the tile is parked
      tr.ev(3, k, 0);
    }

    // ---- pull + cell backward, all threads; items in batches of 2
#pragma unroll 1
    for (int i0 = threadIdx.x; i0 < n_items; i0 += 2 * BC_THREADS) {
      float4 g4[2][4], c4[2], cp4[2], dh4[2], dc4[2], dhr[2];
      int kind[2];                                                     // 0: nothing, 1: pad row (dZ = 0), 2: cell backward
      long ro[2]; int uu[2];
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int item = i0 + b * BC_THREADS;
        const int row = item / HQ, u = (item % HQ) * 4, grow = m0 + rank * ROWS_OWN + row;
        kind[b] = item >= n_items ? 0 : grow >= n_k ? 1 : 2;
        ro[b] = grow; uu[b] = u;
        if (kind[b] == 2) {
          const float* z = Zd + (o_k + grow) * 4 * H + u;
#pragma unroll
          for (int a = 0; a < 4; a++) g4[b][a] = *reinterpret_cast<const float4*>(z + a * H);
          c4[b] = *reinterpret_cast<const float4*>(Cd + (o_k + grow) * H + u);
          cp4[b] = k > 0 ? *reinterpret_cast<const float4*>(Cd + ((long)s_off[k - 1] + grow) * H + u) : make_float4(0.f, 0.f, 0.f, 0.f);
          dh4[b] = *reinterpret_cast<const float4*>(dHd + (o_k + grow) * H + u);
          dc4[b] = *reinterpret_cast<const float4*>(dcc + (long)grow * H + u);
        }
      }
#pragma unroll
      for (int b = 0; b < 2; b++) {
        dhr[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int item = i0 + b * BC_THREADS;
        if (has_mma && item < n_items) {
          const uint32_t la = base + (uint32_t)(((rank * ROWS_OWN + item / HQ) * BC_PARK_LD) + (item % HQ) * 4) * 4;
          float4 p[BC_CS];
#pragma unroll
          for (int s = 0; s < BC_CS; s++) p[s] = ld_cluster_f4(mapa_shared(la, (uint32_t)s));
#pragma unroll
          for (int s = 0; s < BC_CS; s++) { dhr[b].x += p[s].x; dhr[b].y += p[s].y; dhr[b].z += p[s].z; dhr[b].w += p[s].w; }
        }
      }
#pragma unroll
      for (int b = 0; b < 2; b++) {
        if (kind[b] == 0) continue;
        float* z = Zd + (o_k + ro[b]) * 4 * H + uu[b];
        if (kind[b] == 1) {
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(z) = z4; *reinterpret_cast<float4*>(z + H) = z4;
          *reinterpret_cast<float4*>(z + 2 * H) = z4; *reinterpret_cast<float4*>(z + 3 * H) = z4;
          continue;
        }
        if (ro[b] < n_kp1) { dh4[b].x += dhr[b].x; dh4[b].y += dhr[b].y; dh4[b].z += dhr[b].z; dh4[b].w += dhr[b].w; }
        const float *si = &g4[b][0].x, *tj = &g4[b][1].x, *sf = &g4[b][2].x, *so = &g4[b][3].x, *c = &c4[b].x, *cp = &cp4[b].x,
                    *dh = &dh4[b].x, *dc = &dc4[b].x;
        float di[4], dj[4], df[4], dgo[4], dcp[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const CellGrad cg_ = lstm_cell_bwd(si[j], tj[j], sf[j], so[j], c[j], cp[j], dh[j], dc[j]);
          di[j] = maybe_round(cg_.di, g.round_ops); dj[j] = maybe_round(cg_.dj, g.round_ops); df[j] = maybe_round(cg_.df, g.round_ops);
          dgo[j] = maybe_round(cg_.dg_o, g.round_ops); dcp[j] = cg_.dc_prev;
        }
        *reinterpret_cast<float4*>(z) = make_float4(di[0], di[1], di[2], di[3]);
        *reinterpret_cast<float4*>(z + H) = make_float4(dj[0], dj[1], dj[2], dj[3]);
        *reinterpret_cast<float4*>(z + 2 * H) = make_float4(df[0], df[1], df[2], df[3]);
        *reinterpret_cast<float4*>(z + 3 * H) = make_float4(dgo[0], dgo[1], dgo[2], dgo[3]);
        *reinterpret_cast<float4*>(dcc + ro[b] * H + uu[b]) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
      }
    }
    tr.ev(4, k, 0);
    if (k > 0) {
      // S2/S3 merged: my pulls are done (the ring may be refilled) and my dZ_k rows are visible to the async proxy of every
      // CTA of the cluster (their TMA loads of step k-1)
      fence_async_all();
      cluster_arrive(); cluster_wait();
      tr.ev(5, k, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_acc) : "memory");
  }
  cluster_arrive(); cluster_wait();                                    // nobody still reads my parked tile
}

}  // namespace icl
