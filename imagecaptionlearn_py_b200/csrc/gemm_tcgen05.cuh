// TF32 GEMM on the 5th-gen tensor cores (sm_100a): TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory
// -> tcgen05.mma.kind::tf32 with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
//   C[M,N] = epilogue( A * B ),  fp32 in HBM on both sides (kind::tf32 reads the fp32 containers directly).
//   A: K-major ([M,K], k contiguous) or MN-major ([K,M], m contiguous); B likewise ([N,K] or [K,N]).
// Both majors are needed without transposes: forward layers are (K, MN), input-gradient GEMMs are (K, K) and the
// time-batched weight-gradient GEMMs (contraction over tokens) are (MN, MN).
//
// One 128 x BN output tile per CTA (BN = 128 or 256), BK = 32 fp32 (one 128-byte swizzle row), STAGES-deep mbarrier
// pipeline.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-9 = epilogue.
// Configurations (picked per call by tcgen05_gemm_launch):
//   <128,3>  97 KB smem, two CTAs per SM: many-tile GEMMs, one CTA's epilogue overlaps the other's main loop
//   <128,6> 193 KB smem, one CTA per SM: few tiles with a long K loop -- latency-bound, so a deeper ring
//   <64,4>   97 KB / <64,8> 193 KB: narrow tiles for few-tile problems (the head layers): more CTAs share the streaming
//   <256,2>  97 KB / <256,4> 193 KB: wide tiles halve the L2->SM bytes per MMA cycle (TF32 tiles of 128x128 need ~128 B/clk
//            per SM, more than one SM can ingest); used when N >= 512.
// Epilogue: TMEM -> registers (one accumulator row per thread) -> a per-warp 32x33 staging tile in the (by then idle)
// pipeline buffers -> row-wise, so every global access of the fused epilogue (bias, aux, C) is a coalesced 128-byte line.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <map>
#include <tuple>

#include "icl_kernels.cuh"

namespace icl {

constexpr int TG_BM = 128, TG_BK = 32;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 4;                          // 16 KB per A stage
constexpr int TG_EPW = 2;             // epilogue warps per TMEM lane quarter (they take alternating 32-column chunks)
constexpr int TG_THREADS = 64 + 128 * TG_EPW;   // warp 0 TMA, warp 1 MMA, then the epilogue warps
constexpr int tg_smem(int BN, int STAGES) { return STAGES * (TG_A_BYTES + BN * TG_BK * 4) + 1024 /*align*/ + 256 /*barriers*/; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 24); it++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives TMEM lane (base+i), columns [c, c+32)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t r[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (tcgen05 "SmemDescriptor"): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
//  K-major tile  [rows][32 fp32]: 8-row groups are 1024 B apart (SBO); LBO unused for swizzled K-major (1).
//  MN-major tile [mn/32][k][32 fp32]: 128-byte MN blocks are LBO apart, 8-k groups 1024 B apart (SBO).
//  32-bit MN-major operands have ONE legal swizzled layout: SWIZZLE_128B_BASE32B (=1), i.e. 32-byte chunks XOR-ed with
//  (k-row % 4); its TMA twin is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Atom = 4 k-rows x 128 B, so an 8-deep MMA spans two
//  atoms SBO = 512 B apart; 128-byte MN blocks are LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
// MN-major 32-bit operands: SWIZZLE_128B_BASE32B descriptors (layout 1), SBO = 512 B between 4-deep k atoms, LBO = 4 KB
// between 128-byte MN blocks ([mn/32][k][32 fp32] tiles), 1 KB per 8-deep MMA step; TMA twin: SWIZZLE_128B_ATOM_32B.
constexpr uint32_t MN_LAYOUT = 1, MN_SBO = 512, MN_LBO = 32 * TG_BK * 4, MN_KSTEP = 1024;
// Instruction descriptor: c_format F32 (1) [4,6); a/b format [7,10)/[10,13) (F16=0, BF16=1, TF32=2);
// a_major [15], b_major [16] (1 = MN-major); N>>3 [17,23); M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int fmt, bool a_mn, bool b_mn, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// F16: both operands are fp16, K-major, k-blocks of 64 halves (the same 128-byte rows / descriptors / 32-byte MMA steps as TF32):
// tcgen05.mma kind::f16 with fp32 accumulation -- for contractions whose operands fit fp16's 10-bit mantissa + range exactly as
// well as TF32's (the input projection: embeddings x W_ih), at half the operand bytes and twice the MMA rate.  g.K counts halves.
template <bool A_MN, bool B_MN, int BN, int STAGES, bool F16 = false>
__global__ void __launch_bounds__(TG_THREADS) k_gemm_tcgen05(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  constexpr int B_BYTES = BN * TG_BK * 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t sA = base, sB = base + STAGES * TG_A_BYTES;
  const uint32_t bars = sB + STAGES * B_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tmem_full = bars + 16 * STAGES, tmem_slot = tmem_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TG_BM, n0 = blockIdx.x * BN;
  // split-K (gridDim.z > 1): this CTA contracts k-blocks [kb0, kb0+num_kb) and adds its partial tile to C with red.global
  constexpr int KB_ELEMS = F16 ? 2 * TG_BK : TG_BK;                      // elements per 128-byte k-block row
  static_assert(!F16 || (!A_MN && !B_MN), "the fp16 path is K-major only");
  const int total_kb = (g.K + KB_ELEMS - 1) / KB_ELEMS;
  const int kb_per = (total_kb + gridDim.z - 1) / gridDim.z;
  const int kb0 = blockIdx.z * kb_per;
  const int num_kb = max(0, min(kb_per, total_kb - kb0));
  if (num_kb == 0) return;                                             // uniform per CTA, before any barrier / TMEM allocation

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the previous
  // kernel in the stream; its results are visible after the wait.  No-ops when launched without the PDL attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {                                                   // ---- TMA producer
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
        const uint32_t fb = full0 + 8 * s;
        mbar_expect_tx(fb, TG_A_BYTES + B_BYTES);                      // OOB parts of a box are zero-filled and still counted
        const uint32_t a = sA + s * TG_A_BYTES, b = sB + s * B_BYTES;
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < TG_BM / 32; j++) tma_load_2d(a + j * (32 * TG_BK * 4), &tmA, m0 + 32 * j, (kb0 + kb) * TG_BK, fb);
        } else {
          tma_load_2d(a, &tmA, (kb0 + kb) * KB_ELEMS, m0, fb);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 32; j++) tma_load_2d(b + j * (32 * TG_BK * 4), &tmB, n0 + 32 * j, (kb0 + kb) * TG_BK, fb);
        } else {
          tma_load_2d(b, &tmB, (kb0 + kb) * KB_ELEMS, n0, fb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                                   // ---- MMA issuer (one thread)
      constexpr uint32_t idesc = make_idesc(F16 ? 0 : 2, A_MN, B_MN, TG_BM, BN);
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t a = sA + s * TG_A_BYTES, b = sB + s * B_BYTES;
#pragma unroll
        for (int kk = 0; kk < TG_BK / 8; kk++) {                       // UMMA_K = 8 for tf32
          const uint64_t ad = A_MN ? make_smem_desc(a + kk * MN_KSTEP, MN_LBO, MN_SBO, MN_LAYOUT) : make_smem_desc(a + kk * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_smem_desc(b + kk * MN_KSTEP, MN_LBO, MN_SBO, MN_LAYOUT) : make_smem_desc(b + kk * 32, 16, 1024);
          if (F16) tc_mma_f16(tmem_acc, ad, bd, idesc, (kb | kk) != 0);
          else tc_mma_tf32(tmem_acc, ad, bd, idesc, (kb | kk) != 0);
        }
        tc_commit(empty0 + 8 * s);                                     // frees the smem stage when these MMAs retire
      }
      tc_commit(tmem_full);
    }
  } else {                                                             // ---- epilogue: TMEM -> registers -> smem -> HBM
    mbar_wait(tmem_full, 0);                                           // every MMA has retired: the stage buffers are idle
    tc_fence_after();
    const int q = warp & 3;                                            // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int ehalf = (warp - 2) >> 2;                                 // which of the quarter's two warps
    if constexpr (F16) {
      // Output path of the input projection (plain epilogue + bias, dense C): the 128 x BN fp32 tile leaves as 32x32 TMA boxes.
      // Each warp stages a chunk in the 128B-swizzled layout of tmC (16-byte piece j of row r at r*128 + ((j ^ (r & 7)) << 4)),
      // two buffers per warp in the idle operand ring, so the store of chunk i overlaps the TMEM read of chunk i+1.  (The scalar
      // row loop below issues one 128-byte store per row and chunk: 1024 LSU stores per tile.)
      const uint32_t sbuf = base + (uint32_t)(warp - 2) * 8192;
      const float* bias = g.epi.bias;
      int ci = 0;
#pragma unroll 1
      for (int c = ehalf * 32; c < BN; c += 32 * TG_EPW, ci++) {
        if (n0 + c >= g.N) break;
        const uint32_t buf = sbuf + (ci & 1) * 4096;
        if (ci >= 2) {                                                 // the store issued two chunks ago has read this buffer
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        uint32_t r[32];
        tc_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + c, r);
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias && n0 + c + j4 * 4 < g.N) b4 = *reinterpret_cast<const float4*>(bias + n0 + c + j4 * 4);
          const float x0 = __uint_as_float(r[j4 * 4]) + b4.x, x1 = __uint_as_float(r[j4 * 4 + 1]) + b4.y,
                      x2 = __uint_as_float(r[j4 * 4 + 2]) + b4.z, x3 = __uint_as_float(r[j4 * 4 + 3]) + b4.w;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + lane * 128 + ((j4 ^ (lane & 7)) << 4)), "f"(x0), "f"(x1),
                       "f"(x2), "f"(x3) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC), "r"(buf), "r"(n0 + c),
                       "r"(m0 + q * 32) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    } else {
    float* stg = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw))) + (warp - 2) * (32 * 33);
    const int mrow0 = m0 + q * 32;
    // everything the row loop needs lives in registers: kernel parameters sit in the constant bank and a dependent
    // LDCU per use (once per element) made this loop ~10x slower than the MMA main loop
    const Epilogue e = g.epi;
    float* const Cp = g.C;
    const long ldc = g.ldc;
    const int M = g.M, N = g.N;
    const bool splitk = gridDim.z > 1;
    const int rows = min(32, M - mrow0);
    // scalar copies pinned in registers (an empty asm makes them opaque, so they are not re-read from the constant bank)
    int p_mode = e.mode, p_act = e.act, p_round = e.round_out;
    float p_keep = e.drop.keep;
    uint64_t p_seed = e.drop.seed;
    uint32_t p_stream = e.drop.stream;
    long p_gid0 = e.drop.row_gid0, p_ldaux = e.ldaux;
    const float *p_bias = e.bias, *p_aux = e.aux;
    asm volatile("" : "+r"(p_mode), "+r"(p_act), "+r"(p_round), "+f"(p_keep), "+l"(p_seed), "+r"(p_stream), "+l"(p_gid0), "+l"(p_ldaux),
                 "+l"(p_bias), "+l"(p_aux));
    const bool p_drop = p_mode != EPI_PLAIN && p_keep < 1.0f;
    const float p_inv_keep = 1.0f / p_keep;                            // x * (1/keep) instead of an IEEE division per element
    // float4 path: N, ldc (and the aux pitch) multiples of 4 and 16-byte aligned bases
    const bool vec4 = (N & 3) == 0 && (ldc & 3) == 0 && ((uintptr_t)Cp & 15) == 0 && (!e.bias || ((uintptr_t)e.bias & 15) == 0) &&
                      (e.mode != EPI_DACT || ((e.ldaux & 3) == 0 && ((uintptr_t)e.aux & 15) == 0));
#pragma unroll 1
    for (int c = ehalf * 32; c < BN; c += 32 * TG_EPW) {
      if (n0 + c >= N) break;
      uint32_t r[32];
      tc_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + c, r);
#pragma unroll
      for (int j = 0; j < 32; j++) stg[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      const int n = n0 + c + lane;
      if (n < N && rows > 0) {
        float* cp = Cp + (long)mrow0 * ldc + n;
        if (splitk) {                                                  // C pre-zeroed by the host
#pragma unroll 8
          for (int rr = 0; rr < rows; rr++) atomicAdd(cp + (long)rr * ldc, stg[rr * 33 + lane]);
        } else if (e.mode == EPI_PLAIN) {
          const float bv = e.bias ? e.bias[n] : 0.0f;
          if (e.round_out) {
#pragma unroll 8
            for (int rr = 0; rr < rows; rr++) cp[(long)rr * ldc] = tf32_rna(stg[rr * 33 + lane] + bv);
          } else {
#pragma unroll 8
            for (int rr = 0; rr < rows; rr++) cp[(long)rr * ldc] = stg[rr * 33 + lane] + bv;
          }
        } else if (!vec4) {
#pragma unroll 4
          for (int rr = 0; rr < rows; rr++) cp[(long)rr * ldc] = epilogue_apply(e, stg[rr * 33 + lane], mrow0 + rr, n, N);
        }
      }
      if (vec4 && !splitk && e.mode != EPI_PLAIN) {
        // fused activation / dropout epilogues: lane = (row % 4, 4-column group) so one Philox call serves four outputs
        const int cg = lane & 7, rsub = lane >> 3, n4 = n0 + c + cg * 4;
        if (n4 < N) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p_bias) b4 = *reinterpret_cast<const float4*>(p_bias + n4);
          // fully unrolled: eight independent Philox chains per thread hide each other's latency (one warp per scheduler)
#pragma unroll
          for (int rb = 0; rb < 32; rb += 4) {
            const int rr = rb + rsub;
            if (rr < rows) {
              const long m = mrow0 + rr;
              const float* sp = stg + rr * 33 + cg * 4;
              float x[4] = {sp[0] + b4.x, sp[1] + b4.y, sp[2] + b4.z, sp[3] + b4.w};
              float mk[4] = {1.f, 1.f, 1.f, 1.f};
              if (p_drop) drop4(p_seed, p_stream, (uint64_t)((p_gid0 + m) * N + n4) >> 2, p_keep, mk);
              if (p_mode == EPI_BIAS_ACT_DROP) {
#pragma unroll
                for (int j = 0; j < 4; j++) { x[j] = act_fwd(x[j], p_act); if (p_drop) x[j] = x[j] * p_inv_keep * mk[j]; }
              } else {                                                 // EPI_DACT
                const float4 y4 = *reinterpret_cast<const float4*>(p_aux + m * p_ldaux + n4);
                const float y[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                  float a = y[j];
                  if (p_drop) { a = y[j] * p_keep; x[j] = x[j] * p_inv_keep * mk[j]; }
                  x[j] *= act_bwd_from_out(a, p_act);
                }
              }
              if (p_round) {
#pragma unroll
                for (int j = 0; j < 4; j++) x[j] = tf32_rna(x[j]);
              }
              *reinterpret_cast<float4*>(Cp + m * ldc + n4) = make_float4(x[0], x[1], x[2], x[3]);
            }
          }
        }
      }
      __syncwarp();
    }
    }   // !F16
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(BN) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side: tensor maps + launch
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

template <bool A_MN, bool B_MN> static cudaError_t tg_set_attrs() {
  cudaError_t e = cudaFuncSetAttribute(k_gemm_tcgen05<A_MN, B_MN, 128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(128, 3));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tcgen05<A_MN, B_MN, 128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(128, 6));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tcgen05<A_MN, B_MN, 256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(256, 2));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tcgen05<A_MN, B_MN, 256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(256, 4));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tcgen05<A_MN, B_MN, 64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(64, 4));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tcgen05<A_MN, B_MN, 64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(64, 8));
  return e;
}
static int tcgen05_gemm_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
  g_encode = (PFN_encodeTiled)fn;
  cudaError_t e = tg_set_attrs<false, false>();
  if (e == cudaSuccess) e = tg_set_attrs<false, true>();
  if (e == cudaSuccess) e = tg_set_attrs<true, false>();
  if (e == cudaSuccess) e = tg_set_attrs<true, true>();
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tcgen05<false, false, 256, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(256, 2));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tcgen05<false, false, 256, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(256, 4));
  return e == cudaSuccess ? 0 : -2;
}

struct TmaCache {
  std::map<std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t, int>, CUtensorMap> maps;
  // 2-D fp32 tensor: dim0 (contiguous) x dim1 with row pitch ld (floats); box = box0 x box1, zero OOB fill
  int get(const float* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box0, uint32_t box1, int swizzle, CUtensorMap* out) {
    auto key = std::make_tuple((const void*)ptr, dim0, dim1, ld, box0, box1, swizzle);
    auto it = maps.find(key);
    if (it != maps.end()) { *out = it->second; return 0; }
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld * 4};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)r;
    if (maps.size() > 4096) maps.clear();
    maps[key] = tm;
    *out = tm;
    return 0;
  }
  // 3-D fp32 tensor (uncached): dims {d0, d1, d2}, byte strides of dims 1 and 2, box {b0, b1, b2}; OOB reads give zeros, OOB writes are dropped
  static int get3(const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes, uint64_t s2_bytes, uint32_t b0, uint32_t b1,
                  uint32_t b2, int swizzle, CUtensorMap* out) {
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1_bytes, s2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
  }
  // 2-D fp16 tensor (uncached: built once at create): dim0 x dim1 halves, row pitch ld halves
  static int get16(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box0, uint32_t box1, int swizzle, CUtensorMap* out) {
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
  }
};

static bool tcgen05_gemm_supported(const GemmArgs& g, bool a_mn, bool b_mn) {
  // TMA: 16-byte aligned base and row pitch
  if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15) || (g.lda & 3) || (g.ldb & 3)) return false;
  if (g.N < 1 || g.M < 1 || g.K < 1) return false;
  (void)a_mn; (void)b_mn;
  return true;
}

// pdl: launch with programmatic stream serialization so the kernel's prologue overlaps its predecessor's tail
template <bool A_MN, bool B_MN, int BN, int STAGES>
static void tg_launch(dim3 grid, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, bool pdl) {
  if (!pdl) { k_gemm_tcgen05<A_MN, B_MN, BN, STAGES><<<grid, TG_THREADS, tg_smem(BN, STAGES), st>>>(ta, tb, ta, g); return; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(TG_THREADS); cfg.dynamicSmemBytes = tg_smem(BN, STAGES); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k_gemm_tcgen05<A_MN, B_MN, BN, STAGES>, ta, tb, ta, g);      // tmC is only used by the fp16 variant
}
template <bool A_MN, bool B_MN>
static void tg_dispatch(int BN, bool deep, dim3 grid, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g,
                        bool pdl) {
  if (BN == 256) { if (deep) tg_launch<A_MN, B_MN, 256, 4>(grid, st, ta, tb, g, pdl); else tg_launch<A_MN, B_MN, 256, 2>(grid, st, ta, tb, g, pdl); }
  else if (BN == 64) { if (deep) tg_launch<A_MN, B_MN, 64, 8>(grid, st, ta, tb, g, pdl); else tg_launch<A_MN, B_MN, 64, 4>(grid, st, ta, tb, g, pdl); }
  else { if (deep) tg_launch<A_MN, B_MN, 128, 6>(grid, st, ta, tb, g, pdl); else tg_launch<A_MN, B_MN, 128, 3>(grid, st, ta, tb, g, pdl); }
}

// fp16 K-major GEMM with prebuilt tensor maps (A: [M, K] halves box {64,128}; B: [N, K] halves box {64,256}); 128x256 tiles
// C: fp32 [rows, N] box {32, 32} SWIZZLE_128B (the tile leaves through TMA stores; rows beyond the tensor are clipped)
static int tcgen05_gemm_f16_launch(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmArgs& g) {
  dim3 grid((g.N + 255) / 256, (g.M + TG_BM - 1) / TG_BM, 1);
  if ((long)grid.x * grid.y <= 148) k_gemm_tcgen05<false, false, 256, 4, true><<<grid, TG_THREADS, tg_smem(256, 4), st>>>(ta, tb, tc, g);
  else k_gemm_tcgen05<false, false, 256, 2, true><<<grid, TG_THREADS, tg_smem(256, 2), st>>>(ta, tb, tc, g);
  return cudaGetLastError() == cudaSuccess ? 0 : 3000;
}

// splits == 0: choose a split-K factor for few-tile / long-K problems (the caller must then accept a red.global epilogue)
static int tcgen05_gemm_launch(TmaCache& cache, cudaStream_t st, bool a_mn, bool b_mn, const GemmArgs& g, int splits = 1,
                               bool pdl = false) {
  // tile width: the widest of 256 / 128 / 64 that still yields about a wave of CTAs.  The head GEMMs (M = batch, N <= 512) are
  // bound by what ONE CTA can stream through its ring, not by the tensor pipe: 2048x1456x512 as 32 CTAs of 128x256 took 57 us
  const int rows_t = (g.M + TG_BM - 1) / TG_BM;
  auto tiles_for = [&](int bn) { return (long)rows_t * ((g.N + bn - 1) / bn) * splits; };
  int BN = g.N >= 512 ? 256 : 128;
  if (getenv("ICL_GEMM_NO_BN64") == nullptr) {
    if (BN == 256 && tiles_for(256) < 100) BN = 128;
    if (BN == 128 && tiles_for(128) < 100 && g.N > 64) BN = 64;
  }
  CUtensorMap ta, tb;
  int r;
  const int swk = (int)CU_TENSOR_MAP_SWIZZLE_128B, swmn = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (a_mn) r = cache.get(g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda, 32, TG_BK, swmn, &ta);
  else r = cache.get(g.A, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.lda, TG_BK, TG_BM, swk, &ta);
  if (r) return 1000 + r;
  if (b_mn) r = cache.get(g.B, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldb, 32, TG_BK, swmn, &tb);
  else r = cache.get(g.B, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldb, TG_BK, BN, swk, &tb);
  if (r) return 2000 + r;
  dim3 grid((g.N + BN - 1) / BN, (g.M + TG_BM - 1) / TG_BM, splits);
  const bool deep = (long)grid.x * grid.y * grid.z <= 148;             // one CTA per SM anyway: spend the smem on pipeline depth
  if (!a_mn && !b_mn) tg_dispatch<false, false>(BN, deep, grid, st, ta, tb, g, pdl);
  else if (!a_mn && b_mn) tg_dispatch<false, true>(BN, deep, grid, st, ta, tb, g, pdl);
  else if (a_mn && b_mn) tg_dispatch<true, true>(BN, deep, grid, st, ta, tb, g, pdl);
  else tg_dispatch<true, false>(BN, deep, grid, st, ta, tb, g, pdl);
  return cudaGetLastError() == cudaSuccess ? 0 : 3000;
}

}  // namespace icl
