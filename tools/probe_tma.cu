// Bring-up probe: what one SM's TMA unit delivers from L2-resident data, by box shape and by the number of boxes in flight.
//   mode 0: 2-D tensor map, fp16 [rows][360], box {64, 128} SWIZZLE_128B   (k_rec_fwd16's h tile: 128 rows x 128 B, pitch 720 B)
//   mode 1: 2-D tensor map, fp16 [rows][64],  box {64, 128} SWIZZLE_128B   (the same box over contiguous rows, pitch 128 B)
//   mode 2: 1-D bulk copy of 16 KB (cp.async.bulk.shared::cluster.global)
//   mode 3: 2-D tensor map, fp32 [rows][1200], box {20, 32} no swizzle     (k_rec_fwd16's x-projection boxes: 32 rows x 80 B)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_tma.bin tools/probe_tma.cu -lcuda ; run on a B200.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_test(uint32_t bar, uint32_t parity) {          // non-blocking test_wait in a spin loop
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity) {          // try_wait with a 20 ns suspend-time hint
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 20;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm),
               "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Args { int mode, depth, boxes, box_bytes, n_units, stage_bytes, lanes, wait, shared; const char* base; long long* out; };

// one thread per CTA drives the TMA unit: `depth` boxes in flight, `boxes` boxes in total; unit u of the data = one box worth
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int w = a.lanes ? (int)threadIdx.x : (int)(threadIdx.x >> 5), nw = a.lanes ? a.lanes : (int)(blockDim.x >> 5);
  const uint32_t region = (196 * 1024 / nw) & ~1023u;
  const uint32_t bars = base + 196 * 1024 + 64 * w;
  const uint32_t mybase = base + w * region;
  if (a.lanes ? (int)threadIdx.x < a.lanes : (threadIdx.x & 31) == 0) {
    for (int s = 0; s < 8; s++) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t0 = clock64();
    int issued = 0, done = 0;
    while (done < a.boxes) {
      while (issued < a.boxes && issued - done < a.depth) {
        const int s = issued % a.depth;
        const uint32_t bar = bars + 8 * s, dst = mybase + s * a.stage_bytes;
        const int u = a.shared ? (int)(((long)(blockIdx.x / a.shared) * 7919 + (long)issued * 6 + w) % a.n_units)      // groups of `shared` CTAs read the SAME boxes
                               : (int)(((long)(blockIdx.x * 4 + w) * 7919 + (long)issued * 131) % a.n_units);     // spread over the resident region
        mbar_expect_tx(bar, a.box_bytes);
        if (a.mode == 0) tma_load_2d(dst, &tm, (u % 5) * 64, (u / 5) * 128, bar);
        else if (a.mode == 1) tma_load_2d(dst, &tm, 0, u * 128, bar);
        else if (a.mode == 2) bulk_load_1d(dst, a.base + (long)u * a.box_bytes, a.box_bytes, bar);
        else tma_load_2d(dst, &tm, (u % 60) * 20, (u / 60) * 32, bar);
        issued++;
      }
      const int s = done % a.depth;
      if (a.wait == 0) mbar_wait(bars + 8 * s, (done / a.depth) & 1);
      else if (a.wait == 1) mbar_wait_test(bars + 8 * s, (done / a.depth) & 1);
      else mbar_wait_hint(bars + 8 * s, (done / a.depth) & 1);
      done++;
    }
    if (w < 4) a.out[blockIdx.x * 4 + w] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn encode = (EncodeFn)fn;
  const size_t bytes = 64u << 20;                        // L2-resident working set
  char* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
  long long* out; CK(cudaMalloc(&out, 148 * 4 * 8));
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const double ghz = pr.clockRate * 1e-6;
  printf("device %s, %d SMs, clock %.3f GHz (clock64 ticks are converted with it)\n", pr.name, pr.multiProcessorCount, ghz);
  const char* names[4] = {"2-D box {64 h,128 rows} pitch 720 B SW128", "2-D box {64 h,128 rows} pitch 128 B SW128", "1-D bulk 16 KB", "2-D box {20 f,32 rows} pitch 4800 B"};
  struct Case { int mode, box_bytes; };
  const Case cases[] = {{0, 16384}, {2, 16384}};
  for (const Case& c : cases) {
    const int mode = c.mode;
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    Args a; a.mode = mode; a.base = buf; a.out = out; a.boxes = 300; a.box_bytes = c.box_bytes; a.stage_bytes = (c.box_bytes + 1023) & ~1023;
    if (mode != 2) {
      cuuint64_t dims[2], strides[1]; cuuint32_t box[2], es[2] = {1, 1};
      CUtensorMapDataType dt = mode == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
      const size_t pitch = mode == 0 ? 720 : mode == 1 ? 128 : 4800;
      const size_t rows = bytes / pitch;
      dims[0] = mode == 0 ? 360 : mode == 1 ? 64 : 1200; dims[1] = rows; strides[0] = pitch;
      box[0] = mode == 3 ? 20 : 64; box[1] = mode == 3 ? 32 : 128;
      CUresult r = encode(&tm, dt, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          mode == 3 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed mode %d: %d\n", mode, (int)r); return 1; }
      a.n_units = mode == 0 ? (int)(rows / 128) * 5 : mode == 1 ? (int)(rows / 128) : (int)(rows / 32) * 60;
    } else a.n_units = (int)(bytes / c.box_bytes);
    for (int ctas : {120}) {
      for (int shared : {0, 4, 19, 120})
      for (int wait : {0})
      for (int issuers : {-6}) {
        for (int depth : {1, 2}) {
          const int ni = issuers < 0 ? -issuers : issuers;
          if ((long)ni * depth * a.stage_bytes > 196 * 1024) continue;
          a.depth = depth; a.lanes = issuers < 0 ? ni : 0; a.wait = wait; a.shared = shared;
          for (int rep = 0; rep < 2; rep++) { k_probe<<<ctas, issuers < 0 ? 32 : 32 * issuers, 200 * 1024>>>(tm, a); CK(cudaDeviceSynchronize()); }
          std::vector<long long> h(ctas * 4);
          CK(cudaMemcpy(h.data(), out, ctas * 4 * 8, cudaMemcpyDeviceToHost));
          double mx = 0;
          for (int i = 0; i < ctas; i++) for (int w = 0; w < ni && w < 4; w++) mx = h[i * 4 + w] > mx ? h[i * 4 + w] : mx;
          const double ns_box = mx / ghz / (a.boxes * ni);
          printf("%-42s %6d B same-address group %3d wait %d ctas %3d issuers %d depth %d : %7.1f ns per box, %6.1f GB/s per SM, %6.0f GB/s total\n", mode == 2 ? "1-D bulk" : names[mode],
                 a.box_bytes, shared, wait, ctas, issuers, depth, ns_box, a.box_bytes / ns_box, ctas * a.box_bytes / ns_box);
        }
      }
    }
  }
  return 0;
}
