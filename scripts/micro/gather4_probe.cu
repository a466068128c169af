// Probe of cp.async.bulk.tensor.2d.tile::gather4 on sm_100a: box shape accepted by cuTensorMapEncodeTiled, placement of
// the 4 gathered rows in shared memory, 128B-swizzle behaviour and out-of-bounds (negative / >= rows) row indices.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int4 rows, int col0, float* out, int dst_row_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  float* tile = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = -777.f;   // 8 KB sentinel
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512u) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem + dst_row_off * 128)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(col0), "r"(rows.x), "r"(rows.y),
        "r"(rows.z), "r"(rows.w)
        : "memory");
  }
  uint32_t done = 0;
  int spins = 0;
  while (!done && spins < (1 << 22)) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    ++spins;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = tile[i];
  if (threadIdx.x == 0) out[2048] = done ? 1.f : 0.f;
}

int main() {
  const int N = 1000, C = 96;
  std::vector<float> h(size_t(N) * C);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < C; ++c) h[size_t(r) * C + c] = r + c / 1000.f;   // value encodes (row, col)
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 2049 * 4);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(sym);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  for (int boxrows : {1, 4}) {
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {cuuint64_t(C), cuuint64_t(N)};
    const cuuint64_t gstride[1] = {cuuint64_t(C) * 4};
    const cuuint32_t box[2] = {32, cuuint32_t(boxrows)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== box rows %d: encode rc=%d\n", boxrows, int(cr));
    if (cr != CUDA_SUCCESS) continue;
    struct Case { int4 rows; int col0; int dst_row; const char* name; };
    Case cases[] = {{{5, 17, 999, 3}, 0, 0, "valid rows, col 0, dst row 0"},
                    {{5, 17, 999, 3}, 64, 4, "valid rows, col 64 (last block), dst row 4"},
                    {{5, -1, 1000, 7}, 32, 0, "rows {5,-1,1000,7}: negative and == N"},
                    {{2000000, 6, -5, 8}, 0, 8, "rows {2e6,6,-5,8}, dst row 8"}};
    for (auto& cs : cases) {
      cudaMemset(out, 0, 2049 * 4);
      probe<<<1, 128, 16 * 1024>>>(tmap, cs.rows, cs.col0, out, cs.dst_row);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(2049);
      cudaMemcpy(o.data(), out, 2049 * 4, cudaMemcpyDeviceToHost);
      printf("-- %s: sync=%s barrier_done=%g\n", cs.name, cudaGetErrorString(e), o[2048]);
      if (e != cudaSuccess) return 1;
      // report every 16-byte chunk that changed: tile row = idx/32, chunk = (idx%32)/4
      for (int row = 0; row < 16; ++row) {
        bool any = false;
        for (int c = 0; c < 32; ++c) any |= o[row * 32 + c] != -777.f;
        if (!any) continue;
        printf("   smem row %2d:", row);
        for (int ch = 0; ch < 8; ++ch) {
          float v = o[row * 32 + ch * 4];
          if (v == -777.f) printf("   [ -- ]");
          else printf(" [%7.3f]", v);   // integer part = source row, fraction*1000 = source column of the chunk's first element
        }
        printf("\n");
      }
    }
  }
  return 0;
}
