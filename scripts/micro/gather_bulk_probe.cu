// Micro-probe (B200): per-SM throughput of a random ROW GATHER global -> shared memory, three ways:
//   mode 0  cp.async 16 B (8 lanes per 128-byte segment, as conv_tc.cu / conv_bx3.cu do)
//   mode 1  cp.async.bulk (1-D bulk copy through the TMA unit), ONE instruction per row of row_bytes, one thread per row
//   mode 2  as mode 1 but rows flagged missing issue nothing (the zero-fill case costs no operation)
// 148 CTAs x (128 producer threads + 1 consumer warp); a stage = 128 rows x row_bytes; the consumer only recycles stages.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bulk_probe gather_bulk_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t n) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory"); }
__device__ __forceinline__ void cp_async_arrive(uint64_t* b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int STAGES_MAX = 16;

__global__ void __launch_bounds__(288, 1) probe(const uint8_t* __restrict__ feats, const int32_t* __restrict__ idx, int n_stages_total,
                                               int row_bytes, int pitch, int stages, int mode, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES_MAX], empty[STAGES_MAX];
  const int tid = threadIdx.x;
  const int stage_bytes = 128 * pitch;
  const int pw = (mode == 3 || mode == 4) ? 8 : 4;          // producer warps
  const int ptn = pw * 32;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + s, (mode == 4 || mode == 5) ? pw : ptn); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int32_t* my_idx = idx + size_t(blockIdx.x) * n_stages_total * 128;
  if (tid < ptn && mode >= 3 && mode < 6) {
    // the conv kernels' structure: 8 (or 4) producer warps, 8 lanes per 128-byte row segment
    const int chunk = tid & 7, rbase = tid >> 3;
    const int rows_per_pass = ptn / 8, passes = 128 / rows_per_pass;
    const int segs = row_bytes / 128;
    int s = 0; uint32_t ph = 0;
    int pending = 0;
    for (int it = 0; it < n_stages_total; ++it) {
      int32_t rows[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rows[i] = i < passes ? __ldg(my_idx + it * 128 + rbase + rows_per_pass * i) : -1;
      mbar_wait(empty + s, ph ^ 1);
      const uint32_t base = smem_u32(smem) + s * stage_bytes;
      for (int sg = 0; sg < segs; ++sg) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < passes) {
            const int r = rbase + rows_per_pass * i;
            const bool ok = rows[i] >= 0;
            cp_async16(base + r * pitch + sg * 128 + ((chunk ^ (r & 7)) << 4), feats + size_t(ok ? rows[i] : 0) * row_bytes + sg * 128 + chunk * 16, ok ? 16u : 0u);
          }
        }
      }
      if (mode == 3) {
        cp_async_arrive(full + s);
      } else {
        // commit groups, keep 2 in flight, one elected arrival per warp for the stage that has landed
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++pending;
        if (pending > 2) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");
          __syncwarp();
          int sd = s - 2; if (sd < 0) sd += stages;
          if ((tid & 31) == 0) mbar_arrive(full + sd);
          --pending;
        }
      }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    if (mode != 3) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      for (int j = pending; j >= 1; --j) { int sd = s - j; while (sd < 0) sd += stages; if ((tid & 31) == 0) mbar_arrive(full + sd); }
    }
  } else if (tid < 128 && (mode < 3 || mode == 6)) {
    int s = 0; uint32_t ph = 0;
    if (mode == 0 || mode == 6) {
      const int chunk = tid & 7, rbase = tid >> 3;   // 16 rows per pass, 8 passes
      const int segs = row_bytes / 128;
      for (int it = 0; it < n_stages_total; ++it) {
        int32_t rows[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rows[i] = __ldg(my_idx + it * 128 + rbase + 16 * i);
        mbar_wait(empty + s, ph ^ 1);
        const uint32_t base = smem_u32(smem) + s * stage_bytes;
        for (int sg = 0; sg < segs; ++sg) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rbase + 16 * i;
            const bool ok = rows[i] >= 0;
            if (mode == 6) {   // missing rows issue NOTHING (the consumer zero-fills from the index instead)
              if (ok) cp_async16(base + r * pitch + sg * 128 + ((chunk ^ (r & 7)) << 4), feats + size_t(rows[i]) * row_bytes + sg * 128 + chunk * 16, 16u);
            } else {
              cp_async16(base + r * pitch + sg * 128 + ((chunk ^ (r & 7)) << 4), feats + size_t(ok ? rows[i] : 0) * row_bytes + sg * 128 + chunk * 16, ok ? 16u : 0u);
            }
          }
        }
        cp_async_arrive(full + s);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    } else {
      for (int it = 0; it < n_stages_total; ++it) {
        const int32_t row = __ldg(my_idx + it * 128 + tid);
        mbar_wait(empty + s, ph ^ 1);
        const uint32_t dst = smem_u32(smem) + s * stage_bytes + tid * pitch;
        if (row >= 0 || mode == 1) {
          mbar_expect_tx(full + s, uint32_t(row_bytes));
          bulk_copy(dst, feats + size_t(row >= 0 ? row : 0) * row_bytes, uint32_t(row_bytes), full + s);
        } else {
          mbar_arrive(full + s);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (tid >= 256) {
    int s = 0; uint32_t ph = 0; uint32_t acc = 0;
    for (int it = 0; it < n_stages_total; ++it) {
      mbar_wait(full + s, ph);
      acc += smem[s * stage_bytes + (tid - 256) * 4];
      __syncwarp();
      if (tid == 256) mbar_arrive(empty + s);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    if (acc == 0xFFFFFFFFu) sink[0] = acc;
  }
}

int main(int argc, char** argv) {
  const int n_rows = 150000;
  const int iters = 400;   // stages per CTA
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int row_bytes : {128, 384, 512}) {
    uint8_t* feats; cudaMalloc(&feats, size_t(n_rows) * row_bytes); cudaMemset(feats, 1, size_t(n_rows) * row_bytes);
    for (double fill : {1.0, 0.55}) {
      std::vector<int32_t> h(size_t(148) * iters * 128);
      srand(1);
      for (auto& v : h) v = (rand() / double(RAND_MAX) < fill) ? rand() % n_rows : -1;
      int32_t* idx; cudaMalloc(&idx, h.size() * 4); cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
      uint32_t* sink; cudaMalloc(&sink, 4);
      for (int mode = 0; mode < 7; ++mode) {
        for (int pad : {0, 16}) {
          if ((mode == 0 || mode >= 3) && pad) continue;
          if (mode >= 3 && mode < 6 && fill < 1.0) continue;
          if (mode == 1 || mode == 2 || mode == 4 || mode == 5) continue;
          const int pitch = row_bytes + pad;
          int stages = (200 * 1024) / (128 * pitch); if (stages > STAGES_MAX) stages = STAGES_MAX;
          for (int st : {stages, 4}) {
            if (st > stages) continue;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            probe<<<148, 288, st * 128 * pitch>>>(feats, idx, iters, row_bytes, pitch, st, mode, sink);
            cudaEventRecord(a);
            probe<<<148, 288, st * 128 * pitch>>>(feats, idx, iters, row_bytes, pitch, st, mode, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            const double slots = double(iters) * 128;              // per SM
            const double cyc = ms * 1e-3 * 1.965e9;
            printf("row_bytes %3d fill %.2f mode %d pad %2d stages %2d: %7.3f ms  %6.1f cyc/128-row stage  %5.1f B/clk/SM (all slots)  %5.1f B/clk/SM (real)  %s\n",
                   row_bytes, fill, mode, pad, st, ms, cyc / iters, slots * row_bytes / cyc, slots * row_bytes * fill / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
        }
      }
      cudaFree(idx); cudaFree(sink);
    }
    cudaFree(feats);
  }
  return 0;
}
