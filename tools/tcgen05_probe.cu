// tcgen05_probe.cu - checks csrc/so100_tc.cuh on the GPU: D[128 x N] = A[128 x K] B[N x K]^T with tcgen05.mma kind::tf32
// (3xTF32 split, operands in the no-swizzle K-major layout, accumulator in TMEM, tcgen05.ld epilogue) against an fp64 host
// product.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tcgen05_probe tools/tcgen05_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <cmath>
#include <vector>

#include "../so100_mujoco_rl_b200/csrc/so100_tc.cuh"

template <int N, int K, int SPLIT>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* Ah = reinterpret_cast<float*>(smem);                    // 128 x K
  float* Al = Ah + 128 * K;
  float* Bh = Al + 128 * K;                                      // N x K
  float* Bl = Bh + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::bar_init(&bar, 1); tc::bar_init_fence(); }
  if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e % K;
    float hi, lo;
    tc::split(A[e], hi, lo);
    if (!SPLIT) { hi = A[e]; lo = 0.0f; }
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Ah) + tc::op_offset<128>(r, k)) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Al) + tc::op_offset<128>(r, k)) = lo;
  }
  for (int e = tid; e < N * K; e += 128) {
    const int r = e / K, k = e % K;
    float hi, lo;
    tc::split(B[e], hi, lo);
    if (!SPLIT) { hi = B[e]; lo = 0.0f; }
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Bh) + tc::op_offset<N>(r, k)) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Bl) + tc::op_offset<N>(r, k)) = lo;
  }
  tc::fence_smem_to_mma();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  uint32_t parity = 0;
  for (int rep = 0; rep < reps; rep++) {
    if (tid == 0) {
      const uint32_t idesc = tc::idesc_tf32(128, N);
      const uint32_t ah = tc::smem_u32(Ah), al = tc::smem_u32(Al), bh = tc::smem_u32(Bh), bl = tc::smem_u32(Bl);
      bool acc = false;
      for (int k0 = 0; k0 < K; k0 += 8) {
        if (SPLIT) {
          tc::mma_tf32(tmem, tc::op_desc<128>(al, k0 / 4), tc::op_desc<N>(bh, k0 / 4), idesc, acc);
          tc::mma_tf32(tmem, tc::op_desc<128>(ah, k0 / 4), tc::op_desc<N>(bl, k0 / 4), idesc, true);
          tc::mma_tf32(tmem, tc::op_desc<128>(ah, k0 / 4), tc::op_desc<N>(bh, k0 / 4), idesc, true);
        } else {
          tc::mma_tf32(tmem, tc::op_desc<128>(ah, k0 / 4), tc::op_desc<N>(bh, k0 / 4), idesc, acc);
        }
        acc = true;
      }
      tc::mma_commit(&bar);
    }
    tc::bar_wait(&bar, parity);
    parity ^= 1;
    tc::fence_after_sync();
  }
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
    for (int i = 0; i < 16; i++) D[tid * N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_free<64>(tmem);
}

template <int N, int K, int SPLIT>
static int run(const char* name) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const int smem = (128 * K + N * K) * 2 * 4;
  cudaFuncSetAttribute(probe_kernel<N, K, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<N, K, SPLIT><<<1, 128, smem>>>(dA, dB, dD, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; r++)
    for (int n = 0; n < N; n++) {
      double s = 0;
      for (int k = 0; k < K; k++) s += (double)A[r * K + k] * (double)B[n * K + k];
      maxerr = fmax(maxerr, fabs(s - (double)D[r * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("%s: N=%d K=%d split=%d  max|err| %.3e  (max|ref| %.3f)  D[0][0..3] = %g %g %g %g\n", name, N, K, SPLIT, maxerr, maxref, D[0], D[1], D[2], D[3]);
  // throughput of the issue loop: many repetitions of the same product
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 2000;
  cudaEventRecord(e0);
  probe_kernel<N, K, SPLIT><<<1, 128, smem>>>(dA, dB, dD, reps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("   one CTA: %.3f us per 128 x %d x %d product (%d MMAs + commit + wait)\n", ms * 1e3 / reps, N, K, (SPLIT ? 3 : 1) * K / 8);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  const double tol = SPLIT ? 2e-5 : 3e-2;
  return maxerr < tol ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<64, 64, 0>("tf32      ");
  bad += run<64, 64, 1>("3xtf32    ");
  bad += run<64, 16, 1>("3xtf32 K16");
  bad += run<16, 64, 1>("3xtf32 N16");
  bad += run<128, 16, 1>("3xtf32 N128");
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
