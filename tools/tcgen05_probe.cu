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

// MN-major operands: D[m][n] = sum_k SA[k][m] SB[k][n] with SA ([KT][128]) and SB ([KT][N]) stored as K-major IMAGES of
// their own shape (rows = k) and read through MN-major descriptors whose two stride fields and K-block step are runtime
// parameters (the probe tries the candidates).  mode bit 0: A MN-major (else A is given K-major as [128][KT]); bit 1: B
// MN-major (else K-major [N][KT]); bit 2: M = 64, all 128 TMEM lanes dumped so that the host can say where the rows live.
__device__ __forceinline__ uint64_t raw_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
template <int N, int KT>
__global__ void __launch_bounds__(128) probe_mn_kernel(const float* __restrict__ SA, const float* __restrict__ SB, float* __restrict__ D, int mode,
                                                       uint32_t lbo, uint32_t sbo_a, uint32_t sbo_b, uint32_t kstep) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* Ah = reinterpret_cast<float*>(smem);  // image [KT][128] (MN-major) or [128][KT] (K-major)
  float* Al = Ah + KT * 128;
  float* Bh = Al + KT * 128;                   // image [KT][N] or [N][KT]
  float* Bl = Bh + KT * N;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool amn = mode & 1, bmn = mode & 2, m64 = mode & 4;
  if (tid == 0) { tc::bar_init(&bar, 1); tc::bar_init_fence(); }
  if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
  for (int e = tid; e < KT * 128; e += 128) {  // SA[k][m]
    float hi, lo;
    tc::split(SA[e], hi, lo);
    const int k = e / 128, m = e % 128;
    const int off = amn ? tc::op_offset<KT>(k, m) : tc::op_offset<128>(m, k);
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Ah) + off) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Al) + off) = lo;
  }
  for (int e = tid; e < KT * N; e += 128) {  // SB[k][n]
    float hi, lo;
    tc::split(SB[e], hi, lo);
    const int k = e / N, n = e % N;
    const int off = bmn ? tc::op_offset<KT>(k, n) : tc::op_offset<N>(n, k);
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Bh) + off) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Bl) + off) = lo;
  }
  tc::fence_smem_to_mma();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(m64 ? 64 : 128, N, amn, bmn);
    const uint32_t ah = tc::smem_u32(Ah), al = tc::smem_u32(Al), bh = tc::smem_u32(Bh), bl = tc::smem_u32(Bl);
    for (int kb = 0; kb < KT / 8; kb++) {
      const uint64_t dah = amn ? raw_desc(ah + kb * kstep, lbo, sbo_a) : tc::op_desc<128>(ah, 2 * kb);
      const uint64_t dal = amn ? raw_desc(al + kb * kstep, lbo, sbo_a) : tc::op_desc<128>(al, 2 * kb);
      const uint64_t dbh = bmn ? raw_desc(bh + kb * kstep, lbo, sbo_b) : tc::op_desc<N>(bh, 2 * kb);
      const uint64_t dbl = bmn ? raw_desc(bl + kb * kstep, lbo, sbo_b) : tc::op_desc<N>(bl, 2 * kb);
      tc::mma_tf32(tmem, dal, dbh, idesc, kb > 0);
      tc::mma_tf32(tmem, dah, dbl, idesc, true);
      tc::mma_tf32(tmem, dah, dbh, idesc, true);
    }
    tc::mma_commit(&bar);
  }
  tc::bar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
    for (int i = 0; i < 16; i++) D[tid * N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_free<64>(tmem);
}

template <int N, int KT>
static int run_mn(const char* name, int mode, uint32_t lbo, uint32_t sbo_a, uint32_t sbo_b, uint32_t kstep) {
  std::vector<float> A(KT * 128), B(KT * N), D(128 * N);
  srand(2);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const int smem = (KT * 128 + KT * N) * 2 * 4;
  cudaFuncSetAttribute(probe_mn_kernel<N, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_mn_kernel<N, KT><<<1, 128, smem>>>(dA, dB, dD, mode, lbo, sbo_a, sbo_b, kstep);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  std::vector<double> R(128 * N);
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < N; n++) {
      double s = 0;
      for (int k = 0; k < KT; k++) s += (double)A[k * 128 + m] * (double)B[k * N + n];
      R[m * N + n] = s;
    }
  double maxerr = 0;
  if (!(mode & 4)) {
    for (int i = 0; i < 128 * N; i++) maxerr = fmax(maxerr, fabs(R[i] - (double)D[i]));
    int nzq = 0;
    for (int i = 0; i < 128 * N; i++) nzq += D[i] != 0.0f;
    if (!(mode & 8) || nzq > 0)
      printf("%s: mode %d N=%d K=%d lbo %u sbo_a %u sbo_b %u kstep %u  max|err| %.3e nonzero %d %s\n", name, mode, N, KT, lbo, sbo_a, sbo_b, kstep, maxerr,
             nzq, maxerr < 2e-5 ? "OK" : "");
    if (maxerr >= 2e-5 && !(mode & 8)) {
      printf("     D[0][0..5] = %g %g %g %g %g %g   ref %g %g %g %g %g %g\n", D[0], D[1], D[2], D[3], D[4], D[5], R[0], R[1], R[2], R[3], R[4], R[5]);
      printf("     D[1][0..3] = %g %g %g %g   ref %g %g %g %g;  D[5][0..1] = %g %g ref %g %g\n", D[N], D[N + 1], D[N + 2], D[N + 3], R[N], R[N + 1], R[N + 2], R[N + 3], D[5 * N], D[5 * N + 1], R[5 * N], R[5 * N + 1]);
      // does D match the product with the image read K-major (i.e. the flag had no effect)?
      int nz = 0;
      for (int i = 0; i < 128 * N; i++) nz += D[i] != 0.0f;
      printf("     nonzero entries of D: %d of %d\n", nz, 128 * N);
    }
  } else {  // where did row m go?
    printf("%s: mode %d N=%d K=%d  row -> TMEM lane:", name, mode, N, KT);
    int found = 0;
    for (int m = 0; m < 64; m++) {
      int where = -1;
      for (int lane = 0; lane < 128 && where < 0; lane++) {
        double err = 0;
        for (int n = 0; n < N; n++) err = fmax(err, fabs(R[m * N + n] - (double)D[lane * N + n]));
        if (err < 2e-5) where = lane;
      }
      if (where >= 0) found++;
      if (m % 8 == 0) printf(" %d->%d", m, where);
    }
    printf("  (%d of 64 rows found)\n", found);
    maxerr = found == 64 ? 0 : 1;
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 2e-5 ? 0 : 1;
}

// A from TMEM: thread r writes row r of A (hi and lo copies) into TMEM columns with tcgen05.st, B stays in shared memory.
template <int N, int K>
__global__ void __launch_bounds__(128) probe_ts_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* Bh = reinterpret_cast<float*>(smem);
  float* Bl = Bh + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::bar_init(&bar, 1); tc::bar_init_fence(); }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
  for (int e = tid; e < N * K; e += 128) {
    float hi, lo;
    tc::split(B[e], hi, lo);
    const int off = tc::op_offset<N>(e / K, e % K);
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Bh) + off) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(Bl) + off) = lo;
  }
  tc::fence_smem_to_mma();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot, row = tmem + ((uint32_t)(32 * warp) << 16);
  constexpr uint32_t kD = 0, kAh = 64, kAl = 64 + K;
  for (int c0 = 0; c0 < K; c0 += 16) {
    float hi[16], lo[16];
    for (int i = 0; i < 16; i++) tc::split(A[tid * K + c0 + i], hi[i], lo[i]);
    tc::tmem_st16(row + kAh + c0, hi);
    tc::tmem_st16(row + kAl + c0, lo);
  }
  tc::tmem_st_wait();
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t idesc = tc::idesc_tf32(128, N);
    const uint32_t bh = tc::smem_u32(Bh), bl = tc::smem_u32(Bl);
    for (int k0 = 0; k0 < K; k0 += 8) {
      tc::mma_tf32_ts(tmem + kD, tmem + kAl + k0, tc::op_desc<N>(bh, k0 / 4), idesc, k0 > 0);
      tc::mma_tf32_ts(tmem + kD, tmem + kAh + k0, tc::op_desc<N>(bl, k0 / 4), idesc, true);
      tc::mma_tf32_ts(tmem + kD, tmem + kAh + k0, tc::op_desc<N>(bh, k0 / 4), idesc, true);
    }
    tc::mma_commit(&bar);
  }
  tc::bar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(row + kD + c0, v);
    for (int i = 0; i < 16; i++) D[tid * N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_free<256>(tmem);
}
template <int N, int K>
static int run_ts(const char* name) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N);
  srand(3);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const int smem = N * K * 2 * 4;
  cudaFuncSetAttribute(probe_ts_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_ts_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int r = 0; r < 128; r++)
    for (int n = 0; n < N; n++) {
      double s = 0;
      for (int k = 0; k < K; k++) s += (double)A[r * K + k] * (double)B[n * K + k];
      maxerr = fmax(maxerr, fabs(s - (double)D[r * N + n]));
    }
  printf("%s: N=%d K=%d  A from TMEM (tcgen05.st), B from shared memory: max|err| %.3e  D[0][0..3] = %g %g %g %g %s\n", name, N, K, maxerr, D[0], D[1],
         D[2], D[3], maxerr < 2e-5 ? "OK" : "");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 2e-5 ? 0 : 1;
}

int main() {
  run_ts<64, 64>("TS 3xtf32 ");
  run_ts<16, 64>("TS N16    ");

  // informational (not part of the verdict): where M = 64 puts its rows, and what MN-major no-swizzle operands return
  run_mn<64, 64>("M64 K-major          ", 4, 0, 0, 0, 0);
  {  // the image of SA [KT = 64][128]: 4-wide chunks of the operand's M/N index are (KT / 8) * 128 = 1024 B apart, 8-row K groups 128 B
    const uint32_t big = 1024, small = 128;
    run_mn<64, 64>("B mn: lbo=K sbo=MN   ", 2, small, big, big, small);
    run_mn<64, 64>("B mn: lbo=MN sbo=K   ", 2, big, small, small, small);
    run_mn<64, 64>("A mn: lbo=K sbo=MN   ", 1, small, big, big, small);
    int any = 0;
    const uint32_t v[] = {16, 32, 64, 128, 256, 512, 1024, 2048};
    for (uint32_t lbo : v)
      for (uint32_t sbo : v) any += run_mn<64, 8>("grid B mn K8", 2 | 8, lbo, sbo, sbo, 128) == 0;
    printf("MN-major no-swizzle tf32, 64 (lbo, sbo) combinations: %d give the product\n", any);
  }
  int bad = 0;
  bad += run<64, 64, 0>("tf32      ");
  bad += run<64, 64, 1>("3xtf32    ");
  bad += run<64, 16, 1>("3xtf32 K16");
  bad += run<16, 64, 1>("3xtf32 N16");
  bad += run<128, 16, 1>("3xtf32 N128");
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
