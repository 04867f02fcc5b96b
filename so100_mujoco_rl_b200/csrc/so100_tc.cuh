// so100_tc.cuh — the handful of sm_100a primitives the PPO kernels need for their tensor-core products: tcgen05.mma
// (kind::tf32, operands in shared memory, accumulator in TMEM), TMEM allocation, tcgen05.ld for the epilogue, mbarrier
// completion.  Hand-written PTX; the bit layouts of the two descriptors follow the PTX ISA's "matrix descriptor" and
// "instruction descriptor" tables (cross-checked against CUTLASS's cute/arch/mma_sm100_desc.hpp, which restates them).
//
// Shared-memory operand layout used here: NO swizzle, K-major, 4-byte elements.  The hardware's unit is the 8 x 16 B
// "core matrix" (8 rows of the M/N dimension x 4 consecutive K elements, rows 16 B apart = 128 contiguous bytes).  A
// ROWS x K operand is stored as
//     offset(r, k) = (k / 4) * (ROWS / 8) * 128  +  (r / 8) * 128  +  (r % 8) * 16  +  (k % 4) * 4      [bytes]
// i.e. all row blocks of one 4-wide K chunk, then the next chunk: a thread that owns row r writes its K values as float4
// stores 16 B apart from its neighbours' (conflict-free), and one MMA (K = 8) reads two consecutive chunks:
//     stride byte offset  (next 8 rows)       = 128
//     leading byte offset (next 4 K elements) = (ROWS / 8) * 128
#pragma once
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, k) of a ROWS x K operand in the layout above
template <int ROWS>
__device__ __forceinline__ int op_offset(int r, int k) {
  return (k >> 2) * (ROWS / 8) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4;
}
// matrix descriptor of the two K chunks starting at chunk `kc0` (one MMA consumes K = 8 = two chunks)
template <int ROWS>
__device__ __forceinline__ uint64_t op_desc(uint32_t base_addr, int kc0) {
  const uint32_t addr = base_addr + (uint32_t)kc0 * (ROWS / 8) * 128;
  constexpr uint64_t lbo = (ROWS / 8) * 128, sbo = 128;
  return (uint64_t)((addr & 0x3FFFFu) >> 4)  // bits [0,14): start address >> 4
         | ((lbo >> 4) << 16)                // bits [16,30): leading dimension byte offset >> 4
         | ((sbo >> 4) << 32)                // bits [32,46): stride dimension byte offset >> 4
         | (1ull << 46);                     // bits [46,48): descriptor version 1 (sm_100); layout type 0 = no swizzle
}
// MN-major operands.  The image above is also the no-swizzle MN-major canonical image of the TRANSPOSED matrix (strides
// exchanged), which would let every backward product of an MLP read the forward pass's buffers as they are - but kind::tf32
// does not take it: with either major bit set and a no-swizzle descriptor the MMA writes zeros, whatever the two stride
// fields say (tools/tcgen05_probe.cu sweeps them; CUTLASS's sm100 builder states the rule: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout").  Only K-major operands are used here.
// instruction descriptor: D fp32, A and B tf32, M x N; a_mn / b_mn: the operand is MN-major (default K-major)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] B[smem]^T : one thread issues for the CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// the same with A read from TMEM (row m of A = TMEM lane m, its K values in consecutive 32-bit columns; K-major only)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool bar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait for the phase with this parity to complete.  A wait of this kernel family lasts microseconds; a barrier that has
// not completed after ~1e8 polls means a lost MMA completion, and the kernel traps instead of hanging the GPU.
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !bar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 27)) __trap();
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes (st.shared by the epilogue threads) -> visible to the async proxy (the MMA's operand reads)
__device__ __forceinline__ void fence_smem_to_mma() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM: NCOLS columns (power of two >= 32) x 128 lanes x 32 bit; one warp allocates and the same warp frees
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_free(uint32_t tmem_base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NCOLS) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane (warp w of the CTA reads lanes 32 w .. 32 w + 31: the caller
// puts (32 w) << 16 into the address); the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  // the registers are named as in/out operands so that no use of them can be scheduled above the wait
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                 "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers; tmem_st_wait() before anything may read them
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
      "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
      "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
      "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 3xTF32 operand split: hi = the 19 bits the tensor core reads, lo = x - hi (exact in fp32)
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

}  // namespace tc
