// Blackwell (sm_100a) primitives used by the tensor-core edge kernels: mbarrier, bulk async copy,
// TMEM allocation, tcgen05.mma (kind::tf32, A from TMEM, B from shared memory), tcgen05.ld/st.
//
// Operand conventions (validated by csrc/tests/tc_gemm_test.cu on hardware):
//   D[m][n] (fp32 accumulator): TMEM lane m, column d_col + n                    (M = 128 rows = 128 lanes)
//   A[m][k] (tf32, from TMEM) : TMEM lane m, column a_col + k
//   B[n][k] (tf32, from smem) : "K-major" = nn.Linear's own [out][in] order, stored as 128-byte-swizzled slabs:
//       slab kb holds k in [32kb, 32kb+32): row n at byte n*128, 16-byte chunk c at ((c ^ (n & 7)) * 16)
//       (slabs are 1024-byte aligned; 8-row groups are 1024 bytes apart = the descriptor's SBO)
//   3xTF32: x = hi + lo with hi = x & 0xffffe000 (what the tensor core reads of x anyway) and lo = x - hi (exact);
//       D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, fp32 accumulation.  Dropped: lo.lo (<= 2^-20 relative).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tmpnn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must not hang the GPU (a hung box is a strike); on timeout, trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 26)) { asm volatile("trap;"); }
    }
}

// ---- bulk async copy global -> shared, completion on an mbarrier ---------------------------------------------------
// one 2-D tile of a tensor map -> shared memory (TMA); completes `bytes of the box` transactions on the mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy accesses (ld/st.shared) before this fence are ordered before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns: thread t of the warp <-> lane (32*(warp%4) + t), registers <-> consecutive columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 format): K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4)      // start address, 16-byte units      bits [0,14)
           | ((uint64_t)1 << 16)                         // leading byte offset (unused here) bits [16,30)
           | ((uint64_t)(1024 >> 4) << 32)               // stride byte offset = 1024         bits [32,46)
           | ((uint64_t)1 << 46)                         // descriptor version (Blackwell)    bits [46,48)
           | ((uint64_t)2 << 61);                        // layout: SWIZZLE_128B              bits [61,64)
}
// Instruction descriptor: D fp32, A/B tf32, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem d] (+)= A[tmem a] . B[smem desc]^T, one K = 8 step; issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same with the descriptor given as (low word, high word).  The issuing thread is a single lane running at
// ~0.2 IPC: every scalar instruction between two MMAs costs ~5 cycles of issue rate (measured: 67 cycles/MMA with
// precomputed descriptors vs 145-190 with per-MMA descriptor arithmetic, csrc/tests/tc_mma_timing.cu), so callers
// keep desc_lo = base + compile-time offset and a constant desc_hi.
__device__ __forceinline__ void mma_tf32_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.eq.u32 p, 1, 1;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void mma_tf32_ts_lh_init(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(   // D = A.B (accumulator overwritten)
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.eq.u32 p, 1, 0;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
}
// One lane of a converged warp, chosen by elect.sync.  ptxas treats a branch on this predicate as a single-thread
// region and keeps warp-uniform operands in uniform registers; a branch on `lane == 0` instead makes it wrap every
// tcgen05.mma in an ELECT + 4x R2UR.BROADCAST "waterfall" loop (~100 cycles per MMA instead of the pipe's 64).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
constexpr uint32_t DESC_SW128_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_sw128_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }

// Arrive on an mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- packed FP32 pairs (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot for two IEEE fp32 operations -----------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// gelu_fast on a pair: the same operation sequence per lane (bit-identical results), 7.5 instructions per element
__device__ __forceinline__ f32x2 gelu_fast2(f32x2 x) {
    float x0, x1;
    upk2(x, x0, x1);
    const f32x2 t = pk2(fminf(fabsf(x0), 5.7f), fminf(fabsf(x1), 5.7f));
    f32x2 p = pk2(4.545353022e-06f, 4.545353022e-06f);
    p = fma2(p, t, pk2(-1.562216312e-05f, -1.562216312e-05f));
    p = fma2(p, t, pk2(-5.642375909e-04f, -5.642375909e-04f));
    p = fma2(p, t, pk2(7.647684310e-03f, 7.647684310e-03f));
    p = fma2(p, t, pk2(-5.292505026e-02f, -5.292505026e-02f));
    p = fma2(p, t, pk2(-4.590552747e-01f, -4.590552747e-01f));
    p = fma2(p, t, pk2(-1.151124954e+00f, -1.151124954e+00f));
    p = mul2(p, t);
    float p0, p1, e0, e1;
    upk2(p, p0, p1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p1));
    // x Phi(x) = relu(x) - (t/2) 2^P(t):  x >= 0: x (1 - e/2);  x < 0: x e / 2
    return fma2(mul2(t, pk2(-0.5f, -0.5f)), pk2(e0, e1), pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }

// byte offset of element (row n, column k) inside a [rows x 128] operand stored as four SW128 slabs of `rows` rows
__host__ __device__ constexpr uint32_t sw128_offset(int rows, int n, int k) {
    return (uint32_t)((k >> 5) * rows * 128 + n * 128 + ((((k & 31) >> 2) ^ (n & 7)) << 4) + ((k & 3) << 2));
}

// GELU with fp32-level absolute accuracy in 12 instructions: Phi(x) = (1 + erf(x/sqrt2))/2 with
// erf(|x|/sqrt2) = 1 - 2^P(|x|), P a degree-7 polynomial (fit in tools/fit_gelu.py; max |err| 2.8e-7 on [-8,8],
// at or below torch's own fp32 erf-GELU in every range).  torch.nn.GELU() (protein_mpnn_utils.py:813,856,888).
__device__ __forceinline__ float gelu_fast(float x) {
    const float t = fminf(fabsf(x), 5.7f);
    float p = 4.545353022e-06f;
    p = fmaf(p, t, -1.562216312e-05f);
    p = fmaf(p, t, -5.642375909e-04f);
    p = fmaf(p, t, 7.647684310e-03f);
    p = fmaf(p, t, -5.292505026e-02f);
    p = fmaf(p, t, -4.590552747e-01f);
    p = fmaf(p, t, -1.151124954e+00f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
    return fmaf(-0.5f * t, e, fmaxf(x, 0.f));     // relu(x) - (t/2) 2^P(t)
}

// ---- shared by the chained-GEMM kernels (tc_edge_kernels.cu, tc_node_kernels.cu) -------------------------------------------
// TMEM columns: two fp32 accumulators and the hi / lo planes of the A operand, 128 columns each.
constexpr uint32_t COL_D0 = 0, COL_D1 = 128, COL_AHI = 256, COL_ALO = 384;

__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }


// 16 values -> columns [16 sc, 16 sc + 16) of the next A operand (hi plane: the tf32 the tensor core would read anyway;
// lo plane: the exact remainder, which the tensor core truncates to tf32 itself), then publish the sub-chunk.
__device__ __forceinline__ void publish_A16(uint32_t lane_addr, int sc, const float (&v)[16], uint64_t* a_ready) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        hi[j] = __float_as_uint(v[j]) & 0xffffe000u;
        hi[j + 1] = __float_as_uint(v[j + 1]) & 0xffffe000u;
        float l0, l1;                                   // lo = v - hi, exact (packed: fma(hi, -1, v))
        upk2(fma2(pk2(__uint_as_float(hi[j]), __uint_as_float(hi[j + 1])), pk2(-1.0f, -1.0f), pk2(v[j], v[j + 1])), l0, l1);
        lo[j] = __float_as_uint(l0);
        lo[j + 1] = __float_as_uint(l1);
    }
    tmem_st16(lane_addr + COL_AHI + sc * 16, hi);
    tmem_st16(lane_addr + COL_ALO + sc * 16, lo);
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(a_ready);
}
__device__ __forceinline__ void load_vec16(float (&v)[16], const float* p) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 t = *reinterpret_cast<const float4*>(p + j * 4);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}


}  // namespace tc
}  // namespace tmpnn
