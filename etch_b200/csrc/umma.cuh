// tcgen05 (5th-gen tensor core) building blocks for sm_100a: TMEM allocation, shared-memory matrix descriptors,
// kind::tf32 MMA issue with a 3xTF32 split (hi*hi + lo*hi + hi*lo keeps fp32-level accuracy), commit / mbarrier wait,
// TMEM -> register loads.  Operand tiles use the canonical K-major NO-SWIZZLE layout
//     element (r, k) of an [R x K] fp32 tile  ->  byte offset (k/4) * (R*16) + r*16 + (k%4)*4
// i.e. core matrices (8 rows x 16 B) are contiguous over rows (SBO = 128 B) and R*16 B apart along K (LBO = R*16 B).
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- issue-side primitives are WARP-COLLECTIVE: all 32 lanes of one (converged) warp call them with warp-uniform
// ---- arguments and exactly one elected lane executes the instruction.  Issuing from a single divergent thread makes
// ---- the compiler wrap every tcgen05.mma in an elect/branch loop (~50 cycles each); the warp-uniform form keeps the
// ---- descriptors in uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
// broadcast lane 0's value so the compiler can treat it as warp-uniform
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// K-major, no swizzle. lbo/sbo in bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// kind::tf32, D = f32, A and B K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (elect_one())
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// all previously issued MMAs arrive on the mbarrier when complete (implies fence::before_thread_sync); warp-collective.
// NOTE: the elected lane must be the one that issued the MMAs -- elect.sync picks the same lane for the same mask.
__device__ __forceinline__ void commit(uint64_t* bar) {
    if (elect_one())
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

// make generic-proxy shared-memory writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; ncols power of two >= 32; the TMEM base address is written to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// 32 lanes x 8 consecutive fp32 columns: thread i of the warp receives row (lane_base + i), columns [col, col+8)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// non-blocking probe of a phase (for the polling control warps)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// 32 lanes x 4 consecutive fp32 columns (no wait: pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// registers -> TMEM: thread i of the warp writes row (lane_base + i), 32 consecutive fp32 columns from `taddr`
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
          "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
          "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
          "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}

// 2-D tiled TMA load (tensor map in kernel-parameter / global memory) completing `bytes` on `bar`; ONE thread calls it
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 1-D bulk copy global -> shared through the TMA engine; completion (bytes) is signalled on `bar` (warp-collective)
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    const uint32_t d = smem_u32(dst_smem), b = smem_u32(bar);
    if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(src_gmem), "r"(bytes), "r"(b)
                     : "memory");
    }
}

// several bulk copies (<= 32 KB each) of one contiguous region, completing ONE phase of `bar` (warp-collective)
__device__ __forceinline__ void bulk_load_region(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    const uint32_t d = smem_u32(dst_smem), b = smem_u32(bar);
    if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        for (uint32_t off = 0; off < bytes; off += 32768) {
            const uint32_t nb = bytes - off < 32768 ? bytes - off : 32768;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(d + off), "l"(reinterpret_cast<const unsigned char*>(src_gmem) + off), "r"(nb), "r"(b)
                         : "memory");
        }
    }
}

// fp32 -> (hi, lo) with hi = x rounded to TF32 (nearest, ties away from zero) and lo = x - hi (exact); the tensor core
// truncates lo to tf32.  Integer form of cvt.rna.tf32.f32 -- same result for every finite input below 2^128 * (1 - 2^-11),
// 2 instructions instead of the 5 the cvt expands to (it adds NaN/Inf handling); etch_b200/models/tc.py uses the same formula
// for the weight operands.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}

// byte offset of element (r, k) in the canonical K-major no-swizzle tile with R rows
__host__ __device__ __forceinline__ constexpr uint32_t tile_off(int r, int k, int R) {
    return (uint32_t)((k >> 2) * (R * 16) + r * 16 + (k & 3) * 4);
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T with the 3xTF32 split; warp-collective (one elected lane issues).
// a_hi/a_lo/b_hi/b_lo: shared-memory byte addresses of canonical tiles ([128 x K] and [N x K]); K % 8 == 0.
// The four descriptors are built once; each k-step only bumps their 14-bit start-address fields (16-byte units).
__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                                  int K, int N, bool accumulate_first) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)N * 16;
    uint64_t ah = make_desc(a_hi, lbo_a, 128), al = make_desc(a_lo, lbo_a, 128);
    uint64_t bh = make_desc(b_hi, lbo_b, 128), bl = make_desc(b_lo, lbo_b, 128);
    const uint64_t da = (uint64_t)((2 * lbo_a) >> 4), db = (uint64_t)((2 * lbo_b) >> 4);
    mma_tf32(tmem_d, ah, bh, idesc, accumulate_first ? 1u : 0u);
    mma_tf32(tmem_d, al, bh, idesc, 1u);
    mma_tf32(tmem_d, ah, bl, idesc, 1u);
    const int nk = K / 8;
#pragma unroll 4
    for (int ks = 1; ks < nk; ++ks) {
        ah += da; al += da; bh += db; bl += db;
        mma_tf32(tmem_d, ah, bh, idesc, 1u);
        mma_tf32(tmem_d, al, bh, idesc, 1u);
        mma_tf32(tmem_d, ah, bl, idesc, 1u);
    }
}

}  // namespace umma
