// InterSO3Conv, third design: one point per tile, neighbour contraction on the FP32 pipes at full register blocking,
// channel mixing on tcgen05 behind it, everything streamed by TMA.
//
//   z[(p,a), o] = sum_{c,k} W[o, c*24+k] * T[a][c][k] + bias[o],   T[a][c][k] = sum_n f[nbr(p,n), a, c] * w[a][k][n],
//   w[a][k][n]  = relu(1 - |g_n - R_a kappa_k|^2 / sigma)
// (reference: vgtk/so3conv/functional.py:286-324,61-67; modules.py:19-39,120-128).
//
// Why this shape.  T of one point is 60 x 32 x 24 fp32 = 184 KB: it fits the register file of ONE CTA (480 threads x 96
// accumulators) but no shared-memory budget, and the contraction over n is 150k independent [24 x nn] x [nn x c] products
// per scan (the weights depend on (p, a), the features on (n, a)), i.e. not a dense GEMM: on tcgen05 it would run at 25 %
// lane utilisation x 3 TF32 passes, which is no faster than the FP32 pipes.  So:
//   * 15 compute warps: thread (a, o, h) owns T[a][8o..8o+7][12h..12h+11].  Per neighbour it needs 8 features (two
//     conflict-free LDS.128 from a 128B-swizzled TMA tile) and 12 weights (three LDS.128) for 96 FMAs.
//   * the neighbour rows arrive by 2-D tiled TMA (box 60 anchors x 32 channels, SWIZZLE_128B) into a 2-3 deep ring of 2-4-neighbour chunks; the
//     per-(a,k,n) weights are generated ONCE per neighbour by all compute threads into a double-buffered tile, two neighbours
//     per packed fma.rn.f32x2 (the geometry scratch stores neighbour pairs as {gx0,gx1,gy0,gy1},{gz0,gz1,gw0,gw1}; same bits as the
//     scalar expression; round 2 A/B 9.43 -> 9.34 ms over the three layers).
//   * when a pass over the neighbours ends the 96 accumulators are parked in TMEM (tcgen05.st, 384 columns) and the warps
//     start the next point at once; while they work on it they pull the parked values back slab by slab (tcgen05.ld), split
//     them into (hi, lo) TF32 and write the canonical A tile of a 48-column K slab.  The control warp streams the matching weight
//     slab (cp.async.bulk) and issues  A_hi x [W_hi; W_lo]  and  A_lo x W_hi  (UMMA M = 64, N = 2*c_out and c_out), so the
//     3xTF32 product reads every operand once; the epilogue adds the two column halves of the accumulator.
// c_in = 64 runs two passes per point (channel halves) that accumulate into the same TMEM tile.
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>

namespace {

constexpr int NA = 60;
constexpr int NK = 24;
constexpr int NPAIRS = NA * NK;              // 1440 (anchor, kernel point) pairs
constexpr int V3_JJ = 1;                     // neighbours per FMA-loop iteration (see the loop: unrolling costs ~60 MOVs per chunk)
#ifndef ETCH_V3_SLEEP_NS
#define ETCH_V3_SLEEP_NS 256
#endif
constexpr unsigned V3_SLEEP = ETCH_V3_SLEEP_NS;  // ns the idle control warp sleeps between polls (A/B: tools/v3_sleep_ab.sh; 64 / 128 / 512 measured equal)
constexpr int NBR_SLOT = 8192;               // one neighbour tile: 60 rows x 128 B, padded to the 1 KB swizzle atom
constexpr int NBR_TX = NA * 128;             // bytes one TMA box delivers
constexpr int CT = 480;                      // compute threads (15 warps)
constexpr int NTHREADS = 512;
constexpr int SLAB_K = 48;                   // K columns per slab: 4 octets x 12 kernel points
constexpr int NSLAB = 16;                    // slabs per pass: 8 channel slots x 2 kernel-point halves
constexpr int A_LBO = 1056;                  // K-chunk stride of the A slab (64 rows x 16 B + 32 B: conflict-free stores)
constexpr int A_HALF = (SLAB_K / 4) * A_LBO; // one (hi or lo) A slab
constexpr int A_SLOT = 2 * A_HALF;
constexpr int GRING = 4;                     // per-point geometry ring depth
constexpr int PARK_COL = 128;                // TMEM: [0, 2*c_out) accumulator, [128, 512) parked T

template <int CIN, int COUT, int NN>
struct V3Cfg {
    // neighbours per chunk / chunk-ring depth: 4 x 2 where the shared-memory budget allows it (c_out = 32), else 2 x 3
    static constexpr int NB = COUT == 32 ? 4 : 2;
    static constexpr int RING = COUT == 32 ? 2 : 3;
    static constexpr int NPASS = CIN / 32;
    static constexpr int NCHUNK = NN / NB;
    static constexpr int SLAB_EVERY = NCHUNK >= NSLAB ? NCHUNK / NSLAB : 1;   // a slab group every SLAB_EVERY chunks ...
    static constexpr int SLAB_GROUP = NCHUNK >= NSLAB ? 1 : NSLAB / NCHUNK;   // ... of SLAB_GROUP slabs
    static constexpr uint32_t F_BYTES = RING * NB * NBR_SLOT;
    static constexpr uint32_t W_BYTES = 2 * NB * NPAIRS * 4;
    static constexpr uint32_t A_BYTES = 2 * A_SLOT;
    static constexpr uint32_t WSLAB = 2 * COUT * SLAB_K * 4;
    static constexpr uint32_t KRS_BYTES = NPAIRS * 16;
    static constexpr uint32_t Z_BYTES = 64 * COUT * 4;
    static constexpr uint32_t G_BYTES = GRING * NN * 16;
    static constexpr uint32_t NBR_BYTES = GRING * NN * 4;
    static constexpr uint32_t STAT_BYTES = COUT * 16;
    static constexpr size_t smem = 1024 + F_BYTES + W_BYTES + A_BYTES + 2 * WSLAB + KRS_BYTES + Z_BYTES + G_BYTES + NBR_BYTES + STAT_BYTES;
    static_assert(SLAB_GROUP * (NCHUNK / SLAB_EVERY) == NSLAB, "slab schedule covers a pass");
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static_assert(2 * COUT <= PARK_COL, "accumulator columns");
};

__device__ __forceinline__ uint64_t v3_pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void v3_unpk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t v3_fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t v3_add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// relative neighbour positions, once per layer:  g4[b][p][n] = {g, 1 - |g|^2 / sigma}
__global__ void inter_geom_kernel(const float* __restrict__ xyz, const int* __restrict__ sample_idx, const int* __restrict__ nbr,
                                  int q, int P, int NN, float inv_sigma, float4* __restrict__ g4) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * NN) return;
    const int p = i / NN;
    const float* X = xyz + (size_t)b * 3 * q;
    const int c = __ldg(sample_idx + (size_t)b * P + p);
    const int k = __ldg(nbr + (size_t)b * P * NN + i);
    const float gx = __ldg(X + k) - __ldg(X + c);
    const float gy = __ldg(X + q + k) - __ldg(X + q + c);
    const float gz = __ldg(X + 2 * (size_t)q + k) - __ldg(X + 2 * (size_t)q + c);
    // pair layout for the packed weight generation: neighbours (2m, 2m+1) -> {gx0, gx1, gy0, gy1}, {gz0, gz1, gw0, gw1}
    float* gp = reinterpret_cast<float*>(g4 + (size_t)b * P * NN + (i & ~1));
    const int e = i & 1;
    gp[e] = gx; gp[2 + e] = gy; gp[4 + e] = gz; gp[6 + e] = 1.0f - (gx * gx + gy * gy + gz * gz) * inv_sigma;
}

template <int CIN, int COUT, int NN>
__global__ void __launch_bounds__(NTHREADS, 1) inter_conv_v3_kernel(
    const __grid_constant__ CUtensorMap tmap,   // features as a 2-D tensor [B*q*60 rows][CIN], box 60 x 32, SWIZZLE_128B
    const float4* __restrict__ g4,              // [B,P,NN]
    const int* __restrict__ nbr,                // [B,P,NN]
    const float4* __restrict__ krs,             // [60,24] {2/sigma * R_a k, |R_a k|^2/sigma}
    const float* __restrict__ Wc,               // [NPASS*16][12][2*COUT][4]  slabs: rows [W_hi; W_lo], canonical K-major tiles
    const float* __restrict__ bias,
    int B, int q, int P,
    float* __restrict__ zraw, double* __restrict__ stats)
{
    using Cfg = V3Cfg<CIN, COUT, NN>;
    constexpr int NB = Cfg::NB, RING = Cfg::RING, NPASS = Cfg::NPASS, NCHUNK = Cfg::NCHUNK, SLAB_EVERY = Cfg::SLAB_EVERY, SLAB_GROUP = Cfg::SLAB_GROUP;
    constexpr uint32_t WSLAB = Cfg::WSLAB;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((1024u - (umma::smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzle atoms need 1 KB alignment
    unsigned char* s_f = base;                                             // [RING][NB][8192]
    float* s_w = reinterpret_cast<float*>(s_f + Cfg::F_BYTES);             // [2][NB][1440]
    unsigned char* s_A = reinterpret_cast<unsigned char*>(s_w) + Cfg::W_BYTES;   // [2][hi | lo]
    unsigned char* s_W = s_A + Cfg::A_BYTES;                               // [2][WSLAB]
    float4* s_krs = reinterpret_cast<float4*>(s_W + 2 * WSLAB);            // [1440]
    float* s_z = reinterpret_cast<float*>(s_krs + NPAIRS);                 // [64][COUT]
    float4* s_g = reinterpret_cast<float4*>(s_z + 64 * COUT);              // [GRING][NN]
    int* s_nbr = reinterpret_cast<int*>(s_g + GRING * NN);                 // [GRING][NN]
    double* s_stat = reinterpret_cast<double*>(s_nbr + GRING * NN);        // [COUT][2]
    __shared__ uint64_t f_full[RING], f_free[RING], g_full[GRING], a_full[2], a_free[2], w_full[2], w_free[2], acc_full, d_free, c_done;
    __shared__ uint32_t tmem_base;

    // Scan-major schedule: every CTA takes its share of scan 0, then of scan 1, ... so that at any time the whole grid gathers
    // from ONE scan's activations (19-38 MB: L2 resident) instead of all B of them (154 MB at B = 8: 3x the algorithmic DRAM
    // traffic in the profile of the scan-parallel grid).
    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31;
    const long long total_pts = (long long)B * P;      // global point gp = blockIdx.x + li * gridDim.x  ->  scan gp / P, point gp % P
    const int npts = blockIdx.x < total_pts ? (int)((total_pts - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const uint32_t NIT = (uint32_t)npts * NPASS;

    for (int i = tid; i < NPAIRS; i += NTHREADS) s_krs[i] = __ldg(krs + i);
    for (int i = tid; i < 2 * COUT; i += NTHREADS) s_stat[i] = 0.0;
    if (wp == 15) umma::tmem_alloc(&tmem_base, 512);
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) { umma::mbar_init(&f_full[i], 1); umma::mbar_init(&f_free[i], 15); }
        for (int i = 0; i < GRING; ++i) umma::mbar_init(&g_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&a_full[i], 8); umma::mbar_init(&a_free[i], 1);
            umma::mbar_init(&w_full[i], 1); umma::mbar_init(&w_free[i], 1);
        }
        umma::mbar_init(&acc_full, 1);
        umma::mbar_init(&d_free, 4);
        umma::mbar_init(&c_done, 15);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);

    if (wp == 15) {
        // =========================== control warp: TMA producer + MMA issuer (polling) ===========================
        const uint32_t total_chunks = NIT * NCHUNK, total_slabs = NIT * NSLAB;
        uint32_t next_f = 0, next_wl = 0, next_mma = 0, next_g = 0;
        const uint32_t idesc1 = umma::make_idesc_tf32(64, 2 * COUT), idesc2 = umma::make_idesc_tf32(64, COUT);
        while (next_f < total_chunks || next_mma < total_slabs) {
            bool progressed = false;
            // ---- tensor-core work first: it frees both the A ring and the weight ring ----
            if (next_mma < total_slabs && next_mma < next_wl) {
                const uint32_t sl = next_mma & 1, j = next_mma / NSLAB, s = next_mma % NSLAB;
                const uint32_t pass = j % NPASS, pt = j / NPASS;
                const bool first = (pass == 0 && s == 0);
                bool ready = umma::mbar_test(&a_full[sl], (next_mma >> 1) & 1) && umma::mbar_test(&w_full[sl], (next_mma >> 1) & 1);
                if (ready && first && pt >= 1) ready = umma::mbar_test(&d_free, (pt - 1) & 1);
                ready = __all_sync(0xffffffffu, ready);
                if (ready) {
                    umma::fence_after_sync();
                    const uint32_t a_hi = umma::smem_u32(s_A + sl * A_SLOT), a_lo = a_hi + A_HALF;
                    const uint32_t bb = umma::smem_u32(s_W + sl * WSLAB);
                    constexpr uint32_t lbo_b = 2 * COUT * 16;
                    uint64_t dah = umma::make_desc(a_hi, A_LBO, 128), dal = umma::make_desc(a_lo, A_LBO, 128);
                    uint64_t db = umma::make_desc(bb, lbo_b, 128);
#pragma unroll
                    for (int ks = 0; ks < SLAB_K / 8; ++ks) {
                        umma::mma_tf32(tmem, dah, db, idesc1, (first && ks == 0) ? 0u : 1u);
                        umma::mma_tf32(tmem, dal, db, idesc2, 1u);
                        dah += (uint64_t)((2 * A_LBO) >> 4); dal += (uint64_t)((2 * A_LBO) >> 4); db += (uint64_t)((2 * lbo_b) >> 4);
                    }
                    umma::commit(&a_free[sl]);
                    umma::commit(&w_free[sl]);
                    if (pass == NPASS - 1 && s == NSLAB - 1) umma::commit(&acc_full);
                    ++next_mma;
                    progressed = true;
                }
            }
            // ---- weight slab ring (2 deep) ----
            if (next_wl < total_slabs && next_wl < next_mma + 2) {
                const uint32_t sl = next_wl & 1;
                if (next_wl < 2 || __all_sync(0xffffffffu, umma::mbar_test(&w_free[sl], ((next_wl >> 1) - 1) & 1))) {
                    const uint32_t j = next_wl / NSLAB, s = next_wl % NSLAB;
                    umma::bulk_load(s_W + sl * WSLAB, reinterpret_cast<const unsigned char*>(Wc) + (size_t)((j % NPASS) * NSLAB + s) * WSLAB,
                                    WSLAB, &w_full[sl]);
                    ++next_wl;
                    progressed = true;
                }
            }
            // ---- per-point geometry (neighbour ids + relative positions), up to two points ahead of the feature stream ----
            {
                const uint32_t fpt = next_f < total_chunks ? (next_f / NCHUNK) / NPASS : (uint32_t)npts;
                if (next_g < (uint32_t)npts && next_g <= fpt + 2) {
                    const uint32_t gsl = next_g % GRING;
                    const size_t off = ((size_t)blockIdx.x + (size_t)next_g * gridDim.x) * NN;   // (b * P + p) * NN
                    const uint32_t bar = umma::smem_u32(&g_full[gsl]);
                    if (umma::elect_one()) {
                        umma::mbar_expect_tx(bar, NN * 20);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(umma::smem_u32(s_g + gsl * NN)), "l"(g4 + off), "r"(NN * 16), "r"(bar) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(umma::smem_u32(s_nbr + gsl * NN)), "l"(nbr + off), "r"(NN * 4), "r"(bar) : "memory");
                    }
                    __syncwarp();
                    ++next_g;
                    progressed = true;
                }
            }
            // ---- neighbour feature chunks (RING deep) ----
            if (next_f < total_chunks) {
                const uint32_t sl = next_f % RING, it = next_f / NCHUNK, c = next_f % NCHUNK;
                const uint32_t pass = it % NPASS, pt = it / NPASS;
                bool ready = pt < next_g && umma::mbar_test(&g_full[pt % GRING], (pt / GRING) & 1);
                if (ready && next_f >= RING) ready = umma::mbar_test(&f_free[sl], ((next_f / RING) - 1) & 1);
                ready = __all_sync(0xffffffffu, ready);
                if (ready) {
                    const int* nb = s_nbr + (pt % GRING) * NN + c * NB;
                    const uint32_t bar = umma::smem_u32(&f_full[sl]);
                    const uint32_t dst = umma::smem_u32(s_f + sl * NB * NBR_SLOT);
                    int rows[NB];
                    const int b = (int)(((size_t)blockIdx.x + (size_t)pt * gridDim.x) / (size_t)P);
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) rows[jj] = (b * q + nb[jj]) * NA;
                    if (umma::elect_one()) {
                        umma::mbar_expect_tx(bar, NB * NBR_TX);
#pragma unroll
                        for (int jj = 0; jj < NB; ++jj) umma::tma_load_2d(dst + jj * NBR_SLOT, &tmap, (int)pass * 32, rows[jj], bar);
                    }
                    __syncwarp();
                    ++next_f;
                    progressed = true;
                }
            }
            if (!progressed) __nanosleep(V3_SLEEP);   // do not steal issue slots from the compute warps of this scheduler
        }
    } else {
        // =========================== compute warps ===========================
        const int t = tid;                               // 0..479
        const int h = t / 240, rem = t - h * 240, a = rem >> 2, o = rem & 3;
        const uint32_t offA = (uint32_t)(a * 128 + (((2 * o) ^ (a & 7)) << 4));
        const uint32_t offB = (uint32_t)(a * 128 + (((2 * o + 1) ^ (a & 7)) << 4));
        const uint32_t w_off = (uint32_t)(a * NK + h * 12);                  // floats
        const uint32_t tlane = (uint32_t)((wp & 3) * 32) << 16;
        const uint32_t park = tmem + tlane + PARK_COL + (uint32_t)(wp >> 2) * 96;
        const uint32_t a_st = (uint32_t)((3 * o) * A_LBO + a * 16);
        float acc[8][12];
        uint32_t gc = 0, gs = 0;
        const uint32_t total_chunks = NIT * NCHUNK;

        auto wgen = [&](uint32_t chunk) {
            const uint32_t it2 = chunk / NCHUNK, c2 = chunk % NCHUNK, pt2 = it2 / NPASS;
            if (c2 == 0 && (it2 % NPASS) == 0) umma::mbar_wait(&g_full[pt2 % GRING], (pt2 / GRING) & 1);
            const float4* gsrc = s_g + (pt2 % GRING) * NN + c2 * NB;
            float* wdst = s_w + (chunk & 1) * (NB * NPAIRS);
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int pr = t + u * CT;
                const float4 kq = s_krs[pr];
                const uint64_t kx = v3_pk(kq.x, kq.x), ky = v3_pk(kq.y, kq.y), kz = v3_pk(kq.z, kq.z), kw = v3_pk(-kq.w, -kq.w);
#pragma unroll
                for (int jp = 0; jp < NB / 2; ++jp) {
                    const ulonglong2 ga = *reinterpret_cast<const ulonglong2*>(gsrc + 2 * jp);       // {gx0,gx1}, {gy0,gy1}
                    const ulonglong2 gb = *reinterpret_cast<const ulonglong2*>(gsrc + 2 * jp + 1);   // {gz0,gz1}, {gw0,gw1}
                    // the scalar expression fmaf(g.x, kq.x, fmaf(g.y, kq.y, fmaf(g.z, kq.z, g.w - kq.w))) for two neighbours at once
                    uint64_t v = v3_add2(gb.y, kw);
                    v = v3_fma2(gb.x, kz, v); v = v3_fma2(ga.y, ky, v); v = v3_fma2(ga.x, kx, v);
                    float w0, w1;
                    v3_unpk(v, w0, w1);
                    wdst[(2 * jp) * NPAIRS + pr] = fmaxf(w0, 0.f);
                    wdst[(2 * jp + 1) * NPAIRS + pr] = fmaxf(w1, 0.f);
                }
            }
        };

        auto slab_step = [&](uint32_t s) {
            const uint32_t cc = s >> 1, hh = s & 1;
            const bool mine = hh == 0 ? (wp <= 7) : (wp >= 7);
            if (mine) {
                const uint32_t sl = gs & 1;
                if (gs >= 2) umma::mbar_wait(&a_free[sl], ((gs >> 1) - 1) & 1);
                float v[12];
                umma::tmem_ld4_nowait(park + cc * 12, v);
                umma::tmem_ld4_nowait(park + cc * 12 + 4, v + 4);
                umma::tmem_ld4_nowait(park + cc * 12 + 8, v + 8);
                umma::tmem_wait_ld();
                if ((uint32_t)h == hh) {
                    unsigned char* dh = s_A + sl * A_SLOT + a_st;
#pragma unroll
                    for (int j4 = 0; j4 < 3; ++j4) {
                        float4 hi, lo;
                        umma::split_tf32(v[j4 * 4 + 0], hi.x, lo.x); umma::split_tf32(v[j4 * 4 + 1], hi.y, lo.y);
                        umma::split_tf32(v[j4 * 4 + 2], hi.z, lo.z); umma::split_tf32(v[j4 * 4 + 3], hi.w, lo.w);
                        *reinterpret_cast<float4*>(dh + j4 * A_LBO) = hi;
                        *reinterpret_cast<float4*>(dh + A_HALF + j4 * A_LBO) = lo;
                    }
                }
                umma::fence_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&a_full[sl]);
            }
            ++gs;
        };

        // accumulator of local point `pi` -> +bias -> global memory and InstanceNorm statistics (warps 0-3)
        auto epilogue = [&](uint32_t pi) {
            umma::mbar_wait(&acc_full, pi & 1);
            umma::fence_after_sync();
            const int row = wp * 16 + lane;
            const bool ok = lane < 16 && row < NA;
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 8) {
                float v[8], u[8];
                umma::tmem_ld8(tmem + tlane + c0, v);
                umma::tmem_ld8(tmem + tlane + COUT + c0, u);
                if (ok) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
                    *reinterpret_cast<float4*>(s_z + row * COUT + (((c0 >> 2) ^ (row & 7)) << 2)) = make_float4(v[0] + u[0] + b0.x, v[1] + u[1] + b0.y, v[2] + u[2] + b0.z, v[3] + u[3] + b0.w);
                    *reinterpret_cast<float4*>(s_z + row * COUT + ((((c0 >> 2) + 1) ^ (row & 7)) << 2)) = make_float4(v[4] + u[4] + b1.x, v[5] + u[5] + b1.y, v[6] + u[6] + b1.z, v[7] + u[7] + b1.w);
                }
            }
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&d_free);
            bar_sync_named(2, 128);
            const size_t gp = (size_t)blockIdx.x + (size_t)pi * gridDim.x, b = gp / (size_t)P;
            float4* dst = reinterpret_cast<float4*>(zraw + gp * NA * COUT);
            for (int i = t; i < NA * COUT / 4; i += 128) {   // the tile is XOR-swizzled by (row & 7): its row stride is a multiple of 128 B
                const int r = i / (COUT / 4), ch = i % (COUT / 4);
                dst[i] = reinterpret_cast<const float4*>(s_z)[r * (COUT / 4) + (ch ^ (r & 7))];
            }
            if (t < COUT) {
                float s = 0.f, ss = 0.f;
                for (int r = 0; r < NA; ++r) { const float x = s_z[r * COUT + ((((t >> 2) ^ (r & 7)) << 2) | (t & 3))]; s += x; ss = fmaf(x, x, ss); }
                s_stat[2 * t] += (double)s; s_stat[2 * t + 1] += (double)ss;
                if (pi + 1 == (uint32_t)npts || (gp + gridDim.x) / (size_t)P != b) {   // last point of this scan for this CTA: flush its statistics
                    atomicAdd(stats + (b * COUT + t) * 2, s_stat[2 * t]);
                    atomicAdd(stats + (b * COUT + t) * 2 + 1, s_stat[2 * t + 1]);
                    s_stat[2 * t] = 0.0; s_stat[2 * t + 1] = 0.0;
                }
            }
            bar_sync_named(2, 128);
        };

        if (NIT > 0) wgen(0);
        bar_sync_named(1, CT);                       // weight tile of chunk 0 complete
        for (uint32_t it = 0; it <= NIT; ++it) {
            const bool work = it < NIT, parked = it >= 1;
            if (parked && wp < 4) {
                const uint32_t j = it - 1;
                if (j % NPASS == 0 && j / NPASS >= 1) epilogue(j / NPASS - 1);
            }
            if (work) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int i = 0; i < 12; ++i) acc[c][i] = 0.f;
            }
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                // slab work first: it depends on nobody, so a warp that left the previous chunk early spends its skew here
                if (parked && (c % SLAB_EVERY) == 0) {
#pragma unroll
                    for (int sg = 0; sg < SLAB_GROUP; ++sg) slab_step((uint32_t)(c / SLAB_EVERY) * SLAB_GROUP + sg);
                }
                if (work) {
                    // every warp has left chunk gc-1 (arrive/wait split instead of a CTA barrier): its weight tile may be
                    // overwritten, and the tile of chunk gc, written before that, is complete
                    if (gc >= 1) umma::mbar_wait(&c_done, (gc - 1) & 1);
                    if (gc + 1 < total_chunks) wgen(gc + 1);
                    const uint32_t sl = gc % RING;
                    umma::mbar_wait(&f_full[sl], (gc / RING) & 1);
                    const unsigned char* fs = s_f + sl * NB * NBR_SLOT;
                    const float* ws = s_w + (gc & 1) * (NB * NPAIRS) + w_off;
#pragma unroll(V3_JJ)            // 1, not NB: ptxas renames the 96 accumulators across an unrolled body and pays ~60 MOVs per chunk
                    for (int jj = 0; jj < NB; ++jj) {
                        const float4 fa = *reinterpret_cast<const float4*>(fs + jj * NBR_SLOT + offA);
                        const float4 fb = *reinterpret_cast<const float4*>(fs + jj * NBR_SLOT + offB);
                        const float4 w0 = *reinterpret_cast<const float4*>(ws + jj * NPAIRS);
                        const float4 w1 = *reinterpret_cast<const float4*>(ws + jj * NPAIRS + 4);
                        const float4 w2 = *reinterpret_cast<const float4*>(ws + jj * NPAIRS + 8);
                        const float fv[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
                        const float wv[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
                        for (int cc = 0; cc < 8; ++cc)
#pragma unroll
                            for (int i = 0; i < 12; ++i) acc[cc][i] = fmaf(fv[cc], wv[i], acc[cc][i]);
                    }
                    __syncwarp();
                    if (lane == 0) { umma::mbar_arrive(&f_free[sl]); umma::mbar_arrive(&c_done); }
                    ++gc;
                }
            }
            if (work) {
                const float* av = &acc[0][0];
                umma::tmem_st32(park, av);
                umma::tmem_st32(park + 32, av + 32);
                umma::tmem_st32(park + 64, av + 64);
                umma::tmem_wait_st();
            }
        }
        if (npts > 0 && wp < 4) epilogue((uint32_t)npts - 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (wp == 15) umma::tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <int CIN, int COUT, int NN>
int launch_inter_v3(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs, const float* Wc,
                    const float* bias, int B, int q, int P, float sigma, float* g4, float* zraw, double* stats, cudaStream_t stream) {
    using Cfg = V3Cfg<CIN, COUT, NN>;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)CIN, (cuuint64_t)B * q * NA};
    const cuuint64_t gstr[1] = {(cuuint64_t)CIN * 4};
    const cuuint32_t box[2] = {32, NA};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return (int)cudaErrorInvalidValue;
    {
        dim3 grid((unsigned)etch_cdiv(P * NN, 256), (unsigned)B);
        inter_geom_kernel<<<grid, 256, 0, stream>>>(xyz, sample_idx, nbr, q, P, NN, 1.0f / sigma, reinterpret_cast<float4*>(g4));
    }
    auto kern = inter_conv_v3_kernel<CIN, COUT, NN>;
    ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
    int g = etch_sm_budget();
    if (g > P) g = P;
    dim3 grid((unsigned)g);
    kern<<<grid, NTHREADS, Cfg::smem, stream>>>(tmap, reinterpret_cast<const float4*>(g4), nbr, reinterpret_cast<const float4*>(krs), Wc, bias,
                                                B, q, P, zraw, stats);
    ETCH_RETURN_LAST();
}

}  // namespace

// InterSO3Conv (c_in in {32,64}), one point per tile.  Wc = [cin/32*16][12][2*cout][4]: per (pass, channel slot cc, kernel-point
// half hh) the 48-column slab  K'' = o*12 + i  <->  W[.][(32*pass + 8*o + cc)*24 + 12*hh + i], rows [W_hi; W_lo], canonical tiles
// (etch_b200/models/encoder.py::_inter_slabs_v3).  g4 = caller-owned scratch [B,P,nn,4] fp32.
ETCH_API int etch_so3_inter_conv_v3(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs,
                                    const float* Wc, const float* bias, int B, int q, int P, int nn, int cin, int cout, float sigma,
                                    float* g4, float* zraw, double* stats, cudaStream_t stream) {
    if (!xyz || !feat || !sample_idx || !nbr || !krs || !Wc || !bias || !g4 || !zraw || !stats || B <= 0 || P <= 0) return ETCH_EINVAL;
#define CASE(ci, co, n) \
    if (cin == ci && cout == co && nn == n) return launch_inter_v3<ci, co, n>(xyz, feat, sample_idx, nbr, krs, Wc, bias, B, q, P, sigma, g4, zraw, stats, stream);
    CASE(32, 32, 32) CASE(32, 64, 64) CASE(64, 64, 32)
#undef CASE
    return ETCH_EINVAL;
}
