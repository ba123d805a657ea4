// Dense vertex forward on the 5th-generation tensor cores (sm_100a, tcgen05 + TMEM + TMA).
//
//   pose_offsets[b, 3n+c] = sum_k  pose_feature[b,k] * posedirs[k, 3n+c]        (lbs.py:192-195)
//
// is the one dense contraction of the path: M = frames in flight, N = 20670, K = 207.  It is skinny in M
// (<= a few hundred frames), so it is bound by streaming the 17 MB of posedirs from L2/HBM once per 128-frame
// tile, not by tensor throughput; the kernel is therefore organised around keeping the A operand (the frames'
// pose features, 112 KB; only the live 8-row groups are fetched when few frames are active) RESIDENT in shared memory
// for the CTA's whole life while posedirs tiles stream through a TMA ring:
//
//   warp 0   TMA producer   A once (7 boxes 128x32, or 8-row boxes), then B boxes 96x32 through a 6-stage mbarrier ring
//   warp 1   MMA issuer     tcgen05.mma.cta_group::1.kind::tf32  M=128 N=96 K=8, fp32 accumulators in TMEM,
//                           two accumulator buffers (2 x 96 columns) so tile i+1 runs under the epilogue of tile i
//   warp 2   TMEM alloc / dealloc
//   warps 4-7  epilogue     tcgen05.ld (lane = frame, 96 columns = 32 vertices x 3), pose offsets stored TRANSPOSED
//                           (frame fastest) so that a warp's 32 frames form one 128-byte line per column
//
// The rest of the vertex forward (template + shape blend shapes in fp32, linear blend skinning, box partials) is
// skin_kernel / skin_small_kernel below: fused into this epilogue it ran in 128 threads and cost 10x the contraction.
//
// Precision: operands are pre-rounded to TF32 (cvt.rna), accumulation is fp32.  Pose offsets are a small
// correction (<= a few % of a vertex coordinate), so their 2^-11 relative rounding stays < 1e-5 relative on the
// vertices -- an order of magnitude inside the 1e-4 parity bar (tests/test_gpu_tc.py checks against the fp32
// SIMT kernel and the oracle).  The template and the shape blend shapes are NOT sent through TF32.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include "mvs_internal.cuh"

namespace mvs {

constexpr int kTcBM = 128;            // frames per CTA (UMMA M)
constexpr int kTcBN = kTileC;         // 96 columns = 32 vertices (UMMA N)
constexpr int kTcBK = 32;             // floats per 128-byte swizzle row
constexpr int kTcKCh = kFeatPad / kTcBK;   // 7 K chunks
constexpr int kTcStages = 6;
constexpr int kTcABytes = kTcBM * kTcBK * 4;   // 16 KB
constexpr int kTcBBytes = kTcBN * kTcBK * 4;   // 12 KB
constexpr int kTcThreads = 256;
constexpr size_t kTcSmem = 1024 /*align slack*/ + (size_t)kTcKCh * kTcABytes + (size_t)kTcStages * kTcBBytes + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol bug must not hang the GPU -- after ~2 s the kernel raises the error flag and carries on
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
    uint32_t done = 0;
    for (long long it = 0; it < (1ll << 26); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    if (err_flag) atomicExch(err_flag, 1);
    return false;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, 128-byte swizzle, rows of 128 B packed 8 per 1024 B: LBO = 1 (ignored), SBO = 1024 B, version 1
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 96
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

#ifdef MVS_PHASE_DBG
__device__ long long g_tc_clk[32];      // clock64 stamps of CTA (0,0): see scripts/phase_times.py
#define TC_MARK(i) do { if (blockIdx.x == 0 && blockIdx.y == 0) g_tc_clk[i] = clock64(); } while (0)
#else
#define TC_MARK(i) do {} while (0)
#endif

__global__ void __launch_bounds__(kTcThreads, 1)
posedirs_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a8,
                        const __grid_constant__ CUtensorMap map_b, int ldA, int ncols,
                        const int* __restrict__ na_ptr, int cta_slots, int ntiles, float* __restrict__ poffT,
                        int* __restrict__ err_flag) {
    pdl_wait();
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;                                        // 7 x [128 rows][128 B], 1024-aligned
    unsigned char* sB = sA + (size_t)kTcKCh * kTcABytes;             // 6 x [96 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)kTcStages * kTcBBytes);
    uint64_t* a_full = bars;                      // [1]
    uint64_t* b_full = bars + 1;                  // [stages]
    uint64_t* b_empty = b_full + kTcStages;       // [stages]
    uint64_t* t_full = b_empty + kTcStages;       // [2]
    uint64_t* t_empty = t_full + 2;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int na = *na_ptr;
    const int m0 = blockIdx.y * kTcBM;
    if (m0 >= na) return;
    // the launch provides cta_slots CTAs per M tile of the FULL batch; with fewer active M tiles the survivors
    // spread the vertex tiles over all of them (the tail of a stage runs one M tile on every SM)
    const int active_m = (na + kTcBM - 1) / kTcBM;
    const int ctas_n = max(1, min((int)gridDim.x, cta_slots / active_m));
    if ((int)blockIdx.x >= ctas_n) return;
    const int tiles_per_cta = (ntiles + ctas_n - 1) / ctas_n;
    const int tile_begin = blockIdx.x * tiles_per_cta;
    const int tile_end = min(tile_begin + tiles_per_cta, ntiles);
    if (tile_begin >= tile_end) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) TC_MARK(0);

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_MARK(1);

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer
            // A operand.  Every CTA of the M tile reads it, so with 148 CTAs the 112 KB tile costs as much L2 traffic
            // as posedirs itself; at the tail of a fit only a few of its 128 rows are live, and those are fetched as
            // 8-row boxes (one swizzle atom each: same shared-memory image).  Rows that are not loaded hold whatever
            // was there: their accumulator rows are never read.
            const int live_rows = min(kTcBM, na - m0);
            if (live_rows <= 32) {
                const int groups = (live_rows + 7) / 8;
                mbar_expect_tx(a_full, kTcKCh * groups * 8 * kTcBK * 4);
                for (int kc = 0; kc < kTcKCh; ++kc)
                    for (int gq = 0; gq < groups; ++gq)
                        tma_load_2d(sA + (size_t)kc * kTcABytes + (size_t)gq * 1024, &map_a8, kc * kTcBK, m0 + 8 * gq, a_full);
            } else {
                mbar_expect_tx(a_full, kTcKCh * kTcABytes);
                for (int kc = 0; kc < kTcKCh; ++kc) tma_load_2d(sA + (size_t)kc * kTcABytes, &map_a, kc * kTcBK, m0, a_full);
            }
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                for (int kc = 0; kc < kTcKCh; ++kc) {
                    if (!mbar_wait(&b_empty[stage], phase ^ 1, err_flag)) return;
                    mbar_expect_tx(&b_full[stage], kTcBBytes);
                    tma_load_2d(sB + (size_t)stage * kTcBBytes, &map_b, kc * kTcBK, tile * kTcBN, &b_full[stage]);
                    if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer
            constexpr uint32_t idesc = umma_idesc_tf32(kTcBM, kTcBN);
            if (!mbar_wait(a_full, 0, err_flag)) return;
            TC_MARK(2);
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t tphase[2] = {0, 0};
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                if (!mbar_wait(&t_empty[buf], tphase[buf] ^ 1, err_flag)) return;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kTcBN);
                for (int kc = 0; kc < kTcKCh; ++kc) {
                    if (!mbar_wait(&b_full[stage], phase, err_flag)) return;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(sA + (size_t)kc * kTcABytes);
                    const uint32_t b_addr = smem_u32(sB + (size_t)stage * kTcBBytes);
#pragma unroll
                    for (int k = 0; k < kTcBK / 8; ++k)
                        umma_tf32(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                                  (kc | k) ? 1u : 0u);
                    umma_commit(&b_empty[stage]);                    // frees the B slot when these MMAs retire
                    if (++stage == kTcStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&t_full[buf]);                           // accumulator of this tile complete
                TC_MARK(3 + (tile - tile_begin));
                tphase[buf] ^= 1;
                buf ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: lane = frame row of the M tile; pose offsets go out TRANSPOSED
        //                  poffT[column][slot] so that a warp's 32 frames form one 128-byte store
        const int row = 32 * (warp & 3) + lane;
        const int slot = m0 + row;
        const bool live = slot < na;
        int buf = 0; uint32_t tphase[2] = {0, 0};
        for (int tile = tile_begin; tile < tile_end; ++tile) {
            if (!mbar_wait(&t_full[buf], tphase[buf], err_flag)) return;
            if (threadIdx.x == 128) TC_MARK(8 + 2 * (tile - tile_begin));
            tphase[buf] ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(buf * kTcBN);
            uint32_t acc[kTcBN];
            tmem_ld32(taddr, acc);
            tmem_ld32(taddr + 32, acc + 32);
            tmem_ld32(taddr + 64, acc + 64);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&t_empty[buf]);                              // accumulator copied out: the MMA warp may reuse it
            buf ^= 1;
            if (live) {
                const int c0 = tile * kTcBN;
#pragma unroll
                for (int c = 0; c < kTcBN; ++c)
                    if (c0 + c < ncols) poffT[(size_t)(c0 + c) * ldA + slot] = __uint_as_float(acc[c]);
            }
            if (threadIdx.x == 128) TC_MARK(9 + 2 * (tile - tile_begin));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    if (threadIdx.x == 0) TC_MARK(20);
}

// One vertex of one frame: v_posed = template + shapedirs.betas + pose offset, then linear blend skinning
// (lbs.py:179,203,207-220).  skin_kernel (lane = frame) and skin_small_kernel (lane = vertex) both call this with every
// operation spelled out, so the two produce the same bits: a frame's result must not depend on which of the two
// kernels its batch size selected.  A(j, a12) loads the 12 entries (row-major 3x4) of joint j's transform for this frame.
template <class AFn>
__device__ __forceinline__ void skin_vertex(const float* __restrict__ st /* [3][11]: shapedirs row | template */,
                                            const float* beta, const float* poff, const int* __restrict__ ell_j,
                                            const float* __restrict__ ell_w, int KW, size_t n, AFn A, float* vp, float* vv) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float a = st[11 * c + kBetas];
#pragma unroll
        for (int l = 0; l < kBetas; ++l) a = fmaf(st[11 * c + l], beta[l], a);
        vp[c] = __fadd_rn(a, poff[c]);
    }
    float T[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) T[c] = 0.f;
    // padding entries have weight 0 and a valid joint: fma(0, a, T) == T exactly, so no branch (and no dependent
    // load behind it) is needed
#pragma unroll 4
    for (int e = 0; e < KW; ++e) {
        const float w = ell_w[n * KW + e];
        float a12[12];
        A(ell_j[n * KW + e], a12);
#pragma unroll
        for (int c = 0; c < 12; ++c) T[c] = fmaf(w, a12[c], T[c]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
        vv[r] = __fadd_rn(fmaf(T[4 * r + 2], vp[2], fmaf(T[4 * r + 1], vp[1], __fmul_rn(T[4 * r], vp[0]))), T[4 * r + 3]);
}

// ------------------------------------------------------------------------------------------------ skinning
// v_posed = v_template + shapedirs.betas (fp32) + pose offsets (tensor cores), then linear blend skinning
// (lbs.py:179,203,207-220).  CTA = 32 frames x 64 vertices; lane = frame.  The 32 frames' skinning transforms
// (288 floats each) sit in shared memory frame-fastest, so the 48 reads per (frame, vertex) are conflict-free LDS
// instead of global loads; outputs are transposed through shared memory into coalesced row stores.  Also emits the
// per-chunk bounding-box partial of every frame for the SDF kernels.
constexpr int kSkinV = 64;
constexpr int kSkinThreads = 256;
constexpr int kSkinOutLd = 3 * kSkinV + 1;     // 193
constexpr int kSkinEllMax = 8;                 // skinning weights per vertex staged in shared memory (SMPL: 4)
constexpr size_t kSkinSmem =
    (size_t)(kSkinFloats * 32 + kBetas * 32 + 2 * 32 * kSkinOutLd + 8 * 32 * 6 * 2 + kSkinV * 33 + 2 * kSkinV * kSkinEllMax) * sizeof(float);

__global__ void __launch_bounds__(kSkinThreads)
skin_kernel(const float* __restrict__ poffT, const float* __restrict__ ST, const float* __restrict__ Phi,
            const float* __restrict__ At, int ldA, const int* __restrict__ ell_j, const float* __restrict__ ell_w, int KW,
            int N, const int* __restrict__ na_ptr, int cta_slots, float* __restrict__ vposed, float* __restrict__ verts,
            float* __restrict__ bboxp) {
    pdl_wait();
    extern __shared__ __align__(16) float sk[];
    float* As = sk;                                   // [24 joints][32 lanes][12]
    float* Bs = As + kSkinFloats * 32;                // [10][32]
    float* Ovp = Bs + kBetas * 32;                    // [32][193]
    float* Ov = Ovp + 32 * kSkinOutLd;                // [32][193]
    float* Bb = Ov + 32 * kSkinOutLd;                 // [8 warps][32 lanes][6] values, then [..][6] indices
    float* Sts = Bb + 8 * 32 * 6 * 2;                 // [64][33] shapedirs rows | template of this CTA's vertices
    float* Ews = Sts + kSkinV * 33;                   // [64][KW] skinning weights of the chunk
    int* Ejs = reinterpret_cast<int*>(Ews + kSkinV * kSkinEllMax);   // [64][KW] their joints
    const bool ell_smem = KW <= kSkinEllMax;
    const int na = *na_ptr;
    const int f0 = blockIdx.y * 32;
    if (f0 >= na) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slotc = min(f0 + lane, na - 1);
    // chunks per CTA from the number of ACTIVE frame groups: one chunk each while the active CTAs fit in one wave
    // (latency at the tail of a stage), several once they do not (the 36 KB of transforms are then loaded once)
    const int nchunks = (N + kSkinV - 1) / kSkinV;
    const int per_cta = (nchunks * ((na + 31) / 32) + cta_slots - 1) / cta_slots;
    const int nx = (nchunks + per_cta - 1) / per_cta;
    if ((int)blockIdx.x >= nx) return;
    // the 32 frames' transforms and shape coefficients are loaded once and reused for every vertex chunk of this CTA
    // layout [joint][lane][12]: a lane's 3x4 transform is three conflict-free LDS.128 (lane stride 48 B)
    for (int e = tid; e < kSkinFloats * 32; e += kSkinThreads) {
        const int jc = e >> 5, ln = e & 31;
        As[((jc / 12) * 32 + ln) * 12 + jc % 12] = At[(size_t)jc * ldA + min(f0 + ln, na - 1)];
    }
    for (int e = tid; e < kBetas * 32; e += kSkinThreads)
        Bs[e] = Phi[(size_t)min(f0 + (e & 31), na - 1) * kFeatPad + kPoseBasis + (e >> 5)];
    constexpr int kVw = kSkinV / 8;                                       // vertices per warp and chunk
    float beta[kBetas];
    bool have_beta = false;
    for (int ch = blockIdx.x; ch < nchunks; ch += nx) {
        const int v0 = ch * kSkinV;
        // pose offsets of this warp's vertices: 24 independent coalesced loads in flight before anything waits
        float pf[kVw][3];
#pragma unroll
        for (int i = 0; i < kVw; ++i) {
            const int n = v0 + warp * kVw + i;
#pragma unroll
            for (int c = 0; c < 3; ++c) pf[i][c] = n < N ? poffT[(size_t)(3 * n + c) * ldA + slotc] : 0.f;
        }
        __syncthreads();                                                  // previous chunk's staging buffers are free
        for (int e = tid; e < kSkinV * 33; e += kSkinThreads) Sts[e] = (v0 * 33 + e < N * 33) ? ST[(size_t)v0 * 33 + e] : 0.f;
        if (ell_smem)
            for (int e = tid; e < kSkinV * KW; e += kSkinThreads) {
                const bool in = (size_t)v0 * KW + e < (size_t)N * KW;
                Ews[e] = in ? ell_w[(size_t)v0 * KW + e] : 0.f;
                Ejs[e] = in ? ell_j[(size_t)v0 * KW + e] : 0;
            }
        __syncthreads();
        if (!have_beta) {
#pragma unroll
            for (int l = 0; l < kBetas; ++l) beta[l] = Bs[l * 32 + lane];
            have_beta = true;
        }
        float blo[3] = {3e38f, 3e38f, 3e38f}, bhi[3] = {-3e38f, -3e38f, -3e38f};
        int bilo[3] = {0, 0, 0}, bihi[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < kVw; ++i) {
            const int li = warp * kVw + i;
            const int n = v0 + li;
            if (n < N) {
                float vp[3], vvv[3];
                skin_vertex(Sts + li * 33, beta, pf[i], ell_smem ? Ejs : ell_j, ell_smem ? Ews : ell_w, KW,
                            ell_smem ? (size_t)li : (size_t)n,
                            [&](int j, float* a12) {
                                const float4* q = reinterpret_cast<const float4*>(As + (j * 32 + lane) * 12);
                                const float4 q0 = q[0], q1 = q[1], q2 = q[2];
                                a12[0] = q0.x; a12[1] = q0.y; a12[2] = q0.z; a12[3] = q0.w; a12[4] = q1.x; a12[5] = q1.y;
                                a12[6] = q1.z; a12[7] = q1.w; a12[8] = q2.x; a12[9] = q2.y; a12[10] = q2.z; a12[11] = q2.w;
                            }, vp, vvv);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float vv = vvv[r];
                    Ovp[lane * kSkinOutLd + 3 * li + r] = vp[r];
                    Ov[lane * kSkinOutLd + 3 * li + r] = vv;
                    if (vv < blo[r]) { blo[r] = vv; bilo[r] = n; }       // strict: ties keep the lowest vertex index
                    if (vv > bhi[r]) { bhi[r] = vv; bihi[r] = n; }
                }
            }
        }
        int* Bi = reinterpret_cast<int*>(Bb + 8 * 32 * 6);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Bb[(warp * 32 + lane) * 6 + r] = blo[r]; Bb[(warp * 32 + lane) * 6 + 3 + r] = bhi[r];
            Bi[(warp * 32 + lane) * 6 + r] = bilo[r]; Bi[(warp * 32 + lane) * 6 + 3 + r] = bihi[r];
        }
        __syncthreads();
        const int ncol = 3 * min(kSkinV, N - v0);                         // contiguous floats of this chunk per frame
        for (int fl = warp; fl < 32; fl += 8) {                           // a warp stores one frame's row segment at a time
            const int slot = f0 + fl;
            if (slot >= na) break;
            const size_t off = ((size_t)slot * N + v0) * 3;
            for (int col = lane; col < ncol; col += 32) {
                vposed[off + col] = Ovp[fl * kSkinOutLd + col];
                verts[off + col] = Ov[fl * kSkinOutLd + col];
            }
        }
        if (bboxp && warp == 0 && f0 + lane < na) {       // fold the 8 warps' vertex groups (ascending vertex index)
            float lo[3], hi[3];
            int ilo[3], ihi[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) { lo[r] = Bb[lane * 6 + r]; hi[r] = Bb[lane * 6 + 3 + r]; ilo[r] = Bi[lane * 6 + r]; ihi[r] = Bi[lane * 6 + 3 + r]; }
            for (int w2 = 1; w2 < 8; ++w2)
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float l2 = Bb[(w2 * 32 + lane) * 6 + r], h2 = Bb[(w2 * 32 + lane) * 6 + 3 + r];
                    if (l2 < lo[r]) { lo[r] = l2; ilo[r] = Bi[(w2 * 32 + lane) * 6 + r]; }
                    if (h2 > hi[r]) { hi[r] = h2; ihi[r] = Bi[(w2 * 32 + lane) * 6 + 3 + r]; }
                }
            float* bp = bboxp + ((size_t)(f0 + lane) * nchunks + ch) * 12;
#pragma unroll
            for (int r = 0; r < 3; ++r) { bp[r] = lo[r]; bp[3 + r] = hi[r]; bp[6 + r] = __int_as_float(ilo[r]); bp[9 + r] = __int_as_float(ihi[r]); }
        }
    }
}

// The same for a handful of frames (the straggler tail of a fit): with lane = frame almost every lane of skin_kernel
// idles and its fixed costs (36 KB of transforms per CTA, staging, three barriers per chunk) are pure latency.  Here
// lane = vertex, a CTA owns one 64-vertex chunk (= one box partial) and loops over the <= kSkinSmallMax frames.
constexpr int kSkinSmallMax = 8;
constexpr int kSkinSmallThreads = 64;
static_assert(kSkinSmallThreads == kSkinV, "one thread per vertex of a box chunk");

__global__ void __launch_bounds__(kSkinSmallThreads)
skin_small_kernel(const float* __restrict__ poffT, const float* __restrict__ ST, const float* __restrict__ Phi,
                  const float* __restrict__ At, int ldA, const int* __restrict__ ell_j, const float* __restrict__ ell_w, int KW,
                  int N, const int* __restrict__ na_ptr, float* __restrict__ vposed, float* __restrict__ verts,
                  float* __restrict__ bboxp) {
    __shared__ float As[kSkinSmallMax][kSkinFloats];
    __shared__ float Bs[kSkinSmallMax][kBetas];
    __shared__ float s_b[kSkinSmallMax][6];
    __shared__ int s_i[kSkinSmallMax][6];
    pdl_wait();
    const int na = min(*na_ptr, kSkinSmallMax);
    if (na <= 0) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int chunk = blockIdx.x, nchunks = gridDim.x;
    const int n = chunk * kSkinSmallThreads + t;
    for (int e = t; e < na * kSkinFloats; e += kSkinSmallThreads) As[e / kSkinFloats][e % kSkinFloats] = At[(size_t)(e % kSkinFloats) * ldA + e / kSkinFloats];
    for (int e = t; e < na * kBetas; e += kSkinSmallThreads) Bs[e / kBetas][e % kBetas] = Phi[(size_t)(e / kBetas) * kFeatPad + kPoseBasis + e % kBetas];
    float st[33];
    if (n < N) {
#pragma unroll
        for (int q = 0; q < 33; ++q) st[q] = ST[(size_t)n * 33 + q];
    }
    __syncthreads();
    for (int f = 0; f < na; ++f) {
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        int ilo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, ihi[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
        if (n < N) {
            float poff[3], vp[3], vv[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) poff[c] = poffT[(size_t)(3 * n + c) * ldA + f];
            skin_vertex(st, Bs[f], poff, ell_j, ell_w, KW, (size_t)n,
                        [&](int j, float* a12) {
#pragma unroll
                            for (int c = 0; c < 12; ++c) a12[c] = As[f][j * 12 + c];
                        }, vp, vv);
            const size_t off = ((size_t)f * N + n) * 3;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                vposed[off + r] = vp[r]; verts[off + r] = vv[r];
                lo[r] = vv[r]; hi[r] = vv[r]; ilo[r] = n; ihi[r] = n;
            }
        }
        if (bboxp) {               // box of the chunk: extreme value, ties -> lowest vertex index (as skin_kernel)
#pragma unroll
            for (int r = 0; r < 3; ++r) { warp_argmin(lo[r], ilo[r]); warp_argmax(hi[r], ihi[r]); }
            if (warp == 1 && lane == 0) {
#pragma unroll
                for (int r = 0; r < 3; ++r) { s_b[f][r] = lo[r]; s_b[f][3 + r] = hi[r]; s_i[f][r] = ilo[r]; s_i[f][3 + r] = ihi[r]; }
            }
            __syncthreads();
            if (warp == 0 && lane == 0) {
                float* bp = bboxp + ((size_t)f * nchunks + chunk) * 12;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float l2 = s_b[f][r], h2 = s_b[f][3 + r];
                    const int il2 = s_i[f][r], ih2 = s_i[f][3 + r];
                    if (l2 < lo[r] || (l2 == lo[r] && il2 < ilo[r])) { lo[r] = l2; ilo[r] = il2; }
                    if (h2 > hi[r] || (h2 == hi[r] && ih2 < ihi[r])) { hi[r] = h2; ihi[r] = ih2; }
                    bp[r] = lo[r]; bp[3 + r] = hi[r]; bp[6 + r] = __int_as_float(ilo[r]); bp[9 + r] = __int_as_float(ihi[r]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct TcState {
    CUtensorMap map_a, map_a8, map_b;
    float* Qtc = nullptr;       // [3N][224] TF32-rounded posedirs rows (columns >= 207 zero)
    float* poffT = nullptr;     // [3N][ldA] pose offsets, frame fastest (output of the tensor-core contraction)
    int* err = nullptr;
    bool ready = false;
};

static float round_tf32(float x) {                      // cvt.rna.tf32.f32: nearest, ties away from zero
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

static int encode_map(mvs_ctx* ctx, CUtensorMap* map, const float* ptr, uint64_t rows, uint32_t box_rows) {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            return set_error(ctx, MVS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    cuuint64_t dims[2] = {(cuuint64_t)kFeatPad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kFeatPad * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kTcBK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(ctx, MVS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MVS_OK;
}

// posedirs copy for the tensor path (called from mvs_set_model with the host Qk rows)
int tc_upload_model(mvs_ctx* ctx, const float* Qk_host) {
    TcState* T = new TcState();
    ctx->tc = T;
    const size_t n = (size_t)3 * ctx->m.N * kFeatPad;
    std::vector<float> q(n);
    for (size_t i = 0; i < n; ++i) q[i] = (i % kFeatPad) < (size_t)kPoseBasis ? round_tf32(Qk_host[i]) : 0.f;
    int rc = dev_upload(ctx, &T->Qtc, q.data(), n);
    if (rc) return rc;
    if ((rc = dev_alloc(ctx, &T->err, 1))) return rc;
    MVS_CUDA_OK(ctx, cudaMemset(T->err, 0, sizeof(int)));
    return MVS_OK;
}

bool tc_available(const mvs_ctx* ctx) {
    return ctx->exec_mode != 1 && ctx->tc != nullptr && ctx->ws.PhiTc != nullptr;
}

int launch_vertex_fwd_tc(mvs_ctx* ctx, cudaStream_t st) {
    TcState* T = static_cast<TcState*>(ctx->tc);
    Workspace& w = ctx->ws;
    const DevModel& m = ctx->m;
    if (!T->ready) {
        int rc;
        if ((rc = encode_map(ctx, &T->map_a, w.PhiTc, (uint64_t)w.ldA, kTcBM))) return rc;
        if ((rc = encode_map(ctx, &T->map_a8, w.PhiTc, (uint64_t)w.ldA, 8))) return rc;
        if ((rc = encode_map(ctx, &T->map_b, T->Qtc, (uint64_t)3 * m.N, kTcBN))) return rc;
        if ((rc = dev_alloc(ctx, &T->poffT, (size_t)3 * m.N * w.ldA))) return rc;
        MVS_CUDA_OK(ctx, cudaFuncSetAttribute(posedirs_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        MVS_CUDA_OK(ctx, cudaFuncSetAttribute(skin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSkinSmem));
        T->ready = true;
    }
    const int ntiles = (m.N + kTileV - 1) / kTileV;
    const int nb = w.na_bound > 0 ? w.na_bound : w.B;             // upper bound of the active-frame count
    const int mtiles = (nb + kTcBM - 1) / kTcBM;
    dim3 grid(std::min(ctx->sm_count, ntiles), mtiles);          // surplus CTAs exit: the kernel splits the tiles from *na
    MVS_LAUNCH(ctx, KID_VERTEX_FWD_TC, st,
               MVS_CUDA_OK(ctx, launch_pdl(posedirs_gemm_tc_kernel, grid, dim3(kTcThreads), kTcSmem, st, T->map_a, T->map_a8, T->map_b, w.ldA,
                                           3 * m.N, (const int*)w.na, ctx->sm_count, ntiles, T->poffT, T->err)));
    const int nchunks = (m.N + kSkinV - 1) / kSkinV, fgroups = (nb + 31) / 32;
    if (nb <= kSkinSmallMax) {                          // straggler tail: lane = vertex, one chunk per CTA
        MVS_LAUNCH(ctx, KID_SKIN, st,
                   MVS_CUDA_OK(ctx, launch_pdl(skin_small_kernel, dim3(nchunks), dim3(kSkinSmallThreads), 0, st,
                                               (const float*)T->poffT, (const float*)m.ST, (const float*)w.Phi, (const float*)w.At,
                                               w.ldA, (const int*)m.ell_j, (const float*)m.ell_w, m.KW, m.N, (const int*)w.na,
                                               w.vposed, w.verts, w.bboxp)));
    } else {
        // each CTA keeps its 32 frames' transforms in shared memory and walks 1..3 chunks (sized from *na in the kernel)
        dim3 g2(nchunks, fgroups);                      // surplus CTAs exit
        MVS_LAUNCH(ctx, KID_SKIN, st,
                   MVS_CUDA_OK(ctx, launch_pdl(skin_kernel, g2, dim3(kSkinThreads), kSkinSmem, st, (const float*)T->poffT,
                                               (const float*)m.ST, (const float*)w.Phi, (const float*)w.At, w.ldA,
                                               (const int*)m.ell_j, (const float*)m.ell_w, m.KW, m.N, (const int*)w.na,
                                               2 * ctx->sm_count, w.vposed, w.verts, w.bboxp)));
    }
    MVS_CUDA_OK(ctx, cudaGetLastError());
    return MVS_OK;
}

int tc_check_error(mvs_ctx* ctx) {          // host-side check after a synchronisation point
    TcState* T = static_cast<TcState*>(ctx->tc);
    if (!T || !T->ready) return MVS_OK;
    int e = 0;
    MVS_CUDA_OK(ctx, cudaMemcpy(&e, T->err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) return set_error(ctx, MVS_ERR_CUDA, "vertex_fwd_tc_kernel: mbarrier wait timed out (pipeline protocol error)");
    return MVS_OK;
}

}  // namespace mvs

#ifdef MVS_PHASE_DBG
extern "C" int mvs_debug_tc_clocks(long long* out, int n) {       // debug builds only; not part of the ABI
    long long h[32];
    if (cudaMemcpyFromSymbol(h, mvs::g_tc_clk, sizeof(h)) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 32; ++i) out[i] = h[i];
    return 0;
}
#endif
