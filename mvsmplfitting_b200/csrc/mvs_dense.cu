// Dense rounds as ONE persistent cooperative kernel (regime B of DESIGN.md section 4: stages with the SDF term on).
//
// The multi-kernel path runs a round as four dependent launches (posedirs_gemm_tc -> skin -> sdf_fused -> frame_step)
// enqueued by the host, plus a compaction every 8 rounds with a host read-back of the active count.  A fit is 500-600
// such rounds and more than half of them run with <= 30 of 256 frames active, so the step is rounds x (launch gaps +
// per-kernel prologues), and the host sits in the loop.  Here the whole run is one launch:
//
//   grid = (resident CTAs per SM) x SMs, launched cooperatively; every CTA loops over the rounds; inside a round the four
//   phases are the SAME device bodies the multi-kernel path runs (mvs_tc_dev.cuh / mvs_sdf_dev.cuh / mvs_resident_dev.cuh:
//   a frame's arithmetic is bit-identical in both paths), distributed over the CTAs as strided work items and separated
//   by grid barriers (one atomic + acquire spin each, ~1.5 us, instead of a kernel boundary);
//   the tcgen05 contraction keeps its TMEM allocation, its mbarrier ring and their phase bits across rounds; the A operand
//   (pose features of 128 frames) now streams through the ring with the posedirs tiles (A chunk 16 KB + B chunk 12 KB per
//   stage), which brings the phase's shared memory to 84 KB and lets TWO CTAs share an SM (296 frame workers);
//   compaction happens EVERY round and for free: a frame that needs another evaluation takes the next slot of the next
//   round's list with one atomicAdd (a frame's arithmetic does not depend on its slot); the kernel ends when a round
//   finds the list empty -- no host read-back, no launch between rounds.
//
// Data that other CTAs rewrite between rounds (pose offsets, vertices, feature rows, transforms, SDF partials, optimiser
// state) is read with plain loads after the grid barrier's gpu-scope acquire (which also invalidates L1); only model
// constants use the read-only path.  TMA reads the feature rows through the async proxy: writers and the producer thread
// bracket the barrier with fence.proxy.async.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "mvs_internal.cuh"
#ifdef MVS_PHASE_DBG
// debug build (libmvsmpl_dbg.so, scripts/dense_marks.py): thread 0 of CTA 0 accumulates the cycles between consecutive marks
namespace mvs {
__device__ long long g_dense_clk[64], g_dense_cnt[64], g_dense_last;
__device__ __forceinline__ void dense_mark(int i) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long c = clock64();
    g_dense_clk[i] += c - g_dense_last; g_dense_cnt[i] += 1; g_dense_last = c;
}
}
#define PHASE_MARK(i) mvs::dense_mark(i)
#define DMARK(i) mvs::dense_mark(i)
#else
#define DMARK(i) do {} while (0)
#endif
#include "mvs_lbfgs_core.cuh"
#include "mvs_tc_dev.cuh"
#include "mvs_sdf_dev.cuh"
#include "mvs_resident_dev.cuh"

namespace mvs {

constexpr int kDrThreads = 256;
constexpr int kDrStages = 3;                                  // ring stages of (A chunk | B chunk)
constexpr int kDrStageBytes = kTcABytes + kTcBBytes;          // 28 KB
constexpr int kDrPhases = 6;

struct DenseCtl {                     // device bookkeeping of one persistent run (zeroed by the host before the launch)
    unsigned bar;                     // grid-barrier arrivals, monotonic
    int err;                          // 1: a barrier / mbarrier wait timed out (protocol error), 2: round limit hit
    int na[2];                        // active slots of the even / odd rounds
    long long rounds;                 // rounds executed
    unsigned long long phase_ns[kDrPhases];   // CTA 0's wall time per phase: gemm, skin, sdf, frame; [4] = sum of the active counts; [5] total
};

struct DenseParams {
    DenseCtl* ctl;
    int* fidx[2];                     // slot -> frame of the even / odd rounds
    long long max_rounds;
    int ncols, ntiles;                // 3N, vertex tiles of 32
    float* poffT;
    SkinArgs skin;
    SdfFusedArgs sdf;
    FrameStepArgs fs;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_i32(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All threads of all CTAs.  `target` is this thread's running count of expected arrivals (only thread 0's copy is used).
// Bounded spin: a protocol error must not hang the GPU -- the error flag is raised and every CTA falls through.
__device__ __forceinline__ void grid_barrier(DenseCtl* ctl, unsigned& target) {
    asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy writes (feature rows) before later TMA reads
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(&ctl->bar, 1u);
        // RELAXED polls: an acquire load makes the SM drop its whole L1 (CCTL.IVALL) on every poll, which starves the
        // sibling CTA that is still working; one acquire fence after the last poll does the same job once
        long long spins = 0;
        while ((int)(ld_relaxed_u32(&ctl->bar) - target) < 0) {
            if (++spins > (1ll << 24)) {                       // ~ seconds
                if (ld_relaxed_i32(&ctl->err) == 0) atomicExch(&ctl->err, 1);
                break;
            }
            if ((spins & 1023) == 0 && ld_relaxed_i32(&ctl->err) != 0) break;
            __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ phase 1: contraction
struct GemmPipe {                     // per-thread copies; each role advances its own
    int stage; uint32_t phase;        // smem ring (producer and MMA issuer)
    int buf; uint32_t tphase[2];      // TMEM accumulator double buffer (MMA issuer and epilogue)
};

__device__ __noinline__ void gemm_phase(unsigned char* ring, uint64_t* full, uint64_t* empty, uint64_t* t_full, uint64_t* t_empty,
                                           const uint32_t tmem_base, GemmPipe& pp, const CUtensorMap* map_a, const CUtensorMap* map_a8,
                                           const CUtensorMap* map_b, const int na, const int ntiles, const int ncols, const int ldA,
                                           float* poffT, int* err_flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nmt = (na + kTcBM - 1) / kTcBM;
    const int nitems = ntiles * nmt;
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int mt = item / ntiles, vt = item - mt * ntiles, m0 = mt * kTcBM;
                const int live_rows = min(kTcBM, na - m0);
                const int groups = (live_rows + 7) / 8;
                const bool small = live_rows <= 32;           // few live rows: fetch them as 8-row boxes (same smem image)
                const uint32_t bytes = (small ? groups * 8 * kTcBK * 4 : kTcABytes) + kTcBBytes;
                for (int kc = 0; kc < kTcKCh; ++kc) {
                    if (!mbar_wait(&empty[pp.stage], pp.phase ^ 1, err_flag)) break;
                    unsigned char* sA = ring + (size_t)pp.stage * kDrStageBytes;
                    unsigned char* sB = sA + kTcABytes;
                    mbar_expect_tx(&full[pp.stage], bytes);
                    if (small) {
                        for (int gq = 0; gq < groups; ++gq) tma_load_2d(sA + (size_t)gq * 1024, map_a8, kc * kTcBK, m0 + 8 * gq, &full[pp.stage]);
                    } else {
                        tma_load_2d(sA, map_a, kc * kTcBK, m0, &full[pp.stage]);
                    }
                    tma_load_2d(sB, map_b, kc * kTcBK, vt * kTcBN, &full[pp.stage]);
                    if (++pp.stage == kDrStages) { pp.stage = 0; pp.phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kTcBM, kTcBN);
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                if (!mbar_wait(&t_empty[pp.buf], pp.tphase[pp.buf] ^ 1, err_flag)) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(pp.buf * kTcBN);
                bool ok = true;
                for (int kc = 0; kc < kTcKCh; ++kc) {
                    if (!mbar_wait(&full[pp.stage], pp.phase, err_flag)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(ring + (size_t)pp.stage * kDrStageBytes);
                    const uint32_t b_addr = a_addr + kTcABytes;
#pragma unroll
                    for (int k = 0; k < kTcBK / 8; ++k)
                        umma_tf32(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc, (kc | k) ? 1u : 0u);
                    umma_commit(&empty[pp.stage]);
                    if (++pp.stage == kDrStages) { pp.stage = 0; pp.phase ^= 1; }
                }
                if (!ok) break;
                umma_commit(&t_full[pp.buf]);
                pp.tphase[pp.buf] ^= 1;
                pp.buf ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int row = 32 * (warp & 3) + lane;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int mt = item / ntiles, vt = item - mt * ntiles;
            const int slot = mt * kTcBM + row;
            const bool live = slot < na;
            if (!mbar_wait(&t_full[pp.buf], pp.tphase[pp.buf], err_flag)) break;
            pp.tphase[pp.buf] ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(pp.buf * kTcBN);
            const int c0 = vt * kTcBN;
#pragma unroll
            for (int part = 0; part < 3; ++part) {            // 32 columns at a time: 32 live registers instead of 96
                uint32_t acc[32];
                tmem_ld32(taddr + 32 * part, acc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (part == 2) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&t_empty[pp.buf]);             // accumulator copied out: the MMA warp may reuse it
                }
                if (live) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int col = c0 + 32 * part + c;
                        if (col < ncols) poffT[(size_t)col * ldA + slot] = __uint_as_float(acc[c]);
                    }
                }
            }
            pp.buf ^= 1;
        }
    }
    __syncthreads();
}

// Every phase is a separate (non-inlined) function: each gets its own register allocation under the 128-register cap
// of two CTAs per SM instead of sharing one allocation across the whole round loop.
__device__ __noinline__ void phase_skin_small(SkinSmallSmem* sm, const SkinArgs& ar, int na, int nchunks) {
    const int grp = threadIdx.x >> 6;
    for (int ch = blockIdx.x * 4 + grp; ch < nchunks; ch += gridDim.x * 4)
        skin_small_body(sm[grp], ar, na, ch, nchunks, threadIdx.x & 63, 1 + grp);
}
__device__ __noinline__ void phase_skin(float* sk, const SkinArgs& ar, int na, int nchunks) {
    const int nfg = (na + 31) / 32;
    const int per_cta = (nchunks * nfg + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nx = (nchunks + per_cta - 1) / per_cta;
    for (int item = blockIdx.x; item < nx * nfg; item += gridDim.x) skin_body(sk, ar, na, (int)gridDim.x, item % nx, item / nx);
}
__device__ __noinline__ void phase_sdf(SdfFusedSmem& sm, const SdfFusedArgs& ar, int na, int nblocks) {
    const int passes = sdf_passes_for(na, nblocks, (int)gridDim.x);
    const int nparts = (nblocks + passes - 1) / passes;
    for (int item = blockIdx.x; item < nparts * na; item += gridDim.x) sdf_fused_body(sm, ar, na, item / nparts, item % nparts, passes);
}
__device__ __noinline__ void phase_frame(unsigned char* smem, const FrameStepArgs& fa, int na, const int* fidx, int* na_next, int* fidx_next) {
    for (int slot = blockIdx.x; slot < na; slot += gridDim.x) {
        frame_step_body(smem, fa, slot, fidx[slot], na_next, fidx_next);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kDrThreads, 2)
dense_rounds_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a8,
                    const __grid_constant__ CUtensorMap map_b, const __grid_constant__ DenseParams P) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ uint64_t s_bars[2 * kDrStages + 4];
    __shared__ uint32_t s_tmem;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = s_bars;
    uint64_t* empty = s_bars + kDrStages;
    uint64_t* t_full = empty + kDrStages;
    uint64_t* t_empty = t_full + 2;
    DenseCtl* ctl = P.ctl;
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDrStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;

    GemmPipe pp{0, 0u, 0, {0u, 0u}};
    unsigned bar_target = 0;
    const int N = P.skin.N, ldA = P.skin.ldA;
    const int nchunks = (N + kSkinV - 1) / kSkinV;
    const int nblocks = (N + kSdfFThreads - 1) / kSdfFThreads;
    const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t_prev = timer ? globaltimer_ns() : 0ull, t_start = t_prev;
    long long round = 0;
    for (;; ++round) {
        const int par = (int)(round & 1);
        const int na = ld_relaxed_i32(&ctl->na[par]);
        if (na <= 0 || ld_relaxed_i32(&ctl->err) != 0) break;
        if (round >= P.max_rounds) { if (timer) atomicExch(&ctl->err, 2); break; }
        const int* fidx = P.fidx[par];
        if (timer) ctl->phase_ns[4] += (unsigned long long)na;        // frame-rounds (sum of the active counts)

        // ---- phase 1: pose offsets of the active frames (tcgen05)
        DMARK(39);
        gemm_phase(smem, full, empty, t_full, t_empty, tmem_base, pp, &map_a, &map_a8, &map_b, na, P.ntiles, P.ncols, ldA, P.poffT, &ctl->err);
        DMARK(40);
        grid_barrier(ctl, bar_target);
        DMARK(41);
        if (timer) {
            ctl->na[par ^ 1] = 0;              // every CTA read it at the top of the previous round; refilled by phase 4
            const unsigned long long t = globaltimer_ns(); ctl->phase_ns[0] += t - t_prev; t_prev = t;
        }

        // ---- phase 2: skinning + per-chunk boxes (straggler tail: lane = vertex, four 64-thread groups per CTA)
        if (na <= kSkinSmallMax) phase_skin_small(reinterpret_cast<SkinSmallSmem*>(smem), P.skin, na, nchunks);
        else phase_skin(reinterpret_cast<float*>(smem), P.skin, na, nchunks);
        DMARK(42);
        grid_barrier(ctl, bar_target);
        DMARK(43);
        if (timer) { const unsigned long long t = globaltimer_ns(); ctl->phase_ns[1] += t - t_prev; t_prev = t; }

        // ---- phase 3: SDF samples + sparse adjoint partials
        phase_sdf(*reinterpret_cast<SdfFusedSmem*>(smem), P.sdf, na, nblocks);
        DMARK(44);
        grid_barrier(ctl, bar_target);
        DMARK(45);
        if (timer) { const unsigned long long t = globaltimer_ns(); ctl->phase_ns[2] += t - t_prev; t_prev = t; }

        // ---- phase 4: per frame -- closure adjoint, priors, L-BFGS step, next pose forward into next round's slot
        phase_frame(smem, P.fs, na, fidx, &ctl->na[par ^ 1], P.fidx[par ^ 1]);
        DMARK(46);
        grid_barrier(ctl, bar_target);
        DMARK(47);
        if (timer) { const unsigned long long t = globaltimer_ns(); ctl->phase_ns[3] += t - t_prev; t_prev = t; }
    }
    if (timer) { ctl->rounds = round; ctl->phase_ns[5] += globaltimer_ns() - t_start; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

__global__ void dense_list_init_kernel(int* fidx0, int B, DenseCtl* ctl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) fidx0[i] = i;
    if (i == 0) { ctl->na[0] = B; ctl->na[1] = 0; }
}

// ------------------------------------------------------------------------------------------------ host side
struct DenseState {
    DenseCtl* ctl = nullptr;
    DenseCtl* ctl_host = nullptr;     // pinned
    int* fidx[2] = {nullptr, nullptr};
    int grid = 0;
    size_t smem = 0;
    unsigned long long phase_ns_acc[kDrPhases] = {0};
    long long rounds_acc = 0;
};

const void* tc_maps(mvs_ctx* ctx);             // mvs_tc.cu
int* tc_err_flag(mvs_ctx* ctx);

bool dense_persistent_available(const mvs_ctx* ctx) {
    // OPT-IN (MVS_DENSE_PERSISTENT=1).  Measured on B200 (profiles/r02_dense_persistent.md): bit-identical to the four-launch
    // rounds and free of host involvement, but 1.3-1.6x SLOWER per round -- every phase is a dependent-latency chain
    // (pipeline fill, box fold, optimiser recursion) that a grid barrier does not shorten, programmatic dependent launch
    // already hides the launch gaps of the four-launch rounds, and the persistent kernel pays for holding the largest
    // phase's shared memory (110 KB: one CTA per SM, almost no L1) during ALL phases.
    static const bool on = getenv("MVS_DENSE_PERSISTENT") != nullptr && atoi(getenv("MVS_DENSE_PERSISTENT")) != 0;
    return on && ctx->exec_mode != 1 && ctx->tc != nullptr;
}

// One persistent run over the active frames of `S` (initialised by lbfgs_init_kernel: x_eval = parameters, phase =
// STEP_ENTRY).  cfg.step_mode == 2 with max_rounds = 1 evaluates one closure only (mvs_closure in exec mode 3).
int run_dense_persistent(mvs_ctx* ctx, float* params_dev, const void* lbfgs_state, const void* lbfgs_cfg, int nstages,
                         long long max_rounds, cudaStream_t st) {
    Workspace& w = ctx->ws;
    const DevModel& m = ctx->m;
    const LbfgsState& L = *static_cast<const LbfgsState*>(lbfgs_state);
    int rc;
    if ((rc = tc_prepare(ctx))) return rc;
    if ((rc = ensure_sdf_fused_ws(ctx))) return rc;
    DenseState* D = static_cast<DenseState*>(ctx->dense);
    if (!D) {
        D = new DenseState();
        ctx->dense = D;
        unsigned char* raw = nullptr;
        if ((rc = dev_alloc(ctx, &raw, sizeof(DenseCtl)))) return rc;
        D->ctl = reinterpret_cast<DenseCtl*>(raw);
        if ((rc = dev_alloc(ctx, &D->fidx[0], (size_t)w.B))) return rc;
        if ((rc = dev_alloc(ctx, &D->fidx[1], (size_t)w.B))) return rc;
        MVS_CUDA_OK(ctx, cudaMallocHost(&D->ctl_host, sizeof(DenseCtl)));
    }
    FrameStepArgs fa = make_frame_step_args(ctx, params_dev, lbfgs_state, lbfgs_cfg, nstages);
    const size_t smem_frame = frame_step_smem(L.H, fa.with_vposer != 0);
    const size_t smem_gemm = (size_t)kDrStages * kDrStageBytes;
    const size_t smem_skin = kSkinSmemNoVp;    // v_posed is not stored in this path
    size_t smem = smem_frame;
    if (smem_gemm > smem) smem = smem_gemm;
    if (smem_skin > smem) smem = smem_skin;
    if (sizeof(SdfFusedSmem) > smem) smem = sizeof(SdfFusedSmem);
    if (4 * sizeof(SkinSmallSmem) > smem) smem = 4 * sizeof(SkinSmallSmem);
    smem += 1024;                              // alignment slack of the swizzled ring
    if (D->smem != smem) {
        MVS_CUDA_OK(ctx, cudaFuncSetAttribute(dense_rounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        MVS_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_rounds_kernel, kDrThreads, smem));
        if (per_sm < 1) return set_error(ctx, MVS_ERR_CUDA, "dense_rounds_kernel does not fit an SM (%zu bytes of shared memory)", smem);
        if (per_sm > 2) per_sm = 2;            // TMEM: 2 x 256 columns per SM
        if (getenv("MVS_DENSE_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(getenv("MVS_DENSE_CTAS_PER_SM"))));   // experiments
        D->grid = per_sm * ctx->sm_count;
        D->smem = smem;
    }
    const CUtensorMap* maps = static_cast<const CUtensorMap*>(tc_maps(ctx));
    if (!maps) return set_error(ctx, MVS_ERR_CUDA, "tensor maps unavailable");
    DenseParams P;
    P.ctl = D->ctl; P.fidx[0] = D->fidx[0]; P.fidx[1] = D->fidx[1]; P.max_rounds = max_rounds;
    P.ncols = 3 * m.N; P.ntiles = (m.N + kTileV - 1) / kTileV; P.poffT = tc_poffT(ctx);
    P.skin = SkinArgs{P.poffT, m.ST, w.Phi, w.At, w.ldA, m.ell_j, m.ell_w, m.KW, m.N, nullptr, w.verts, w.bboxp};
    P.sdf = make_sdf_fused_args(ctx);
    P.fs = fa;
    MVS_CUDA_OK(ctx, cudaMemsetAsync(D->ctl, 0, sizeof(DenseCtl), st));
    MVS_LAUNCH(ctx, KID_MISC, st, dense_list_init_kernel<<<(w.B + 255) / 256, 256, 0, st>>>(D->fidx[0], w.B, D->ctl));
    CUtensorMap ma = maps[0], ma8 = maps[1], mb = maps[2];
    void* args[] = {&ma, &ma8, &mb, &P};
    MVS_LAUNCH(ctx, KID_DENSE_ROUNDS, st,
               MVS_CUDA_OK(ctx, cudaLaunchCooperativeKernel((const void*)dense_rounds_kernel, dim3(D->grid), dim3(kDrThreads), args, smem, st)));
    MVS_CUDA_OK(ctx, cudaMemcpyAsync(D->ctl_host, D->ctl, sizeof(DenseCtl), cudaMemcpyDeviceToHost, st));
    MVS_CUDA_OK(ctx, cudaStreamSynchronize(st));
    MVS_CUDA_OK(ctx, cudaGetLastError());
    for (int k = 0; k < kDrPhases; ++k) D->phase_ns_acc[k] += D->ctl_host->phase_ns[k];
    D->rounds_acc += D->ctl_host->rounds;
    if (D->ctl_host->err == 1) return set_error(ctx, MVS_ERR_CUDA, "dense_rounds_kernel: a barrier wait timed out (protocol error)");
    if (D->ctl_host->err == 2) return set_error(ctx, MVS_ERR_CUDA, "dense_rounds_kernel: round limit reached with frames still active");
    return MVS_OK;
}

long long dense_last_rounds(const mvs_ctx* ctx) {
    const DenseState* D = static_cast<const DenseState*>(ctx->dense);
    return D && D->ctl_host ? D->ctl_host->rounds : 0;
}

}  // namespace mvs

using namespace mvs;

// measurement support (bench.py): accumulated wall time of the phases of the persistent dense rounds as seen by CTA 0
// (ns: gemm, skin, sdf, frame; [4] = sum over rounds of the active-frame counts; [5] = whole kernel, ns) and the number of
// rounds, since the last call; resets the counters.
extern "C" int mvs_dense_phase_times(mvs_ctx* ctx, double* ns6, long long* rounds) {
    if (!ctx) return set_error(nullptr, MVS_ERR_INVALID, "ctx is NULL");
    DenseState* D = static_cast<DenseState*>(ctx->dense);
    for (int k = 0; k < kDrPhases; ++k) { if (ns6) ns6[k] = D ? (double)D->phase_ns_acc[k] : 0.0; if (D) D->phase_ns_acc[k] = 0; }
    if (rounds) *rounds = D ? D->rounds_acc : 0;
    if (D) D->rounds_acc = 0;
    return MVS_OK;
}

#ifdef MVS_PHASE_DBG
extern "C" int mvs_debug_dense_clocks(long long* clk, long long* cnt, int n) {       // debug builds only; not part of the ABI
    long long h[64], c[64], z[64] = {0};
    if (cudaMemcpyFromSymbol(h, mvs::g_dense_clk, sizeof(h)) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(c, mvs::g_dense_cnt, sizeof(c)) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 64; ++i) { clk[i] = h[i]; cnt[i] = c[i]; }
    cudaMemcpyToSymbol(mvs::g_dense_clk, z, sizeof(z)); cudaMemcpyToSymbol(mvs::g_dense_cnt, z, sizeof(z));
    return 0;
}
#endif
