// Device-side building blocks of the frame-resident kernels (mvs_resident.cu) shared with the persistent dense-round
// kernel (mvs_dense.cu): shared-memory layout of one frame, level-parallel kinematic chain, VPoser decoder, the closure,
// the pose forward of the next trial point, and the per-frame half of a dense round (frame_step_body).
// A translation unit may define PHASE_MARK(i) before including this file (clock instrumentation); default: nothing.
#pragma once
#include <cstddef>
#include "mvs_internal.cuh"
#include "mvs_lbfgs_core.cuh"
#include "mvs_tc_dev.cuh"
#ifndef PHASE_MARK
#define PHASE_MARK(i) do {} while (0)
#endif

namespace mvs {

constexpr int kResThreads = 256;
constexpr int kResMaxSup = 96;            // support vertices (SMPL: 86)
constexpr int kResMaxSupJ = 512;          // (support vertex, joint) pairs with a non-zero skinning weight (SMPL: <= 344)
constexpr int kResMaxVK = 272;            // views x keypoints (16 x 17)
constexpr int kResMaxM = 8;               // GMM components

// Depth-level schedule of the kinematic tree: joints of one level only depend on the level above (forward) / below
// (adjoint), so a level is processed by parallel lane groups and the 23-step serial chain becomes <= 9 level steps
// for SMPL.  Children are listed in DESCENDING index order: gathering them in that order reproduces the summation
// order of the plain "for j = 23 .. 1" reverse sweep bit for bit.
struct ChainSched {
    int nlev;
    unsigned char lev_ptr[kJoints + 1], lev_j[kJoints];
    unsigned char ch_ptr[kJoints + 1], ch_j[kJoints];
};

struct ResidentModel {                    // device pointers + sizes, by value
    const float* Qk; const float* Jt; const float* JS;
    const int* ell_j; const float* ell_w; int KW;
    int K, nsup;
    const int* kp_ptr; const int* kp_spos; const float* kp_w; const int* kp_chain;
    const int* sup; const int* sup_ptr; const int* sup_k; const float* sup_w;
    const int* supj_ptr; const int* supj_i; const float* supj_w;      // per joint: (support index, weight)
    int M; const float* gmm_means; const float* gmm_prec; const float* gmm_lognllw;
    const float* anchor; const float* anchor_w;     // sequence mode (mvs_set_anchor)
    const float *vp_w1, *vp_w2, *vp_w3, *vp_w1t, *vp_w2t, *vp_w3t, *vp_b1, *vp_b2, *vp_b3;      // VPoser decoder
    Parents par;
    ChainSched cs;
};

struct ResidentSmem {
    // pose
    float x[88], R[216], J[72], Gam[216], g[72], A[288], Phi[224];
    // support vertices
    float vp[kResMaxSup * 3], v[kResMaxSup * 3], dv[kResMaxSup * 3], dvp[kResMaxSup * 3];
    int rowbase[kResMaxSup * 3];          // Qk row of column (3 i + c)
    // keypoints / data term
    float q[kMaxKeypoints * 3], dq[kMaxKeypoints * 3], contrib[kResMaxVK * 3], lterm[kResMaxVK], vsum[kMaxViews];
    // adjoint
    float dA[288], dPhi[224], dR[216], dGam[216], dg[72], dJ[72], grad[88];
    // priors
    float gm_diff[kResMaxM * 69], gm_y[kResMaxM * 69], gm_ll[kResMaxM];
    float red[kResThreads];
    float sc[16];                          // [0] data loss, [1] |theta|^2, [2] total loss
    // optimiser vectors (lbfgs_resident_kernel only)
    float lx[88], lg[88], ld[88], lprev_g[88], lx_init[88], lg_prev[88], lbg0[88], lbg1[88], lx_eval[88], lg_new[88];
    float ro[128], al[128];
    float gram[kGramFloats];               // s_i . y_j within 8-slot blocks of the history ring (blocked two-loop recursion)
    float tl_scratch[96];
    // level schedule of the kinematic tree (copy of ResidentModel::cs / par: shared-memory latency instead of
    // dependent constant-bank loads inside the level loops)
    int nlev, lev_ptr[kJoints + 1], lev_j[kJoints], ch_ptr[kJoints + 1], ch_j[kJoints], par[kJoints];
    // per joint: the support vertices it skins (copy of supj_*: the dA loop of P7 is a chain of dependent list reads)
    int sj_ptr[kJoints + 1], sj_i[kResMaxSupJ], sj_on;
    float sj_w[kResMaxSupJ];
    // dense regime: frame scalars of the SDF term and its box-extreme vertex list (frame_step_kernel)
    int ext_n[8];
    float ext_d[24];
    float sdf_sc[4];                       // [0] cg / scale, [1] pen loss, [2] number of extreme entries
    int fl[64], nfl;                       // 256-vertex blocks with a penetration adjoint partial
    LossParams lp;                         // this frame's CURRENT stage (multi-stage kernels)
    FrameScalars fs;
};

constexpr int kPoseCacheFloats = 216 + 72 + 216 + 72 + 288;      // R | J | Gam | g | A
static_assert(offsetof(ResidentSmem, A) - offsetof(ResidentSmem, R) == (216 + 72 + 216 + 72) * sizeof(float), "pose cache layout");
static_assert(offsetof(ResidentSmem, gram) % 16 == 0, "cp.async 16-byte staging of the block Gram");
static_assert(sizeof(ResidentSmem) + 2 * 100 * kParams * sizeof(float) <= 112 * 1024, "two frame CTAs per SM (227 KB)");

// Inputs of the dense regime (SDF term on): the skinning kernel already produced every vertex of the frame, and
// sdf_fused_kernel the unit-factor adjoint of the vertices inside the penetration cone (per 1024-vertex part).
struct DenseIn {
    const float* verts;       // [N][3] this frame's skinned vertices, pre-transl (NULL: sparse regime, recompute the support vertices)
    const float* poffT;       // [3N][ldA] pose offsets of the tensor-core contraction; this frame is column `slot`
    const float* ST;          // [N][33] shapedirs rows | template: v_posed = vposed_of(ST row, betas, pose offset)
    int ldA, slot;
    const int* extra_n;       // [n_extra] box-extreme vertices that carry the gradient through the box centre / scale
    const float* extra_d;     // [n_extra][3] d pen / d vertex
    int n_extra;
    float pen_loss;
    const float* Wd;          // [N][24] dense skinning weights (adjoint of the extra vertices)
    const float* part;        // [nblocks][512] this frame's unit-factor partial adjoints, or NULL (no penetration)
    const int* fl;            // [nfl] blocks that wrote partials, ascending
    int nfl;
    float factor;             // cg / scale: d pen / d (sum of samples) over the box scale
    bool pose_ready;          // R, J, Gam, g, A of this trial point are already in shared memory (frame_step's pose cache)
};

// ------------------------------------------------------------------------------------------------
// Kinematic chain by depth levels (lbs.py:348-374).  Every ROLE (rotation part, translation part, ...) has its own
// warps so that no warp executes more than one branch body: a level costs one short dependent chain plus a barrier.
// All kResThreads threads must call; ends with a barrier.
__device__ __forceinline__ void chain_fwd_levels(ResidentSmem& S) {
    const int t = threadIdx.x;
    if (t < 9) S.Gam[t] = S.x[kOffScale] * S.R[t];                               // lbs.py:348
    else if (t >= 64 && t < 67) S.g[t - 64] = S.J[t - 64];
    __syncthreads();
    for (int L = 1; L < S.nlev; ++L) {
        const int beg = S.lev_ptr[L], cnt = S.lev_ptr[L + 1] - beg;
        if (t < 63) {                                        // threads 0..62: Gam_j = Gam_p R_j, 9 lanes per joint
            const int l = t % 9;
            for (int q = t / 9; q < cnt; q += 7) {
                const int j = S.lev_j[beg + q], p = S.par[j];
                const float* Gp = &S.Gam[9 * p];
                const float* Rj = &S.R[9 * j];
                const int r = l / 3, c = l % 3;
                S.Gam[9 * j + l] = Gp[3 * r] * Rj[c] + Gp[3 * r + 1] * Rj[3 + c] + Gp[3 * r + 2] * Rj[6 + c];
            }
        } else if (t >= 64 && t < 64 + 30) {                 // threads 64..93: g_j = Gam_p (J_j - J_p) + g_p, 3 lanes per joint
            const int u = t - 64, r = u % 3;
            for (int q = u / 3; q < cnt; q += 10) {
                const int j = S.lev_j[beg + q], p = S.par[j];
                const float* Gp = &S.Gam[9 * p];
                const float r0 = S.J[3 * j] - S.J[3 * p], r1 = S.J[3 * j + 1] - S.J[3 * p + 1], r2 = S.J[3 * j + 2] - S.J[3 * p + 2];
                S.g[3 * j + r] = (Gp[3 * r] * r0 + Gp[3 * r + 1] * r1 + Gp[3 * r + 2] * r2) + S.g[3 * p + r];
            }
        }
        __syncthreads();
    }
}

// Adjoint of the chain, deepest level first.  Per level L (parents p with children j at level L + 1, already final):
//   role A (threads 0..62)     dGam_p += sum_j dGam_j R_j^T + dg_j rel_j^T      9 lanes per parent
//   role B (threads 64..126)   dR_j = Gam_p^T dGam_j                             9 lanes per child
//   role C (threads 128..157)  drel_j = Gam_p^T dg_j;  dJ_j += drel_j, dJ_p -= sum_j drel_j     3 lanes per parent
//   role D (threads 160..189)  dg_p += sum_j dg_j                                3 lanes per parent
// Children are gathered in descending index order = the summation order of the plain j = 23..1 sweep.
__device__ __forceinline__ void chain_bwd_levels(ResidentSmem& S) {
    const int t = threadIdx.x;
    for (int L = S.nlev - 2; L >= 0; --L) {
        const int beg = S.lev_ptr[L], cnt = S.lev_ptr[L + 1] - beg;
        if (t < 63) {
            const int l = t % 9, r = l / 3, c = l % 3;
            for (int q = t / 9; q < cnt; q += 7) {
                const int p = S.lev_j[beg + q];
                const int c0 = S.ch_ptr[p], c1 = S.ch_ptr[p + 1];
                if (c0 == c1) continue;
                float acc = S.dGam[9 * p + l];
                for (int ci = c0; ci < c1; ++ci) {
                    const int j = S.ch_j[ci];
                    const float* dGj = &S.dGam[9 * j];
                    const float* Rj = &S.R[9 * j];
                    const float rel = S.J[3 * j + c] - S.J[3 * p + c];
                    acc += (dGj[3 * r] * Rj[3 * c] + dGj[3 * r + 1] * Rj[3 * c + 1] + dGj[3 * r + 2] * Rj[3 * c + 2]) + S.dg[3 * j + r] * rel;
                }
                S.dGam[9 * p + l] = acc;
            }
        } else if (t >= 64 && t < 64 + 63) {
            const int u = t - 64, e = u % 9, r = e / 3, c = e % 3;
            const int cbeg = S.lev_ptr[L + 1], ccnt = S.lev_ptr[L + 2] - cbeg;
            for (int q = u / 9; q < ccnt; q += 7) {
                const int j = S.lev_j[cbeg + q], p = S.par[j];
                const float* Gp = &S.Gam[9 * p];
                const float* dGj = &S.dGam[9 * j];
                S.dR[9 * j + e] = Gp[r] * dGj[c] + Gp[3 + r] * dGj[3 + c] + Gp[6 + r] * dGj[6 + c];
            }
        } else if (t >= 128 && t < 128 + 30) {
            const int u = t - 128, r = u % 3;
            for (int q = u / 3; q < cnt; q += 10) {
                const int p = S.lev_j[beg + q];
                const int c0 = S.ch_ptr[p], c1 = S.ch_ptr[p + 1];
                if (c0 == c1) continue;
                const float* Gp = &S.Gam[9 * p];
                float acc = S.dJ[3 * p + r];
                for (int ci = c0; ci < c1; ++ci) {
                    const int j = S.ch_j[ci];
                    const float* dgj = &S.dg[3 * j];
                    const float out = Gp[r] * dgj[0] + Gp[3 + r] * dgj[1] + Gp[6 + r] * dgj[2];
                    S.dJ[3 * j + r] += out;
                    acc -= out;
                }
                S.dJ[3 * p + r] = acc;
            }
        } else if (t >= 160 && t < 160 + 30) {
            const int u = t - 160, r = u % 3;
            for (int q = u / 3; q < cnt; q += 10) {
                const int p = S.lev_j[beg + q];
                const int c0 = S.ch_ptr[p], c1 = S.ch_ptr[p + 1];
                if (c0 == c1) continue;
                float acc = S.dg[3 * p + r];
                for (int ci = c0; ci < c1; ++ci) acc += S.dg[3 * S.ch_j[ci] + r];
                S.dg[3 * p + r] = acc;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// VPoser decoder on the device (use_vposer = 2): body_pose = VPoser.decode(z, 'aa') (model/VPoser.py:218-232,
// fitting.py:121-123) and its adjoint.  z = S.x[kOffPose .. kOffPose + 32).  The scratch lives in dynamic shared memory
// behind the curvature history (only kernels launched for this mode reserve it).
constexpr int kVpH = 512, kVpZ = 32, kVpO = 138;
struct VposerSmem {
    float h1[kVpH], h2[kVpH], o6[kVpO + 6], th[72], dth[72], dh[kVpH], red[256];
};

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : 0.2f * v; }

// all threads; ends with a barrier.  Leaves W.th = decoded axis-angle pose (69) and zeroes W.dth.
static __device__ void vposer_decode(const ResidentSmem& S, const ResidentModel& m, VposerSmem& W) {
    const int t = threadIdx.x;
    const float* z = &S.x[kOffPose];
#pragma unroll
    for (int o = t; o < kVpH; o += kResThreads) {
        float a = m.vp_b1[o];
#pragma unroll 8
        for (int i = 0; i < kVpZ; ++i) a = fmaf(__ldg(m.vp_w1t + (size_t)i * kVpH + o), z[i], a);
        W.h1[o] = lrelu(a);
    }
    __syncthreads();
    {
        float a0 = m.vp_b2[t], a1 = m.vp_b2[t + kResThreads];
#pragma unroll 32
        for (int i = 0; i < kVpH; ++i) {
            const float hi = W.h1[i];
            a0 = fmaf(__ldg(m.vp_w2t + (size_t)i * kVpH + t), hi, a0);
            a1 = fmaf(__ldg(m.vp_w2t + (size_t)i * kVpH + t + kResThreads), hi, a1);
        }
        W.h2[t] = lrelu(a0); W.h2[t + kResThreads] = lrelu(a1);
    }
    __syncthreads();
    if (t < kVpO) {
        float a = m.vp_b3[t];
#pragma unroll 32
        for (int i = 0; i < kVpH; ++i) a = fmaf(__ldg(m.vp_w3t + (size_t)i * kVpO + t), W.h2[i], a);
        W.o6[t] = a;
    }
    if (t < 72) W.dth[t] = 0.f;
    __syncthreads();
    if (t < kJoints - 1) {
        Cont6dState<float> st;
        cont6d_to_aa_fwd(&W.o6[6 * t], &W.th[3 * t], st);
    }
    __syncthreads();
}

// all threads; adjoint of vposer_decode: W.dth (d loss / d decoded pose, 69) -> dz[32] in W.red[0..31].  Ends with a barrier.
static __device__ void vposer_decode_bwd(const ResidentModel& m, VposerSmem& W) {
    const int t = threadIdx.x;
    if (t < kJoints - 1) {
        Cont6dState<float> st;
        float aa[3], d6[6];
        cont6d_to_aa_fwd(&W.o6[6 * t], aa, st);               // recompute the intermediates (cheaper than keeping 23 states)
        cont6d_to_aa_bwd(st, &W.dth[3 * t], d6);
#pragma unroll
        for (int c = 0; c < 6; ++c) W.red[6 * t + c] = d6[c];                     // d o6 (138 <= 256)
    }
    __syncthreads();
    {   // d h2 = lrelu'(h2) .* (W3^T d o6)
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 23
        for (int o = 0; o < kVpO; ++o) {
            const float d = W.red[o];
            a0 = fmaf(__ldg(m.vp_w3 + (size_t)o * kVpH + t), d, a0);
            a1 = fmaf(__ldg(m.vp_w3 + (size_t)o * kVpH + t + kResThreads), d, a1);
        }
        W.dh[t] = (W.h2[t] > 0.f ? 1.f : 0.2f) * a0;
        W.dh[t + kResThreads] = (W.h2[t + kResThreads] > 0.f ? 1.f : 0.2f) * a1;
    }
    __syncthreads();
    {   // d h1 = lrelu'(h1) .* (W2^T d h2)   (result overwrites h2, no longer needed)
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 32
        for (int o = 0; o < kVpH; ++o) {
            const float d = W.dh[o];
            a0 = fmaf(__ldg(m.vp_w2 + (size_t)o * kVpH + t), d, a0);
            a1 = fmaf(__ldg(m.vp_w2 + (size_t)o * kVpH + t + kResThreads), d, a1);
        }
        W.h2[t] = (W.h1[t] > 0.f ? 1.f : 0.2f) * a0;
        W.h2[t + kResThreads] = (W.h1[t + kResThreads] > 0.f ? 1.f : 0.2f) * a1;
    }
    __syncthreads();
    {   // d z = W1^T d h1: 8 partial sums of 64 terms per latent entry
        const int k = t & 31, part = t >> 5;
        float a = 0.f;
#pragma unroll 8
        for (int o = part * 64; o < part * 64 + 64; ++o) a = fmaf(__ldg(m.vp_w1 + (size_t)o * kVpZ + k), W.h2[o], a);
        W.dh[t] = a;
    }
    __syncthreads();
    if (t < kVpZ) {
        float a = 0.f;
#pragma unroll
        for (int part = 0; part < 8; ++part) a += W.dh[32 * part + t];
        W.red[t] = a;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// One closure evaluation at S.x (already loaded).  Writes S.sc[2] = total loss and S.lg_new = gradient
// (frozen segments zeroed).  All kResThreads threads must call it.
static __device__ void resident_closure(ResidentSmem& S, const ResidentModel& m, const CamSet& cams, const LossParams& lp,
                                 const float* __restrict__ gt_uv, const float* __restrict__ conf,
                                 const float* __restrict__ joint_w, int B, int b, bool have_grad,
                                 float* __restrict__ joints_out, float* __restrict__ proj_out, const DenseIn& din,
                                 VposerSmem* W = nullptr) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nsup = m.nsup, K = m.K, V = cams.num_views, ncol = 3 * nsup;
    // use_vposer == 2: the body pose is decoded from the latent code in the pose slot (fitting.py:121-123)
    const bool vp2 = (lp.use_vposer == 2) && W != nullptr;
    if (vp2) vposer_decode(S, m, *W);
    const float* theta = vp2 ? W->th : &S.x[kOffPose];          // the 69 body-pose entries the model sees

    // ---- P1 Rodrigues, rest joints (skipped when the caller restored them: the previous frame_step computed the pose
    //      forward of exactly this trial point for the dense kernels and parked it in the per-frame pose cache)
    if (!din.pose_ready) {
        if (t < kJoints) rodrigues_fwd(t == 0 ? &S.x[kOffOrient] : &theta[3 * (t - 1)], &S.R[9 * t]);
        else if (t >= 32 && t < 32 + 72) {
            const int jc = t - 32;
            float a = m.Jt[jc];
#pragma unroll
            for (int l = 0; l < kBetas; ++l) a = fmaf(m.JS[jc * kBetas + l], S.x[kOffBetas + l], a);
            S.J[jc] = a;
        }
    }
    for (int i = t; i < kJoints * 9; i += kResThreads) { S.dGam[i] = 0.f; S.dR[i] = 0.f; }
    for (int i = t; i < kJoints * 3; i += kResThreads) { S.dg[i] = 0.f; S.dJ[i] = 0.f; }
    for (int i = t; i < kParams; i += kResThreads) S.grad[i] = 0.f;
    __syncthreads();
    PHASE_MARK(2);
    if (!din.pose_ready) {
        // ---- P2 kinematic chain, level-parallel
        chain_fwd_levels(S);
        PHASE_MARK(3);
        // ---- P3 skinning transforms and the feature row
        if (t < kJoints) make_skin_transform(&S.Gam[9 * t], &S.g[3 * t], &S.J[3 * t], &S.A[12 * t]);
        else if (t >= 32) {
            const int k = t - 32;                                                      // 224 threads, 224 entries
            float v;
            if (k < kPoseBasis) v = S.R[9 + k] - (((k % 9) % 4 == 0) ? 1.0f : 0.0f);
            else if (k < kPoseBasis + kBetas) v = S.x[kOffBetas + k - kPoseBasis];
            else v = (k == kFeat - 1) ? 1.0f : 0.0f;
            S.Phi[k] = v;
        }
        __syncthreads();
    }
    PHASE_MARK(4);
    // ---- P4 v_posed for the support columns: warp per Qk row, 2 x LDG.128 per lane
    if (din.verts) {                  // dense regime: the vertex kernels already have them (v_posed: bit for bit skin_vertex's)
        for (int col = t; col < ncol; col += kResThreads) {
            const int n = m.sup[col / 3], c = col % 3;
            S.vp[col] = vposed_of(din.ST + (size_t)n * 33 + 11 * c, &S.x[kOffBetas], din.poffT[(size_t)(3 * n + c) * din.ldA + din.slot]);
            S.v[col] = din.verts[3 * n + c];
        }
        // the box-extreme vertices of the penetration term ride along as support entries nsup .. nsup + n_extra - 1
        if (t < 3 * din.n_extra) {
            const int n = din.extra_n[t / 3], c = t % 3;
            S.vp[ncol + t] = vposed_of(din.ST + (size_t)n * 33 + 11 * c, &S.x[kOffBetas], din.poffT[(size_t)(3 * n + c) * din.ldA + din.slot]);
            S.dv[ncol + t] = din.extra_d[t];
            S.rowbase[ncol + t] = 3 * n + c;
        }
    } else {
        // warp per Qk row (896 B = 2 x LDG.128 per lane), FOUR rows in flight per warp: the 231 KB slice comes from L2
        // and the phase is pure latency, so the depth of the load queue is what sets its length
        const float4 ph0 = *reinterpret_cast<const float4*>(&S.Phi[4 * lane]);
        const float4 ph1 = lane < 24 ? *reinterpret_cast<const float4*>(&S.Phi[128 + 4 * lane]) : make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr int kW = kResThreads / 32;
        for (int col0 = warp; col0 < ncol; col0 += 4 * kW) {
            float4 a[4], bq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int col = col0 + u * kW;
                const float* r = m.Qk + (size_t)S.rowbase[col < ncol ? col : col0] * kFeatPad;
                a[u] = __ldg(reinterpret_cast<const float4*>(r + 4 * lane));
                bq[u] = lane < 24 ? __ldg(reinterpret_cast<const float4*>(r + 128 + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float p = a[u].x * ph0.x;
                p = fmaf(a[u].y, ph0.y, p); p = fmaf(a[u].z, ph0.z, p); p = fmaf(a[u].w, ph0.w, p);
                p = fmaf(bq[u].x, ph1.x, p); p = fmaf(bq[u].y, ph1.y, p); p = fmaf(bq[u].z, ph1.z, p); p = fmaf(bq[u].w, ph1.w, p);
                p = warp_sum(p);
                const int col = col0 + u * kW;
                if (lane == 0 && col < ncol) S.vp[col] = p;
            }
        }
    }
    __syncthreads();
    PHASE_MARK(5);
    // ---- P5 linear blend skinning of the support vertices
    if (!din.verts && t < nsup) {
        const int n = m.sup[t];
        float T[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) T[c] = 0.f;
        for (int e = 0; e < m.KW; ++e) {
            const float w = m.ell_w[(size_t)n * m.KW + e];
            if (w != 0.f) {
                const float* Aj = &S.A[12 * m.ell_j[(size_t)n * m.KW + e]];
#pragma unroll
                for (int c = 0; c < 12; ++c) T[c] = fmaf(w, Aj[c], T[c]);
            }
        }
        const float p0 = S.vp[3 * t], p1 = S.vp[3 * t + 1], p2 = S.vp[3 * t + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) S.v[3 * t + r] = T[4 * r] * p0 + T[4 * r + 1] * p1 + T[4 * r + 2] * p2 + T[4 * r + 3];
    }
    __syncthreads();
    PHASE_MARK(6);
    // ---- P6 keypoints, projection, GMoF data term and its adjoint down to the support vertices
    if (t < K) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int e = m.kp_ptr[t]; e < m.kp_ptr[t + 1]; ++e) {
            const int pos = m.kp_spos[e];
            const float w = m.kp_w[e];
            a0 = fmaf(w, S.v[3 * pos], a0); a1 = fmaf(w, S.v[3 * pos + 1], a1); a2 = fmaf(w, S.v[3 * pos + 2], a2);
        }
        const int cj = m.kp_chain[t];
        if (cj >= 0) { a0 += S.g[3 * cj]; a1 += S.g[3 * cj + 1]; a2 += S.g[3 * cj + 2]; }
        S.q[3 * t] = a0 + S.x[kOffTransl]; S.q[3 * t + 1] = a1 + S.x[kOffTransl + 1]; S.q[3 * t + 2] = a2 + S.x[kOffTransl + 2];
        if (joints_out) {
            joints_out[((size_t)b * K + t) * 3] = S.q[3 * t]; joints_out[((size_t)b * K + t) * 3 + 1] = S.q[3 * t + 1];
            joints_out[((size_t)b * K + t) * 3 + 2] = S.q[3 * t + 2];
        }
    }
    __syncthreads();
    PHASE_MARK(7);
    {
        const float rho2 = lp.rho * lp.rho, dw2 = lp.data_weight * lp.data_weight;
        for (int idx = t; idx < V * K; idx += kResThreads) {
            const int v = idx / K, k = idx % K;
            float xc[3], uv[2];
            project_fwd(cams.cam[v], &S.q[3 * k], xc, uv);
            const size_t o = ((size_t)v * B + b) * K + k;
            if (proj_out) { proj_out[2 * o] = uv[0]; proj_out[2 * o + 1] = uv[1]; }
            float w = joint_w[k];
            if (lp.use_joints_conf) w *= conf[o];
            const float w2 = w * w;
            float d0, d1;
            const float g0 = gmof(gt_uv[2 * o] - uv[0], rho2, &d0);
            const float g1 = gmof(gt_uv[2 * o + 1] - uv[1], rho2, &d1);
            S.lterm[idx] = w2 * g0 + w2 * g1;
            const float duv[2] = {-(w2 * d0) * dw2, -(w2 * d1) * dw2};
            float dqv[3] = {0.f, 0.f, 0.f};
            project_bwd(cams.cam[v], xc, duv, dqv);
            S.contrib[3 * idx] = dqv[0]; S.contrib[3 * idx + 1] = dqv[1]; S.contrib[3 * idx + 2] = dqv[2];
        }
        __syncthreads();
        PHASE_MARK(8);
        if (t < K) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int v = 0; v < V; ++v) {
                a0 += S.contrib[3 * (v * K + t)]; a1 += S.contrib[3 * (v * K + t) + 1]; a2 += S.contrib[3 * (v * K + t) + 2];
            }
            S.dq[3 * t] = a0; S.dq[3 * t + 1] = a1; S.dq[3 * t + 2] = a2;
        } else if (t >= 32 && t < 32 + V) {
            const int v = t - 32;
            float a = 0.f;
            for (int k = 0; k < K; ++k) a += S.lterm[v * K + k];
            S.vsum[v] = a * dw2;
        }
        __syncthreads();
        PHASE_MARK(9);
        if (t == 0) {
            float a = 0.f;
            for (int v = 0; v < V; ++v) a += S.vsum[v];
            S.sc[0] = a;
        }
        if (t >= 32 && t < 35) {
            float a = 0.f;
            for (int k = 0; k < K; ++k) a += S.dq[3 * k + t - 32];
            S.grad[kOffTransl + t - 32] = a;
        }
        if (t >= 64 && t < 64 + kJoints) {
            const int j = t - 64;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int k = 0; k < K; ++k)
                if (m.kp_chain[k] == j) { a0 += S.dq[3 * k]; a1 += S.dq[3 * k + 1]; a2 += S.dq[3 * k + 2]; }
            S.dg[3 * j] = a0; S.dg[3 * j + 1] = a1; S.dg[3 * j + 2] = a2;
        }
        if (t >= 96 && t < 96 + nsup) {
            const int i = t - 96;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int e = m.sup_ptr[i]; e < m.sup_ptr[i + 1]; ++e) {
                const int k = m.sup_k[e];
                const float w = m.sup_w[e];
                a0 = fmaf(w, S.dq[3 * k], a0); a1 = fmaf(w, S.dq[3 * k + 1], a1); a2 = fmaf(w, S.dq[3 * k + 2], a2);
            }
            S.dv[3 * i] = a0; S.dv[3 * i + 1] = a1; S.dv[3 * i + 2] = a2;
        }
    }
    __syncthreads();
    PHASE_MARK(10);
    if (have_grad) {
        // ---- P7 adjoint of skinning: dvp = T3x3^T dv ; dA_j = sum_i W[i,j] [dv (x) vp | dv]
        const int nx = din.n_extra, ncol_all = ncol + 3 * nx;
        if (t < nsup + nx) {
            const int n = t < nsup ? m.sup[t] : din.extra_n[t - nsup];
            float G[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) G[c] = 0.f;
            for (int e = 0; e < m.KW; ++e) {
                const float w = m.ell_w[(size_t)n * m.KW + e];
                if (w != 0.f) {
                    const float* Aj = &S.A[12 * m.ell_j[(size_t)n * m.KW + e]];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) G[3 * r + c] = fmaf(w, Aj[4 * r + c], G[3 * r + c]);
                }
            }
            const float d0 = S.dv[3 * t], d1 = S.dv[3 * t + 1], d2 = S.dv[3 * t + 2];
            S.dvp[3 * t] = G[0] * d0 + G[3] * d1 + G[6] * d2;
            S.dvp[3 * t + 1] = G[1] * d0 + G[4] * d1 + G[7] * d2;
            S.dvp[3 * t + 2] = G[2] * d0 + G[5] * d1 + G[8] * d2;
        }
        for (int e = t; e < kSkinFloats; e += kResThreads) {
            const int j = e / 12, r = (e % 12) / 4, c = e % 4;
            float a = 0.f;
            if (S.sj_on) {
#pragma unroll 4
                for (int q2 = S.sj_ptr[j]; q2 < S.sj_ptr[j + 1]; ++q2) {
                    const int i = S.sj_i[q2];
                    const float wd = S.sj_w[q2] * S.dv[3 * i + r];
                    a = (c < 3) ? fmaf(wd, S.vp[3 * i + c], a) : a + wd;
                }
            } else {
#pragma unroll 4
                for (int q2 = m.supj_ptr[j]; q2 < m.supj_ptr[j + 1]; ++q2) {
                    const int i = m.supj_i[q2];
                    const float wd = m.supj_w[q2] * S.dv[3 * i + r];
                    a = (c < 3) ? fmaf(wd, S.vp[3 * i + c], a) : a + wd;
                }
            }
            for (int k = 0; k < nx; ++k) {                     // box-extreme vertices: generic (dense-W) joint ownership
                const float w = __ldg(din.Wd + (size_t)din.extra_n[k] * kJoints + j);
                if (w != 0.f) {
                    const float wd = w * S.dv[3 * (nsup + k) + r];
                    a = (c < 3) ? fmaf(wd, S.vp[3 * (nsup + k) + c], a) : a + wd;
                }
            }
            if (din.part) {                                    // penetration gradient: unit-factor block partials
                float pa = 0.f;                                // fixed summation order (ascending block index)
                for (int q = 0; q < din.nfl; ++q) pa += din.part[(size_t)din.fl[q] * kPartFloats + e];
                a = fmaf(pa, din.factor, a);
            }
            S.dA[e] = a;
        }
        __syncthreads();
        PHASE_MARK(11);
        // ---- P8 dPhi[k] = sum_col dvp[col] Qk[row(col)][k]   (thread per k, coalesced rows)
        if (t < kFeatPad) {
            float a = 0.f;
            int col = 0;
            for (; col + 32 <= ncol_all; col += 32) {          // 32 independent L2 loads in flight per thread: the phase is as long
                float qv[32];                                  // as its round trips to L2 (258 rows: 8 instead of 21)
#pragma unroll
                for (int u = 0; u < 32; ++u) qv[u] = __ldg(m.Qk + (size_t)S.rowbase[col + u] * kFeatPad + t);
#pragma unroll
                for (int u = 0; u < 32; ++u) a = fmaf(S.dvp[col + u], qv[u], a);
            }
            for (; col + 8 <= ncol_all; col += 8) {
                float qv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) qv[u] = __ldg(m.Qk + (size_t)S.rowbase[col + u] * kFeatPad + t);
#pragma unroll
                for (int u = 0; u < 8; ++u) a = fmaf(S.dvp[col + u], qv[u], a);
            }
            for (; col < ncol_all; ++col) a = fmaf(S.dvp[col], __ldg(m.Qk + (size_t)S.rowbase[col] * kFeatPad + t), a);
            if (din.part) {
                float pa = 0.f;
                for (int q = 0; q < din.nfl; ++q) pa += din.part[(size_t)din.fl[q] * kPartFloats + kSkinFloats + t];
                a = fmaf(pa, din.factor, a);
            }
            S.dPhi[t] = a;
        }
        __syncthreads();
        PHASE_MARK(12);
        // ---- P9 adjoint of the skinning transforms
        if (t < kJoints)
            skin_transform_bwd(&S.dA[12 * t], &S.Gam[9 * t], &S.J[3 * t], &S.dGam[9 * t], &S.dg[3 * t], &S.dJ[3 * t]);
        __syncthreads();
        PHASE_MARK(14);
        // reverse sweep over the tree, level-parallel
        chain_bwd_levels(S);
        if (warp == 0) {
            const float sc = S.x[kOffScale];
            if (lane < 9) S.dR[lane] = sc * S.dGam[lane];
            if (lane == 0) {
                float ds = 0.f;
#pragma unroll
                for (int i = 0; i < 9; ++i) ds = fmaf(S.dGam[i], S.R[i], ds);
                S.grad[kOffScale] = ds;
            }
            if (lane < 3) S.dJ[lane] += S.dg[lane];
        }
        __syncthreads();
        PHASE_MARK(15);
        // ---- P10 pose-feature adjoint, Rodrigues adjoint, shape gradient
        if (t < kPoseBasis) S.dR[9 + t] += S.dPhi[t];
        __syncthreads();
        PHASE_MARK(16);
        if (t < kJoints) {
            float dr[3] = {0.f, 0.f, 0.f};
            rodrigues_bwd(t == 0 ? &S.x[kOffOrient] : &theta[3 * (t - 1)], &S.dR[9 * t], dr);
            float* go = (vp2 && t > 0) ? &W->dth[3 * (t - 1)] : &S.grad[kOffOrient + 3 * t];     // decoded pose: adjoint goes on
            go[0] = dr[0]; go[1] = dr[1]; go[2] = dr[2];
        } else if (t >= 32 && t < 32 + kBetas) {
            const int l = t - 32;
            float a = S.dPhi[kPoseBasis + l];
            for (int jc = 0; jc < kJoints * 3; ++jc) a = fmaf(m.JS[jc * kBetas + l], S.dJ[jc], a);
            S.grad[kOffBetas + l] = a;
        }
    }
    // ---- P11 priors (fitting.py:327-350), all GMM components in parallel
    const float bpw = lp.body_pose_weight, bpw2 = bpw * bpw;
    const int M = m.M;
    if (!lp.use_vposer && lp.body_prior == MVS_PRIOR_GMM)
        for (int e = t; e < M * 69; e += kResThreads) S.gm_diff[e] = theta[e % 69] - m.gmm_means[e];
    if (vp2) S.red[t] = (t < kVpZ) ? S.x[kOffPose + t] * S.x[kOffPose + t] : 0.f;      // |z|^2 (fitting.py:327-329)
    else S.red[t] = (t < 69) ? theta[t] * theta[t] : 0.f;
    __syncthreads();
    PHASE_MARK(17);
    if (!lp.use_vposer && lp.body_prior == MVS_PRIOR_GMM) {
        for (int e = t; e < M * 69; e += kResThreads) {
            const int mm = e / 69, i = e % 69;
            const float* P = m.gmm_prec + (size_t)mm * 69 * 69 + i;
            const float* df = &S.gm_diff[mm * 69];
            float y = 0.f;
#pragma unroll
            for (int j = 0; j < 69; ++j) y = fmaf(__ldg(P + j * 69), df[j], y);       // all 69 loads of the row in flight
            S.gm_y[e] = y;
        }
    }
    if (t == 0) { float a = 0.f; for (int i = 0; i < 69; ++i) a += S.red[i]; S.sc[1] = a; }
    __syncthreads();
    PHASE_MARK(18);
    if (!lp.use_vposer && lp.body_prior == MVS_PRIOR_GMM && warp < M) {
        float p = 0.f;
        for (int i = lane; i < 69; i += 32) p = fmaf(S.gm_y[warp * 69 + i], S.gm_diff[warp * 69 + i], p);
        p = warp_sum(p);
        if (lane == 0) S.gm_ll[warp] = 0.5f * p - m.gmm_lognllw[warp];
    }
    __syncthreads();
    PHASE_MARK(19);
    float pprior = 0.f, l2extra = 0.f;
    if (vp2) pprior = S.sc[1] * bpw2;
    if (!lp.use_vposer) {
        float gs = bpw2;
        int best_m = -1;
        if (lp.body_prior == MVS_PRIOR_GMM) {
            float best = 3.0e38f;
            for (int mm = 0; mm < M; ++mm) if (S.gm_ll[mm] < best) { best = S.gm_ll[mm]; best_m = mm; }
            pprior = best * bpw2;
        } else if (lp.body_prior == MVS_PRIOR_L2) {
            pprior = S.sc[1] * bpw2;
        }
        if (pprior > 5e4f) { pprior = 0.f; gs = 0.f; }
        const float w4 = (bpw * 4.f) * (bpw * 4.f);
        l2extra = S.sc[1] * w4;
        if (have_grad && t < 69) {
            float gth = 2.f * theta[t] * w4;
            if (lp.body_prior == MVS_PRIOR_GMM && best_m >= 0) gth = fmaf(gs, S.gm_y[best_m * 69 + t], gth);
            else if (lp.body_prior == MVS_PRIOR_L2) gth = fmaf(gs, 2.f * theta[t], gth);
            S.grad[kOffPose + t] += gth;
        }
    }
    float shape_loss = 0.f;
    if (!lp.fix_shape) {
        const float sw2 = lp.shape_weight * lp.shape_weight;
        float a = 0.f;
        for (int l = 0; l < kBetas; ++l) a = fmaf(S.x[l], S.x[l], a);
        shape_loss = a * sw2;
        if (have_grad && t >= 96 && t < 96 + kBetas) S.grad[kOffBetas + t - 96] += 2.f * S.x[t - 96] * sw2;
    }
    __syncthreads();
    PHASE_MARK(20);
    float angle = 0.f;
    {
        const int idx[4] = {52, 55, 9, 12};
        const float sg[4] = {1.f, -1.f, -1.f, -1.f};
        float ev[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float e = mvs_exp(theta[idx[i]] * sg[i]); ev[i] = e * e; angle += ev[i]; }
        angle *= lp.bending_prior_weight;
        float gs = lp.bending_prior_weight;
        if (angle > 1e4f && !lp.use_vposer) { angle = 0.f; gs = 0.f; }
        if (have_grad && t == 0) {
            float* gth = vp2 ? W->dth : &S.grad[kOffPose];
#pragma unroll
            for (int i = 0; i < 4; ++i) gth[idx[i]] += 2.f * sg[i] * ev[i] * gs;
        }
    }
    __syncthreads();
    PHASE_MARK(21);
    if (vp2 && have_grad) {               // d loss / d decoded pose -> d loss / d z, plus the prior's 2 bpw^2 z
        vposer_decode_bwd(m, *W);
        if (t < 69) S.grad[kOffPose + t] = (t < kVpZ) ? fmaf(2.f * bpw2, S.x[kOffPose + t], W->red[t]) : 0.f;
        __syncthreads();
    }
    float anchor_loss = 0.f;
    if (lp.anchor_on) {                                   // sequence mode: sum_i w_i (x_i - a_i)^2
        float dif = 0.f, wt = 0.f;
        if (t < kParams) { dif = S.x[t] - m.anchor[(size_t)b * kParams + t]; wt = m.anchor_w[(size_t)b * kParams + t]; }
        S.red[t] = wt * dif * dif;
        if (have_grad && t < kParams) S.grad[t] += 2.f * wt * dif;
        __syncthreads();
        if (t == 0) { float a = 0.f; for (int i = 0; i < kParams; ++i) a += S.red[i]; S.sc[3] = a; }
        __syncthreads();
        anchor_loss = S.sc[3];
    }
    // ---- P12 total (same order as fitting.py:411-413) and the masked gradient
    if (t == 0) {
        float total = S.sc[0];
        total += (pprior + l2extra);
        total += shape_loss;
        total += angle;
        if (din.verts) total += din.pen_loss;
        total += anchor_loss;
        S.sc[2] = total;
    }
    if (have_grad) {
        for (int i = t; i < kParams; i += kResThreads) {
            const int seg = i < kOffOrient ? 0 : i < kOffPose ? 1 : i < kOffTransl ? 2 : i < kOffScale ? 3 : 4;
            S.lg_new[i] = ((lp.frozen_mask >> seg) & 1u) ? 0.f : S.grad[i];
        }
    }
    __syncthreads();
    PHASE_MARK(22);
}

__device__ __forceinline__ void resident_setup(ResidentSmem& S, const ResidentModel& m) {
    for (int col = threadIdx.x; col < 3 * m.nsup; col += kResThreads) S.rowbase[col] = 3 * m.sup[col / 3] + col % 3;
    const int t = threadIdx.x;
    if (t < kJoints) { S.lev_j[t] = m.cs.lev_j[t]; S.ch_j[t] = m.cs.ch_j[t]; S.par[t] = m.par.p[t]; }
    else if (t >= 32 && t < 32 + kJoints + 1) { S.lev_ptr[t - 32] = m.cs.lev_ptr[t - 32]; S.ch_ptr[t - 32] = m.cs.ch_ptr[t - 32]; }
    else if (t == 64) S.nlev = m.cs.nlev;
    const int nsj = m.supj_ptr[kJoints];
    if (t == 65) S.sj_on = nsj <= kResMaxSupJ ? 1 : 0;
    if (nsj <= kResMaxSupJ) {
        if (t >= 96 && t < 96 + kJoints + 1) S.sj_ptr[t - 96] = m.supj_ptr[t - 96];
        for (int q = t; q < nsj; q += kResThreads) { S.sj_i[q] = m.supj_i[q]; S.sj_w[q] = m.supj_w[q]; }
    }
}

// A frame that finished stage k of a multi-stage run starts stage k + 1 at once, with a fresh optimiser (the reference
// builds one optimiser per stage, fitting.py / fit_single_frame) and the parameters it has: frames are independent
// problems, so nothing makes a fast frame wait for the slowest one at a stage boundary.
__device__ __forceinline__ void next_stage_scalars(FrameScalars& s) {
    const long long it = s.iters, ev = s.evals;
    const int st = s.stage + 1, nacc = s.nan_acc + s.nan_flag, end = s.stage_end;
    memset(&s, 0, sizeof(s));
    s.stage_end = end;
    s.H_diag = 1.f;
    s.phase = PH_STEP_ENTRY;
    s.final_loss = __int_as_float(0x7fc00000);
    s.iters = it; s.evals = ev; s.stage = st; s.nan_acc = nacc;
}

// Pose forward of the trial point in S.x for the dense kernels of the next round: feature row (fp32 and TF32), skinning
// transforms (frame fastest), translation -- all addressed by the slot -- and the per-frame pose cache.  With vp2 the
// body pose is decoded from the latent code first.  All threads; no trailing barrier.
__device__ __forceinline__ void next_pose_forward(ResidentSmem& S, const ResidentModel& m, VposerSmem* W, bool vp2, int slot, int b,
                                                  float* __restrict__ Phi, float* __restrict__ PhiTc, float* __restrict__ At,
                                                  int ldA, float* __restrict__ slot_tr, float* __restrict__ pose_cache,
                                                  int* __restrict__ pose_valid) {
    const int t = threadIdx.x;
    if (vp2) vposer_decode(S, m, *W);
    const float* theta = vp2 ? W->th : &S.x[kOffPose];
    if (t < kJoints) rodrigues_fwd(t == 0 ? &S.x[kOffOrient] : &theta[3 * (t - 1)], &S.R[9 * t]);
    else if (t >= 32 && t < 32 + 72) {
        const int jc = t - 32;
        float a = m.Jt[jc];
#pragma unroll
        for (int l = 0; l < kBetas; ++l) a = fmaf(m.JS[jc * kBetas + l], S.x[kOffBetas + l], a);
        S.J[jc] = a;
    }
    __syncthreads();
    PHASE_MARK(26);
    chain_fwd_levels(S);
    PHASE_MARK(27);
    if (t < kJoints) {
        float* A = &S.A[12 * t];
        make_skin_transform(&S.Gam[9 * t], &S.g[3 * t], &S.J[3 * t], A);
#pragma unroll
        for (int c = 0; c < 12; ++c) At[(size_t)(t * 12 + c) * ldA + slot] = A[c];
        if (t < 3) slot_tr[4 * slot + t] = S.x[kOffTransl + t];
    } else if (t >= 32) {
        const int k = t - 32;
        float v;
        if (k < kPoseBasis) v = S.R[9 + k] - (((k % 9) % 4 == 0) ? 1.0f : 0.0f);
        else if (k < kPoseBasis + kBetas) v = S.x[kOffBetas + k - kPoseBasis];
        else v = (k == kFeat - 1) ? 1.0f : 0.0f;
        Phi[(size_t)slot * kFeatPad + k] = v;
        if (PhiTc) {                 // A operand of the tensor-core contraction: v = hi + lo, both TF32 (tf32_split)
            float hi = 0.f, lo = 0.f;
            if (k < kPoseBasis) tf32_split(v, hi, lo);
            PhiTc[(size_t)slot * kFeatPad + k] = hi;
            PhiTc[((size_t)ldA + slot) * kFeatPad + k] = lo;
        }
    }
    __syncthreads();
    if (pose_cache) {   // park R | J | Gam | g | A for the next round's closure adjoint
        const float* src = &S.R[0];
        float* dst = pose_cache + (size_t)b * kPoseCacheFloats;
        for (int i = t; i < kPoseCacheFloats; i += kResThreads) dst[i] = src[i];
        if (t == 0) pose_valid[b] = 1;
    }
}

// Arguments of frame_step_body.  Everything except the model / camera / detection constants is rewritten every round by
// other CTAs of the persistent dense-round kernel: plain pointers.
struct FrameStepArgs {
    ResidentModel m; CamSet cams; const LossParams* lp_tab; int nstages; LbfgsCfg cfg; LbfgsState L; float* params;
    const float* gt_uv; const float* conf; const float* joint_w; int B, N;
    const float* verts_ws; const float* poffT; const float* ST;
    const float* parts5; const float* part; const int* pflag; const FrameBox* box; const float* Wd;
    float* Phi; float* PhiTc; float* At; int ldA; float* slot_tr; float* pose_cache; int* pose_valid; int with_vposer;
};

// One round of the dense regime for one frame: consumes the dense vertices + the SDF gradient list of this
// frame's trial point, finishes the closure (keypoint term, adjoint, priors), advances the frame's L-BFGS state
// machine and -- if the frame needs another evaluation -- runs the pose forward of the NEXT trial point and
// writes its feature row / skinning transforms for the next dense vertex launch.
// `smem_raw`: sizeof(ResidentSmem) + 2 H 86 floats (+ VposerSmem).  All kResThreads threads call; no trailing barrier.
__device__ __forceinline__ void frame_step_body(unsigned char* smem_raw, const FrameStepArgs& a, const int slot, const int b,
                                                int* na_next = nullptr, int* fidx_next = nullptr) {
    ResidentSmem& S = *reinterpret_cast<ResidentSmem*>(smem_raw);
    const ResidentModel& m = a.m; const CamSet& cams = a.cams; const LossParams* lp_tab = a.lp_tab; const int nstages = a.nstages;
    const LbfgsCfg& cfg = a.cfg; const LbfgsState& L = a.L; float* params = a.params;
    const float* __restrict__ gt_uv = a.gt_uv; const float* __restrict__ conf = a.conf; const float* __restrict__ joint_w = a.joint_w;
    const int B = a.B, N = a.N, ldA = a.ldA, with_vposer = a.with_vposer;
    const float* verts_ws = a.verts_ws; const float* parts5 = a.parts5; const float* part = a.part; const int* pflag = a.pflag;
    const FrameBox* box = a.box; const float* __restrict__ Wd = a.Wd;
    float* Phi = a.Phi; float* PhiTc = a.PhiTc; float* At = a.At; float* slot_tr = a.slot_tr; float* pose_cache = a.pose_cache;
    int* pose_valid = a.pose_valid;
    const FrameScalars fs0 = L.sc[b];
    if (fs0.phase == PH_DONE) return;
    PHASE_MARK(0);
    const int nparts = (N + 255) / 256;                         // sdf_fused_kernel emits per 256-vertex block
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    resident_setup(S, m);
    if (t < (int)(sizeof(LossParams) / 4))          // the loss parameters of the stage THIS frame is in
        reinterpret_cast<int*>(&S.lp)[t] = reinterpret_cast<const int*>(&lp_tab[fs0.stage])[t];
    float* x_eval = L.x_eval + (size_t)b * kParams;
    for (int i = t; i < kParams; i += kResThreads) S.x[i] = x_eval[i];
    // Stage this frame's curvature history (<= 2 x 100 x 86 floats) into shared memory with cp.async while the
    // closure runs: the two-loop recursion is ~200 DEPENDENT dot products, at L2 latency it costs ~80 us, from
    // shared memory ~5 us.
    float* hy = reinterpret_cast<float*>(smem_raw + sizeof(ResidentSmem));
    float* hs = hy + (size_t)L.H * kParams;
    VposerSmem* W = with_vposer ? reinterpret_cast<VposerSmem*>(hs + (size_t)L.H * kParams) : nullptr;
    {
        const int hl = fs0.hist_len;
        const float* gy = L.hist_y + (size_t)b * L.H * kParams;
        const float* gs = L.hist_s + (size_t)b * L.H * kParams;
        // ring buffer: with hl < H the live rows are [0, hl); when full all rows are live.  8-byte packets
        // (rows are 344 B: 8-byte aligned)
        const int total = hl * kParams / 2;
        for (int i = t; i < total; i += kResThreads) {
            const unsigned sy = (unsigned)__cvta_generic_to_shared(hy + 2 * i);
            const unsigned ss = (unsigned)__cvta_generic_to_shared(hs + 2 * i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sy), "l"(gy + 2 * i));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ss), "l"(gs + 2 * i));
        }
        for (int i = t; i < hl; i += kResThreads) S.ro[i] = L.ro[(size_t)b * L.H + i];
        {   // block Gram of the live part of the ring
            const float* gg = L.gram + (size_t)b * kGramFloats;
            const int nfl4 = ((hl == L.H ? L.H : hl) + 7) / 8 * 16;      // float4 packets
            for (int i = t; i < nfl4; i += kResThreads) {
                const unsigned sg = (unsigned)__cvta_generic_to_shared(S.gram + 4 * i);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sg), "l"(gg + 4 * i));
            }
        }
        asm volatile("cp.async.commit_group;");
        // the optimiser's seven 86-vectors: one coalesced read now, one write-back after the step (every dot
        // product / axpy of the state machine then runs out of shared memory instead of ~30 dependent L2 round trips)
        for (int i = t; i < kParams; i += kResThreads) {
            S.lx[i] = params[(size_t)b * kParams + i];
            S.lg[i] = L.g[(size_t)b * kParams + i];
            S.ld[i] = L.d[(size_t)b * kParams + i];
            S.lprev_g[i] = L.prev_g[(size_t)b * kParams + i];
            S.lx_init[i] = L.x_init[(size_t)b * kParams + i];
            S.lg_prev[i] = L.g_prev[(size_t)b * kParams + i];
            S.lbg0[i] = L.bg[(size_t)b * 2 * kParams + i];
            S.lbg1[i] = L.bg[(size_t)b * 2 * kParams + kParams + i];
        }
    }
    // pose forward of THIS trial point: computed by the previous round's frame_step (for the dense kernels), restored here
    // (loaded unconditionally, next to the validity flag instead of behind it: an invalid cache is simply overwritten
    // by P1-P3 of the closure)
    const bool pose_ready = pose_valid[b] != 0;
    {
        float* dst = &S.R[0];                         // R | J | Gam | g | A are contiguous (864 floats)
        const float* src = pose_cache + (size_t)b * kPoseCacheFloats;
        for (int i = t; i < kPoseCacheFloats; i += kResThreads) dst[i] = src[i];
    }
    if (warp == 7) {
        // frame scalars of the penetration term (fitting.py:386-392): total of the sampled values -> loss, the
        // factor of the listed vertices' adjoint, and the <= 6 box-extreme vertices that carry the gradient through
        // the box centre (mean of min / max vertex) and scale (0.6 x the largest extent)
        float a = 0.f;
        if (lane < 5)
            for (int p = 0; p < nparts; ++p) a += parts5[((size_t)slot * nparts + p) * 5 + lane];
        float tot[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) tot[q] = __shfl_sync(0xffffffffu, a, q);
        if (lane == 0) {
            const FrameBox fb = box[slot];
            const float coll_w = lp_tab[fs0.stage].coll_loss_weight;
            const float wsum = coll_w * tot[0];                 // coll_loss_weight * cur_loss.sum() / 1
            const float cg = 2.f * wsum * coll_w;               // d pen / d (sum of samples)
            const float inv_s = 1.f / fb.scale;
            int cnt = 0;
            if (cg != 0.f) {
                const float dscale = -cg * tot[4] * inv_s;      // local = (v - c) / s  ->  d local / d s = -local / s
                for (int c = 0; c < 3; ++c) {
                    const float dcentre = -cg * tot[1 + c] * inv_s;
                    float dl = 0.5f * dcentre, dh = 0.5f * dcentre;
                    if (c == fb.cmax) { dh += 0.6f * dscale; dl -= 0.6f * dscale; }
                    if (dl != 0.f) { S.ext_n[cnt] = fb.ilo[c]; S.ext_d[3 * cnt] = 0.f; S.ext_d[3 * cnt + 1] = 0.f; S.ext_d[3 * cnt + 2] = 0.f; S.ext_d[3 * cnt + c] = dl; ++cnt; }
                    if (dh != 0.f) { S.ext_n[cnt] = fb.ihi[c]; S.ext_d[3 * cnt] = 0.f; S.ext_d[3 * cnt + 1] = 0.f; S.ext_d[3 * cnt + 2] = 0.f; S.ext_d[3 * cnt + c] = dh; ++cnt; }
                }
            }
            S.sdf_sc[0] = cg * inv_s;
            S.sdf_sc[1] = wsum * wsum;                          // fitting.py:391-392
            S.sdf_sc[2] = (float)cnt;
        }
        // blocks whose vertices carry a sample gradient (ascending), so the adjoint phases loop over those only
        int nfl = 0;
        for (int p0 = 0; p0 < nparts; p0 += 32) {
            const int p = p0 + lane;
            const bool f = p < nparts && pflag[(size_t)slot * nparts + p] != 0;
            const unsigned mk = __ballot_sync(0xffffffffu, f);
            if (f) { const int pos = nfl + __popc(mk & ((1u << lane) - 1u)); if (pos < 64) S.fl[pos] = p; }
            nfl += __popc(mk);
        }
        if (lane == 0) S.nfl = nfl < 64 ? nfl : 64;
    }
    __syncthreads();
    PHASE_MARK(1);
    DenseIn din;
    din.verts = verts_ws + (size_t)slot * N * 3;
    din.poffT = a.poffT; din.ST = a.ST; din.ldA = ldA; din.slot = slot;
    din.extra_n = S.ext_n;
    din.extra_d = S.ext_d;
    din.n_extra = (int)S.sdf_sc[2];
    din.pen_loss = S.sdf_sc[1];
    din.Wd = Wd;
    din.factor = S.sdf_sc[0];
    din.part = (din.factor != 0.f && S.nfl > 0) ? part + (size_t)slot * nparts * kPartFloats : nullptr;
    din.fl = S.fl;
    din.nfl = S.nfl;
    din.pose_ready = pose_ready;
    resident_closure(S, m, cams, S.lp, gt_uv, conf, joint_w, B, b, true, nullptr, nullptr, din, W);
    for (int i = t; i < kParams; i += kResThreads) L.g_eval[(size_t)b * kParams + i] = S.lg_new[i];
    asm volatile("cp.async.wait_all;");
    __syncthreads();
    PHASE_MARK(23);
    if (cfg.step_mode == 2) {                          // closure only (mvs_closure in exec mode 3): no optimiser step
        if (t == 0) L.loss_eval[b] = S.sc[2];
        return;
    }
    if (warp == 0) {
        FrameScalars s = fs0;
        LbfgsPtrs P{S.lx, S.lg, S.ld, S.lprev_g, S.lx_init, S.lg_prev, S.lbg0, S.lbg1, hy, hs, S.ro, S.al, S.lx_eval,
                    S.lg_new, L.H, S.gram, S.tl_scratch};
        lbfgs_advance_core(s, P, S.sc[2], cfg, lane, S.lp.use_vposer == 2 ? 32 : kOffTransl - kOffPose);
        __syncwarp();
        if (s.phase == PH_DONE && s.stage + 1 < s.stage_end) {  // this frame moves on to its next stage
            next_stage_scalars(s);
            VLOOP(i) S.lx_eval[i] = S.lx[i];
            __syncwarp();
        }
        VLOOP(i) {
            params[(size_t)b * kParams + i] = S.lx[i];
            L.g[(size_t)b * kParams + i] = S.lg[i];
            L.d[(size_t)b * kParams + i] = S.ld[i];
            L.prev_g[(size_t)b * kParams + i] = S.lprev_g[i];
            L.x_init[(size_t)b * kParams + i] = S.lx_init[i];
            L.g_prev[(size_t)b * kParams + i] = S.lg_prev[i];
            L.bg[(size_t)b * 2 * kParams + i] = S.lbg0[i];
            L.bg[(size_t)b * 2 * kParams + kParams + i] = S.lbg1[i];
            x_eval[i] = S.lx_eval[i];
        }
        if (s.pushed_slot >= 0) {                       // write the new curvature pair through to global memory
            const int wslot = s.pushed_slot;
            VLOOP(i) {
                L.hist_y[((size_t)b * L.H + wslot) * kParams + i] = hy[(size_t)wslot * kParams + i];
                L.hist_s[((size_t)b * L.H + wslot) * kParams + i] = hs[(size_t)wslot * kParams + i];
            }
            if (lane == 0) L.ro[(size_t)b * L.H + wslot] = S.ro[wslot];
            float* gg = L.gram + (size_t)b * kGramFloats + (wslot >> 3) * 64;
            gg[lane] = S.gram[(wslot >> 3) * 64 + lane]; gg[lane + 32] = S.gram[(wslot >> 3) * 64 + lane + 32];
        }
        if (lane == 0) { L.sc[b] = s; S.fs = s; }
    }
    __syncthreads();
    PHASE_MARK(24);
    if (S.fs.phase == PH_DONE) return;
    // pose forward of the next trial point -> Phi row and skinning transforms for the next vertex launch.  The persistent
    // dense-round kernel hands out next round's slots here (atomic counter = compaction every round; a frame's arithmetic
    // does not depend on the slot it lands in), the multi-kernel path keeps the slot until its next compaction.
    if (na_next) {
        if (t == 0) { const int ns = atomicAdd(na_next, 1); fidx_next[ns] = b; S.nfl = ns; }
        __syncthreads();
    }
    const int slot_next = na_next ? S.nfl : slot;
    for (int i = t; i < kParams; i += kResThreads) S.x[i] = S.lx_eval[i];
    __syncthreads();
    PHASE_MARK(25);
    {
        const bool vp2n = W != nullptr && lp_tab[S.fs.stage].use_vposer == 2;     // the stage the NEXT evaluation belongs to
        next_pose_forward(S, m, W, vp2n, slot_next, b, Phi, PhiTc, At, ldA, slot_tr, pose_cache, pose_valid);
    }
    PHASE_MARK(28);
}

// mvs_resident.cu (host)
FrameStepArgs make_frame_step_args(mvs_ctx* ctx, float* params_dev, const void* lbfgs_state, const void* lbfgs_cfg, int nstages);
size_t frame_step_smem(int H, bool with_vposer);

}  // namespace mvs
