// Frame-resident execution of the sparse regime (no vertices requested, no SDF term):
// ONE CTA owns ONE frame and runs the whole closure -- and, in lbfgs_resident_kernel, the whole
// L-BFGS / strong-Wolfe optimisation of that frame -- out of shared memory, without returning to
// the host or to other kernels between evaluations.
//
// Why: the loss of the sparse regime depends on only 86 of the 6890 vertices, so one
// closure is ~0.1 MFLOP + two passes over a 231 KB L2-resident slice of Qk.  Spread over five
// kernels per evaluation and ~900 rounds per fit, the batched path is bound by launch + dependent
// load latency (215 us per round measured, profiles/r01_ncu_summary.md).  Frames are independent
// optimisation problems, so nothing forces them into lock-step rounds: here every frame advances
// at its own pace, the optimiser state (7 vectors, 100-pair history = 69 KB) lives in shared
// memory, and a 256-frame stage is ONE kernel launch.
//
// The arithmetic is the same set of building blocks (mvs_math.cuh) in the same data flow as the
// batched kernels of mvs_closure.cu; only reduction orders differ (warp-tree vs serial sums).
#include "mvs_internal.cuh"
#include "mvs_lbfgs_core.cuh"

namespace mvs {

constexpr int kResThreads = 256;
constexpr int kResMaxSup = 96;            // support vertices (SMPL: 86)
constexpr int kResMaxVK = 272;            // views x keypoints (16 x 17)
constexpr int kResMaxM = 8;               // GMM components

struct ResidentModel {                    // device pointers + sizes, by value
    const float* Qk; const float* Jt; const float* JS;
    const int* ell_j; const float* ell_w; int KW;
    int K, nsup;
    const int* kp_ptr; const int* kp_spos; const float* kp_w; const int* kp_chain;
    const int* sup; const int* sup_ptr; const int* sup_k; const float* sup_w;
    const int* supj_ptr; const int* supj_i; const float* supj_w;      // per joint: (support index, weight)
    int M; const float* gmm_means; const float* gmm_prec; const float* gmm_lognllw;
    const float* anchor; const float* anchor_w;     // sequence mode (mvs_set_anchor)
    Parents par;
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
    FrameScalars fs;
};

// Inputs of the dense regime (SDF term on): the dense vertex kernel already produced every vertex of the frame,
// and the SDF kernel a compact list of vertices with a non-zero penetration gradient.
struct DenseIn {
    const float* vposed;      // [N][3] this frame's v_posed (NULL: sparse regime, recompute the support vertices)
    const float* verts;       // [N][3] skinned, pre-transl
    const int* extra_n;       // [n_extra] vertex ids
    const float* extra_d;     // [n_extra][3] d pen / d vertex
    int n_extra;
    float pen_loss;
    const float* Wd;          // [N][24] dense skinning weights (adjoint of the extra vertices)
    const float* part;        // [strip][ldA][512] partial adjoints of the dense SDF gradient (vertex_bwd), or NULL
    int nstrips, ldA, slot;
    const int* strip_active;  // [strip][ftiles]
    int ftiles;
};

// ------------------------------------------------------------------------------------------------
// One closure evaluation at S.x (already loaded).  Writes S.sc[2] = total loss and S.lg_new = gradient
// (frozen segments zeroed).  All kResThreads threads must call it.
__device__ void resident_closure(ResidentSmem& S, const ResidentModel& m, const CamSet& cams, const LossParams& lp,
                                 const float* __restrict__ gt_uv, const float* __restrict__ conf,
                                 const float* __restrict__ joint_w, int B, int b, bool have_grad,
                                 float* __restrict__ joints_out, float* __restrict__ proj_out, const DenseIn& din) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nsup = m.nsup, K = m.K, V = cams.num_views, ncol = 3 * nsup;

    // ---- P1 Rodrigues, rest joints
    if (t < kJoints) rodrigues_fwd(&S.x[kOffOrient + 3 * t], &S.R[9 * t]);
    else if (t >= 32 && t < 32 + 72) {
        const int jc = t - 32;
        float a = m.Jt[jc];
#pragma unroll
        for (int l = 0; l < kBetas; ++l) a = fmaf(m.JS[jc * kBetas + l], S.x[kOffBetas + l], a);
        S.J[jc] = a;
    }
    for (int i = t; i < kJoints * 9; i += kResThreads) { S.dGam[i] = 0.f; S.dR[i] = 0.f; }
    for (int i = t; i < kJoints * 3; i += kResThreads) { S.dg[i] = 0.f; S.dJ[i] = 0.f; }
    for (int i = t; i < kParams; i += kResThreads) S.grad[i] = 0.f;
    __syncthreads();
    // ---- P2 kinematic chain (warp 0, 12 lanes per joint, joints in index order: parents[j] < j)
    if (warp == 0) {
        const float sc = S.x[kOffScale];
        if (lane < 9) S.Gam[lane] = sc * S.R[lane];                               // lbs.py:348
        if (lane < 3) S.g[lane] = S.J[lane];
        __syncwarp();
        for (int j = 1; j < kJoints; ++j) {
            const int p = m.par.p[j];
            if (lane < 12) {
                const float* Gp = &S.Gam[9 * p];
                if (lane < 9) {
                    const int r = lane / 3, c = lane % 3;
                    const float* Rj = &S.R[9 * j];
                    S.Gam[9 * j + lane] = Gp[3 * r] * Rj[c] + Gp[3 * r + 1] * Rj[3 + c] + Gp[3 * r + 2] * Rj[6 + c];
                } else {
                    const int r = lane - 9;
                    const float r0 = S.J[3 * j] - S.J[3 * p], r1 = S.J[3 * j + 1] - S.J[3 * p + 1], r2 = S.J[3 * j + 2] - S.J[3 * p + 2];
                    S.g[3 * j + r] = (Gp[3 * r] * r0 + Gp[3 * r + 1] * r1 + Gp[3 * r + 2] * r2) + S.g[3 * p + r];
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
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
    // ---- P4 v_posed for the support columns: warp per Qk row, 2 x LDG.128 per lane
    if (din.vposed) {                 // dense regime: the vertex kernel already has them
        for (int col = t; col < ncol; col += kResThreads) {
            const int n = m.sup[col / 3];
            S.vp[col] = din.vposed[3 * n + col % 3];
            S.v[col] = din.verts[3 * n + col % 3];
        }
    } else {
        const float4 ph0 = *reinterpret_cast<const float4*>(&S.Phi[4 * lane]);
        const float4 ph1 = lane < 24 ? *reinterpret_cast<const float4*>(&S.Phi[128 + 4 * lane]) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int col0 = warp; col0 < ncol; col0 += 2 * (kResThreads / 32)) {
            const int col1 = col0 + kResThreads / 32;
            const float* r0 = m.Qk + (size_t)S.rowbase[col0] * kFeatPad;
            const float* r1 = m.Qk + (size_t)S.rowbase[col1 < ncol ? col1 : col0] * kFeatPad;
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(r0 + 4 * lane));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(r1 + 4 * lane));
            float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = a1;
            if (lane < 24) {
                a1 = __ldg(reinterpret_cast<const float4*>(r0 + 128 + 4 * lane));
                b1 = __ldg(reinterpret_cast<const float4*>(r1 + 128 + 4 * lane));
            }
            float p0 = a0.x * ph0.x;
            p0 = fmaf(a0.y, ph0.y, p0); p0 = fmaf(a0.z, ph0.z, p0); p0 = fmaf(a0.w, ph0.w, p0);
            p0 = fmaf(a1.x, ph1.x, p0); p0 = fmaf(a1.y, ph1.y, p0); p0 = fmaf(a1.z, ph1.z, p0); p0 = fmaf(a1.w, ph1.w, p0);
            float p1 = b0.x * ph0.x;
            p1 = fmaf(b0.y, ph0.y, p1); p1 = fmaf(b0.z, ph0.z, p1); p1 = fmaf(b0.w, ph0.w, p1);
            p1 = fmaf(b1.x, ph1.x, p1); p1 = fmaf(b1.y, ph1.y, p1); p1 = fmaf(b1.z, ph1.z, p1); p1 = fmaf(b1.w, ph1.w, p1);
            p0 = warp_sum(p0);
            p1 = warp_sum(p1);
            if (lane == 0) { S.vp[col0] = p0; if (col1 < ncol) S.vp[col1] = p1; }
        }
    }
    __syncthreads();
    // ---- P5 linear blend skinning of the support vertices
    if (!din.vposed && t < nsup) {
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
    if (have_grad) {
        // ---- P7 adjoint of skinning: dvp = T3x3^T dv ; dA_j = sum_i W[i,j] [dv (x) vp | dv]
        if (t < nsup) {
            const int n = m.sup[t];
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
            for (int q2 = m.supj_ptr[j]; q2 < m.supj_ptr[j + 1]; ++q2) {
                const int i = m.supj_i[q2];
                const float wd = m.supj_w[q2] * S.dv[3 * i + r];
                a = (c < 3) ? fmaf(wd, S.vp[3 * i + c], a) : a + wd;
            }
            S.dA[e] = a;
        }
        __syncthreads();
        // ---- P8 dPhi[k] = sum_col dvp[col] Qk[row(col)][k]   (thread per k, coalesced rows)
        if (t < kFeatPad) {
            float a = 0.f;
            int col = 0;
            for (; col + 4 <= ncol; col += 4) {
                const float q0 = __ldg(m.Qk + (size_t)S.rowbase[col] * kFeatPad + t);
                const float q1 = __ldg(m.Qk + (size_t)S.rowbase[col + 1] * kFeatPad + t);
                const float q2 = __ldg(m.Qk + (size_t)S.rowbase[col + 2] * kFeatPad + t);
                const float q3 = __ldg(m.Qk + (size_t)S.rowbase[col + 3] * kFeatPad + t);
                a = fmaf(S.dvp[col], q0, a); a = fmaf(S.dvp[col + 1], q1, a);
                a = fmaf(S.dvp[col + 2], q2, a); a = fmaf(S.dvp[col + 3], q3, a);
            }
            for (; col < ncol; ++col) a = fmaf(S.dvp[col], __ldg(m.Qk + (size_t)S.rowbase[col] * kFeatPad + t), a);
            S.dPhi[t] = a;
        }
        __syncthreads();
        // ---- P8b dense regime: vertices with a penetration gradient, in chunks of kResMaxSup (reusing vp/dv/dvp);
        //      same adjoint as P7/P8 with generic (dense-W) joint ownership -- deterministic, no atomics
        for (int e0 = 0; e0 < din.n_extra; e0 += kResMaxSup) {
            const int cnt = min(kResMaxSup, din.n_extra - e0);
            for (int col = t; col < 3 * cnt; col += kResThreads) {
                const int n = din.extra_n[e0 + col / 3];
                S.rowbase[col] = 3 * n + col % 3;
                S.vp[col] = din.vposed[3 * n + col % 3];
                S.dv[col] = din.extra_d[(size_t)(e0 + col / 3) * 3 + col % 3];
            }
            __syncthreads();
            if (t < cnt) {
                const int n = din.extra_n[e0 + t];
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
                for (int i = 0; i < cnt; ++i) {
                    const float w = __ldg(din.Wd + (size_t)din.extra_n[e0 + i] * kJoints + j);
                    if (w != 0.f) {
                        const float wd = w * S.dv[3 * i + r];
                        a = (c < 3) ? fmaf(wd, S.vp[3 * i + c], a) : a + wd;
                    }
                }
                S.dA[e] += a;
            }
            __syncthreads();
            if (t < kFeatPad) {
                float a = 0.f;
                for (int col = 0; col < 3 * cnt; ++col) a = fmaf(S.dvp[col], __ldg(m.Qk + (size_t)S.rowbase[col] * kFeatPad + t), a);
                S.dPhi[t] += a;
            }
            __syncthreads();
        }
        if (din.part) {                   // dense SDF gradient: partial adjoints from the batched vertex kernel
            for (int e = t; e < kPartFloats; e += kResThreads) {
                // fixed summation order (strip 0,1,2,...) with 8 loads in flight: the strips are independent L2 reads
                float a = 0.f;
                const float* pp = din.part + (size_t)din.slot * kPartFloats + e;
                const size_t stride = (size_t)din.ldA * kPartFloats;
                const int* fl = din.strip_active + din.slot / kTileF;
                int sidx = 0;
                for (; sidx + 8 <= din.nstrips; sidx += 8) {       // strips that met no active tile wrote nothing
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = fl[(sidx + u) * din.ftiles] ? pp[(size_t)(sidx + u) * stride] : 0.f;
#pragma unroll
                    for (int u = 0; u < 8; ++u) a += v[u];
                }
                for (; sidx < din.nstrips; ++sidx) a += fl[sidx * din.ftiles] ? pp[(size_t)sidx * stride] : 0.f;
                if (e < kSkinFloats) S.dA[e] += a; else S.dPhi[e - kSkinFloats] += a;
            }
            __syncthreads();
        }
        if (din.n_extra > 0) {            // restore the support-list row map for the next evaluation
            for (int col = t; col < ncol; col += kResThreads) S.rowbase[col] = 3 * m.sup[col / 3] + col % 3;
        }
        // ---- P9 adjoint of the skinning transforms
        if (t < kJoints)
            skin_transform_bwd(&S.dA[12 * t], &S.Gam[9 * t], &S.J[3 * t], &S.dGam[9 * t], &S.dg[3 * t], &S.dJ[3 * t]);
        __syncthreads();
        // reverse sweep over the tree: 24 lanes per joint
        if (warp == 0) {
            for (int j = kJoints - 1; j >= 1; --j) {
                const int p = m.par.p[j];
                const float* Gp = &S.Gam[9 * p];
                const float* dGj = &S.dGam[9 * j];
                const float* dgj = &S.dg[3 * j];
                float out = 0.f;
                if (lane < 9) {                // dGp += dGam_j R_j^T + dg_j rel^T
                    const int r = lane / 3, c = lane % 3;
                    const float* Rj = &S.R[9 * j];
                    const float rel = S.J[3 * j + c] - S.J[3 * p + c];
                    out = (dGj[3 * r] * Rj[3 * c] + dGj[3 * r + 1] * Rj[3 * c + 1] + dGj[3 * r + 2] * Rj[3 * c + 2]) + dgj[r] * rel;
                } else if (lane < 18) {        // dR_j = Gp^T dGam_j
                    const int e = lane - 9, r = e / 3, c = e % 3;
                    out = Gp[r] * dGj[c] + Gp[3 + r] * dGj[3 + c] + Gp[6 + r] * dGj[6 + c];
                } else if (lane < 21) {        // drel = Gp^T dg_j
                    const int r = lane - 18;
                    out = Gp[r] * dgj[0] + Gp[3 + r] * dgj[1] + Gp[6 + r] * dgj[2];
                } else if (lane < 24) {
                    out = dgj[lane - 21];
                }
                __syncwarp();
                if (lane < 9) S.dGam[9 * p + lane] += out;
                else if (lane < 18) S.dR[9 * j + lane - 9] = out;
                else if (lane < 21) { S.dJ[3 * j + lane - 18] += out; S.dJ[3 * p + lane - 18] -= out; }
                else if (lane < 24) S.dg[3 * p + lane - 21] += out;
                __syncwarp();
            }
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
        // ---- P10 pose-feature adjoint, Rodrigues adjoint, shape gradient
        if (t < kPoseBasis) S.dR[9 + t] += S.dPhi[t];
        __syncthreads();
        if (t < kJoints) {
            float dr[3] = {0.f, 0.f, 0.f};
            rodrigues_bwd(&S.x[kOffOrient + 3 * t], &S.dR[9 * t], dr);
            S.grad[kOffOrient + 3 * t] = dr[0]; S.grad[kOffOrient + 3 * t + 1] = dr[1]; S.grad[kOffOrient + 3 * t + 2] = dr[2];
        } else if (t >= 32 && t < 32 + kBetas) {
            const int l = t - 32;
            float a = S.dPhi[kPoseBasis + l];
            for (int jc = 0; jc < kJoints * 3; ++jc) a = fmaf(m.JS[jc * kBetas + l], S.dJ[jc], a);
            S.grad[kOffBetas + l] = a;
        }
    }
    // ---- P11 priors (fitting.py:327-350), all GMM components in parallel
    const float bpw = lp.body_pose_weight, bpw2 = bpw * bpw;
    const float* theta = &S.x[kOffPose];
    const int M = m.M;
    if (!lp.use_vposer && lp.body_prior == MVS_PRIOR_GMM)
        for (int e = t; e < M * 69; e += kResThreads) S.gm_diff[e] = theta[e % 69] - m.gmm_means[e];
    S.red[t] = (t < 69) ? theta[t] * theta[t] : 0.f;
    __syncthreads();
    if (!lp.use_vposer && lp.body_prior == MVS_PRIOR_GMM) {
        for (int e = t; e < M * 69; e += kResThreads) {
            const int mm = e / 69, i = e % 69;
            const float* P = m.gmm_prec + (size_t)mm * 69 * 69 + i;
            const float* df = &S.gm_diff[mm * 69];
            float y = 0.f;
#pragma unroll 3
            for (int j = 0; j < 69; ++j) y = fmaf(__ldg(P + j * 69), df[j], y);
            S.gm_y[e] = y;
        }
    }
    if (t == 0) { float a = 0.f; for (int i = 0; i < 69; ++i) a += S.red[i]; S.sc[1] = a; }
    __syncthreads();
    if (!lp.use_vposer && lp.body_prior == MVS_PRIOR_GMM && warp < M) {
        float p = 0.f;
        for (int i = lane; i < 69; i += 32) p = fmaf(S.gm_y[warp * 69 + i], S.gm_diff[warp * 69 + i], p);
        p = warp_sum(p);
        if (lane == 0) S.gm_ll[warp] = 0.5f * p - m.gmm_lognllw[warp];
    }
    __syncthreads();
    float pprior = 0.f, l2extra = 0.f;
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
#pragma unroll
            for (int i = 0; i < 4; ++i) S.grad[kOffPose + idx[i]] += 2.f * sg[i] * ev[i] * gs;
        }
    }
    __syncthreads();
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
        if (din.vposed) total += din.pen_loss;
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
}

__device__ __forceinline__ void resident_setup(ResidentSmem& S, const ResidentModel& m) {
    for (int col = threadIdx.x; col < 3 * m.nsup; col += kResThreads) S.rowbase[col] = 3 * m.sup[col / 3] + col % 3;
}

// ------------------------------------------------------------------------------------------------ single closure
__global__ void __launch_bounds__(kResThreads, 2)
closure_resident_kernel(ResidentModel m, CamSet cams, LossParams lp, const float* __restrict__ x,
                        const int* __restrict__ fidx, const int* __restrict__ na_ptr, const float* __restrict__ gt_uv,
                        const float* __restrict__ conf, const float* __restrict__ joint_w, int B, int have_grad,
                        float* __restrict__ loss_out, float* __restrict__ grad_out, float* __restrict__ joints_out,
                        float* __restrict__ proj_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ResidentSmem& S = *reinterpret_cast<ResidentSmem*>(smem_raw);
    const int slot = blockIdx.x;
    if (slot >= *na_ptr) return;
    const int b = fidx[slot];
    resident_setup(S, m);
    for (int i = threadIdx.x; i < kParams; i += kResThreads) S.x[i] = x[(size_t)b * kParams + i];
    __syncthreads();
    resident_closure(S, m, cams, lp, gt_uv, conf, joint_w, B, b, have_grad != 0, joints_out, proj_out, DenseIn{});
    if (threadIdx.x == 0) loss_out[b] = S.sc[2];
    if (have_grad)
        for (int i = threadIdx.x; i < kParams; i += kResThreads) grad_out[(size_t)b * kParams + i] = S.lg_new[i];
}

// ------------------------------------------------------------------------------------------------ whole stage
__global__ void __launch_bounds__(kResThreads, 2)
lbfgs_resident_kernel(ResidentModel m, CamSet cams, LossParams lp, LbfgsCfg cfg, float* __restrict__ params,
                      const float* __restrict__ gt_uv, const float* __restrict__ conf,
                      const float* __restrict__ joint_w, int B, int H, FrameScalars* __restrict__ sc_out,
                      float* __restrict__ last_grad_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ResidentSmem& S = *reinterpret_cast<ResidentSmem*>(smem_raw);
    float* hy = reinterpret_cast<float*>(smem_raw + sizeof(ResidentSmem));
    float* hs = hy + (size_t)H * kParams;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (b >= B) return;
    resident_setup(S, m);
    for (int i = t; i < kParams; i += kResThreads) { const float v = params[(size_t)b * kParams + i]; S.lx[i] = v; S.lx_eval[i] = v; }
    if (t == 0) {
        FrameScalars s;
        memset(&s, 0, sizeof(s));
        s.phase = PH_STEP_ENTRY;
        s.H_diag = 1.f;
        s.final_loss = __int_as_float(0x7fc00000);
        S.fs = s;
    }
    __syncthreads();
    LbfgsPtrs P{S.lx, S.lg, S.ld, S.lprev_g, S.lx_init, S.lg_prev, S.lbg0, S.lbg1, hy, hs, S.ro, S.al, S.lx_eval, S.lg_new, H};
    // hard bound on closure evaluations per frame (never reached by a terminating line search; a guard against
    // hanging the GPU): every outer step costs at most 1 + max_eval + max_iter evaluations
    const long long eval_cap = (long long)cfg.max_outer * (cfg.max_eval + cfg.max_iter + 2) + 8;
    while (true) {
        for (int i = t; i < kParams; i += kResThreads) S.x[i] = S.lx_eval[i];
        __syncthreads();
        resident_closure(S, m, cams, lp, gt_uv, conf, joint_w, B, b, true, nullptr, nullptr, DenseIn{});
        if (warp == 0) {
            FrameScalars s = S.fs;
            lbfgs_advance_core(s, P, S.sc[2], cfg, lane);
            if (s.evals >= eval_cap && s.phase != PH_DONE) { s.phase = PH_DONE; s.nan_flag = 1; }
            if (lane == 0) S.fs = s;
        }
        __syncthreads();
        if (S.fs.phase == PH_DONE) break;
    }
    for (int i = t; i < kParams; i += kResThreads) {
        params[(size_t)b * kParams + i] = S.lx[i];
        if (last_grad_out) last_grad_out[(size_t)b * kParams + i] = S.lg_new[i];
    }
    if (t == 0) sc_out[b] = S.fs;
}

// ------------------------------------------------------------------------------------------------ dense regime
// One round of the dense regime for one frame: consumes the dense vertices + the SDF gradient list of this
// frame's trial point, finishes the closure (keypoint term, adjoint, priors), advances the frame's L-BFGS state
// machine and -- if the frame needs another evaluation -- runs the pose forward of the NEXT trial point and
// writes its feature row / skinning transforms for the next dense vertex launch.
__global__ void __launch_bounds__(kResThreads, 2)
frame_step_kernel(ResidentModel m, CamSet cams, LossParams lp, LbfgsCfg cfg, LbfgsState L, float* __restrict__ params,
                  const int* __restrict__ fidx, const int* __restrict__ na_ptr, const float* __restrict__ gt_uv,
                  const float* __restrict__ conf, const float* __restrict__ joint_w, int B, int N,
                  const float* __restrict__ vposed_ws, const float* __restrict__ verts_ws,
                  const int* __restrict__ list_n, const float* __restrict__ list_d, const int* __restrict__ list_count,
                  const float* __restrict__ pen_loss, const float* __restrict__ Wd, float* __restrict__ Phi,
                  float* __restrict__ PhiTc, float* __restrict__ At, int ldA, const float* __restrict__ part, int nstrips,
                  const float* __restrict__ sdf_scal, const int* __restrict__ strip_active, int ftiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ResidentSmem& S = *reinterpret_cast<ResidentSmem*>(smem_raw);
    const int slot = blockIdx.x;
    if (slot >= *na_ptr) return;
    const int b = fidx[slot], t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (L.sc[b].phase == PH_DONE) return;
    resident_setup(S, m);
    float* x_eval = L.x_eval + (size_t)b * kParams;
    for (int i = t; i < kParams; i += kResThreads) S.x[i] = x_eval[i];
    // Stage this frame's curvature history (<= 2 x 100 x 86 floats) into shared memory with cp.async while the
    // closure runs: the two-loop recursion is ~200 DEPENDENT dot products, at L2 latency it costs ~80 us, from
    // shared memory ~5 us.
    float* hy = reinterpret_cast<float*>(smem_raw + sizeof(ResidentSmem));
    float* hs = hy + (size_t)L.H * kParams;
    {
        const int hl = L.sc[b].hist_len;
        const float* gy = L.hist_y + (size_t)b * L.H * kParams;
        const float* gs = L.hist_s + (size_t)b * L.H * kParams;
        const int nvec = (hl == L.H ? L.H : hl) * kParams / 2;       // 8-byte packets (rows are 344 B: 8-byte aligned)
        const int total = (hl == L.H) ? nvec : hl * kParams / 2;
        // ring buffer: with hl < H the live rows are [0, hl); when full all rows are live
        for (int i = t; i < total; i += kResThreads) {
            const unsigned sy = (unsigned)__cvta_generic_to_shared(hy + 2 * i);
            const unsigned ss = (unsigned)__cvta_generic_to_shared(hs + 2 * i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sy), "l"(gy + 2 * i));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ss), "l"(gs + 2 * i));
        }
        for (int i = t; i < hl; i += kResThreads) S.ro[i] = L.ro[(size_t)b * L.H + i];
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
    __syncthreads();
    DenseIn din;
    din.vposed = vposed_ws + (size_t)slot * N * 3;
    din.verts = verts_ws + (size_t)slot * N * 3;
    din.extra_n = list_n + (size_t)slot * N;
    din.extra_d = list_d + (size_t)slot * N * 3;
    din.n_extra = list_count[slot];
    din.pen_loss = pen_loss[slot];
    din.Wd = Wd;
    din.part = (part && sdf_scal[4 * slot] != 0.f) ? part : nullptr;      // frames without penetration have no partials
    din.nstrips = nstrips; din.ldA = ldA; din.slot = slot;
    din.strip_active = strip_active; din.ftiles = ftiles;
    resident_closure(S, m, cams, lp, gt_uv, conf, joint_w, B, b, true, nullptr, nullptr, din);
    for (int i = t; i < kParams; i += kResThreads) L.g_eval[(size_t)b * kParams + i] = S.lg_new[i];
    asm volatile("cp.async.wait_all;");
    __syncthreads();
    if (warp == 0) {
        FrameScalars s = L.sc[b];
        LbfgsPtrs P{S.lx, S.lg, S.ld, S.lprev_g, S.lx_init, S.lg_prev, S.lbg0, S.lbg1, hy, hs, S.ro, S.al, S.lx_eval,
                    S.lg_new, L.H};
        lbfgs_advance_core(s, P, S.sc[2], cfg, lane);
        __syncwarp();
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
        }
        if (lane == 0) { L.sc[b] = s; S.fs = s; }
    }
    __syncthreads();
    if (S.fs.phase == PH_DONE) return;
    // pose forward of the next trial point -> Phi row and skinning transforms for the next vertex launch
    for (int i = t; i < kParams; i += kResThreads) S.x[i] = S.lx_eval[i];
    __syncthreads();
    if (t < kJoints) rodrigues_fwd(&S.x[kOffOrient + 3 * t], &S.R[9 * t]);
    else if (t >= 32 && t < 32 + 72) {
        const int jc = t - 32;
        float a = m.Jt[jc];
#pragma unroll
        for (int l = 0; l < kBetas; ++l) a = fmaf(m.JS[jc * kBetas + l], S.x[kOffBetas + l], a);
        S.J[jc] = a;
    }
    __syncthreads();
    if (warp == 0) {
        const float sc = S.x[kOffScale];
        if (lane < 9) S.Gam[lane] = sc * S.R[lane];
        if (lane < 3) S.g[lane] = S.J[lane];
        __syncwarp();
        for (int j = 1; j < kJoints; ++j) {
            const int p = m.par.p[j];
            if (lane < 12) {
                const float* Gp = &S.Gam[9 * p];
                if (lane < 9) {
                    const int r = lane / 3, c = lane % 3;
                    const float* Rj = &S.R[9 * j];
                    S.Gam[9 * j + lane] = Gp[3 * r] * Rj[c] + Gp[3 * r + 1] * Rj[3 + c] + Gp[3 * r + 2] * Rj[6 + c];
                } else {
                    const int r = lane - 9;
                    const float r0 = S.J[3 * j] - S.J[3 * p], r1 = S.J[3 * j + 1] - S.J[3 * p + 1], r2 = S.J[3 * j + 2] - S.J[3 * p + 2];
                    S.g[3 * j + r] = (Gp[3 * r] * r0 + Gp[3 * r + 1] * r1 + Gp[3 * r + 2] * r2) + S.g[3 * p + r];
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (t < kJoints) {
        float A[12];
        make_skin_transform(&S.Gam[9 * t], &S.g[3 * t], &S.J[3 * t], A);
#pragma unroll
        for (int c = 0; c < 12; ++c) At[(size_t)(t * 12 + c) * ldA + slot] = A[c];
    } else if (t >= 32) {
        const int k = t - 32;
        float v;
        if (k < kPoseBasis) v = S.R[9 + k] - (((k % 9) % 4 == 0) ? 1.0f : 0.0f);
        else if (k < kPoseBasis + kBetas) v = S.x[kOffBetas + k - kPoseBasis];
        else v = (k == kFeat - 1) ? 1.0f : 0.0f;
        Phi[(size_t)slot * kFeatPad + k] = v;
        if (PhiTc) {
            float r = 0.f;
            if (k < kPoseBasis) { unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v)); r = __uint_as_float(u); }
            PhiTc[(size_t)slot * kFeatPad + k] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
static bool resident_supported(const mvs_ctx* ctx) {
    const DevModel& m = ctx->m;
    return ctx->exec_mode != 1 && m.supj_ptr != nullptr && m.nsup <= kResMaxSup && m.K <= kMaxKeypoints &&
           ctx->cams.num_views * m.K <= kResMaxVK && m.M <= kResMaxM;
}

static ResidentModel make_resident_model(const mvs_ctx* ctx) {
    const DevModel& m = ctx->m;
    ResidentModel r;
    r.Qk = m.Qk; r.Jt = m.Jt; r.JS = m.JS; r.ell_j = m.ell_j; r.ell_w = m.ell_w; r.KW = m.KW;
    r.K = m.K; r.nsup = m.nsup; r.kp_ptr = m.kp_ptr; r.kp_spos = m.kp_spos; r.kp_w = m.kp_w; r.kp_chain = m.kp_chain;
    r.sup = m.sup; r.sup_ptr = m.sup_ptr; r.sup_k = m.sup_k; r.sup_w = m.sup_w;
    r.supj_ptr = m.supj_ptr; r.supj_i = m.supj_i; r.supj_w = m.supj_w;
    r.M = m.M; r.gmm_means = m.gmm_means; r.gmm_prec = m.gmm_prec; r.gmm_lognllw = m.gmm_lognllw;
    r.anchor = ctx->ws.anchor; r.anchor_w = ctx->ws.anchor_w;
    r.par = ctx->parents;
    return r;
}

bool resident_closure_available(const mvs_ctx* ctx) { return resident_supported(ctx); }

int launch_closure_resident(mvs_ctx* ctx, const float* x_dev, float* loss_dev, float* grad_dev, float* joints_dev,
                            float* proj_dev, cudaStream_t st) {
    Workspace& w = ctx->ws;
    const size_t smem = sizeof(ResidentSmem);
    if (!ctx->attr_done_res) {
        MVS_CUDA_OK(ctx, cudaFuncSetAttribute(closure_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_done_res = true;
    }
    MVS_LAUNCH(ctx, KID_RESIDENT_CLOSURE, st,
               closure_resident_kernel<<<w.B, kResThreads, smem, st>>>(make_resident_model(ctx), ctx->cams, ctx->loss, x_dev,
                                                                       w.fidx, w.na, w.gt_uv, w.conf, w.joint_w, w.B,
                                                                       grad_dev ? 1 : 0, loss_dev ? loss_dev : w.loss_scratch,
                                                                       grad_dev, joints_dev, proj_dev));
    MVS_CUDA_OK(ctx, cudaGetLastError());
    return MVS_OK;
}

bool resident_lbfgs_available(const mvs_ctx* ctx, int H) {
    const LossParams& lp = ctx->loss;
    const bool sdf_on = lp.interpenetration && lp.coll_loss_weight > 0.f;
    return resident_supported(ctx) && !sdf_on && H <= 100;
}

int launch_lbfgs_resident(mvs_ctx* ctx, float* params_dev, const void* cfg_ptr, int H, void* sc_out, float* last_grad_dev,
                          cudaStream_t st) {
    Workspace& w = ctx->ws;
    const LbfgsCfg& cfg = *static_cast<const LbfgsCfg*>(cfg_ptr);
    const size_t smem = sizeof(ResidentSmem) + (size_t)2 * H * kParams * sizeof(float);
    if (ctx->attr_res_lbfgs_smem < (int)smem) {
        MVS_CUDA_OK(ctx, cudaFuncSetAttribute(lbfgs_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_res_lbfgs_smem = (int)smem;
    }
    MVS_LAUNCH(ctx, KID_RESIDENT_LBFGS, st,
               lbfgs_resident_kernel<<<w.B, kResThreads, smem, st>>>(make_resident_model(ctx), ctx->cams, ctx->loss, cfg,
                                                                     params_dev, w.gt_uv, w.conf, w.joint_w, w.B, H,
                                                                     static_cast<FrameScalars*>(sc_out), last_grad_dev));
    MVS_CUDA_OK(ctx, cudaGetLastError());
    return MVS_OK;
}

bool hybrid_available(const mvs_ctx* ctx) {
    const LossParams& lp = ctx->loss;
    const bool sdf_on = lp.interpenetration && lp.coll_loss_weight > 0.f;
    return resident_supported(ctx) && sdf_on;
}

int launch_frame_step(mvs_ctx* ctx, float* params_dev, const void* lbfgs_state, const void* lbfgs_cfg, int nstrips,
                      cudaStream_t st) {
    Workspace& w = ctx->ws;
    const DevModel& dm = ctx->m;
    const LbfgsState& L = *static_cast<const LbfgsState*>(lbfgs_state);
    const LbfgsCfg& cfg = *static_cast<const LbfgsCfg*>(lbfgs_cfg);
    const size_t smem = sizeof(ResidentSmem) + (size_t)2 * L.H * kParams * sizeof(float);
    if (!ctx->attr_done_step) {
        MVS_CUDA_OK(ctx, cudaFuncSetAttribute(frame_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_done_step = true;
    }
    MVS_LAUNCH(ctx, KID_FRAME_STEP, st,
               frame_step_kernel<<<w.B, kResThreads, smem, st>>>(make_resident_model(ctx), ctx->cams, ctx->loss, cfg, L, params_dev,
                                                                 w.fidx, w.na, w.gt_uv, w.conf, w.joint_w, w.B, dm.N, w.vposed,
                                                                 w.verts, w.sdf_list_n, w.sdf_list_d, w.sdf_list_count, w.pen_loss,
                                                                 dm.Wd, w.Phi, w.PhiTc, w.At, w.ldA, nstrips > 0 ? w.part : nullptr, nstrips,
                                                                 w.sdf_scal, w.strip_active, (w.B + kTileF - 1) / kTileF));
    MVS_CUDA_OK(ctx, cudaGetLastError());
    return MVS_OK;
}

}  // namespace mvs
