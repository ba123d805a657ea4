// Batched, device-resident L-BFGS with strong-Wolfe cubic line search: one independent
// optimisation problem per frame, advanced as a per-frame state machine between batched
// closure evaluations.  Replaces (reference paths under code/):
//   optimizers/lbfgs_ls.py:11-36   _cubic_interpolate
//   optimizers/lbfgs_ls.py:39-167  _strong_Wolfe
//   optimizers/lbfgs_ls.py:256-445 LBFGS.step   (history 100, persistent state across step() calls)
//   utils/fitting.py:71-142        FittingMonitor.run_fitting (NaN/Inf guard, ftol, gtol)
//
// The reference runs this loop in Python with a blocking float(closure()) per evaluation
// (lbfgs_ls.py:251,281).  Here no scalar leaves the GPU: every "round" evaluates the closure
// at each active frame's trial point, then lbfgs_advance_kernel (one warp per frame) consumes
// (f, g), walks the frame's state machine up to its next closure request and writes the next
// trial point; finished frames are compacted out of the active list.  The host only polls the
// active count every few rounds.
//
// Arithmetic mirrors the reference's types: dot products, step lengths and directional
// derivatives are fp32 (0-dim fp32 tensors there), losses are Python floats there and are
// compared / subtracted in double here.
#include <math.h>
#include <stdio.h>

#include "mvs_internal.cuh"

namespace mvs {

enum Phase { PH_STEP_ENTRY = 0, PH_LS_BRACKET = 1, PH_LS_ZOOM = 2, PH_DONE = 3 };

struct FrameScalars {
    int phase, outer_n, state_n_iter, n_iter, current_evals;
    int ls_iter, ls_evals, ls_first, low, high, done, insuf, have_prev;
    int hist_len, hist_head, nan_flag;
    float t, H_diag, gtd0, t_prev, gtd_prev, d_norm;
    float bt[2], bgtd[2], bf[2];
    float loss, prev_loss, orig_loss, f_prev, mon_prev_loss, final_loss;
    long long iters, evals;
};

struct LbfgsState {
    int B = 0, H = 0;
    float *g = nullptr, *d = nullptr, *prev_g = nullptr, *x_init = nullptr, *g_prev = nullptr, *bg = nullptr;
    float *hist_y = nullptr, *hist_s = nullptr, *ro = nullptr;
    float *x_eval = nullptr, *loss_eval = nullptr, *g_eval = nullptr;
    FrameScalars* sc = nullptr;
    long long* totals = nullptr;     // [4] iters, evals, nan frames, real rounds
    float *x_fit = nullptr, *final_fit = nullptr;   // mvs_fit_host staging
    int* na_host = nullptr;          // pinned
};

struct LbfgsCfg {
    int max_outer, max_iter, max_eval, history, max_ls;
    int step_mode;                   // 1: exactly one LBFGS.step() per call (the outer run_fitting loop is the caller's)
    float lr, tol_grad, tol_change;
    double ftol, gtol;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// each lane owns elements lane, lane+32, lane+64 (< 86) of every 86-vector
#define VL(c, i) _Pragma("unroll") for (int c = 0; c < 3; ++c) for (int i = lane + 32 * c; i < kParams; i = kParams)
#define VLOOP(i) VL(_c, i)

__device__ __forceinline__ float vdot(const float* a, const float* b, int lane) {
    float p = 0.f;
    VLOOP(i) p = fmaf(a[i], b[i], p);
    return warp_sum(p);
}
__device__ __forceinline__ float vabsmax(const float* a, int lane) {
    float p = 0.f;
    VLOOP(i) p = fmaxf(p, fabsf(a[i]));
    return warp_max(p);
}

// lbfgs_ls.py:11-36.  Python's min(max(pos, lo), hi) keeps `pos` when a comparison with NaN fails.
__device__ float cubic_interpolate(float x1, float f1, float g1, float x2, float f2, float g2, bool has_bounds,
                                   float lo, float hi) {
    if (!has_bounds) { if (x1 <= x2) { lo = x1; hi = x2; } else { lo = x2; hi = x1; } }
    const float num = (float)(3.0 * ((double)f1 - (double)f2));
    const float d1 = (g1 + g2) - num / (x1 - x2);
    const float d2sq = d1 * d1 - g1 * g2;
    if (d2sq >= 0.f) {
        const float d2 = sqrtf(d2sq);
        float pos;
        if (x1 <= x2) pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.f * d2));
        else pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.f * d2));
        float r = pos;
        if (lo > r) r = lo;
        if (hi < r) r = hi;
        return r;
    }
    return (lo + hi) / 2.f;
}

// reset != 0: fresh optimiser (new stage).  reset == 0: next step() of the same optimiser: history, d, t,
// H_diag, prev_flat_grad, prev_loss and n_iter persist (lbfgs_ls.py:292-300,436-443).
__global__ void lbfgs_init_kernel(LbfgsState S, const float* __restrict__ params, int B, int reset) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= B) return;
    VLOOP(i) S.x_eval[(size_t)b * kParams + i] = params[(size_t)b * kParams + i];
    if (lane == 0) {
        FrameScalars s;
        if (reset) {
            memset(&s, 0, sizeof(s));
            s.H_diag = 1.f;
        } else {
            s = S.sc[b];
            s.outer_n = 0; s.have_prev = 0; s.nan_flag = 0; s.iters = 0; s.evals = 0;
        }
        s.phase = PH_STEP_ENTRY;
        s.final_loss = __int_as_float(0x7fc00000);
        S.sc[b] = s;
    }
}

__global__ void __launch_bounds__(32)
lbfgs_advance_kernel(LbfgsState S, LbfgsCfg cfg, float* __restrict__ params, const int* __restrict__ fidx,
                     const int* __restrict__ na_ptr) {
    const int slot = blockIdx.x;
    if (slot >= *na_ptr) return;
    const int b = fidx[slot];
    const int lane = threadIdx.x;
    __shared__ float al[128];
    FrameScalars s = S.sc[b];
    float* x = params + (size_t)b * kParams;
    float* g = S.g + (size_t)b * kParams;
    float* d = S.d + (size_t)b * kParams;
    float* prev_g = S.prev_g + (size_t)b * kParams;
    float* x_init = S.x_init + (size_t)b * kParams;
    float* g_prev = S.g_prev + (size_t)b * kParams;
    float* bg0 = S.bg + (size_t)b * 2 * kParams;
    float* bg1 = bg0 + kParams;
    float* hy = S.hist_y + (size_t)b * S.H * kParams;
    float* hs = S.hist_s + (size_t)b * S.H * kParams;
    float* ro = S.ro + (size_t)b * S.H;
    float* x_eval = S.x_eval + (size_t)b * kParams;
    const float* g_new = S.g_eval + (size_t)b * kParams;
    const float f_new = S.loss_eval[b];
    const float c1 = 1e-4f, c2 = 0.9f;

    s.evals++;
    enum { L_STEP_ENTRY, L_LS_BRACKET, L_LS_ZOOM, L_ITER_BEGIN, L_ZOOM_INIT, L_ZOOM_HEAD, L_LS_FINISH, L_STEP_END,
           L_REQ_LS, L_FINISH, L_EXIT };
    int label = s.phase == PH_STEP_ENTRY ? L_STEP_ENTRY : (s.phase == PH_LS_BRACKET ? L_LS_BRACKET : L_LS_ZOOM);
    float gtd_new = 0.f;

    while (label != L_EXIT) {
        switch (label) {
        case L_STEP_ENTRY: {                                   // lbfgs_ls.py:279-290
            s.orig_loss = f_new;
            s.loss = f_new;
            s.current_evals = 1;
            VLOOP(i) g[i] = g_new[i];
            __syncwarp();
            if (vabsmax(g, lane) <= cfg.tol_grad) { label = L_STEP_END; break; }
            s.n_iter = 0;
            label = L_ITER_BEGIN;
            break;
        }
        case L_ITER_BEGIN: {                                   // lbfgs_ls.py:304-379
            s.n_iter++;
            s.state_n_iter++;
            s.iters++;
            if (s.state_n_iter == 1) {
                VLOOP(i) d[i] = -g[i];
                s.hist_len = 0; s.hist_head = 0;
                s.H_diag = 1.f;
            } else {
                // y = g - prev_g ; s = d * t
                float py = 0.f, pyy = 0.f;
                const int slot_new = (s.hist_head + s.hist_len) % S.H;
                float yv[3] = {0.f, 0.f, 0.f}, sv[3] = {0.f, 0.f, 0.f};
                VL(c, i) {
                    yv[c] = g[i] - prev_g[i];
                    sv[c] = d[i] * s.t;
                    py = fmaf(yv[c], sv[c], py);
                    pyy = fmaf(yv[c], yv[c], pyy);
                }
                const float ys = warp_sum(py), yy = warp_sum(pyy);
                if (ys > 1e-10f) {
                    int w = slot_new;
                    if (s.hist_len == S.H) { w = s.hist_head; s.hist_head = (s.hist_head + 1) % S.H; }   // drop the oldest
                    else s.hist_len++;
                    VL(c, i) { hy[(size_t)w * kParams + i] = yv[c]; hs[(size_t)w * kParams + i] = sv[c]; }
                    if (lane == 0) ro[w] = 1.f / ys;
                    s.H_diag = ys / yy;
                }
                __syncwarp();
                // two-loop recursion (collapsed to one buffer like the reference), q lives in registers
                float q[3] = {0.f, 0.f, 0.f};
                VL(c, i) q[c] = -g[i];
                for (int k = s.hist_len - 1; k >= 0; --k) {
                    const int w = (s.hist_head + k) % S.H;
                    float p = 0.f;
                    VL(c, i) p = fmaf(hs[(size_t)w * kParams + i], q[c], p);
                    const float a = warp_sum(p) * ro[w];
                    if (lane == 0) al[k] = a;
                    VL(c, i) q[c] = fmaf(-a, hy[(size_t)w * kParams + i], q[c]);
                }
                __syncwarp();
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) q[cc] *= s.H_diag;
                for (int k = 0; k < s.hist_len; ++k) {
                    const int w = (s.hist_head + k) % S.H;
                    float p = 0.f;
                    VL(c, i) p = fmaf(hy[(size_t)w * kParams + i], q[c], p);
                    const float be = warp_sum(p) * ro[w];
                    const float coef = al[k] - be;
                    VL(c, i) q[c] = fmaf(coef, hs[(size_t)w * kParams + i], q[c]);
                }
                VL(c, i) d[i] = q[c];
            }
            VLOOP(i) prev_g[i] = g[i];
            s.prev_loss = s.loss;
            __syncwarp();
            if (s.state_n_iter == 1) {
                float p = 0.f;
                VLOOP(i) p += fabsf(g[i]);
                const float inv = 1.f / warp_sum(p);
                s.t = (inv < 1.f ? inv : 1.f) * cfg.lr;        // min(1., 1./|g|_1) * lr
            } else {
                s.t = cfg.lr;
            }
            const float gtd = vdot(g, d, lane);
            if (gtd > -cfg.tol_change) { label = L_STEP_END; break; }
            // ---- _strong_Wolfe prologue (lbfgs_ls.py:42-52)
            VLOOP(i) { x_init[i] = x[i]; g_prev[i] = g[i]; }
            s.d_norm = vabsmax(d, lane);
            s.gtd0 = gtd;
            s.t_prev = 0.f; s.f_prev = s.loss; s.gtd_prev = gtd;
            s.ls_iter = 0; s.ls_evals = 0; s.ls_first = 1; s.done = 0; s.insuf = 0;
            s.phase = PH_LS_BRACKET;
            label = L_REQ_LS;
            break;
        }
        case L_LS_BRACKET: {                                   // lbfgs_ls.py:53-101
            s.ls_evals++;
            gtd_new = vdot(g_new, d, lane);
            if (!s.ls_first) {
                s.ls_iter++;
                if (s.ls_iter == cfg.max_ls) {                 // lbfgs_ls.py:97-101
                    s.bt[0] = 0.f; s.bt[1] = s.t; s.bf[0] = s.loss; s.bf[1] = f_new;
                    s.bgtd[0] = s.gtd0; s.bgtd[1] = gtd_new;
                    VLOOP(i) { bg0[i] = g[i]; bg1[i] = g_new[i]; }
                    label = L_ZOOM_INIT;
                    break;
                }
            }
            s.ls_first = 0;
            const float armijo = s.loss + (c1 * s.t) * s.gtd0;
            if (f_new > armijo || (s.ls_iter > 1 && f_new >= s.f_prev)) {
                s.bt[0] = s.t_prev; s.bt[1] = s.t; s.bf[0] = s.f_prev; s.bf[1] = f_new;
                s.bgtd[0] = s.gtd_prev; s.bgtd[1] = gtd_new;
                VLOOP(i) { bg0[i] = g_prev[i]; bg1[i] = g_new[i]; }
                label = L_ZOOM_INIT;
                break;
            }
            if (fabsf(gtd_new) <= -c2 * s.gtd0) {
                s.bt[0] = s.t; s.bt[1] = s.t; s.bf[0] = f_new; s.bf[1] = f_new; s.bgtd[0] = gtd_new; s.bgtd[1] = gtd_new;
                VLOOP(i) bg0[i] = g_new[i];
                s.done = 1; s.low = 0; s.high = 1;
                label = L_LS_FINISH;
                break;
            }
            if (gtd_new >= 0.f) {
                s.bt[0] = s.t_prev; s.bt[1] = s.t; s.bf[0] = s.f_prev; s.bf[1] = f_new;
                s.bgtd[0] = s.gtd_prev; s.bgtd[1] = gtd_new;
                VLOOP(i) { bg0[i] = g_prev[i]; bg1[i] = g_new[i]; }
                label = L_ZOOM_INIT;
                break;
            }
            {   // interpolate (lbfgs_ls.py:82-95)
                const float min_step = s.t + 0.01f * (s.t - s.t_prev);
                const float max_step = s.t * 10.f;
                const float tmp = s.t;
                s.t = cubic_interpolate(s.t_prev, s.f_prev, s.gtd_prev, s.t, f_new, gtd_new, true, min_step, max_step);
                s.t_prev = tmp; s.f_prev = f_new; s.gtd_prev = gtd_new;
                VLOOP(i) g_prev[i] = g_new[i];
                label = L_REQ_LS;
            }
            break;
        }
        case L_ZOOM_INIT: {                                    // lbfgs_ls.py:106-108
            s.insuf = 0;
            if (s.bf[0] <= s.bf[1]) { s.low = 0; s.high = 1; } else { s.low = 1; s.high = 0; }
            s.phase = PH_LS_ZOOM;
            label = L_ZOOM_HEAD;
            break;
        }
        case L_ZOOM_HEAD: {                                    // lbfgs_ls.py:109-129
            if (s.done || s.ls_iter >= cfg.max_iter) { label = L_LS_FINISH; break; }
            float t = cubic_interpolate(s.bt[0], s.bf[0], s.bgtd[0], s.bt[1], s.bf[1], s.bgtd[1], false, 0.f, 0.f);
            const float bmax = (s.bt[1] > s.bt[0]) ? s.bt[1] : s.bt[0];
            const float bmin = (s.bt[1] < s.bt[0]) ? s.bt[1] : s.bt[0];
            const float eps = 0.1f * (bmax - bmin);
            const float da = bmax - t, db = t - bmin;
            if (((db < da) ? db : da) < eps) {
                if (s.insuf || t >= bmax || t <= bmin) {
                    t = (fabsf(t - bmax) < fabsf(t - bmin)) ? bmax - eps : bmin + eps;
                    s.insuf = 0;
                } else {
                    s.insuf = 1;
                }
            } else {
                s.insuf = 0;
            }
            s.t = t;
            label = L_REQ_LS;
            break;
        }
        case L_LS_ZOOM: {                                      // lbfgs_ls.py:131-161
            s.ls_evals++;
            gtd_new = vdot(g_new, d, lane);
            s.ls_iter++;
            const float armijo = s.loss + (c1 * s.t) * s.gtd0;
            float* bgl = s.low ? bg1 : bg0;
            float* bgh = s.high ? bg1 : bg0;
            if (f_new > armijo || f_new >= s.bf[s.low]) {
                s.bt[s.high] = s.t; s.bf[s.high] = f_new; s.bgtd[s.high] = gtd_new;
                VLOOP(i) bgh[i] = g_new[i];
                if (s.bf[0] <= s.bf[1]) { s.low = 0; s.high = 1; } else { s.low = 1; s.high = 0; }
            } else {
                if (fabsf(gtd_new) <= -c2 * s.gtd0) {
                    s.done = 1;
                } else if (gtd_new * (s.bt[s.high] - s.bt[s.low]) >= 0.f) {
                    s.bt[s.high] = s.bt[s.low]; s.bf[s.high] = s.bf[s.low]; s.bgtd[s.high] = s.bgtd[s.low];
                    VLOOP(i) bgh[i] = bgl[i];
                }
                s.bt[s.low] = s.t; s.bf[s.low] = f_new; s.bgtd[s.low] = gtd_new;
                VLOOP(i) bgl[i] = g_new[i];
            }
            __syncwarp();
            if (fabsf(s.bt[1] - s.bt[0]) * s.d_norm < cfg.tol_change) { label = L_LS_FINISH; break; }
            label = L_ZOOM_HEAD;
            break;
        }
        case L_LS_FINISH: {                                    // lbfgs_ls.py:163-167, 399-434
            const float* bgl = s.low ? bg1 : bg0;
            s.t = s.bt[s.low];
            s.loss = s.bf[s.low];
            VLOOP(i) { g[i] = bgl[i]; x[i] = fmaf(s.t, d[i], x_init[i]); }
            __syncwarp();
            const bool opt_cond = vabsmax(g, lane) <= cfg.tol_grad;
            s.current_evals += s.ls_evals;
            float p = 0.f;
            VLOOP(i) p = fmaxf(p, fabsf(d[i] * s.t));
            const float step_max = warp_max(p);
            if (s.n_iter == cfg.max_iter || s.current_evals >= cfg.max_eval || opt_cond || step_max <= cfg.tol_change ||
                fabs((double)s.loss - (double)s.prev_loss) < (double)cfg.tol_change) {
                label = L_STEP_END;
                break;
            }
            label = L_ITER_BEGIN;
            break;
        }
        case L_STEP_END: {                                     // fitting.py:99-140
            const float ret = s.orig_loss;
            if (cfg.step_mode) {                               // step() returns the loss at entry (lbfgs_ls.py:445)
                s.final_loss = ret;
                s.phase = PH_DONE;
                label = L_EXIT;
                break;
            }
            if (isnan(ret) || isinf(ret)) { s.nan_flag = 1; label = L_FINISH; break; }
            if (s.outer_n > 0 && s.have_prev && cfg.ftol > 0.0) {
                const double pl = (double)s.mon_prev_loss, cl = (double)ret;
                const double den = fmax(fmax(fabs(pl), fabs(cl)), 1.0);
                if ((pl - cl) / den <= cfg.ftol) { label = L_FINISH; break; }
            }
            {   // all(abs(max(grad_tensor)) < gtol) on the grads left by the LAST closure call (fitting.py:115-117)
                const int seg_a[5] = {kOffBetas, kOffOrient, kOffPose, kOffTransl, kOffScale};
                const int seg_e[5] = {kOffOrient, kOffPose, kOffTransl, kOffScale, kParams};
                bool all_small = true;
                for (int sg = 0; sg < 5; ++sg) {
                    float m = -3.0e38f;
                    VLOOP(i) if (i >= seg_a[sg] && i < seg_e[sg]) m = fmaxf(m, g_new[i]);
                    m = warp_max(m);
                    if (!(fabs((double)m) < cfg.gtol)) all_small = false;
                }
                if (all_small) { label = L_FINISH; break; }
            }
            s.mon_prev_loss = ret;
            s.have_prev = 1;
            s.outer_n++;
            if (s.outer_n >= cfg.max_outer) { label = L_FINISH; break; }
            VLOOP(i) x_eval[i] = x[i];
            s.phase = PH_STEP_ENTRY;
            label = L_EXIT;
            break;
        }
        case L_REQ_LS: {                                       // lbfgs_ls.py:249-254: x + t*d, evaluated by the next round
            VLOOP(i) x_eval[i] = fmaf(s.t, d[i], x_init[i]);
            label = L_EXIT;
            break;
        }
        case L_FINISH: {
            s.final_loss = s.have_prev ? s.mon_prev_loss : __int_as_float(0x7fc00000);
            s.phase = PH_DONE;
            label = L_EXIT;
            break;
        }
        default: label = L_EXIT; break;
        }
    }
    __syncwarp();
    if (lane == 0) S.sc[b] = s;
}

// Ordered compaction of the frames that still need closure evaluations.
__global__ void __launch_bounds__(1024) lbfgs_compact_kernel(LbfgsState S, int B, int* __restrict__ fidx, int* __restrict__ na) {
    if (threadIdx.x == 0 && *na > 0) S.totals[3] += 1;     // rounds that actually evaluated something
    __shared__ int wsum[32];
    __shared__ int base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < B; b0 += 1024) {
        const int b = b0 + tid;
        const int act = (b < B && S.sc[b].phase != PH_DONE) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, act);
        const int pre = __popc(m & ((1u << lane) - 1));
        if (lane == 0) wsum[wid] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < wid; ++w) off += wsum[w];
        if (act) fidx[off + pre] = b;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += wsum[w]; base += tot; }
        __syncthreads();
    }
    if (tid == 0) *na = base;
}

__global__ void lbfgs_finalize_kernel(LbfgsState S, int B, float* __restrict__ final_loss, long long* __restrict__ totals) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const FrameScalars s = S.sc[b];
    if (final_loss) final_loss[b] = s.final_loss;
    atomicAdd((unsigned long long*)&totals[0], (unsigned long long)s.iters);
    atomicAdd((unsigned long long*)&totals[1], (unsigned long long)s.evals);
    atomicAdd((unsigned long long*)&totals[2], (unsigned long long)s.nan_flag);
}

__global__ void iota2_kernel(int* p, int n, int* na) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
    if (i == 0) *na = n;
}

static int ensure_state(mvs_ctx* ctx, int H) {
    LbfgsState* S = static_cast<LbfgsState*>(ctx->lbfgs);
    const int B = ctx->ws.B;
    if (S && S->B == B && S->H == H) return MVS_OK;
    if (S) return set_error(ctx, MVS_ERR_INVALID, "mvs_lbfgs_run: history_size changed between calls (%d -> %d)", S->H, H);
    S = new LbfgsState();
    S->B = B; S->H = H;
    int rc;
    const size_t V = (size_t)B * kParams;
    if ((rc = dev_alloc(ctx, &S->g, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->d, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->prev_g, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->x_init, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->g_prev, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->bg, 2 * V))) return rc;
    if ((rc = dev_alloc(ctx, &S->hist_y, V * H))) return rc;
    if ((rc = dev_alloc(ctx, &S->hist_s, V * H))) return rc;
    if ((rc = dev_alloc(ctx, &S->ro, (size_t)B * H))) return rc;
    if ((rc = dev_alloc(ctx, &S->x_eval, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->loss_eval, B))) return rc;
    if ((rc = dev_alloc(ctx, &S->g_eval, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->x_fit, V))) return rc;
    if ((rc = dev_alloc(ctx, &S->final_fit, B))) return rc;
    unsigned char* raw = nullptr;
    if ((rc = dev_alloc(ctx, &raw, (size_t)B * sizeof(FrameScalars)))) return rc;
    S->sc = reinterpret_cast<FrameScalars*>(raw);
    unsigned char* raw2 = nullptr;
    if ((rc = dev_alloc(ctx, &raw2, 4 * sizeof(long long)))) return rc;
    S->totals = reinterpret_cast<long long*>(raw2);
    MVS_CUDA_OK(ctx, cudaMallocHost(&S->na_host, 8 * sizeof(long long)));
    ctx->lbfgs = S;
    return MVS_OK;
}

static int run_stage(mvs_ctx* ctx, float* params_dev, float* final_loss_dev, const mvs_lbfgs_config* c,
                     mvs_lbfgs_stats* stats, cudaStream_t st, int step_mode = 0, int reset = 1,
                     float* last_grad_dev = nullptr) {
    const int B = ctx->ws.B;
    const int H = c->history_size > 0 ? c->history_size : 100;
    if (H > 128) return set_error(ctx, MVS_ERR_INVALID, "mvs_lbfgs_run: history_size must be <= 128");
    int rc = ensure_state(ctx, H);
    if (rc) return rc;
    LbfgsState& S = *static_cast<LbfgsState*>(ctx->lbfgs);
    LbfgsCfg cfg;
    cfg.max_outer = c->max_outer > 0 ? c->max_outer : 30;
    cfg.max_iter = c->max_iter > 0 ? c->max_iter : 30;
    cfg.max_eval = c->max_eval > 0 ? c->max_eval : cfg.max_iter * 5 / 4;
    cfg.history = H;
    cfg.max_ls = 25;                                            // _strong_Wolfe default (lbfgs_ls.py:41)
    cfg.step_mode = step_mode;
    cfg.lr = c->lr > 0.f ? c->lr : 1.f;
    cfg.tol_grad = c->tolerance_grad; cfg.tol_change = c->tolerance_change;
    cfg.ftol = (double)c->ftol; cfg.gtol = (double)c->gtol;

    Workspace& w = ctx->ws;
    MVS_LAUNCH(ctx, KID_MISC, st, iota2_kernel<<<(B + 255) / 256, 256, 0, st>>>(w.fidx, B, w.na));
    MVS_LAUNCH(ctx, KID_MISC, st, lbfgs_init_kernel<<<B, 32, 0, st>>>(S, params_dev, B, reset));
    MVS_CUDA_OK(ctx, cudaMemsetAsync(S.totals, 0, 4 * sizeof(long long), st));
    const int chunk = 8;
    const long long max_rounds = (long long)cfg.max_outer * (cfg.max_eval + cfg.max_iter + 2) + 8;
    long long rounds = 0;
    int na_host = B;
    while (na_host > 0 && rounds < max_rounds) {
        for (int r = 0; r < chunk; ++r) {
            rc = launch_closure(ctx, S.x_eval, S.loss_eval, S.g_eval, nullptr, nullptr, nullptr, st);
            if (rc) return rc;
            MVS_LAUNCH(ctx, KID_LBFGS_ADVANCE, st, lbfgs_advance_kernel<<<B, 32, 0, st>>>(S, cfg, params_dev, w.fidx, w.na));
            MVS_LAUNCH(ctx, KID_LBFGS_COMPACT, st, lbfgs_compact_kernel<<<1, 1024, 0, st>>>(S, B, w.fidx, w.na));
            ++rounds;
        }
        MVS_CUDA_OK(ctx, cudaMemcpyAsync(S.na_host, w.na, sizeof(int), cudaMemcpyDeviceToHost, st));
        MVS_CUDA_OK(ctx, cudaStreamSynchronize(st));
        na_host = *S.na_host;
    }
    MVS_LAUNCH(ctx, KID_MISC, st, lbfgs_finalize_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, B, final_loss_dev, S.totals));
    if (last_grad_dev)
        MVS_CUDA_OK(ctx, cudaMemcpyAsync(last_grad_dev, S.g_eval, (size_t)B * kParams * sizeof(float), cudaMemcpyDeviceToDevice, st));
    long long* tot_host = reinterpret_cast<long long*>(S.na_host) + 1;
    MVS_CUDA_OK(ctx, cudaMemcpyAsync(tot_host, S.totals, 4 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    // leave the context usable for plain closures: full active list again
    MVS_LAUNCH(ctx, KID_MISC, st, iota2_kernel<<<(B + 255) / 256, 256, 0, st>>>(w.fidx, B, w.na));
    MVS_CUDA_OK(ctx, cudaStreamSynchronize(st));
    if (stats) {
        stats->frame_iterations += tot_host[0];
        stats->frame_evals += tot_host[1];
        stats->frames_nan += (int)tot_host[2];
        stats->rounds += (int)tot_host[3];
    }
    MVS_CUDA_OK(ctx, cudaGetLastError());
    return MVS_OK;
}

}  // namespace mvs

using namespace mvs;

extern "C" {

int mvs_lbfgs_run(mvs_ctx* ctx, float* params_dev, float* final_loss_dev, const mvs_lbfgs_config* cfg,
                  mvs_lbfgs_stats* stats, void* stream) {
    if (!ctx) return set_error(nullptr, MVS_ERR_INVALID, "ctx is NULL");
    if (!(ctx->have_model && ctx->have_cams && ctx->have_kp && ctx->have_loss && ctx->ws.B > 0))
        return set_error(ctx, MVS_ERR_INVALID, "mvs_lbfgs_run: model, cameras, batch, keypoints and loss config must be set first");
    if (!params_dev || !cfg) return set_error(ctx, MVS_ERR_INVALID, "mvs_lbfgs_run: NULL argument");
    MVS_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    return run_stage(ctx, params_dev, final_loss_dev, cfg, stats, (cudaStream_t)stream);
}

int mvs_lbfgs_step(mvs_ctx* ctx, float* params_dev, float* loss_dev, float* last_grad_dev, const mvs_lbfgs_config* cfg,
                   int reset, mvs_lbfgs_stats* stats, void* stream) {
    if (!ctx) return set_error(nullptr, MVS_ERR_INVALID, "ctx is NULL");
    if (!(ctx->have_model && ctx->have_cams && ctx->have_kp && ctx->have_loss && ctx->ws.B > 0))
        return set_error(ctx, MVS_ERR_INVALID, "mvs_lbfgs_step: model, cameras, batch, keypoints and loss config must be set first");
    if (!params_dev || !cfg) return set_error(ctx, MVS_ERR_INVALID, "mvs_lbfgs_step: NULL argument");
    if (!reset && !ctx->lbfgs) return set_error(ctx, MVS_ERR_INVALID, "mvs_lbfgs_step: reset=0 before any step");
    MVS_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    return run_stage(ctx, params_dev, loss_dev, cfg, stats, (cudaStream_t)stream, 1, reset ? 1 : 0, last_grad_dev);
}

int mvs_fit_host(mvs_ctx* ctx, float* params_host, const float* gt_uv_host, const float* conf_host,
                 const float* joint_weights_host, int n_stages, const mvs_loss_config* stage_cfgs,
                 const mvs_lbfgs_config* opt_cfg, float* final_loss_host, mvs_lbfgs_stats* stats, void* stream) {
    if (!ctx) return set_error(nullptr, MVS_ERR_INVALID, "ctx is NULL");
    if (!(ctx->have_model && ctx->have_cams && ctx->ws.B > 0))
        return set_error(ctx, MVS_ERR_INVALID, "mvs_fit_host: model, cameras and batch must be set first");
    if (!params_host || !gt_uv_host || !conf_host || !joint_weights_host || n_stages <= 0 || !stage_cfgs || !opt_cfg)
        return set_error(ctx, MVS_ERR_INVALID, "mvs_fit_host: NULL / empty argument");
    MVS_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int B = ctx->ws.B;
    int rc = mvs_set_keypoints(ctx, gt_uv_host, conf_host, joint_weights_host, 0, stream);
    if (rc) return rc;
    {
        const int H = opt_cfg->history_size > 0 ? opt_cfg->history_size : 100;
        if (H > 128) return set_error(ctx, MVS_ERR_INVALID, "mvs_fit_host: history_size must be <= 128");
        rc = ensure_state(ctx, H);
        if (rc) return rc;
    }
    LbfgsState& S = *static_cast<LbfgsState*>(ctx->lbfgs);
    float* x = S.x_fit;                              // [B][86] device copy of the parameters
    MVS_CUDA_OK(ctx, cudaMemcpyAsync(x, params_host, (size_t)B * kParams * sizeof(float), cudaMemcpyHostToDevice, st));
    if (stats) memset(stats, 0, sizeof(*stats));
    for (int sidx = 0; sidx < n_stages; ++sidx) {
        rc = mvs_set_loss_config(ctx, &stage_cfgs[sidx]);
        if (rc) return rc;
        rc = run_stage(ctx, x, S.final_fit, opt_cfg, stats, st);
        if (rc) return rc;
    }
    MVS_CUDA_OK(ctx, cudaMemcpyAsync(params_host, x, (size_t)B * kParams * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (final_loss_host)
        MVS_CUDA_OK(ctx, cudaMemcpyAsync(final_loss_host, S.final_fit, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, st));
    MVS_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return MVS_OK;
}

}  // extern "C"
