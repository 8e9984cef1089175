// ict_kernel_pipe.cu — K2p, the software-pipelined production form of SetPose + TrackPose.
//
// In k_track_fast (one track per CTA) every Gauss-Newton iteration ends in a serial section — cross-warp sums,
// 6x6 solve, additive se(3) update, exp, stop rule — during which seven of eight warps wait at a barrier; ncu put
// 48 % of all warp stall samples on those barriers, and skipping the serial work (profiles/tools/serial_cost.sh)
// showed it costs 26 % of the kernel.  The serial section of one track cannot overlap its own pixel work (the next
// gather needs the new pose), but it can overlap ANOTHER track's:
//
//   * a CTA owns TWO track slots (2 x 49 KB of shared-memory template) and nine warps: eight pixel warps and one
//     solver warp;
//   * phases alternate between the slots: in phase X the pixel warps do slot X's pixel step (template gather +
//     Hessian partials, or one iteration's gather + residual + J^T r partials) while the solver warp does slot
//     1-X's pending serial step (Hessian sums + LU, or solve + update + stop rule, or finish + fetch next track);
//     ONE barrier per phase instead of two;
//   * CTAs are persistent (two per SM) and a slot that finishes a track takes the next one from a global ticket, so
//     tracks that converge early do not leave their partner alone.
//
// Per-pixel and per-track arithmetic is that of k_track_fast (factorised SD sums, tree reductions in a fixed order,
// reciprocal back substitution, series exp); results differ from it only through the order of the per-warp partial
// sums.  Everything a pixel step needs beyond the template lives in the slot's small shared state, written by the
// solver warp one phase earlier; points are read from global memory (L1-resident broadcasts).
#include "ict_kernels.cuh"
#include "ict_device.cuh"

namespace ict {

void count_launch_external();

enum { PAR_NONE = 0, PAR_PRE = 1, PAR_ITER = 2 };

struct PipeSlot {
  float G0[12];            // pose set by SetPose: reference reprojection, frozen for the track
  float G[12];             // current pose
  float p[6];
  float sum[8];
  float dp[6];
  float Hsum[24];
  float part[8 * 24];      // per pixel-warp partial sums of the step just done
  Lu6 lu;
  float normdp, normdp_init;
  int track;               // -1: slot idle
  int level;
  int it;
  int par;                 // the pixel step to run next
  int pending;             // 1: that pixel step has run and its serial step has not yet
  int first;               // 1 while on the track's first level (ResetOdometer semantics for invisible points)
  int trace_n;
  long long npix;
};

#define NPW 8              // pixel warps; warp NPW is the solver warp

template <int PSZ>
struct PipeCfg {
  static constexpr int N = PSZ * PSZ;
  static constexpr int KT = (N / 32 < 16) ? N / 32 : 16;
  static constexpr int GE = 32 * KT;
  static constexpr int GPP = N / GE;
};

// ---- pixel step PRE: template gather + SD coefficients + Hessian partials (odometer.cpp:268-334) -----------------
template <int PSZ>
__device__ __forceinline__ void pipe_pre(const TrackParams& prm, PipeSlot& S, float* s_ref, float* s_gx, float* s_gy,
                                         float* s_coef, int warp, int lane) {
  using C = PipeCfg<PSZ>;
  const ict_optparam& op = prm.op;
  const int t = S.track, sl = S.level;
  const int64_t off = prm.pt_off[t];
  const int n_in = (int)(prm.pt_off[t + 1] - off);
  const int P = min(n_in, op.maxpttrack);
  const float* __restrict__ q = prm.pt3d + 3 * off;
  const int rf = prm.ref_frame ? prm.ref_frame[t] : prm.fixed_ref;
  const FrameDesc* fr = prm.frames + rf;
  const float* __restrict__ Iref = fr->I[sl];
  const float* __restrict__ Dxr = fr->dx[sl];
  const float* __restrict__ Dyr = fr->dy[sl];
  const float fx = prm.cam.fx[sl], fy = prm.cam.fy[sl], cx = prm.cam.cx[sl], cy = prm.cam.cy[sl];
  const float swo = prm.cam.swo[sl], sho = prm.cam.sho[sl];
  const int width = prm.cam.width[sl];
  const bool first = S.first != 0;
  float G0[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) G0[k] = S.G0[k];

  const int G_all = P * C::GPP;
  const int gpw = (G_all + NPW - 1) / NPW;
  const int g_lo = min(warp * gpw, G_all), g_hi = min(g_lo + gpw, G_all);

  float acc[21];
#pragma unroll
  for (int k = 0; k < 21; ++k) acc[k] = 0.0f;
  int cur = -1;
  bool vis = false;
  PatchPlace pl = {0, 0.f, 0.f, 0.f, 0.f};
  float cf[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) cf[k] = 0.0f;
  float sxx = 0.0f, sxy = 0.0f, syy = 0.0f;
  for (int g = g_lo; g < g_hi; ++g) {
    const int i = g / C::GPP, gp = g - i * C::GPP;
    if (i != cur) {
      fold_hessian(acc, cf, sxx, sxy, syy);
      sxx = sxy = syy = 0.0f;
      cur = i;
      const float X = __ldg(q + i), Y = __ldg(q + n_in + i), Z = __ldg(q + 2 * (int64_t)n_in + i);
      const float xc = G0[0] * X + G0[1] * Y + G0[2] * Z + G0[3];     // project_pt_save_rotated, pose.cpp:400-488
      const float yc = G0[4] * X + G0[5] * Y + G0[6] * Z + G0[7];
      const float zc = G0[8] * X + G0[9] * Y + G0[10] * Z + G0[11];
      const float mx = (xc / zc) * fx + cx, my = (yc / zc) * fy + cy;
      vis = (mx >= 0) & (my >= 0) & (mx <= swo) & (my <= sho);          // odometer.cpp:273-275
      if (first && gp == 0 && lane == 0 && prm.pt2d_out) {              // Get2DPoints(): pt2d[lv_l]
        const int l = op.lv_l;
        prm.pt2d_out[2 * off + i] = (xc / zc) * prm.cam.fx[l] + prm.cam.cx[l];
        prm.pt2d_out[2 * off + n_in + i] = (yc / zc) * prm.cam.fy[l] + prm.cam.cy[l];
      }
      if (vis) {
        pl = patch_place(mx, my, PSZ / 2, width);
        sd_coefs(xc, yc, zc, fx, fy, cf);
      } else if (first) {
#pragma unroll
        for (int k = 0; k < 10; ++k) cf[k] = 0.0f;                      // ResetOdometer state
      } else {
#pragma unroll
        for (int k = 0; k < 10; ++k) cf[k] = s_coef[i * 10 + k];        // stale coefficients
      }
      if (gp == 0 && (vis || first)) {                                  // persists for later steps
#pragma unroll
        for (int k = 0; k < 10; ++k)
          if (lane == k) s_coef[i * 10 + k] = cf[k];
      }
    }
    const int ebase = i * C::N + gp * C::GE + lane;
    if (vis) {
      if (PSZ == 32) {
        const int a0 = pl.base + (gp * C::KT) * width + lane;
        float ci = __ldg(Iref + a0 - width), di = __ldg(Iref + a0 - width - 1);
        float cx_ = __ldg(Dxr + a0 - width), dx_ = __ldg(Dxr + a0 - width - 1);
        float cy_ = __ldg(Dyr + a0 - width), dy_ = __ldg(Dyr + a0 - width - 1);
#pragma unroll
        for (int j = 0; j < C::KT; ++j) {
          const int a = a0 + j * width;
          const float ai = __ldg(Iref + a), bi = __ldg(Iref + a - 1);
          const float ax = __ldg(Dxr + a), bx = __ldg(Dxr + a - 1);
          const float ay = __ldg(Dyr + a), by = __ldg(Dyr + a - 1);
          const float vr = ((pl.w0 * ai + pl.w1 * bi) + pl.w2 * ci) + pl.w3 * di;
          const float vx = ((pl.w0 * ax + pl.w1 * bx) + pl.w2 * cx_) + pl.w3 * dx_;
          const float vy = ((pl.w0 * ay + pl.w1 * by) + pl.w2 * cy_) + pl.w3 * dy_;
          s_ref[ebase + 32 * j] = vr;
          s_gx[ebase + 32 * j] = vx;
          s_gy[ebase + 32 * j] = vy;
          sxx = sxx + vx * vx;
          sxy = sxy + vx * vy;
          syy = syy + vy * vy;
          ci = ai; di = bi; cx_ = ax; dx_ = bx; cy_ = ay; dy_ = by;
        }
      } else {
#pragma unroll
        for (int j = 0; j < C::KT; ++j) {
          const int qq = gp * C::GE + 32 * j + lane, r = qq / PSZ, c = qq - r * PSZ;
          const int a = pl.base + r * width + c;
          const float vr = bilin4(Iref, a, width, pl.w0, pl.w1, pl.w2, pl.w3);
          const float vx = bilin4(Dxr, a, width, pl.w0, pl.w1, pl.w2, pl.w3);
          const float vy = bilin4(Dyr, a, width, pl.w0, pl.w1, pl.w2, pl.w3);
          s_ref[ebase + 32 * j] = vr;
          s_gx[ebase + 32 * j] = vx;
          s_gy[ebase + 32 * j] = vy;
          sxx = sxx + vx * vx;
          sxy = sxy + vx * vy;
          syy = syy + vy * vy;
        }
      }
    } else if (first) {
#pragma unroll
      for (int j = 0; j < C::KT; ++j) { s_ref[ebase + 32 * j] = 0.0f; s_gx[ebase + 32 * j] = 0.0f; s_gy[ebase + 32 * j] = 0.0f; }
    } else {
#pragma unroll
      for (int j = 0; j < C::KT; ++j) {   // stale template of an earlier level still counts in H
        const float vx = s_gx[ebase + 32 * j], vy = s_gy[ebase + 32 * j];
        sxx = sxx + vx * vx;
        sxy = sxy + vx * vy;
        syy = syy + vy * vy;
      }
    }
  }
  fold_hessian(acc, cf, sxx, sxy, syy);
  warp_sum_store<21>(acc, &S.part[warp * 24]);
}

// ---- pixel step ITER: project, new-frame patch, residual, J^T r partials (odometer.cpp:352-404) -----------------
template <int PSZ>
__device__ __forceinline__ void pipe_iter(const TrackParams& prm, PipeSlot& S, const float* s_ref, const float* s_gx,
                                          const float* s_gy, const float* s_coef, int warp, int lane) {
  using C = PipeCfg<PSZ>;
  const ict_optparam& op = prm.op;
  const int t = S.track, sl = S.level;
  const int64_t off = prm.pt_off[t];
  const int n_in = (int)(prm.pt_off[t + 1] - off);
  const int P = min(n_in, op.maxpttrack);
  const float* __restrict__ q = prm.pt3d + 3 * off;
  const int nf = prm.new_frame ? prm.new_frame[t] : prm.fixed_new;
  const float* __restrict__ Inew = prm.frames[nf].I[sl];
  const float fx = prm.cam.fx[sl], fy = prm.cam.fy[sl], cx = prm.cam.cx[sl], cy = prm.cam.cy[sl];
  const float swo = prm.cam.swo[sl], sho = prm.cam.sho[sl];
  const int width = prm.cam.width[sl];
  float Gm[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) Gm[k] = S.G[k];

  const int G_all = P * C::GPP;
  const int gpw = (G_all + NPW - 1) / NPW;
  const int g_lo = min(warp * gpw, G_all), g_hi = min(g_lo + gpw, G_all);

  float acc[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = 0.0f;
  int nvis = 0;
  int cur = -1;
  bool vis = false;
  PatchPlace pl = {0, 0.f, 0.f, 0.f, 0.f};
  float cf[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) cf[k] = 0.0f;
  float ax = 0.0f, ay = 0.0f;
  for (int g = g_lo; g < g_hi; ++g) {
    const int i = g / C::GPP, gp = g - i * C::GPP;
    if (i != cur) {
      fold_jtr(acc, cf, ax, ay);
      ax = ay = 0.0f;
      cur = i;
      const float X = __ldg(q + i), Y = __ldg(q + n_in + i), Z = __ldg(q + 2 * (int64_t)n_in + i);
      const float tx = Gm[0] * X + Gm[1] * Y + Gm[2] * Z + Gm[3];      // project_pt, pose.cpp:307-397
      const float ty = Gm[4] * X + Gm[5] * Y + Gm[6] * Z + Gm[7];
      const float tz = Gm[8] * X + Gm[9] * Y + Gm[10] * Z + Gm[11];
      const float mx = (tx / tz) * fx + cx, my = (ty / tz) * fy + cy;
      vis = (mx >= 0) & (my >= 0) & (mx <= swo) & (my <= sho);          // odometer.cpp:369-371
      if (vis) {
        pl = patch_place(mx, my, PSZ / 2, width);
#pragma unroll
        for (int k = 0; k < 10; ++k) cf[k] = s_coef[i * 10 + k];
        nvis += (gp == 0);
      }
    }
    if (!vis) continue;
    const int ebase = i * C::N + gp * C::GE + lane;
    if (PSZ == 32) {
      const int a0 = pl.base + (gp * C::KT) * width + lane;
      float c_ = __ldg(Inew + a0 - width), d_ = __ldg(Inew + a0 - width - 1);
#pragma unroll
      for (int j = 0; j < C::KT; ++j) {
        const float a_ = __ldg(Inew + a0 + j * width), b_ = __ldg(Inew + a0 + j * width - 1);
        const float pn = ((pl.w0 * a_ + pl.w1 * b_) + pl.w2 * c_) + pl.w3 * d_;
        c_ = a_; d_ = b_;
        const float pd = s_ref[ebase + 32 * j] - pn;                    // pdiff, odometer.cpp:381
        ax = ax + s_gx[ebase + 32 * j] * pd;
        ay = ay + s_gy[ebase + 32 * j] * pd;
      }
    } else {
#pragma unroll
      for (int j = 0; j < C::KT; ++j) {
        const int qq = gp * C::GE + 32 * j + lane, r = qq / PSZ, c = qq - r * PSZ;
        const float pn = bilin4(Inew, pl.base + r * width + c, width, pl.w0, pl.w1, pl.w2, pl.w3);
        const float pd = s_ref[ebase + 32 * j] - pn;
        ax = ax + s_gx[ebase + 32 * j] * pd;
        ay = ay + s_gy[ebase + 32 * j] * pd;
      }
    }
  }
  fold_jtr(acc, cf, ax, ay);
  warp_sum6_store(acc, &S.part[warp * 24]);
  if (lane == 0) S.part[warp * 24 + 6] = (float)nvis;
}

// ---- solver warp -----------------------------------------------------------------------------------------------
// take the next track from the global ticket and run setpose_se3 (pose.cpp:25-76); lane 0 only
__device__ __forceinline__ void pipe_fetch(const TrackParams& prm, PipeSlot& S, int* ticket) {
  const int k = atomicAdd(ticket, 1);
  if (k >= prm.T) {
    S.track = -1;
    S.par = PAR_NONE;
    S.pending = 0;
    return;
  }
  const int t = k + prm.t0;
  const ict_optparam& op = prm.op;
  float p[6], G[12];
  setpose_se3(prm.p_in + 6 * (int64_t)t, op.donorm != 0, prm.norm + 4 * (int64_t)t, prm.norm[4 * (int64_t)t + 3], p, G);
#pragma unroll
  for (int j = 0; j < 6; ++j) S.p[j] = p[j];
#pragma unroll
  for (int j = 0; j < 12; ++j) { S.G[j] = G[j]; S.G0[j] = G[j]; }
  S.track = t;
  S.level = op.lv_f;
  S.it = 0;
  S.par = PAR_PRE;
  S.pending = 0;
  S.first = 1;
  S.trace_n = 0;
  S.npix = 0;
}

// level finished (lane 0): record the iteration count, go one level finer or finish the track and fetch the next
__device__ __forceinline__ void pipe_level_done(const TrackParams& prm, PipeSlot& S, int* ticket) {
  const ict_optparam& op = prm.op;
  const int t = S.track;
  if (prm.iters) prm.iters[(int64_t)t * (op.lv_f - op.lv_l + 1) + (op.lv_f - S.level)] = S.it;
  if (S.level > op.lv_l) {
    S.level -= 1;
    S.first = 0;
    S.par = PAR_PRE;
    return;
  }
  float p[6], G[12];
#pragma unroll
  for (int j = 0; j < 6; ++j) p[j] = S.p[j];
#pragma unroll
  for (int j = 0; j < 12; ++j) G[j] = S.G[j];
  getpose_se3(p, G, op.donorm != 0, prm.norm + 4 * (int64_t)t, prm.norm[4 * (int64_t)t + 3], prm.p_out + 6 * (int64_t)t);
  if (prm.npixres) prm.npixres[t] = S.npix;
  if (prm.trace)
    for (int k = S.trace_n; k < prm.trace_cap; ++k) {
      float* rec = prm.trace + ((int64_t)t * prm.trace_cap + k) * ICT_TRACE_FLOATS;
      for (int j = 0; j < ICT_TRACE_FLOATS; ++j) rec[j] = 0.0f;
      rec[0] = -1.0f;
    }
  pipe_fetch(prm, S, ticket);
}

// the serial step that follows the slot's last pixel step; called by all 32 lanes of the solver warp
template <int PSZ>
__device__ __forceinline__ void pipe_serial(const TrackParams& prm, PipeSlot& S, int* ticket, int lane) {
  const ict_optparam& op = prm.op;
  if (S.track < 0 || !S.pending) return;   // (the first phase finds slot 1 fetched but not yet processed)
  __syncwarp();
  if (lane == 0) S.pending = 0;
  if (S.par == PAR_PRE) {
    // Hessian: fixed-order sum of the pixel warps' partials, then Hes.fullPivLu() (odometer.cpp:428-472, 514)
    if (lane < 21) {
      float v[NPW];
#pragma unroll
      for (int wv = 0; wv < NPW; ++wv) v[wv] = S.part[wv * 24 + lane];
      float s = v[0];
#pragma unroll
      for (int wv = 1; wv < NPW; ++wv) s = s + v[wv];
      S.Hsum[lane] = s;
    }
    __syncwarp();
    lu6_factor_warp(S.Hsum, S.lu);
    if (lane == 0) {
      S.normdp_init = 1e-10f;                      // odometer.cpp:341-342
      S.normdp = 1e-10f;
      S.it = 0;
      const bool cont = (0 < op.maxiter) & ((S.normdp / S.normdp_init) > op.normdp_ratio);
      if (cont) S.par = PAR_ITER; else pipe_level_done(prm, S, ticket);
    }
    return;
  }
  // PAR_ITER: J^T r sums, solve, additive se(3) update, exp, stop rule (odometer.cpp:399-414)
  if (lane < 7) {
    float v[NPW];
#pragma unroll
    for (int wv = 0; wv < NPW; ++wv) v[wv] = S.part[wv * 24 + lane];
    float s = v[0];
#pragma unroll
    for (int wv = 1; wv < NPW; ++wv) s = s + v[wv];
    S.sum[lane] = s;
  }
  __syncwarp();
  if (lane == 0) {
    if (S.lu.rank == 6) lu6_solve_full_rcp(S.lu, S.sum, S.dp); else lu6_solve(S.lu, S.sum, S.dp);
    float dp[6], pr[6], Gr[12];
#pragma unroll
    for (int k = 0; k < 6; ++k) { dp[k] = S.dp[k]; pr[k] = S.p[k] + dp[k]; S.p[k] = pr[k]; }   // addpose_se3
    Gr[3] = Gr[7] = Gr[11] = 0.0f;
    se3_exp_f32_series(Gr, pr);
#pragma unroll
    for (int k = 0; k < 12; ++k) S.G[k] = Gr[k];
    const float normdp = ((fabsf(dp[0]) + fabsf(dp[2])) + (fabsf(dp[1]) + fabsf(dp[3]))) + (fabsf(dp[4]) + fabsf(dp[5]));
    if (S.it == 0) S.normdp_init = normdp;
    S.normdp = normdp;
    const int nv = (int)S.sum[6];
    if (prm.trace && S.trace_n < prm.trace_cap) {
      float* rec = prm.trace + ((int64_t)S.track * prm.trace_cap + S.trace_n++) * ICT_TRACE_FLOATS;
      rec[0] = (float)S.level;
      rec[1] = (float)S.it;
      for (int k = 0; k < 6; ++k) { rec[2 + k] = S.sum[k]; rec[8 + k] = dp[k]; }
      rec[14] = normdp;
      rec[15] = (float)nv;
      for (int k = 16; k < ICT_TRACE_FLOATS; ++k) rec[k] = 0.0f;
    }
    S.npix += (long long)nv * (PSZ * PSZ);
    S.it += 1;
    const bool cont = (S.it < op.maxiter) & ((S.normdp / S.normdp_init) > op.normdp_ratio);   // odometer.cpp:344-346
    if (!cont) pipe_level_done(prm, S, ticket);
  }
}

template <int PSZ>
__global__ void __launch_bounds__(32 * (NPW + 1), 2) k_track_pipe(const TrackParams prm, int* ticket, int emax, int pmax) {
  extern __shared__ __align__(16) float smem[];
  __shared__ PipeSlot SL[2];
  __shared__ int s_alive[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot_floats = 3 * emax + 10 * pmax;

  if (warp == NPW) {
    if (lane == 0) {
      pipe_fetch(prm, SL[0], ticket);
      pipe_fetch(prm, SL[1], ticket);
      s_alive[0] = (SL[0].track >= 0) | (SL[1].track >= 0);
    }
  }
  __syncthreads();
  int X = 0, ph = 0;
  while (s_alive[ph & 1]) {
    if (warp == NPW) {
      pipe_serial<PSZ>(prm, SL[1 - X], ticket, lane);
      __syncwarp();
      if (lane == 0) s_alive[(ph + 1) & 1] = (SL[0].track >= 0) | (SL[1].track >= 0);
    } else {
      PipeSlot& S = SL[X];
      if (S.track >= 0 && S.par != PAR_NONE) {
        float* base = smem + (size_t)X * slot_floats;
        if (S.par == PAR_PRE)
          pipe_pre<PSZ>(prm, S, base, base + emax, base + 2 * emax, base + 3 * emax, warp, lane);
        else
          pipe_iter<PSZ>(prm, S, base, base + emax, base + 2 * emax, base + 3 * emax, warp, lane);
        if (tid == 0) S.pending = 1;
      }
    }
    __syncthreads();
    X ^= 1;
    ++ph;
  }
}

size_t pipe_smem_bytes(const ict_optparam& op, int max_pts) {
  const size_t P = (size_t)(max_pts < op.maxpttrack ? max_pts : op.maxpttrack);
  return sizeof(float) * 2 * (3 * P * op.novals + 10 * P);
}

template <int PSZ>
static cudaError_t launch_pipe_t(const TrackParams& prm, int max_pts, int* ticket, cudaStream_t stream) {
  static bool attr_set = false;
  static int nsm = 0;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_track_pipe<PSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, ICT_TRACK_SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_track_pipe<PSZ>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int dev = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int P = max_pts < prm.op.maxpttrack ? max_pts : prm.op.maxpttrack;
  const size_t smem = pipe_smem_bytes(prm.op, max_pts);
  const int per_sm = smem * 2 + 8192 <= (size_t)(227 * 1024) ? 2 : 1;
  int grid = nsm * per_sm;
  const int need = (prm.T + 1) / 2;
  if (grid > need) grid = need;
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  k_track_pipe<PSZ><<<grid, 32 * (NPW + 1), smem, stream>>>(prm, ticket, P * prm.op.novals, P);
  count_launch_external();
  return cudaGetLastError();
}

cudaError_t launch_track_pipe(const TrackParams& prm, int max_pts, int* ticket, cudaStream_t stream) {
  switch (prm.op.psz) {
    case 8: return launch_pipe_t<8>(prm, max_pts, ticket, stream);
    case 16: return launch_pipe_t<16>(prm, max_pts, ticket, stream);
    case 32: return launch_pipe_t<32>(prm, max_pts, ticket, stream);
    default: return cudaErrorInvalidConfiguration;
  }
}

}  // namespace ict
