// ict_kernels.cuh — shared declarations between the kernels (ict_kernels.cu) and the C ABI (ict_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ictrack.h"
#include "ict_knobs.h"

namespace ict {

// per-level intrinsics, CamClass (camera.cpp:32-43); width = (int)getsw(l) as util_getPatch receives it
struct CamLevels {
  float fx[ICT_MAX_LEVELS], fy[ICT_MAX_LEVELS], cx[ICT_MAX_LEVELS], cy[ICT_MAX_LEVELS];
  float swo[ICT_MAX_LEVELS], sho[ICT_MAX_LEVELS];
  int width[ICT_MAX_LEVELS];
};

// device pointers to the padded level planes of one frame
struct FrameDesc {
  const float* I[ICT_MAX_LEVELS];
  const float* dx[ICT_MAX_LEVELS];
  const float* dy[ICT_MAX_LEVELS];
  const void* tmap;                // device array of CUtensorMap [3 planes: I, dx, dy][ICT_MAX_LEVELS] (128 bytes each):
                                   //    2-D tiled maps of the padded level planes, box 40 x 17 floats (K2r); null when
                                   //    the geometry does not allow TMA (row pitch not a multiple of 16 bytes)
};

struct TrackParams {
  ict_optparam op;
  CamLevels cam;
  const FrameDesc* frames;
  const int* ref_frame;
  const int* new_frame;
  int fixed_ref, fixed_new;        // used when ref_frame/new_frame == nullptr (sequence chains)
  const int64_t* pt_off;           // [T+1]
  const float* pt3d;               // per track: X block, Y block, Z block (n_t each) at 3*pt_off[t]
  const double* norm;              // per track: meanshift[3], varval
  const double* p_in;              // [T*6]
  double* p_out;                   // [T*6]
  int* iters;                      // [T*L] or null
  float* trace;                    // [T*trace_cap*16] or null
  int trace_cap;
  const float* teacher;            // [T*trace_cap*8] or null (tests): teacher forcing for the fast-mode kernels — after
                                   //    iteration record r of track t the pose coefficients become teacher[(t*cap+r)*8 + 0..5]
                                   //    instead of the kernel's own p + delta_p, and the loop continues iff [6] != 0; the
                                   //    trace still records the kernel's OWN J^T r and delta_p.  Needs trace != null.
  long long* npixres;              // [T] or null
  float* pt2d_out;                 // [2*total] or null: reference 2-D points at lv_l (Get2DPoints)
  int T;                           // tracks in this launch
  int t0;                          // first track of this launch (index into the per-track arrays)
  int serial_warp_last;            // profiling knob: run the serial sections on the CTA's last warp instead of warp 0
  int dbg_skip_serial;             // profiling experiment: skip solve/update (results meaningless)
  int v2_lu_setup;                 // K2v2 A/B knob: per-level solve matrix from the warp LU instead of the sweeps
  int force_general;               // 1: always use the general kernel k_track (tests compare the two)
  int seq_n, seq_step;             // K2v8 only: seq_n > 1 runs a whole chain in one launch — step k tracks frame
                                   //    fixed_ref + k*seq_step -> + seq_step from pose p_in + 6*T*k to p_out + 6*T*k
                                   //    (iters + T*L*k, npixres + T*k); 0/1: a single step
  float* state;                    // K2x8 with the tracker's "keep_state": per-track template state carried between calls
  int64_t state_stride;            //    (floats per track: 3*64 + 12 per point of the largest track), null: reset every call
  int state_load;                  //    0: the first call after Set3Dpoints (arrays start zeroed, odometer.cpp:173)
  unsigned robust;                 // ICT_ROBUST_* flags (ictrack.h): opt-in deviations from the reference, fast mode / psz 8 (K2v8)
  int knob_no_k2r;                 // tracker knob "no_k2r": reference-order psz 32 runs K2x even where K2r applies (A/B, tests)
  int knob_seq_launches;           // tracker knob "seq_launches": a chain is one launch per frame step even where K2v8 could loop
  int tma_ok;                      // every frame of the store carries tensor maps (FrameDesc.tmap)
  int r_pcap;                      // K2r: points per track the shared-memory layout is sized for (set by its launcher)
  int sum_mode;                    // 0: fixed-order tree reductions (fast); 1: Eigen-3.3 packet order (bit-exact
                                   //    with the oracle's default model of the reference, ~3x slower)
};

#define ICT_TRACK_SMEM_LIMIT (227 * 1024 - 8192)

// ---- launches (all asynchronous on `stream`) -------------------------------------------------------------------
// K0: util_constructpyramide for `count` frames.  src is float (src_u8 == nullptr) or uint8.
cudaError_t launch_pyramid(const float* src_f32, const unsigned char* src_u8, int count, int w, int h, int lv_f,
                           int pad, float* I, float* dx, float* dy, int64_t plane_floats,
                           const int64_t* level_off, cudaStream_t stream);

// a5: Set3Dpoints for T tracks (double SoA in, float SoA + normalisation out; pts_mut gets the centred doubles)
cudaError_t launch_set_points(int T, const int64_t* pt_off, const double* pts, double* pts_mut, float* pt3d,
                              double* norm, int donorm, int maxpttrack, int max_pts, cudaStream_t stream);

// a6..a18: SetPose + TrackPose, one CTA per track, template + gradients resident in shared memory.
// Returns cudaErrorInvalidConfiguration when a track does not fit (caller then uses the multi-CTA path).
size_t track_smem_bytes(const ict_optparam& op, int max_pts, int sum_mode);
bool track_fits_one_cta(const ict_optparam& op, int max_pts, int sum_mode, int force_general);
bool track_chain_in_one_launch(const ict_optparam& op, int max_pts, int sum_mode, int force_general, int per_frame_launches);
cudaError_t launch_track(const TrackParams& prm, int max_pts, cudaStream_t stream);

// K2v2 (ict_kernel_v2.cu): the production kernel for psz 32 without dopatchnorm, tree sums; launch_track routes
// to it unless ICT_FAST_V1 is set (then k_track_fast, the previous production kernel, runs — for A/B measurements).
size_t v2_smem_bytes(const ict_optparam& op, int max_pts);
cudaError_t launch_track_v2(const TrackParams& prm, int max_pts, cudaStream_t stream);

// K2v8 (ict_kernel_v8.cu): the K2v2 scheme for 8x8 patches (the reference's own configuration), up to 240 points per
// track, with or without dopatchnorm, tree sums; honours TrackParams.seq_n (a chain of frame steps in one launch).
bool v8_supported(const ict_optparam& op, int max_pts);
size_t v8_smem_bytes(const ict_optparam& op, int max_pts);
cudaError_t launch_track_v8(const TrackParams& prm, int max_pts, cudaStream_t stream);

// K2x (ict_kernel_x.cu): reference-order (Eigen packet) sums for psz 32 without dopatchnorm — bit-identical to the
// oracle like k_track<32,2>, with producer warps and one chain warp; launch_track routes sum_mode 1 to it unless
// ICT_EXACT_V1 is set.
size_t kx_smem_bytes(const ict_optparam& op, int max_pts);
cudaError_t launch_track_x(const TrackParams& prm, int max_pts, cudaStream_t stream);

// K2r (ict_kernel_r.cu): reference-order sums for psz 32 without dopatchnorm, up to 8 points per track, with the
// steepest-descent images resident in shared memory and the windows staged by 2-D TMA; needs prm.tma_ok.
bool kr_supported(const ict_optparam& op, int max_pts, int tma_ok);
size_t kr_smem_bytes(const ict_optparam& op, int max_pts);
cudaError_t launch_track_r(const TrackParams& prm, int max_pts, cudaStream_t stream);

// K2x8 (ict_kernel_x8.cu): reference-order sums for 8x8 patches (up to 224 points per track), with or without
// dopatchnorm — bit-identical to the oracle like k_track<8,2|3>.
bool kx8_supported(const ict_optparam& op, int max_pts);
size_t kx8_smem_bytes(const ict_optparam& op, int max_pts);
cudaError_t launch_track_x8(const TrackParams& prm, int max_pts, cudaStream_t stream);

// SetPose only: setpose_se3 + reprojection at lv_l into prm.pt2d_out (one CTA per track)
cudaError_t launch_reproject(const TrackParams& prm, cudaStream_t stream);

// multi-CTA path for one big track (dense alignment: psz=1, millions of points)
struct BigTrackWork;
size_t bigtrack_work_bytes(const ict_optparam& op, int64_t npts);
cudaError_t launch_track_big(const TrackParams& prm, int t, int64_t npts, void* work, cudaStream_t stream);

// NCC hypothesis scoring, run_track_nposes.cpp:271-355
cudaError_t launch_ncc(const ict_optparam& op, const CamLevels& cam, const float* img_b, const float* img_r,
                       const float* img_f, int nback, int nfwd, const int64_t* pt_off, int T, const float* pb,
                       const float* pr, const float* pf, float* out, cudaStream_t stream);

// util_getPatch / util_getPatch_grad for npatch centres (mids: x, y pairs) on one level plane set
cudaError_t launch_get_patches(const float* I, const float* dx, const float* dy, int width, int psz, int pszd2,
                               int patchnorm, int npatch, const float* mids, float* out_I, float* out_dx, float* out_dy,
                               cudaStream_t stream);

// f3: pose hypotheses from minimal 4-point samples + inlier sets (ict_hypotheses.cu); all pointers device memory
cudaError_t launch_hypotheses(const double fc[2], const double cc[2], int npts, const double* pt2d, const double* pt3d,
                              int nsamples, const int* sample, const double p_init[6], double inlthresh, int maxiter,
                              double* pose, int* status, int* ninl, unsigned char* mask, cudaStream_t stream);

int64_t launch_count(int reset);

}  // namespace ict
