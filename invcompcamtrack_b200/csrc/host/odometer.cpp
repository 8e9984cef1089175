// odometer.cpp — host side of OdometerClass: argument marshalling only; the alignment runs in the CUDA kernels of
// libictrack.so (Set3Dpoints -> k_set_points, SetPose -> k_reproject, TrackPose -> k_track).
#include "odometer.h"

#include <cstdio>

namespace CTR {

static void report(const char* where) { std::printf("OdometerClass::%s: %s\n", where, ict_last_error()); }

OdometerClass::OdometerClass(PoseClass* pose_in, const optparam* op_in)
    : pose(pose_in), op(op_in), tracker(nullptr), view(nullptr), scratch(nullptr), nopoints(0), n_in(0) {
  const CamClass* cam = pose->camobj;
  tracker = ict_tracker_create(op, cam->fc(), cam->cc(), cam->wh());
  if (!tracker) report("OdometerClass");
  // like the reference object, keep the patch / steepest-descent arrays between TrackPose calls (they are reset by
  // Set3Dpoints only, odometer.cpp:153, 173); configurations the library cannot do that for fall back in TrackPose
  if (tracker) ict_tracker_set_knob(tracker, "keep_state", 1);
  view = ict_frames_create_view(2, cam->wh()[0], cam->wh()[1], op->lv_f, cam->getpadding());
  if (!view) report("OdometerClass");
  pt2d_lvl.assign(2 * (size_t)op->maxpttrack, 0.0f);
  for (int k = 0; k < 6; ++k) p_cur[k] = 0;
  for (int k = 0; k < ICT_MAX_LEVELS; ++k) iters[k] = -1;
}

OdometerClass::~OdometerClass() {
  ict_tracker_destroy(tracker);
  ict_frames_destroy(view);
  ict_frames_destroy(scratch);
}

void OdometerClass::SetSumOrder(int mode) {
  if (tracker && ict_tracker_set_sum_order(tracker, mode) != ICT_OK) report("SetSumOrder");
}

void OdometerClass::Set3Dpoints(double* pt_in, const int nopoints_in) {
  if (!tracker) return;
  if (ict_tracker_set_optparam(tracker, op) != ICT_OK) report("Set3Dpoints: optparam");   // the caller may have changed *op since construction
  const int64_t off[2] = {0, nopoints_in};
  n_in = nopoints_in;
  nopoints = nopoints_in < op->maxpttrack ? nopoints_in : op->maxpttrack;
  // mutate_caller = 1: with donorm the reference centres the caller's array in place (odometer.cpp:207-212)
  if (ict_tracker_set_points(tracker, 1, off, pt_in, 1) != ICT_OK) report("Set3Dpoints");
}

// Makes view[slot] point at the device pyramid behind the host pointer table I[0..lv_f].
bool OdometerClass::bind_frame(int slot, const float** I, const float** dx, const float** dy) {
  ict_frames* store = nullptr;
  int frame = 0;
  if (util_find_device_pyramid(I[0], &store, &frame)) return ict_frames_alias(view, slot, store, frame) == ICT_OK;
  // planes built elsewhere: copy them level by level into a private store
  const CamClass* cam = pose->camobj;
  const int w = cam->wh()[0], h = cam->wh()[1], pad = cam->getpadding();
  int64_t off[ICT_MAX_LEVELS];
  int sw[ICT_MAX_LEVELS], sh[ICT_MAX_LEVELS];
  const int64_t total = ict_pyramid_layout(w, h, op->lv_f, pad, off, sw, sh);
  if (total < 0) return false;
  if (!scratch) scratch = ict_frames_create(2, w, h, op->lv_f, pad);
  if (!scratch) return false;
  std::vector<float> a((size_t)total), b(dx ? (size_t)total : 0), c(dy ? (size_t)total : 0);
  for (int l = 0; l <= op->lv_f; ++l) {
    const size_t n = (size_t)sw[l] * sh[l];
    std::memcpy(a.data() + off[l], I[l], n * sizeof(float));
    if (dx) std::memcpy(b.data() + off[l], dx[l], n * sizeof(float));
    if (dy) std::memcpy(c.data() + off[l], dy[l], n * sizeof(float));
  }
  if (ict_frames_upload_planes(scratch, slot, a.data(), dx ? b.data() : nullptr, dy ? c.data() : nullptr) != ICT_OK)
    return false;
  return ict_frames_alias(view, slot, scratch, slot) == ICT_OK;
}

void OdometerClass::SetPose(const double* p_in, const float** img_ref_in, const float** img_ref_dx_in,
                            const float** img_ref_dy_in, const float** img_new_in) {
  if (!tracker || !view) return;
  if (ict_tracker_set_optparam(tracker, op) != ICT_OK) report("optparam");
  for (int k = 0; k < 6; ++k) p_cur[k] = p_in[k];
  if (!bind_frame(0, img_ref_in, img_ref_dx_in, img_ref_dy_in) || !bind_frame(1, img_new_in, nullptr, nullptr))
    report("SetPose");
  // reference reprojection at lv_l, valid right after SetPose (run_track_nposes.cpp:217-225 relies on it);
  // the library returns x block, y block at the stride the points came in, Get2DPoints() promises maxpttrack
  std::vector<float> q(2 * (size_t)(n_in > 0 ? n_in : 1));
  if (n_in > 0 && ict_tracker_reproject(tracker, p_cur, q.data()) == ICT_OK) {
    for (int i = 0; i < nopoints; ++i) {
      pt2d_lvl[i] = q[i];
      pt2d_lvl[i + op->maxpttrack] = q[n_in + i];
    }
  } else if (n_in > 0) {
    report("SetPose");
  }
  if (!op->donorm) {   // keep the host-side pose object in step (normalised poses stay on the device)
    Eigen::Vector3d zero;
    pose->setpose_se3(p_in, zero, 1.0);
  }
}

void OdometerClass::TrackPose(double* p_out) {
  if (!tracker || !view) return;
  if (ict_tracker_set_optparam(tracker, op) != ICT_OK) report("optparam");
  const int rf = 0, nf = 1;
  int rc = ict_track_batch(tracker, view, &rf, &nf, p_cur, p_out, iters, nullptr, 0, nullptr);
  if (rc == ICT_ERR_UNSUPPORTED) {   // no carried state for this configuration: every TrackPose starts from zeroed arrays
    ict_tracker_set_knob(tracker, "keep_state", 0);
    rc = ict_track_batch(tracker, view, &rf, &nf, p_cur, p_out, iters, nullptr, 0, nullptr);
  }
  if (rc != ICT_OK) {
    report("TrackPose");
    return;
  }
  if (!op->donorm) {   // hand the result back to the host-side pose object (pose.cpp:79-113 reads cpos_p)
    Eigen::Vector3d zero;
    pose->setpose_se3(p_out, zero, 1.0);
  }
}

}  // namespace CTR
