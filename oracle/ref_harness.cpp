// oracle/ref_harness.cpp -- C ABI over the REFERENCE'S OWN classes, compiled unmodified from
// /root/reference/src against oracle/shim (see oracle/Makefile, target _ref).  TEST INFRASTRUCTURE ONLY.
//
// What runs behind every entry point is the reference's code: createCamera / Camera::compCamParams
// (src/utils.cpp:94-107, src/Camera.h:177-187), MatrixTriangulator / RayTriangulator (src/*.cpp),
// DetectionsContainer and DroneClassifier::classifyDrones (src/DroneClassifier.cpp:96-154).  The harness
// only feeds inputs and taps outputs:
//   * the classifier's assignment decisions (DroneClassifier.cpp:130 and :315-321 are not observable through
//     the reference's interface) are recovered from hidden tags on cv::Point2d / cv::Point3d (shim): detections
//     are tagged with their index, a wrapper Triangulator tags every triangulated point with the call that
//     produced it, and the tag travels with the point into the returned paths;
//   * frame boundaries come from the reference's own progress line (DroneClassifier.cpp:113), read through a
//     streambuf installed on std::cout;
//   * phase 1 (tracking) is told from phase 2 by the cv::norm call of DroneClassifier.cpp:244, which only
//     triangulateWithLastPos makes with the candidate as the left operand.
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <sstream>
#include <streambuf>
#include <string>
#include <vector>

#include "DetectionsContainer.h"
#include "DroneClassifier.h"
#include "MatrixTriangulator.h"
#include "RayTriangulator.h"
#include "utils.h"

namespace cv { namespace tap { void (*on_norm3)(uint64_t, uint64_t, double) = nullptr; } }

namespace {

thread_local std::string g_error;

struct Handle {
  std::vector<const tdr::Camera*> cameras;
  Triangulator* tri = nullptr;
  int mode = 0;
};

// one triangulatePoint call of the current frame
struct Call {
  int8_t comb[32];  // per camera: 0 = not in the subset, k = detection k-1 (from the Point2d tags)
  double err;
  int iters;
};

struct Tap : public Triangulator {
  Triangulator* inner;
  std::map<const tdr::Camera*, int> index;
  std::vector<Call> calls;  // calls of the current frame; call id = base + position + 1
  uint64_t base = 0;
  int64_t solves = 0, lm_iters = 0;
  double min_err_margin = 1e300;  // min |error - error_| over every solve (DroneClassifier.cpp:185, :209, :243)
  double threshold = 0;
  explicit Tap(Triangulator* t) : inner(t) {
    cameras = t->getCameras();
    type_ = t->getType();
    for (size_t i = 0; i < cameras.size(); i++) index[cameras[i]] = (int)i;
    threshold = type_ == "matrix" ? MAX_ERROR_MATRIX : MAX_ERROR_RAY;
  }
  std::pair<cv::Point3d, double> triangulatePoint(std::vector<CamPointPair> images) override {
    std::pair<cv::Point3d, double> r = inner->triangulatePoint(images);
    Call c;
    memset(c.comb, 0, sizeof(c.comb));
    for (const CamPointPair& im : images) c.comb[index.at(im.camera)] = (int8_t)im.point.tag;
    c.err = r.second;
    c.iters = type_ == "ray" ? cv::shim::LMSolverImpl::last_iters() : 0;
    calls.push_back(c);
    solves++;
    lm_iters += c.iters;
    min_err_margin = std::min(min_err_margin, std::fabs(r.second - threshold));
    r.first.tag = base + calls.size();
    return r;
  }
  std::vector<cv::Point3d> triangulatePoints(std::vector<std::vector<cv::Point2d>> points) override {
    return inner->triangulatePoints(points);
  }
};

// std::cout sink: counts the "frame / n" lines of DroneClassifier.cpp:113 and calls back at each
struct LineHook : public std::streambuf {
  std::function<void()> on_line;
  int overflow(int ch) override {
    if (ch == '\n' && on_line) on_line();
    return ch;
  }
};

struct Session {  // state of one ref_classify call, reachable from the taps
  Tap* tap = nullptr;
  std::vector<std::vector<cv::Point3d>>* paths = nullptr;
  const DetectionsContainer* container = nullptr;
  std::vector<size_t> seen;        // path lengths at the last frame boundary
  std::set<uint64_t> phase1;       // call ids accepted by triangulateWithLastPos in the current frame
  int frame = -1;                  // frame being processed
  int n_frames = 0, n_cams = 0, n_drones = 0;
  int8_t* assign = nullptr;
  uint8_t* phase = nullptr;
  double* err = nullptr;
  int64_t n_phase1 = 0, n_phase2 = 0;
  double min_step_margin = 1e300;  // min | |c.point - pos| - MAX_STEP | (DroneClassifier.cpp:244)
  double min_gate_margin = 1e300;  // min | getDistFromRay - MAX_STEP |   (DroneClassifier.cpp:231-233)
};
thread_local Session* g_session = nullptr;

void on_norm3(uint64_t tag_a, uint64_t tag_b, double v) {
  Session* s = g_session;
  if (!s || !s->tap) return;
  (void)tag_b;
  if (tag_a > s->tap->base) {  // left operand triangulated in this frame: the compare of DroneClassifier.cpp:244
    s->min_step_margin = std::min(s->min_step_margin, std::fabs(v - MAX_STEP));
    if (v < MAX_STEP) s->phase1.insert(tag_a);
  }
}

// the points pushed onto the paths since the last boundary belong to frame s->frame
void close_frame(Session* s) {
  if (s->frame >= 0) {
    for (int p = 0; p < s->n_drones; p++) {
      const std::vector<cv::Point3d>& path = (*s->paths)[p];
      if (path.size() == s->seen[p]) continue;
      const cv::Point3d& pt = path.back();
      const uint64_t id = pt.tag;
      const Call& c = s->tap->calls.at((size_t)(id - s->tap->base - 1));
      const size_t o = (size_t)p * s->n_frames + s->frame;
      if (s->assign) for (int k = 0; k < s->n_cams; k++) s->assign[o * s->n_cams + k] = c.comb[k];
      const bool p1 = s->phase1.count(id) != 0;
      if (s->phase) s->phase[o] = p1 ? 1 : 2;
      if (s->err) s->err[o] = c.err;
      (p1 ? s->n_phase1 : s->n_phase2)++;
      s->seen[p] = path.size();
    }
  }
  s->tap->base += s->tap->calls.size();
  s->tap->calls.clear();
  s->phase1.clear();
}

void open_frame(Session* s) {
  close_frame(s);
  s->frame++;
  // margin audit of the ray gate (DroneClassifier.cpp:228-236), with the reference's own getDistFromRay
  for (int p = 0; p < s->n_drones; p++) {
    const std::vector<cv::Point3d>& path = (*s->paths)[p];
    if (path.empty() || !(path.back() != cv::Point3d(0, 0, 0))) continue;
    for (int cam = 0; cam < s->n_cams; cam++)
      for (int det = 0; det < s->container->detCountForCam(cam, s->frame); det++) {
        const double d = Triangulator::getDistFromRay({s->tap->getCamera(cam), s->container->getRecord(cam, s->frame, det)}, path.back());
        s->min_gate_margin = std::min(s->min_gate_margin, std::fabs(d - MAX_STEP));
      }
  }
}

template <typename F>
int guarded(F f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_error.c_str(); }

// mode 0 = MatrixTriangulator, 1 = RayTriangulator (src/main.cpp:54-63); cameras through createCamera (src/utils.cpp:94-107)
void* ref_create(int n_cams, const int* ids, const int* width, const int* height, const double* focal, const double* pos,
                 const double* quat, int mode) {
  try {
    Handle* h = new Handle();
    for (int c = 0; c < n_cams; c++)
      h->cameras.push_back(createCamera(ids[c], (size_t)width[c], (size_t)height[c], focal[c],
                                        (cv::Mat_<double>(3, 1) << pos[3 * c], pos[3 * c + 1], pos[3 * c + 2]),
                                        (cv::Mat_<double>(4, 1) << quat[4 * c], quat[4 * c + 1], quat[4 * c + 2], quat[4 * c + 3])));
    h->mode = mode;
    h->tri = mode == 0 ? (Triangulator*)new MatrixTriangulator(h->cameras) : (Triangulator*)new RayTriangulator(h->cameras);
    return h;
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}

// the same through the reference's own XML loader (src/utils.cpp:46-92)
void* ref_create_xml(const char* path, int mode) {
  try {
    Handle* h = new Handle();
    h->cameras = loadCamerasXML(path);
    h->mode = mode;
    h->tri = mode == 0 ? (Triangulator*)new MatrixTriangulator(h->cameras) : (Triangulator*)new RayTriangulator(h->cameras);
    return h;
  } catch (const std::exception& e) {
    g_error = e.what();
    return nullptr;
  }
}

void ref_destroy(void* hv) {
  Handle* h = (Handle*)hv;
  if (!h) return;
  delete h->tri;
  for (const tdr::Camera* c : h->cameras) delete c;
  delete h;
}

int ref_n_cameras(void* hv) { return (int)((Handle*)hv)->cameras.size(); }

// camera constants as compCamParams left them: P (3x4), K (3x3), E (3x4), fovx, fovy, fx, fy, cx, cy, width, height
int ref_camera(void* hv, int cam, double P[12], double K[9], double E[12], double scal[8]) {
  return guarded([&]() {
    const tdr::Camera* c = ((Handle*)hv)->cameras.at(cam);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 4; j++) { P[i * 4 + j] = c->cameraPerspectiveMatrix.at<double>(i, j); E[i * 4 + j] = c->cameraExtrinsicMatrix.at<double>(i, j); }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) K[i * 3 + j] = c->cameraMatrix.at<double>(i, j);
    const double s[8] = {c->fovx, c->fovy, c->fx, c->fy, (double)c->cx, (double)c->cy, (double)c->width, (double)c->height};
    memcpy(scal, s, sizeof(s));
    return 0;
  });
}

// Triangulator::triangulatePoints (MatrixTriangulator.cpp:70-100 / RayTriangulator.cpp:51-81); xy is [n_point_cams][n_frames][2].
// Returns 0, or 1 with the runtime_error text in ref_last_error().
int ref_triangulate_points(void* hv, const double* xy, int n_point_cams, int64_t n_frames, double* out_xyz) {
  return guarded([&]() {
    std::vector<std::vector<cv::Point2d>> pts((size_t)n_point_cams);
    for (int c = 0; c < n_point_cams; c++) {
      pts[c].resize((size_t)n_frames);
      for (int64_t f = 0; f < n_frames; f++) pts[c][f] = cv::Point2d(xy[2 * ((size_t)c * n_frames + f)], xy[2 * ((size_t)c * n_frames + f) + 1]);
    }
    const std::vector<cv::Point3d> r = ((Handle*)hv)->tri->triangulatePoints(pts);
    for (size_t f = 0; f < r.size(); f++) { out_xyz[3 * f] = r[f].x; out_xyz[3 * f + 1] = r[f].y; out_xyz[3 * f + 2] = r[f].z; }
    return 0;
  });
}

// Triangulator::triangulatePoint (MatrixTriangulator.cpp:3-62 / RayTriangulator.cpp:83-107) on one camera subset
int ref_triangulate_point(void* hv, int n, const int* cam_idx, const double* xy, double X[3], double* err, int* iters) {
  return guarded([&]() {
    Handle* h = (Handle*)hv;
    std::vector<Triangulator::CamPointPair> images;
    for (int i = 0; i < n; i++) images.push_back({h->cameras.at(cam_idx[i]), cv::Point2d(xy[2 * i], xy[2 * i + 1])});
    const std::pair<cv::Point3d, double> r = h->tri->triangulatePoint(images);
    X[0] = r.first.x; X[1] = r.first.y; X[2] = r.first.z;
    if (err) *err = r.second;
    if (iters) *iters = h->mode == 1 ? cv::shim::LMSolverImpl::last_iters() : 0;
    return 0;
  });
}

// Triangulator::getDistFromRay (Triangulator.cpp:57-61)
double ref_dist_from_ray(void* hv, int cam, double x, double y, const double p[3]) {
  Handle* h = (Handle*)hv;
  return Triangulator::getDistFromRay({h->cameras.at(cam), cv::Point2d(x, y)}, cv::Point3d(p[0], p[1], p[2]));
}

// DroneClassifier::classifyDrones (DroneClassifier.cpp:96-154).  Detections in CSR form as in tri_b200.h / tri_oracle.h.
// out_paths [n_drones][n_frames][3]; out_assign [n_drones][n_frames][n_cams] (0 = camera unused, k = detection k-1, -1 = no
// point); out_phase [n_drones][n_frames] (0 none, 1 tracking, 2 re-initialisation); out_err [n_drones][n_frames] the accepted
// combination's error.  stats: solves, lm_iters, phase1, phase2.  margins: min |error - error_|, min ||c.point - pos| - MAX_STEP|,
// min |getDistFromRay - MAX_STEP|.
int ref_classify(void* hv, int n_drones, const int32_t* det_offsets, const double* dets_xy, int n_frames, double* out_paths,
                 int8_t* out_assign, uint8_t* out_phase, double* out_err, int64_t stats[4], double margins[3]) {
  return guarded([&]() {
    Handle* h = (Handle*)hv;
    const int n_cams = (int)h->cameras.size();
    DetectionsContainer container(n_cams);  // DetectionsContainer.cpp:11-17; frames appended as readFiles would leave them
    for (int f = 0; f < n_frames; f++) {
      container.addEmptyFrame();
      for (int c = 0; c < n_cams; c++) {
        const int32_t* o = det_offsets + (size_t)c * (n_frames + 1);
        for (int k = o[f]; k < o[f + 1]; k++) {
          cv::Point2d p(dets_xy[2 * (size_t)k], dets_xy[2 * (size_t)k + 1]);
          p.tag = (uint64_t)(k - o[f] + 1);
          container.addDetectionToCamera(p, c);
        }
      }
    }
    container.n_frames = n_frames;
    Tap tap(h->tri);
    DroneClassifier classifier(&tap, (size_t)n_drones);
    std::vector<std::vector<cv::Point3d>> paths;
    Session s;
    s.tap = &tap; s.paths = &paths; s.container = &container;
    s.n_frames = n_frames; s.n_cams = n_cams; s.n_drones = n_drones;
    s.assign = out_assign; s.phase = out_phase; s.err = out_err;
    s.seen.assign((size_t)n_drones, 0);
    if (out_assign) memset(out_assign, 0xff, (size_t)n_drones * n_frames * n_cams);
    if (out_phase) memset(out_phase, 0, (size_t)n_drones * n_frames);
    if (out_err) for (size_t i = 0; i < (size_t)n_drones * n_frames; i++) out_err[i] = 0;
    LineHook hook;
    hook.on_line = [&]() { open_frame(&s); };
    std::streambuf* old = std::cout.rdbuf(&hook);
    g_session = &s;
    cv::tap::on_norm3 = on_norm3;
    try {
      classifier.classifyDrones(container, paths);
    } catch (...) {
      std::cout.rdbuf(old);
      g_session = nullptr;
      cv::tap::on_norm3 = nullptr;
      throw;
    }
    std::cout.rdbuf(old);
    // the zero insertions of DroneClassifier.cpp:147-153 ran after the last frame: lengths before them are what close_frame needs
    // (an inserted zero sits at a frame without a point, so back() of a path that grew in the last frame is still its point only
    // if the last frame was not empty for it) -- undo by comparing against the counts instead
    {
      // points per path so far (without zeros) + zeros == n_frames; a path grew in the last frame iff its non-zero count exceeds `seen`
      for (int p = 0; p < n_drones; p++) {
        const std::vector<cv::Point3d>& path = paths[p];
        size_t real = 0;
        const cv::Point3d* last = nullptr;
        for (const cv::Point3d& q : path)
          if (q.tag != 0) { real++; last = &q; }
        if (real > s.seen[p] && last) {
          const uint64_t id = last->tag;
          const Call& c = tap.calls.at((size_t)(id - tap.base - 1));
          const size_t o = (size_t)p * n_frames + s.frame;
          if (out_assign) for (int k = 0; k < n_cams; k++) out_assign[o * n_cams + k] = c.comb[k];
          const bool p1 = s.phase1.count(id) != 0;
          if (out_phase) out_phase[o] = p1 ? 1 : 2;
          if (out_err) out_err[o] = c.err;
          (p1 ? s.n_phase1 : s.n_phase2)++;
        }
      }
    }
    g_session = nullptr;
    cv::tap::on_norm3 = nullptr;
    for (int p = 0; p < n_drones; p++) {
      if ((int)paths[p].size() != n_frames) throw std::runtime_error("path length differs from the frame count");
      for (int f = 0; f < n_frames; f++) {
        double* o = out_paths + 3 * ((size_t)p * n_frames + f);
        o[0] = paths[p][f].x; o[1] = paths[p][f].y; o[2] = paths[p][f].z;
      }
    }
    if (stats) { stats[0] = tap.solves; stats[1] = tap.lm_iters; stats[2] = s.n_phase1; stats[3] = s.n_phase2; }
    if (margins) { margins[0] = tap.min_err_margin; margins[1] = s.min_step_margin; margins[2] = s.min_gate_margin; }
    return 0;
  });
}

}  // extern "C"
