/* tri_oracle.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libtri_b200.so and the C++ host above it) never links it.
 *
 * Reference: Grzetan/3D-Reconstruction-Triangulation.  The reference cannot be compiled here
 * (needs OpenCV C++, absent), so this is a restatement; it is pinned against the Python/cv2 twin
 * (oracle/py_twin.py, real OpenCV arithmetic) through tests/golden/ -- see oracle/README.md.
 */
#ifndef TRI_ORACLE_H
#define TRI_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MATRIX = 0, ORC_RAY = 1 };
enum { ORC_OK = 0, ORC_ERR_DIM = 1, ORC_ERR_TOO_FEW = 2, ORC_ERR_CAMERA = 3 };

/* tdr::Camera (src/Camera.h:33-300) reduced to what the hot path reads. */
typedef struct orc_camera {
  int id, width, height, cx, cy;
  double focal, fx, fy, fovx, fovy;
  double pos[3];     /* "tvec": world position, src/utils.cpp:101            */
  double quat[4];    /* "rquat": (w,i,j,k) raw, src/utils.cpp:103            */
  double cam_pos[3]; /* "camPos" = -R*tvec, src/Camera.h:167-170             */
  double K[9];       /* cameraMatrix, src/Camera.h:123-128                   */
  double E[12];      /* cameraExtrinsicMatrix, src/Camera.h:133-153          */
  double P[12];      /* cameraPerspectiveMatrix, src/Camera.h:159-161        */
} orc_camera;

typedef struct orc_stats {
  int64_t nodes, solves, leaves, lm_iters, ties, phase1, phase2;
} orc_stats;

int orc_camera_make(orc_camera* c, int id, int width, int height, double focal, const double pos[3],
                    const double quat[4]);

/* cv::invert(A, DECOMP_SVD) for an m x n (n<=4, m>=n) row-major A; pinv is n x m row-major. */
void orc_pinv_svd(const double* A, int m, int n, double* pinv);
/* cv::Mat::inv() of a 3x3 (closed-form branch of cv::invert); returns 0 for a singular matrix */
int orc_inv3(const double S[9], double T[9]);
/* cv::solve(A, b, DECOMP_EIG) and diag(cv::invert(A, DECOMP_EIG)) for symmetric 3x3 A. */
void orc_eig_solve3(const double A[9], const double b[3], double x[3]);
void orc_eig_inv_diag3(const double A[9], double diag[3]);

void orc_make_ray(const orc_camera* cam, double x, double y, double origin[3], double dir[3]);
double orc_dist_to_ray(const double origin[3], const double dir[3], const double p[3]);
double orc_dist_from_ray(const orc_camera* cam, double x, double y, const double p[3]);

/* Triangulator::triangulatePoint for n (camera index, pixel) pairs; returns the error. */
double orc_matrix_point(const orc_camera* cams, int n, const int* cam_idx, const double* xy, double X[3]);
double orc_ray_point(const orc_camera* cams, int n, const int* cam_idx, const double* xy, double X[3],
                     int* iters);
/* Exact minimiser of sum |d x (p-o)|^2 (not a reference function; independent check). */
void orc_ray_closed_form(const orc_camera* cams, int n, const int* cam_idx, const double* xy, double X[3]);

/* Triangulator::triangulatePoints.  xy is [n_point_cams][n_frames][2] with the (-1,-1) sentinel.
 * allow_too_few = 0 reproduces the throw (returns ORC_ERR_TOO_FEW at the first such frame);
 * = 1 writes (0,0,0) and mask<2 bits instead (synthetic-benchmark convention).
 * nthreads > 1 uses OpenMP over frames (CPU baseline). */
int orc_triangulate_points(const orc_camera* cams, int n_cams, int n_point_cams, int mode, const double* xy,
                           int64_t n_frames, int allow_too_few, double* out_xyz, double* out_err,
                           uint32_t* out_mask, int32_t* out_iters, int nthreads);
/* Same, float2 input (the layout the CUDA engine reads) */
int orc_triangulate_points_f32(const orc_camera* cams, int n_cams, int n_point_cams, int mode,
                               const float* xy, int64_t n_frames, int allow_too_few, double* out_xyz,
                               double* out_err, uint32_t* out_mask, int32_t* out_iters, int nthreads);

/* DroneClassifier::classifyDrones.  Detections in CSR form: det_offsets[cam*(n_frames+1)+f] indexes
 * dets_xy (pairs) ordered [cam][frame][det].  out_paths is [n_drones][n_frames][3];
 * out_assign is [n_drones][n_frames][n_cams] combination indices (0 = camera unused, k = detection
 * k-1), all -1 when the path received no point in that frame; out_phase [n_drones][n_frames] is
 * 0 none / 1 tracking (DroneClassifier.cpp:119-135) / 2 re-initialisation (:140-143). */
int orc_classify(const orc_camera* cams, int n_cams, int mode, int n_drones, const int32_t* det_offsets,
                 const double* dets_xy, int n_frames, double* out_paths, int8_t* out_assign,
                 uint8_t* out_phase, orc_stats* stats);

/* Margin audit of the last orc_classify call on this thread: the smallest |lhs - rhs| over the compares that
 * decide an index -- [0] error vs error_, [1] step vs MAX_STEP, [2] ray gate vs MAX_STEP, [3] error vs error
 * within a priority class, [4] path-tail distances (HUGE_VAL where no such compare happened). */
void orc_last_margins(double out[5]);

/* fillCombinationQueue over one full frame: leaves in DFS order.  Returns the number of leaves
 * (may exceed max_leaves; only max_leaves are written). */
int orc_enumerate_frame(const orc_camera* cams, int n_cams, int mode, const int32_t* det_offsets,
                        const double* dets_xy, int n_frames, int frame, int max_leaves, int8_t* out_comb,
                        double* out_xyz, double* out_err, orc_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
