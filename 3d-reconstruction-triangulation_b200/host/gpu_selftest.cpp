// gpu_selftest.cpp -- exercises the C++ Triangulator adapters on the GPU the way the reference's callers
// do (src/DroneClassifier.cpp:176-181,231-232; src/main.cpp:54-63) and prints the results as text for
// tests/test_host_cpp.py.  usage: host_gpu_selftest <cameras.xml> <csv dir>
#include <cstdio>
#include <stdexcept>

#include "DetectionsContainer.h"
#include "Triangulator.h"
#include "utils.h"

static void dump(const char* tag, const std::vector<cv::Point3d>& pts) {
  for (size_t i = 0; i < pts.size(); i++) std::printf("%s %zu %.17g %.17g %.17g\n", tag, i, pts[i].x, pts[i].y, pts[i].z);
}

int main(int argc, const char** argv) {
  if (argc < 3) return 2;
  std::vector<const tdr::Camera*> cams = loadCamerasXML(argv[1]);
  DetectionsContainer box(argv[2], 0, 7);
  Triangulator* matrix = new MatrixTriangulator(cams);
  Triangulator* ray = new RayTriangulator(cams);
  std::printf("types %s %s %zu\n", matrix->getType().c_str(), ray->getType().c_str(), matrix->getCameras().size());

  std::vector<std::vector<cv::Point2d>> pts = box.getDataForTriangulation();
  dump("matrix", matrix->triangulatePoints(pts));
  dump("ray", ray->triangulatePoints(pts));
  // the same detections shifted off the integer grid: the adapter then packs float2 (+0.25) / double2 (+1e-7)
  // instead of ushort2
  for (double shift : {0.25, 1e-7}) {
    std::vector<std::vector<cv::Point2d>> q = pts;
    for (auto& row : q) for (cv::Point2d& p : row) { p.x += shift; p.y += shift; }
    dump(shift == 0.25 ? "matrix_f32" : "matrix_f64", matrix->triangulatePoints(q));
  }

  // triangulatePoint on a camera subset, the way fillCombinationQueue builds it
  std::vector<Triangulator::CamPointPair> images;
  for (int c = 0; c < box.getCamCount(); c += 2) images.push_back({matrix->getCamera(c), box.getRecord(c, 0, 0)});
  auto one = matrix->triangulatePoint(images);
  std::printf("point_matrix %.17g %.17g %.17g %.17g\n", one.first.x, one.first.y, one.first.z, one.second);
  std::vector<Triangulator::CamPointPair> rimages;
  for (int c = 0; c < box.getCamCount(); c += 2) rimages.push_back({ray->getCamera(c), box.getRecord(c, 0, 0)});
  auto oner = ray->triangulatePoint(rimages);
  std::printf("point_ray %.17g %.17g %.17g %.17g\n", oner.first.x, oner.first.y, oner.first.z, oner.second);
  std::printf("dist %.17g\n", Triangulator::getDistFromRay({matrix->getCamera(1), box.getRecord(1, 0, 0)}, one.first));

  // the reference's exceptions
  try {
    std::vector<std::vector<cv::Point2d>> ragged = pts;
    ragged[1].pop_back();
    matrix->triangulatePoints(ragged);
    std::printf("throw_dim none\n");
  } catch (const std::runtime_error& e) { std::printf("throw_dim %s\n", e.what()); }
  for (Triangulator* t : {matrix, ray}) {
    try {
      std::vector<std::vector<cv::Point2d>> few = pts;
      for (size_t c = 1; c < few.size(); c++) few[c][2] = cv::Point2d(-1, -1);
      t->triangulatePoints(few);
      std::printf("throw_few none\n");
    } catch (const std::runtime_error& e) { std::printf("throw_few %s\n", e.what()); }
  }
  try {
    matrix->triangulatePoint({images[0]});
    std::printf("throw_one none\n");
  } catch (const std::runtime_error& e) { std::printf("throw_one %s\n", e.what()); }
  delete matrix;
  delete ray;
  for (const auto& cam : cams) delete cam;
  return 0;
}
