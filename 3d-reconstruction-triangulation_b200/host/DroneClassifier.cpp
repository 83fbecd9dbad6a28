#include "DroneClassifier.h"

#include <stdexcept>
#include <string>

void DroneClassifier::classifyDrones(const DetectionsContainer& container,
                                     std::vector<std::vector<cv::Point3d>>& triangulatedPoints) {
  const int n_frames = container.getFrameCount(), n_cams = container.getCamCount();
  if (n_cams != (int)triangulator_->getCameras().size())
    throw std::runtime_error("tri_b200: the detections have " + std::to_string(n_cams) + " cameras, the rig has " +
                             std::to_string(triangulator_->getCameras().size()));
  std::vector<int32_t> offsets;
  std::vector<double> xy;
  container.toCSR(offsets, xy);
  std::vector<double> paths(3 * n_drones_ * (size_t)n_frames);
  assign_.assign(n_drones_ * (size_t)n_frames * n_cams, -1);
  phase_.assign(n_drones_ * (size_t)n_frames, 0);
  const int st = tri_classify(triangulator_->engine(), triangulator_->mode(), triangulator_->flags(), (int)n_drones_, offsets.data(),
                              xy.data(), n_frames, paths.data(), assign_.data(), phase_.data(), &stats_);
  if (st != TRI_OK) throw std::runtime_error(std::string("tri_b200: ") + tri_last_error());
  for (size_t d = 0; d < n_drones_; d++) {  // appended like the reference does (src/DroneClassifier.cpp:104-107)
    triangulatedPoints.emplace_back();
    std::vector<cv::Point3d>& path = triangulatedPoints.back();
    path.reserve((size_t)n_frames);
    for (int f = 0; f < n_frames; f++) {
      const double* p = &paths[3 * (d * (size_t)n_frames + f)];
      path.emplace_back(p[0], p[1], p[2]);
    }
  }
}
