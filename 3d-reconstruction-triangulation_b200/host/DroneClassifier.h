// DroneClassifier.h -- the reference's DroneClassifier (src/DroneClassifier.h) as a thin host object
// over tri_classify: same constructor, same classifyDrones signature and result layout
// (triangulatedPoints[drone][frame], (0,0,0) where a path got no point).
#pragma once
#include <cstdint>
#include <vector>

#include "DetectionsContainer.h"
#include "Triangulator.h"

#define MAX_ERROR_MATRIX 1e+5  // enforced on the device (csrc/tri_classify.cu), src/DroneClassifier.h:11-17
#define MAX_ERROR_RAY 120
#define MAX_STEP 200
#define MIN_CAMERAS 2
#define PATH_TAIL 3

class DroneClassifier {
 public:
  DroneClassifier(Triangulator* triangulator, size_t n_drones) : triangulator_(triangulator), n_drones_(n_drones) {}

  void classifyDrones(const DetectionsContainer& container, std::vector<std::vector<cv::Point3d>>& triangulatedPoints);

  // what the reference computes but never returns: per (drone, frame, camera) the chosen detection
  // (0 = camera unused, k = detection k-1, -1 = no point) and the phase (0 none, 1 tracking, 2 re-init)
  const std::vector<int8_t>& assignments() const { return assign_; }
  const std::vector<uint8_t>& phases() const { return phase_; }
  const tri_classify_stats& stats() const { return stats_; }
  void setProgress(bool on) { progress_ = on; }  // the per-frame "frame / n" lines of src/DroneClassifier.cpp:113 (on by default)

 private:
  Triangulator* triangulator_;
  size_t n_drones_;
  std::vector<int8_t> assign_;
  std::vector<uint8_t> phase_;
  tri_classify_stats stats_{};
  bool progress_ = true;
};
