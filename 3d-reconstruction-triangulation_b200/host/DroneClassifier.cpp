#include "DroneClassifier.h"

#include <stdlib.h>

#include <algorithm>
#include <iostream>
#include <stdexcept>
#include <string>

void DroneClassifier::classifyDrones(const DetectionsContainer& container,
                                     std::vector<std::vector<cv::Point3d>>& triangulatedPoints) {
  const int n_frames = container.getFrameCount(), n_cams = container.getCamCount();
  if (n_cams != (int)triangulator_->getCameras().size())
    throw std::runtime_error("tri_b200: the detections have " + std::to_string(n_cams) + " cameras, the rig has " +
                             std::to_string(triangulator_->getCameras().size()));
  // the reference announces every frame on stdout, flushed (src/DroneClassifier.cpp:113); the batched engine classifies the
  // whole sequence in one call, so the lines come first.  setProgress(false) / TRI_B200_QUIET=1 silences them.
  if (progress_ && !getenv("TRI_B200_QUIET"))
    for (int frame = 0; frame < n_frames; frame++) std::cout << frame << " / " << n_frames << std::endl;
  std::vector<int32_t> offsets;
  std::vector<double> xy;
  container.toCSR(offsets, xy);
  std::vector<double> paths(3 * n_drones_ * (size_t)n_frames);
  assign_.assign(n_drones_ * (size_t)n_frames * n_cams, -1);
  phase_.assign(n_drones_ * (size_t)n_frames, 0);
  // TRI_B200_GPUS=N (N > 1): the sequence is frame-sharded over N engines -- candidate generation on every GPU at once,
  // linking along the chain (tri_classify_multi; same result bit for bit).  Engine g sits on GPU g modulo the
  // number of GPUs, so the sharded path can also be exercised on a one-GPU box.
  int n_shards = 1;
  if (const char* v = getenv("TRI_B200_GPUS")) n_shards = std::max(1, std::min(64, atoi(v)));
  int st;
  if (n_shards == 1) {
    st = tri_classify(triangulator_->engine(), triangulator_->mode(), triangulator_->flags(), (int)n_drones_, offsets.data(),
                      xy.data(), n_frames, paths.data(), assign_.data(), phase_.data(), &stats_);
  } else {
    const int n_dev = std::max(1, tri_device_count()), dev0 = tri_engine_device(triangulator_->engine());
    std::vector<tri_camera> desc;
    for (const tdr::Camera* c : triangulator_->getCameras()) desc.push_back(c->describe());
    std::vector<tri_engine*> engines(1, triangulator_->engine());
    st = TRI_OK;
    for (int g = 1; g < n_shards && st == TRI_OK; g++) {
      tri_engine* e = nullptr;
      st = tri_create((int)desc.size(), desc.data(), (dev0 + g) % n_dev, &e);
      if (st == TRI_OK) engines.push_back(e);
    }
    if (st == TRI_OK)
      st = tri_classify_multi(engines.data(), (int)engines.size(), triangulator_->mode(), triangulator_->flags(), (int)n_drones_,
                              offsets.data(), xy.data(), n_frames, paths.data(), assign_.data(), phase_.data(), &stats_);
    const std::string msg = st != TRI_OK ? tri_last_error() : "";
    for (size_t g = 1; g < engines.size(); g++) tri_destroy(engines[g]);
    if (st != TRI_OK) throw std::runtime_error("tri_b200: " + msg);
  }
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
