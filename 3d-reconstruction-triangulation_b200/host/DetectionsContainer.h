// DetectionsContainer.h -- the reference's detection store (src/DetectionsContainer.h) with the same
// CSV semantics (src/DetectionsContainer.cpp:19-92), kept on the host; the kernels consume it through
// toCSR() (classifier) and getDataForTriangulation() (batch API).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "cv_compat.h"

typedef std::vector<cv::Point2d> Detections;
typedef std::vector<Detections> Cameras;

class DetectionsContainer {
 public:
  int n_frames = 0, n_cameras = 0;
  std::vector<Cameras> data;  // [camera][frame][detection]
  // optional: original detection index of each local one, [camera][frame] -> {local index (1-based, 0 = none) -> original}
  std::vector<std::vector<std::map<int, int>>> offsetVector;

  DetectionsContainer(const char* path, int offset, int recordSize, int startFrame = 0, int endFrame = 0);
  explicit DetectionsContainer(int camCount, bool useOffset = false);

  static std::vector<std::string> getFiles(const char* path);
  void readFiles(const std::vector<std::string>& files, int offset, int recordSize, int startFrame, int endFrame);

  std::vector<Detections> getFrame(int i) const;
  int getFrameCount() const { return n_frames; }
  int getCamCount() const { return n_cameras; }
  std::vector<int> getDetectionsCount(int frame) const;
  int detCountForCam(int cam, int frame) const { return (int)data[cam][frame].size(); }
  cv::Point2d getRecord(int camera, int frame, int detection) const { return data[camera][frame][detection]; }
  void addEmptyFrame();
  void addDetectionToCamera(cv::Point2d det, int cam, int originalIndex = -1);
  // local combination indices -> indices of the container the detections were taken from (:173-186)
  std::vector<int> getOriginalCombination(const std::vector<int>& combination, int frame) const;
  std::vector<std::vector<cv::Point2d>> getDataForTriangulation();

  // CSR form for tri_classify: offsets[cam * (n_frames + 1) + frame], xy pairs in [cam][frame][det] order
  void toCSR(std::vector<int32_t>& offsets, std::vector<double>& xy) const;
};
