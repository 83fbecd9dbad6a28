#include "DetectionsContainer.h"

#include "CsvIngest.h"

#include <algorithm>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <stdexcept>

DetectionsContainer::DetectionsContainer(const char* path, int offset, int recordSize, int startFrame, int endFrame) {
  readFiles(getFiles(path), offset, recordSize, startFrame, endFrame);
}

DetectionsContainer::DetectionsContainer(int camCount, bool useOffset) : n_cameras(camCount), data((size_t)camCount) {
  if (useOffset) offsetVector.resize((size_t)camCount);
}

std::vector<std::string> DetectionsContainer::getFiles(const char* path) {
  // every *.csv below `path`, camera order = sorted path order (DetectionsContainer.cpp:78-92)
  std::vector<std::string> files;
  for (const auto& entry : std::filesystem::recursive_directory_iterator(path)) {
    const std::string name = entry.path().u8string();
    if (name.find(".csv") != std::string::npos && name.find(".avi") == std::string::npos) files.push_back(name);
  }
  std::sort(files.begin(), files.end());
  return files;
}

void DetectionsContainer::readFiles(const std::vector<std::string>& files, int offset, int recordSize, int startFrame,
                                    int endFrame) {
  // Row = frame,(x,y,w,h,cx,cy,conf) x k.  Every token goes through stoi (so 0.97 -> 0); the detection is
  // the box centre, fields 5 and 6 of the record; missing frame numbers become empty frames
  // (DetectionsContainer.cpp:19-76).  The parsing itself is CsvIngest.cpp (mmap, one thread per camera).
  ingestDetectionFiles(files, offset, recordSize, startFrame, endFrame, data);
  if (data.size() < 2) throw std::runtime_error("There must be at least 2 cameras");
  for (size_t i = 1; i < data.size(); i++)
    if (data[i - 1].size() != data[i].size()) throw std::runtime_error("Number of frames on all cameras must be the same");
  n_frames = (int)data[0].size();
  n_cameras = (int)data.size();
}

std::vector<Detections> DetectionsContainer::getFrame(int i) const {
  std::vector<Detections> frame;
  for (const Cameras& cam : data) frame.push_back(cam[i]);
  return frame;
}

std::vector<int> DetectionsContainer::getDetectionsCount(int frame) const {
  std::vector<int> n((size_t)n_cameras);
  for (int i = 0; i < n_cameras; i++) n[i] = (int)data[i][frame].size() + 1;  // + "no detection"
  return n;
}

void DetectionsContainer::addEmptyFrame() {  // :121-130
  for (size_t i = 0; i < data.size(); i++) {
    data[i].emplace_back();
    if (!offsetVector.empty()) {
      offsetVector[i].emplace_back();
      offsetVector[i].back()[0] = 0;  // index 0 = "no detection" maps to itself
    }
  }
  n_frames = (int)data[0].size();
}

void DetectionsContainer::addDetectionToCamera(cv::Point2d det, int cam, int originalIndex) {  // :132-139
  data[cam].back().push_back(det);
  if (originalIndex > -1 && !offsetVector.empty()) offsetVector[cam].back()[(int)data[cam].back().size()] = originalIndex + 1;
}

std::vector<int> DetectionsContainer::getOriginalCombination(const std::vector<int>& combination, int frame) const {
  if (offsetVector.empty()) return combination;
  std::vector<int> original;
  for (size_t i = 0; i < combination.size(); i++) original.push_back(offsetVector[i][frame].at(combination[i]));
  return original;
}

std::vector<std::vector<cv::Point2d>> DetectionsContainer::getDataForTriangulation() {
  std::vector<std::vector<cv::Point2d>> result((size_t)n_cameras);
  for (int frame = 0; frame < n_frames; frame++)
    for (int cam = 0; cam < n_cameras; cam++) {
      const size_t size = data[cam][frame].size();
      if (size > 1)
        throw std::runtime_error(
            "Function 'getDataForTriangulation' can be used only for one drone. Each frame can have max one detection");
      result[cam].push_back(size == 1 ? data[cam][frame][0] : cv::Point2d(-1, -1));
    }
  return result;
}

void DetectionsContainer::toCSR(std::vector<int32_t>& offsets, std::vector<double>& xy) const {
  offsets.assign((size_t)n_cameras * (n_frames + 1), 0);
  xy.clear();
  int32_t at = 0;
  for (int cam = 0; cam < n_cameras; cam++) {
    for (int frame = 0; frame < n_frames; frame++) {
      offsets[(size_t)cam * (n_frames + 1) + frame] = at;
      for (const cv::Point2d& p : data[cam][frame]) {
        xy.push_back(p.x);
        xy.push_back(p.y);
        at++;
      }
    }
    offsets[(size_t)cam * (n_frames + 1) + n_frames] = at;
  }
}
