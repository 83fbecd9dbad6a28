// Camera.h -- tdr::Camera reduced to what the triangulation path reads, without OpenCV.
// Mirrors the arithmetic of the reference's src/Camera.h:78-187 and :274-287 (including the truncated
// 57.2958 of :91, the integer principal point of :81-82 and the inverse -- not transpose -- of :134),
// so the projection matrices are the ones the reference would hand to its triangulators.
#pragma once
#include <cmath>
#include <stdexcept>

#include "../../include/tri_b200.h"

namespace tdr {

class Camera {
 public:
  int id = 0;
  int width = 0, height = 0, cx = 0, cy = 0;
  double fx = 0, fy = 0, fovx = 0, fovy = 0;
  double tvec[3] = {0, 0, 0};   // world position (the reference's naming, src/utils.cpp:101)
  double rquat[4] = {1, 0, 0, 0};  // (w,i,j,k) as stored
  double camPos[3] = {0, 0, 0};    // -R * tvec (src/Camera.h:167-170)
  double cameraMatrix[9] = {0};
  double cameraExtrinsicMatrix[12] = {0};
  double cameraPerspectiveMatrix[12] = {0};

  explicit Camera(int id_ = 0) : id(id_) {}

  void compCamParams() {
    constexpr double RAD_TO_DEG = 57.29577951308232087679, DEG_TO_RAD = 0.01745329251994329576;
    if (width == 0 || height == 0) throw std::runtime_error("Camera: width and height must be set");
    cx = (int)std::round(width / 2.0);
    cy = (int)std::round(height / 2.0);
    if (fx == 0) throw std::runtime_error("Camera: focal length must be set");
    fovx = 2 * std::atan(width / (2 * fx)) * 57.2958;
    fovy = 2.0 * std::atan(std::tan(fovx * 0.5 * DEG_TO_RAD) / ((double)width / (double)height)) * RAD_TO_DEG;
    fx = (width / 2.0) / (std::tan((fovx / 2.0) * DEG_TO_RAD));
    fy = (height / 2.0) / (std::tan((fovy / 2.0) * DEG_TO_RAD));
    const double a = rquat[0], b = rquat[1], c = rquat[2], d = rquat[3];
    const double S[9] = {1 - 2 * (c * c + d * d), 2 * (b * c - a * d),     2 * (b * d + a * c),
                         2 * (b * c + a * d),     1 - 2 * (b * b + d * d), 2 * (c * d - a * b),
                         2 * (b * d - a * c),     2 * (c * d + a * b),     1 - 2 * (b * b + c * c)};
    double R[9];
    invert3(S, R);
    for (int i = 0; i < 3; i++) camPos[i] = -(R[i * 3] * tvec[0] + R[i * 3 + 1] * tvec[1] + R[i * 3 + 2] * tvec[2]);
    if (fx == 0 || fy == 0 || cx == 0 || cy == 0) throw std::runtime_error("Camera: intrinsics must be set");
    const double K[9] = {fx, 0, (double)cx, 0, fy, (double)cy, 0, 0, 1};
    for (int i = 0; i < 9; i++) cameraMatrix[i] = K[i];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) cameraExtrinsicMatrix[i * 4 + j] = R[i * 3 + j];
      cameraExtrinsicMatrix[i * 4 + 3] = camPos[i];
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 4; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += K[i * 3 + k] * cameraExtrinsicMatrix[k * 4 + j];
        cameraPerspectiveMatrix[i * 4 + j] = s;
      }
  }

  tri_camera describe() const {  // what the engine reads of this camera (include/tri_b200.h)
    tri_camera t{};
    t.width = width;
    t.height = height;
    t.fovy_deg = fovy;
    for (int i = 0; i < 12; i++) t.P[i] = cameraPerspectiveMatrix[i];
    for (int i = 0; i < 3; i++) t.position[i] = tvec[i];
    for (int i = 0; i < 4; i++) t.quat[i] = rquat[i];
    return t;
  }

 private:
  static void invert3(const double* S, double* T) {  // closed-form 3x3 inverse (cv::Mat::inv on 3x3)
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.) {
      for (int i = 0; i < 9; i++) T[i] = 0;
      return;
    }
    d = 1. / d;
    T[0] = (S[4] * S[8] - S[5] * S[7]) * d; T[1] = (S[2] * S[7] - S[1] * S[8]) * d; T[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    T[3] = (S[5] * S[6] - S[3] * S[8]) * d; T[4] = (S[0] * S[8] - S[2] * S[6]) * d; T[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    T[6] = (S[3] * S[7] - S[4] * S[6]) * d; T[7] = (S[1] * S[6] - S[0] * S[7]) * d; T[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  }
};

}  // namespace tdr
