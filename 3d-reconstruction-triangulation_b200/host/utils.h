// utils.h -- camera XML loader and PLY writer of the reference (src/utils.cpp:46-131), without pugixml.
#pragma once
#include <vector>

#include "Camera.h"
#include "cv_compat.h"

std::vector<const tdr::Camera*> loadCamerasXML(const char* path);
const tdr::Camera* createCamera(int id, size_t width, size_t height, double focalLength, const double position[3],
                                const double quat[4]);
void writeOutputFile(const char* path, const std::vector<cv::Point3d>& triangulatedPoints);
