// oracle/shim: <opencv2/calib3d/calib3d.hpp> as included by the reference's Camera.h -- see ../opencv.hpp
#include "../opencv.hpp"
