// oracle/shim: the tap pointers for binaries that do not install taps (oracle/_ref/ref_main = the reference's main.cpp)
#include <opencv2/opencv.hpp>
namespace cv { namespace tap { void (*on_norm3)(uint64_t, uint64_t, double) = nullptr; } }
