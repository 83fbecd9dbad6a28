// selftest.cpp -- CPU-only checks of the host mirror (no engine calls): prints one line per check,
// driven by tests/test_host_cpp.py.  usage: host_selftest <cameras.xml> <csv dir>
#include <cstdio>

#include "DetectionsContainer.h"
#include "Triangulator.h"
#include "utils.h"

int main(int argc, const char** argv) {
  if (argc < 3) return 2;
  std::vector<const tdr::Camera*> cams = loadCamerasXML(argv[1]);
  std::printf("cameras %zu\n", cams.size());
  for (const tdr::Camera* c : cams) {
    std::printf("cam %d %d %d %.17g %.17g", c->id, c->width, c->height, c->fovy, c->fx);
    for (int i = 0; i < 12; i++) std::printf(" %.17g", c->cameraPerspectiveMatrix[i]);
    std::printf("\n");
  }
  DetectionsContainer box(argv[2], 0, 7);
  std::printf("container %d %d\n", box.getCamCount(), box.getFrameCount());
  for (int cam = 0; cam < box.getCamCount(); cam++)
    for (int f = 0; f < box.getFrameCount(); f++) {
      std::printf("det %d %d %d", cam, f, box.detCountForCam(cam, f));
      for (int d = 0; d < box.detCountForCam(cam, f); d++) std::printf(" %g %g", box.getRecord(cam, f, d).x, box.getRecord(cam, f, d).y);
      std::printf("\n");
    }
  std::vector<int32_t> offs;
  std::vector<double> xy;
  box.toCSR(offs, xy);
  std::printf("csr %zu %zu %d\n", offs.size(), xy.size() / 2, offs.back());
  // the pixel format the batch adapter packs (no engine call): ushort2 for integer detections incl. the sentinel,
  // float2 off the integer grid or out of the 16-bit range, double2 when a value is not a float
  using Pts = std::vector<std::vector<cv::Point2d>>;
  const Pts ints = {{{12, 7}, {-1, -1}, {1919, 1079}}, {{0, 0}, {-1, 5}, {65534, 3}}};
  Pts big = ints, frac = ints, neg = ints, dbl = ints;
  big[1][2].x = 65535; frac[0][0].x = 12.25; neg[0][0].y = -2; dbl[0][0].x = 12.000000001;
  std::printf("pixfmt %u %u %u %u %u %u\n", Triangulator::pixelFormatFor(ints), Triangulator::pixelFormatFor(big),
              Triangulator::pixelFormatFor(frac), Triangulator::pixelFormatFor(neg), Triangulator::pixelFormatFor(dbl),
              Triangulator::pixelFormatFor(Pts()));
  for (const auto& cam : cams) delete cam;
  return 0;
}
