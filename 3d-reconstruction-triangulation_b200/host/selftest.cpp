// selftest.cpp -- CPU-only checks of the host mirror (no engine calls): prints one line per check,
// driven by tests/test_host_cpp.py.  usage: host_selftest <cameras.xml> <csv dir>
#include <cstdio>

#include "DetectionsContainer.h"
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
  for (const auto& cam : cams) delete cam;
  return 0;
}
