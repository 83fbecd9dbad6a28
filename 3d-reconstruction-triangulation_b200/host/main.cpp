// main.cpp -- the reference's CLI (src/main.cpp:15-88) on the B200 engine: same positional arguments,
// same --n_drones / --triangulator flags and defaults, same ./results/drone{i}.ply outputs and the same
// "Execution time" line.  Extra, optional: --device N, --fast-ray (analytic LM instead of the
// trajectory-exact cv::LMSolver emulation), --dump FILE (full-precision paths + assignments, binary).
#include <chrono>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>

#include "DetectionsContainer.h"
#include "DroneClassifier.h"
#include "Triangulator.h"
#include "utils.h"

static std::string OUTPUT_DIR = "./results/";

static void usage() {
  std::cerr << "Usage: 3D-Reconstruction-Triangulation [--n_drones N] [--triangulator matrix|ray] [--device N] [--fast-ray] "
               "[--dump FILE] cameras_path data_path\n";
}

int main(int argc, const char** argv) {
  std::string cameras_path, data_path, kind = "matrix", dump;
  int n_drones = 1, device = 0;
  bool fast_ray = false;
  int positional = 0;
  try {
    for (int i = 1; i < argc; i++) {
      const std::string a = argv[i];
      auto value = [&](const char* name) -> std::string {
        if (i + 1 >= argc) throw std::runtime_error(std::string(name) + ": 1 argument(s) expected. 0 provided.");
        return argv[++i];
      };
      if (a == "--n_drones") n_drones = std::stoi(value("--n_drones"));
      else if (a == "--triangulator") kind = value("--triangulator");
      else if (a == "--device") device = std::stoi(value("--device"));
      else if (a == "--dump") dump = value("--dump");
      else if (a == "--fast-ray") fast_ray = true;
      else if (a == "-h" || a == "--help") { usage(); return 0; }
      else if (positional == 0) { cameras_path = a; positional++; }
      else if (positional == 1) { data_path = a; positional++; }
      else throw std::runtime_error("Maximum number of positional arguments exceeded");
    }
    if (positional < 2) throw std::runtime_error(positional == 0 ? "cameras_path: 1 argument(s) expected. 0 provided." : "data_path: 1 argument(s) expected. 0 provided.");
  } catch (const std::exception& err) {
    std::cerr << err.what() << std::endl;
    usage();
    std::exit(1);
  }

  if (std::filesystem::exists(OUTPUT_DIR)) std::filesystem::remove_all(OUTPUT_DIR);
  std::filesystem::create_directory(OUTPUT_DIR);

  std::vector<const tdr::Camera*> cameras = loadCamerasXML(cameras_path.c_str());

  Triangulator* triangulator;
  if (kind == "matrix") {
    triangulator = new MatrixTriangulator(cameras, device);
  } else if (kind == "ray") {
    triangulator = new RayTriangulator(cameras, device, !fast_ray);
  } else {
    throw std::runtime_error("Invalid --triangulator argument. Allowed options are 'matrix' and 'ray'");
  }

  DroneClassifier classifier(triangulator, (size_t)n_drones);
  DetectionsContainer container(data_path.c_str(), 0, 7);

  auto start = std::chrono::high_resolution_clock::now();
  std::vector<std::vector<cv::Point3d>> triangulatedPoints;
  classifier.classifyDrones(container, triangulatedPoints);
  auto stop = std::chrono::high_resolution_clock::now();
  auto time = std::chrono::duration_cast<std::chrono::microseconds>(stop - start);
  std::cout << "Execution time: " << time.count() * 1e-6 << "s" << std::endl;

  for (int i = 1; i <= n_drones; i++) {
    std::string name = OUTPUT_DIR + "drone" + std::to_string(i) + ".ply";
    writeOutputFile(name.c_str(), triangulatedPoints[i - 1]);
  }
  if (!dump.empty()) {  // int32 n_drones, n_frames, n_cams; double paths[d][f][3]; int8 assign[d][f][c]; uint8 phase[d][f]
    std::ofstream out(dump, std::ios::binary);
    const int32_t hdr[3] = {n_drones, container.getFrameCount(), container.getCamCount()};
    out.write((const char*)hdr, sizeof(hdr));
    for (const auto& path : triangulatedPoints)
      for (const cv::Point3d& p : path) { const double v[3] = {p.x, p.y, p.z}; out.write((const char*)v, sizeof(v)); }
    out.write((const char*)classifier.assignments().data(), (std::streamsize)classifier.assignments().size());
    out.write((const char*)classifier.phases().data(), (std::streamsize)classifier.phases().size());
  }

  delete triangulator;
  for (const auto& cam : cameras) delete cam;
}
