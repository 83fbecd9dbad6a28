// main.cpp -- the reference's CLI (src/main.cpp:15-88) on the B200 engine: same positional arguments,
// same --n_drones / --triangulator flags and defaults, same ./results/drone{i}.ply outputs, the same per-frame
// progress lines and "Execution time" line.  Flags follow the forms the reference's argument parser (p-ranav
// argparse, src/argparse.hpp) accepts: "--flag value" and "--flag=value" anywhere among the positionals, -h/--help and
// -v/--version exit 0, anything else unknown is an error with the usage text and exit code 1.
// Extra, optional: --device N, --fast-ray (closed-form ray solves instead of the trajectory-exact cv::LMSolver
// emulation), --quiet (no per-frame lines), --dump FILE (full-precision paths + assignments, binary).
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
  std::cerr << "Usage: 3D-Reconstruction-Triangulation [--help] [--version] [--n_drones VAR] [--triangulator VAR] [--device VAR] [--fast-ray] "
               "[--quiet] [--dump VAR] cameras_path data_path\n\n"
               "Positional arguments:\n  cameras_path  \tPath to the file with camera data \n  data_path     \tPath to the folder woth detections on each camera \n\n"
               "Optional arguments:\n  -h, --help    \tshows help message and exits \n  -v, --version \tprints version information and exits \n"
               "  --n_drones    \tNumber of drones in the scene [default: 1]\n  --triangulator\tWhich triangulator should be used [default: \"matrix\"]\n";
}

int main(int argc, const char** argv) {
  std::string cameras_path, data_path, kind = "matrix", dump;
  int n_drones = 1, device = 0;
  bool fast_ray = false, quiet = false;
  int positional = 0;
  try {
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i], inline_value;
      bool has_inline = false;
      if (a.rfind("--", 0) == 0) {  // "--flag=value"
        const size_t eq = a.find('=');
        if (eq != std::string::npos) { inline_value = a.substr(eq + 1); a = a.substr(0, eq); has_inline = true; }
      }
      auto value = [&](const char* name) -> std::string {
        if (has_inline) return inline_value;
        if (i + 1 >= argc) throw std::runtime_error(std::string(name) + ": 1 argument(s) expected. 0 provided.");
        return argv[++i];
      };
      auto integer = [&](const char* name) {  // scan<'i', int>: the whole token must be a decimal integer
        const std::string v = value(name);
        size_t used = 0;
        int r = 0;
        try { r = std::stoi(v, &used); } catch (const std::exception&) { used = 0; }
        if (used != v.size() || v.empty()) throw std::runtime_error("pattern '" + v + "' not found");
        return r;
      };
      if (a == "--n_drones") n_drones = integer("--n_drones");
      else if (a == "--triangulator") kind = value("--triangulator");
      else if (a == "--device") device = integer("--device");
      else if (a == "--dump") dump = value("--dump");
      else if (a == "--fast-ray") fast_ray = true;
      else if (a == "--quiet") quiet = true;
      else if (a == "-h" || a == "--help") { usage(); return 0; }
      else if (a == "-v" || a == "--version") { std::cout << "1.0" << std::endl; return 0; }
      else if (a.size() > 1 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) throw std::runtime_error("Unknown argument: " + a);
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
  classifier.setProgress(!quiet);
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
