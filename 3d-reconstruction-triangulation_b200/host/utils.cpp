#include "utils.h"

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace {
// value of attribute `name` inside the tag text `tag` ("" if absent)
std::string attribute(const std::string& tag, const std::string& name) {
  size_t at = 0;
  while ((at = tag.find(name, at)) != std::string::npos) {
    const bool starts = at == 0 || tag[at - 1] == ' ' || tag[at - 1] == '\t' || tag[at - 1] == '\n' || tag[at - 1] == '<';
    size_t eq = at + name.size();
    while (eq < tag.size() && (tag[eq] == ' ' || tag[eq] == '\t')) eq++;
    if (starts && eq < tag.size() && tag[eq] == '=') {
      size_t q = tag.find_first_of("\"'", eq);
      if (q == std::string::npos) return "";
      const size_t end = tag.find(tag[q], q + 1);
      if (end == std::string::npos) return "";
      return tag.substr(q + 1, end - q - 1);
    }
    at += name.size();
  }
  return "";
}
}  // namespace

const tdr::Camera* createCamera(int id, size_t width, size_t height, double focalLength, const double position[3],
                                const double quat[4]) {  // src/utils.cpp:94-107
  tdr::Camera* cam = new tdr::Camera(id);
  cam->width = (int)width;
  cam->height = (int)height;
  cam->fx = focalLength;
  for (int i = 0; i < 3; i++) cam->tvec[i] = position[i];
  for (int i = 0; i < 4; i++) cam->rquat[i] = quat[i];
  cam->compCamParams();
  return cam;
}

std::vector<const tdr::Camera*> loadCamerasXML(const char* path) {
  // Every <Camera> under <Cameras> that has a <ControlFrames><ControlFrame .../>, in document order;
  // width/height = 2 * the integer PRINCIPAL_POINT (src/utils.cpp:46-92).
  std::ifstream in(path);
  if (!in) throw std::runtime_error("Cannot open camera XML file");
  std::stringstream buf;
  buf << in.rdbuf();
  const std::string doc = buf.str();
  std::vector<const tdr::Camera*> cameras;
  size_t at = 0;
  while ((at = doc.find("<Camera", at)) != std::string::npos) {
    const char next = at + 7 < doc.size() ? doc[at + 7] : '\0';
    if (next != ' ' && next != '>' && next != '\t' && next != '\n' && next != '\r') { at += 7; continue; }  // <Cameras>
    const size_t head_end = doc.find('>', at);
    if (head_end == std::string::npos) break;
    const std::string head = doc.substr(at, head_end - at);
    const bool self_closed = head_end > 0 && doc[head_end - 1] == '/';
    size_t body_end = self_closed ? head_end : doc.find("</Camera>", head_end);
    if (body_end == std::string::npos) body_end = doc.size();
    const std::string body = doc.substr(head_end, body_end - head_end);
    at = body_end;
    const size_t cf = body.find("<ControlFrame ");
    if (cf == std::string::npos) continue;
    const std::string frame = body.substr(cf, body.find('>', cf) - cf);
    const int id = std::atoi(attribute(head, "DEVICEID").c_str());
    const double focal = std::strtod(attribute(frame, "FOCAL_LENGTH").c_str(), nullptr);
    int width = 0, height = 0;
    std::stringstream pp(attribute(frame, "PRINCIPAL_POINT"));
    pp >> width >> height;
    width *= 2;
    height *= 2;
    double pos[3] = {0, 0, 0}, q[4] = {0, 0, 0, 0};
    std::stringstream ps(attribute(frame, "POSITION"));
    ps >> pos[0] >> pos[1] >> pos[2];
    std::stringstream qs(attribute(frame, "ORIENTATION"));
    qs >> q[0] >> q[1] >> q[2] >> q[3];
    cameras.push_back(createCamera(id, (size_t)width, (size_t)height, focal, pos, q));
  }
  return cameras;
}

void writeOutputFile(const char* path, const std::vector<cv::Point3d>& triangulatedPoints) {  // src/utils.cpp:109-131
  std::ofstream out(path);
  out << "ply\n";
  out << "format ascii 1.0\n";
  out << "element vertex " << (int)triangulatedPoints.size() << "\n";
  out << "property float x\n";
  out << "property float y\n";
  out << "property float z\n";
  out << "element face " << 0 << "\n";
  out << "property list uchar int vertex_index\n";
  out << "end_header\n";
  for (const cv::Point3d& p : triangulatedPoints) out << p.x << " " << p.y << " " << p.z << "\n";
}
