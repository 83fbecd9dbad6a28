#include "utils.h"

#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace {
// The document with everything a tag scanner must not look into blanked out (same length, so offsets stay valid):
// comments <!-- -->, CDATA sections, processing instructions <? ?> and the DOCTYPE declaration.  pugixml, which the
// reference uses (src/utils.cpp:46-49), skips the same constructs.
std::string strip_non_markup(std::string doc) {
  auto blank = [&](const char* open, const char* close) {
    size_t at = 0;
    while ((at = doc.find(open, at)) != std::string::npos) {
      size_t end = doc.find(close, at + strlen(open));
      end = end == std::string::npos ? doc.size() : end + strlen(close);
      for (size_t i = at; i < end; i++) if (doc[i] != '\n') doc[i] = ' ';
      at = end;
    }
  };
  blank("<!--", "-->");
  blank("<![CDATA[", "]]>");
  blank("<?", "?>");
  blank("<!DOCTYPE", ">");
  return doc;
}
// the five predefined entities and numeric character references of an attribute value
std::string unescape(const std::string& v) {
  std::string out;
  for (size_t i = 0; i < v.size(); i++) {
    if (v[i] != '&') { out += v[i]; continue; }
    const size_t semi = v.find(';', i);
    if (semi == std::string::npos) { out += v[i]; continue; }
    const std::string e = v.substr(i + 1, semi - i - 1);
    if (e == "amp") out += '&'; else if (e == "lt") out += '<'; else if (e == "gt") out += '>';
    else if (e == "quot") out += '"'; else if (e == "apos") out += '\'';
    else if (e.size() > 1 && e[0] == '#') out += (char)(e[1] == 'x' ? std::strtol(e.c_str() + 2, nullptr, 16) : std::strtol(e.c_str() + 1, nullptr, 10));
    else { out += v[i]; continue; }
    i = semi;
  }
  return out;
}
// end of the tag that starts at `at` ('>' outside quoted attribute values), npos if unterminated
size_t tag_end(const std::string& doc, size_t at) {
  char quote = 0;
  for (size_t i = at; i < doc.size(); i++) {
    if (quote) { if (doc[i] == quote) quote = 0; }
    else if (doc[i] == '"' || doc[i] == '\'') quote = doc[i];
    else if (doc[i] == '>') return i;
  }
  return std::string::npos;
}
// value of attribute `name` inside the tag text `tag` ("" if absent); either quote character
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
      return unescape(tag.substr(q + 1, end - q - 1));
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
  const std::string doc = strip_non_markup(buf.str());
  if (doc.find("<Cameras") == std::string::npos && doc.find('<') == std::string::npos) throw std::runtime_error("Cannot open camera XML file");
  std::vector<const tdr::Camera*> cameras;
  size_t at = 0;
  while ((at = doc.find("<Camera", at)) != std::string::npos) {
    const char next = at + 7 < doc.size() ? doc[at + 7] : '\0';
    if (next != ' ' && next != '>' && next != '\t' && next != '\n' && next != '\r') { at += 7; continue; }  // <Cameras>
    const size_t head_end = tag_end(doc, at);
    if (head_end == std::string::npos) break;
    const std::string head = doc.substr(at, head_end - at);
    const bool self_closed = head_end > 0 && doc[head_end - 1] == '/';
    size_t body_end = self_closed ? head_end : doc.find("</Camera>", head_end);
    if (body_end == std::string::npos) body_end = doc.size();
    const std::string body = doc.substr(head_end, body_end - head_end);
    at = body_end;
    size_t cf = 0;  // the first <ControlFrame ...> (not <ControlFrames>)
    while ((cf = body.find("<ControlFrame", cf)) != std::string::npos) {
      const char c2 = cf + 13 < body.size() ? body[cf + 13] : '\0';
      if (c2 == ' ' || c2 == '\t' || c2 == '\n' || c2 == '\r' || c2 == '/' || c2 == '>') break;
      cf += 13;
    }
    if (cf == std::string::npos) continue;
    const size_t cf_end = tag_end(body, cf);
    const std::string frame = body.substr(cf, (cf_end == std::string::npos ? body.size() : cf_end) - cf);
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
