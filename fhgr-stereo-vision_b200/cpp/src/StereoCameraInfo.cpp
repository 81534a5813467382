// StereoCameraInfo.cpp — dependency-free reader / writer of the reference's calibration YAML; see the header.
#include <stereo-vision-gpu/StereoCameraInfo.hpp>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace stereo_vision {
  namespace {
    // ---- writer: the layout of cv::FileStorage (three spaces per level, flow sequences wrapped at ~80 columns)
    std::string num(double v) {
      char buf[40];
      std::snprintf(buf, sizeof buf, "%.17g", v);   // cv::FileStorage prints doubles with 17 significant digits
      std::string s(buf);
      if (s.find_first_of(".enai") == std::string::npos) s += ".";  // "1." like cv::FileStorage
      return s;
    }
    template <std::size_t N>
    void matrix(std::ostringstream& o, char const* key, std::array<double, N> const& a, int rows, int cols) {
      o << "   " << key << ": !!opencv-matrix\n      rows: " << rows << "\n      cols: " << cols << "\n      dt: d\n";
      std::string line = "      data: [ ";
      bool has = false;  // the current line holds at least one value
      for (std::size_t i = 0; i < N; i++) {
        std::string const t = num(a[i]) + (i + 1 < N ? "," : "");
        if (has && line.size() + t.size() > 72) {   // cv::FileStorage wraps flow sequences at its 72-column margin
          while (!line.empty() && line.back() == ' ') line.pop_back();
          o << line << "\n";
          line = "          ";
          has = false;
        }
        line += t + " ";
        has = true;
      }
      o << line << "]\n";
    }

    // ---- reader: a line-oriented parser of the subset of YAML that cv::FileStorage emits
    struct Node {
      std::string scalar;                 // plain scalar or the text of a flow sequence "[ ... ]" (joined over lines)
      std::map<std::string, Node> map;    // nested mapping
      bool is_map = false;
    };
    int indent_of(std::string const& l) {
      int n = 0;
      while (n < static_cast<int>(l.size()) && l[n] == ' ') n++;
      return n;
    }
    std::string trim(std::string s) {
      while (!s.empty() && std::isspace(static_cast<unsigned char>(s.back()))) s.pop_back();
      std::size_t b = 0;
      while (b < s.size() && std::isspace(static_cast<unsigned char>(s[b]))) b++;
      return s.substr(b);
    }
    // parses the mapping whose entries are indented deeper than `parent_indent`, starting at line i
    void parse_map(std::vector<std::string> const& lines, std::size_t& i, int parent_indent, Node& out) {
      out.is_map = true;
      int my_indent = -1;
      while (i < lines.size()) {
        std::string const& raw = lines[i];
        std::string const t = trim(raw);
        if (t.empty() || t[0] == '#' || t[0] == '%' || t == "---" || t == "...") { i++; continue; }
        int const ind = indent_of(raw);
        if (ind <= parent_indent) return;
        if (my_indent < 0) my_indent = ind;
        if (ind < my_indent) return;
        auto const colon = t.find(':');
        if (colon == std::string::npos) throw std::runtime_error("calibration file: expected 'key: value' in line '" + t + "'");
        std::string key = trim(t.substr(0, colon));
        if (key.size() >= 2 && key.front() == '"' && key.back() == '"') key = key.substr(1, key.size() - 2);
        std::string val = trim(t.substr(colon + 1));
        i++;
        Node child;
        if (val.empty() || val.rfind("!!", 0) == 0) {  // nested mapping (optionally tagged, e.g. !!opencv-matrix)
          parse_map(lines, i, ind, child);
        } else if (val[0] == '[') {                      // flow sequence, may continue on the following lines
          while (val.find(']') == std::string::npos && i < lines.size()) val += " " + trim(lines[i++]);
          child.scalar = val;
        } else {
          child.scalar = val;
        }
        out.map[key] = child;
      }
    }
    std::vector<double> numbers(std::string const& seq) {
      std::vector<double> v;
      char const* p = seq.c_str();
      while (*p) {
        if (std::isdigit(static_cast<unsigned char>(*p)) || *p == '-' || *p == '+' || *p == '.') {
          char* e = nullptr;
          double const d = std::strtod(p, &e);
          if (e == p) { p++; continue; }
          v.push_back(d);
          p = e;
        } else {
          p++;
        }
      }
      return v;
    }
    Node const& child(Node const& n, char const* key) {
      auto const it = n.map.find(key);
      if (it == n.map.end()) throw std::runtime_error(std::string("calibration file: missing '") + key + "'");
      return it->second;
    }
    template <std::size_t N>
    void get_matrix(Node const& n, char const* key, std::array<double, N>& out) {
      Node const& m = child(n, key);
      if (!m.is_map) throw std::runtime_error(std::string("calibration file: '") + key + "' is not an !!opencv-matrix");
      auto const rows = numbers(child(m, "rows").scalar), cols = numbers(child(m, "cols").scalar);
      auto const data = numbers(child(m, "data").scalar);
      if (rows.size() != 1 || cols.size() != 1 || static_cast<std::size_t>(rows[0] * cols[0]) != data.size() || data.size() < N)
        throw std::runtime_error(std::string("calibration file: '") + key + "' has an unexpected size");
      for (std::size_t i = 0; i < N; i++) out[i] = data[i];   // dist_coeffs: the first five (k1 k2 p1 p2 k3) of up to 14
    }
  }  // namespace

  auto write(StereoCameraInfo const& x, bool document) -> std::string {
    std::ostringstream o;
    if (document) o << "%YAML:1.0\n---\n";
    o << "calibration:\n   file_version: " << static_cast<int>(kStereoCameraInfoVersion) << "\n   image_size: [ " << x.width << ", " << x.height << " ]\n";
    matrix(o, "camera_matrix_left", x.camera_matrix_left, 3, 3);
    matrix(o, "camera_matrix_right", x.camera_matrix_right, 3, 3);
    matrix(o, "dist_coeffs_left", x.dist_coeffs_left, 1, 5);
    matrix(o, "dist_coeffs_right", x.dist_coeffs_right, 1, 5);
    matrix(o, "R", x.R, 3, 3);
    matrix(o, "T", x.T, 3, 1);
    matrix(o, "E", x.E, 3, 3);
    matrix(o, "F", x.F, 3, 3);
    o << "   reprojection_error: " << num(x.reprojection_error) << "\n";
    return o.str();
  }

  auto read(std::string const& yaml_text) -> StereoCameraInfo {
    std::vector<std::string> lines;
    {
      std::istringstream in(yaml_text);
      for (std::string l; std::getline(in, l);) {
        if (!l.empty() && l.back() == '\r') l.pop_back();
        lines.push_back(l);
      }
    }
    Node root;
    std::size_t i = 0;
    parse_map(lines, i, -1, root);
    auto const it = root.map.find("calibration");
    if (it == root.map.end() || !it->second.is_map || it->second.map.empty())
      throw std::runtime_error("missing camera info in config file!");                       // StereoCameraInfo.hpp:68-70
    Node const& n = it->second;
    auto const ver = numbers(child(n, "file_version").scalar);
    int const version = ver.empty() ? -1 : static_cast<int>(ver[0]);
    if (version != kStereoCameraInfoVersion)                                                 // StereoCameraInfo.hpp:72-76
      throw std::runtime_error("invalid config file version! Expected " + std::to_string(kStereoCameraInfoVersion) + " but found " +
                               std::to_string(version) +
                               "! Please re-generate the calibration file using the matching calibration software version.");
    StereoCameraInfo x;
    auto const size = numbers(child(n, "image_size").scalar);
    if (size.size() != 2) throw std::runtime_error("calibration file: image_size must be [ width, height ]");
    x.width = static_cast<int>(size[0]);
    x.height = static_cast<int>(size[1]);
    get_matrix(n, "camera_matrix_left", x.camera_matrix_left);
    get_matrix(n, "camera_matrix_right", x.camera_matrix_right);
    get_matrix(n, "dist_coeffs_left", x.dist_coeffs_left);
    get_matrix(n, "dist_coeffs_right", x.dist_coeffs_right);
    get_matrix(n, "R", x.R);
    get_matrix(n, "T", x.T);
    get_matrix(n, "E", x.E);
    get_matrix(n, "F", x.F);
    auto const err = numbers(child(n, "reprojection_error").scalar);
    x.reprojection_error = err.empty() ? 0.0 : err[0];
    return x;
  }

  auto LoadStereoCameraInfo(std::filesystem::path const& path) -> StereoCameraInfo {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("missing camera info in config file!");  // cv::FileStorage yields an empty node for a missing file
    std::stringstream ss;
    ss << f.rdbuf();
    return read(ss.str());
  }

  void SaveStereoCameraInfo(std::filesystem::path const& path, StereoCameraInfo const& x) {
    std::ofstream f(path);
    if (!f) throw std::runtime_error("cannot write " + path.string());
    f << write(x, true);
  }
}  // namespace stereo_vision

namespace {
  template <std::size_t N>
  void mat(std::ostream& os, std::array<double, N> const& a, int cols) {  // cv::Mat's default formatter: [a, b, c;\n d, e, f]
    os << "[";
    for (std::size_t i = 0; i < N; i++) os << a[i] << (i + 1 == N ? "" : ((i + 1) % cols == 0 ? ";\n " : ", "));
    os << "]";
  }
}  // namespace

std::ostream& operator<<(std::ostream& os, stereo_vision::StereoCameraInfo const& s) {
  os << "Stereo camera calibration data:" << std::endl
     << "  reprojection error = " << s.reprojection_error << std::endl
     << "  image size = [" << s.width << " x " << s.height << "]" << std::endl
     << "  R = "; mat(os, s.R, 3);
  os << std::endl << "  T = "; mat(os, s.T, 1);
  os << std::endl << "  E = "; mat(os, s.E, 3);
  os << std::endl << "  F = "; mat(os, s.F, 3);
  os << std::endl << std::endl << "  left camera:" << std::endl << "    camera_matrix = "; mat(os, s.camera_matrix_left, 3);
  os << std::endl << "    dist_coeffs = "; mat(os, s.dist_coeffs_left, 5);
  os << std::endl << "  right camera:" << std::endl << "    camera_matrix = "; mat(os, s.camera_matrix_right, 3);
  os << std::endl << "    dist_coeffs = "; mat(os, s.dist_coeffs_right, 5);
  return os << std::endl;
}
