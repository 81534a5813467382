// folder-main — the folder front-end of the reference (apps/stereo-vision/static-image-main.cpp:107-127) on the GPU facade.
//
//   folder-main <folder> [calibration.yml] [pairs per device call]
//
// <folder> holds `left/` and `right/` with one frame per file (and `calibration.yml` unless given), exactly the reference's
// layout.  The reference reads its frames with cv::imread and walks them one pair at a time through Viewer::ProcessImagePair;
// here the pairs go through StereoVision::AnalyzeAndAnnotateImages in batches (svgpu_pipeline_batch), which is where the GPU
// path gets its throughput.  Image decoding is outside the path: frames are binary PPM (P6, 8 bit).
// Prints, per pair, what AnalysisResult's stream operator prints for the two search points (image centre +- a quarter width)
// and an FNV-1a checksum of the disparity map; then the throughput.
#include <stereo-vision-gpu/lib.hpp>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>

namespace fs = std::filesystem;
using namespace stereo_vision;

struct Frame {
  std::vector<std::uint8_t> bgr;
  int width = 0, height = 0;
  [[nodiscard]] Image view() const { return Image{bgr.data(), width, height, 3, static_cast<std::size_t>(width) * 3}; }
};

// binary PPM -> BGR (cv::imread(IMREAD_COLOR) delivers BGR); an unreadable file gives an empty frame, like cv::imread
static Frame ReadPpm(fs::path const& path) {
  Frame f;
  std::ifstream in(path, std::ios::binary);
  std::string magic;
  int maxval = 0;
  auto next_token = [&](auto& v) {
    while (in >> std::ws && in.peek() == '#') in.ignore(1 << 20, '\n');
    in >> v;
  };
  next_token(magic);
  next_token(f.width);
  next_token(f.height);
  next_token(maxval);
  if (!in || magic != "P6" || maxval != 255 || f.width <= 0 || f.height <= 0) return Frame{};
  in.ignore(1);
  f.bgr.resize(static_cast<std::size_t>(f.width) * f.height * 3);
  in.read(reinterpret_cast<char*>(f.bgr.data()), static_cast<std::streamsize>(f.bgr.size()));
  if (!in) return Frame{};
  for (std::size_t i = 0; i + 2 < f.bgr.size(); i += 3) std::swap(f.bgr[i], f.bgr[i + 2]);
  return f;
}

static std::vector<fs::path> SortedFiles(fs::path const& dir) {
  std::vector<fs::path> v;
  for (auto const& e : fs::directory_iterator{dir})
    if (e.is_regular_file()) v.push_back(e.path());
  std::sort(v.begin(), v.end());
  return v;
}

int main(int argc, char** argv) {
  try {
    if (argc < 2) throw std::runtime_error("usage: folder-main <folder> [calibration.yml] [pairs per device call]");
    fs::path const folder = argv[1];
    fs::path const calibration = argc > 2 ? fs::path{argv[2]} : folder / "calibration.yml";
    int const in_flight = argc > 3 ? std::max(1, std::atoi(argv[3])) : 16;
    if (!fs::exists(folder / "left") || !fs::exists(folder / "right"))
      throw std::runtime_error("the image folder must contain a folder called 'left' and one called 'right'!");   // static-image-main.cpp:112
    StereoVision::Settings settings;
    settings.stereo_camera_info = LoadStereoCameraInfo(calibration);
    settings.algorithm = StereoVision::Settings::Algorithm::kSGBM;
    StereoVision const stereo_vis{settings};

    auto const left_files = SortedFiles(folder / "left"), right_files = SortedFiles(folder / "right");
    std::size_t const count = std::min(left_files.size(), right_files.size());   // std::views::zip stops at the shorter one
    std::vector<Frame> left(count), right(count);
    for (std::size_t i = 0; i < count; i++) { left[i] = ReadPpm(left_files[i]); right[i] = ReadPpm(right_files[i]); }
    std::vector<Image> lv(count), rv(count);
    std::vector<std::pair<Point, Point>> points(count);
    int const W = settings.stereo_camera_info->width, H = settings.stereo_camera_info->height;
    for (std::size_t i = 0; i < count; i++) {
      lv[i] = left[i].view();
      rv[i] = right[i].view();
      points[i] = {Point{W / 2 - W / 4, H / 2}, Point{W / 2 + W / 4, H / 2}};
    }
    auto const t0 = std::chrono::steady_clock::now();
    auto const results = stereo_vis.AnalyzeAndAnnotateImages(lv, rv, points, in_flight);
    double const seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (std::size_t i = 0; i < count; i++) {
      std::cout << left_files[i].filename().string() << " / " << right_files[i].filename().string() << ": ";
      if (!results[i]) { std::cout << "error: " << results[i].error().ToString() << std::endl; continue; }
      unsigned long long h = 1469598103934665603ull;
      auto const* p = reinterpret_cast<unsigned char const*>(results[i]->disparity.data.data());
      for (std::size_t k = 0; k < results[i]->disparity.data.size() * 2; k++) { h ^= p[k]; h *= 1099511628211ull; }
      std::cout << "disparity checksum " << h << std::endl << *results[i];
    }
    std::printf("%zu pairs in %.3f s (%.1f pairs/s, %d per device call, first call includes allocation)\n", count, seconds,
                count / std::max(seconds, 1e-9), in_flight);
  } catch (std::exception const& e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
