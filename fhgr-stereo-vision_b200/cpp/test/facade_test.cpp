// facade_test — exercises the C++ facade the way the reference's gtest does (test/stereo-vision-lib-test.cpp:5-13:
// empty images -> kInvalidImage) and, on a GPU box, one full AnalyzeAndAnnotateImage call on raw frames read from a
// binary file written by tests/test_gpu_parity.py; prints the checksums the Python side compares with the oracle.
// With a second file (W, H as int32, then K_l, D_l, K_r, D_r, R, T as f64) the instance is constructed from the
// calibration instead of from precomputed maps, i.e. the constructor does all of src/lib.cpp:29-47 itself.
#include <stereo-vision-gpu/lib.hpp>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace stereo_vision;

template <typename T>
static bool read_vec(std::ifstream& f, std::vector<T>& v, std::size_t n) {
  v.resize(n);
  f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(n * sizeof(T)));
  return static_cast<bool>(f);
}
template <typename T>
static unsigned long long checksum(std::vector<T> const& v) {  // FNV-1a over the bytes
  unsigned long long h = 1469598103934665603ull;
  auto const* p = reinterpret_cast<unsigned char const*>(v.data());
  for (std::size_t i = 0; i < v.size() * sizeof(T); i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: facade_test <case.bin>\n"); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  int hdr[4];  // W, H, px0.x.., packed below
  f.read(reinterpret_cast<char*>(hdr), sizeof hdr);
  int const W = hdr[0], H = hdr[1];
  std::size_t const n = static_cast<std::size_t>(W) * H;
  StereoVision::Settings s;
  s.algorithm = StereoVision::Settings::Algorithm::kSGBM;  // the reference's default is kMatchTemplate (lib.hpp:80)
  auto& r = s.rectification;
  r.width = W; r.height = H;
  std::vector<double> q;
  std::vector<float> bf;
  std::vector<std::uint8_t> left, right;
  if (!read_vec(f, r.map1_left, 2 * n) || !read_vec(f, r.map2_left, n) || !read_vec(f, r.map1_right, 2 * n) || !read_vec(f, r.map2_right, n) ||
      !read_vec(f, q, 16) || !read_vec(f, bf, 2) || !read_vec(f, left, 3 * n) || !read_vec(f, right, 3 * n)) {
    std::fprintf(stderr, "short case file\n");
    return 2;
  }
  std::memcpy(r.Q.data(), q.data(), sizeof(double) * 16);
  r.baseline = bf[0]; r.focal_length = bf[1];
  if (argc > 2) {
    std::ifstream cf(argv[2], std::ios::binary);
    int wh[2];
    cf.read(reinterpret_cast<char*>(wh), sizeof wh);
    std::vector<double> v;
    if (!read_vec(cf, v, 9 + 5 + 9 + 5 + 9 + 3)) { std::fprintf(stderr, "short calibration file\n"); return 2; }
    StereoCameraInfo info;
    info.width = wh[0]; info.height = wh[1];
    double const* p = v.data();
    std::memcpy(info.camera_matrix_left.data(), p, 9 * sizeof(double)); p += 9;
    std::memcpy(info.dist_coeffs_left.data(), p, 5 * sizeof(double)); p += 5;
    std::memcpy(info.camera_matrix_right.data(), p, 9 * sizeof(double)); p += 9;
    std::memcpy(info.dist_coeffs_right.data(), p, 5 * sizeof(double)); p += 5;
    std::memcpy(info.R.data(), p, 9 * sizeof(double)); p += 9;
    std::memcpy(info.T.data(), p, 3 * sizeof(double));
    s.stereo_camera_info = info;
    s.rectification = Rectification{};  // nothing precomputed
  }
  StereoVision const sv{s};

  auto const empty = sv.AnalyzeAndAnnotateImage(Image{}, Image{}, {});
  if (empty.has_value() || !(empty.error() == AnalysisError::kInvalidImage)) { std::puts("FAIL empty-image contract"); return 1; }

  Image const l{left.data(), W, H, 3, static_cast<std::size_t>(W) * 3}, rr{right.data(), W, H, 3, static_cast<std::size_t>(W) * 3};
  auto const res = sv.AnalyzeAndAnnotateImage(l, rr, {Point{hdr[2] % W, hdr[2] / W}, Point{hdr[3] % W, hdr[3] / W}});
  if (!res) { std::printf("FAIL %s\n", res.error().ToString().c_str()); return 1; }
  std::printf("rect_left %llu\nrect_right %llu\ndisp %llu\ndepth %llu\n", checksum(res->left_image.data), checksum(res->right_image.data),
              checksum(res->disparity.data), checksum(res->depth_map->data));
  std::printf("p0 %.9g %.9g %.9g\np1 %.9g %.9g %.9g\n", res->points_3d.first[0], res->points_3d.first[1], res->points_3d.first[2],
              res->points_3d.second[0], res->points_3d.second[1], res->points_3d.second[2]);
  std::cout << *res;
  return 0;
}
