// stereo-vision-gpu/lib.cpp — StereoVision on top of the svgpu C ABI (include/svgpu.h); see lib.hpp.
// Every numerical step of /root/reference/src/lib.cpp's kSGBM path is one call into libsvgpu.so; there is no CPU
// implementation in this file.
#include <stereo-vision-gpu/lib.hpp>

#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "svgpu.h"

namespace stereo_vision {

  auto AnalysisError::ToString() const -> std::string {
    switch (this->value_) {
      case kInvalidImage:
        return "invalid image!";
      case kInvalidSettings:
        return "Invalid settings!";
      default:
        throw std::runtime_error("unknown AnalysisError type!");
    }
  }

  AnalysisError::operator std::string() const { return this->ToString(); }

  std::ostream& operator<<(std::ostream& o, AnalysisResult const& r) {
    auto const& a = r.points_3d.first;
    auto const& b = r.points_3d.second;
    auto const norm = std::sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]));
    return o << "Distance between the two selected points: " << norm << " mm" << std::endl
             << "Distance of first point to camera:        " << a[2] << " mm" << std::endl
             << "Distance of second point to camera:       " << b[2] << " mm" << std::endl;
  }

  namespace {
    // the setter block of src/lib.cpp:248-259
    constexpr svgpu_sgbm_params kReferenceSgbm{/*minDisparity*/ 0, /*numDisparities*/ 16 * 14, /*blockSize*/ 7,
                                               /*P1*/ 8 * 7 * 7, /*P2*/ 32 * 7 * 7, /*disp12MaxDiff*/ 1,
                                               /*preFilterCap*/ 32, /*uniquenessRatio*/ 15, /*speckleWindowSize*/ 0,
                                               /*speckleRange*/ 0, /*mode*/ SV_MODE_HH4};

    void check(svgpu_ctx* c, int rc, char const* what) {
      if (rc != SV_OK) throw std::runtime_error(std::string(what) + ": " + svgpu_last_error(c));
    }
  }  // namespace

  StereoVision::StereoVision(Settings settings) : settings_(std::move(settings)) {
    auto& r = this->settings_.rectification;
    auto const& info = this->settings_.stereo_camera_info;
    if (info) {  // src/lib.cpp:29-47 from the calibration: cv::stereoRectify (host) ...
      std::array<double, 9> R1{}, R2{};
      std::array<double, 12> P1{}, P2{};
      if (svgpu_stereo_rectify(info->camera_matrix_left.data(), info->dist_coeffs_left.data(), info->camera_matrix_right.data(),
                               info->dist_coeffs_right.data(), info->width, info->height, info->R.data(), info->T.data(), R1.data(),
                               R2.data(), P1.data(), P2.data(), r.Q.data()) != SV_OK)
        throw std::runtime_error("stereo_camera_info cannot be rectified (degenerate calibration)");
      r.width = info->width;
      r.height = info->height;
      r.baseline = static_cast<float>(std::abs(info->T[0]));          // src/lib.cpp:283
      r.focal_length = static_cast<float>(info->camera_matrix_left[0]);  // src/lib.cpp:284
      svgpu_ctx* c = nullptr;
      if (svgpu_create(this->settings_.device, r.width, r.height, &c) != SV_OK) throw std::runtime_error("no usable CUDA device for svgpu");
      this->ctx_ = std::shared_ptr<svgpu_ctx>(c, [](svgpu_ctx* p) { svgpu_destroy(p); });
      // ... and the two cv::initUndistortRectifyMap(CV_16SC2), generated on the device
      check(c, svgpu_init_rectify_maps(c, SV_LEFT, info->camera_matrix_left.data(), info->dist_coeffs_left.data(), R1.data(), P1.data()),
            "svgpu_init_rectify_maps(left)");
      check(c, svgpu_init_rectify_maps(c, SV_RIGHT, info->camera_matrix_right.data(), info->dist_coeffs_right.data(), R2.data(), P2.data()),
            "svgpu_init_rectify_maps(right)");
    } else {
      auto const n = static_cast<std::size_t>(r.width) * static_cast<std::size_t>(r.height);
      if (r.width <= 0 || r.height <= 0 || r.map1_left.size() != 2 * n || r.map1_right.size() != 2 * n || r.map2_left.size() != n ||
          r.map2_right.size() != n)
        throw std::runtime_error("rectification maps do not match image_size");
      svgpu_ctx* c = nullptr;
      if (svgpu_create(this->settings_.device, r.width, r.height, &c) != SV_OK) throw std::runtime_error("no usable CUDA device for svgpu");
      this->ctx_ = std::shared_ptr<svgpu_ctx>(c, [](svgpu_ctx* p) { svgpu_destroy(p); });
      check(c, svgpu_set_rectify_maps(c, SV_LEFT, r.map1_left.data(), r.map2_left.data()), "svgpu_set_rectify_maps(left)");
      check(c, svgpu_set_rectify_maps(c, SV_RIGHT, r.map1_right.data(), r.map2_right.data()), "svgpu_set_rectify_maps(right)");
    }
    svgpu_ctx* c = this->ctx_.get();
    check(c, svgpu_set_Q(c, r.Q.data()), "svgpu_set_Q");
    check(c, svgpu_set_depth_consts(c, r.baseline, r.focal_length), "svgpu_set_depth_consts");
    check(c, svgpu_set_sgbm(c, &kReferenceSgbm), "svgpu_set_sgbm");
  }

  auto StereoVision::RescaleAndRectifyImages(Image const& left_image, Image const& right_image) const
      -> std::pair<Buffer<std::uint8_t>, Buffer<std::uint8_t>> {
    auto const& r = this->settings_.rectification;
    svgpu_ctx* c = this->ctx_.get();
    auto one = [&](int side, Image const& in) {
      Buffer<std::uint8_t> out;
      out.width = r.width; out.height = r.height; out.channels = in.channels;
      out.data.resize(static_cast<std::size_t>(r.width) * r.height * in.channels);
      std::uint8_t const* src = in.data;
      int sw = in.width, sh = in.height;
      std::size_t ss = in.stride;
      std::vector<std::uint8_t> scaled;
      if (in.width != r.width) {  // src/lib.cpp:97 compares the width only
        scaled.resize(out.data.size());
        check(c, svgpu_resize(c, in.data, in.width, in.height, in.stride, in.channels, scaled.data(),
                              static_cast<std::size_t>(r.width) * in.channels, SV_MEM_HOST), "svgpu_resize");
        src = scaled.data(); sw = r.width; sh = r.height; ss = static_cast<std::size_t>(r.width) * in.channels;
      }
      check(c, svgpu_remap(c, side, src, sw, sh, ss, in.channels, out.data.data(), static_cast<std::size_t>(r.width) * in.channels,
                           SV_MEM_HOST), "svgpu_remap");
      return out;
    };
    return {one(SV_LEFT, left_image), one(SV_RIGHT, right_image)};
  }

  auto StereoVision::AnalyzeAndAnnotateImage(Image const& left_image, Image const& right_image,
                                             std::pair<Point, Point> const& search_points) const
      -> std::expected<AnalysisResult, AnalysisError> {
    if (left_image.empty() || right_image.empty()) {
      return std::unexpected{AnalysisError::kInvalidImage};
    }
    if (this->settings_.algorithm != Settings::Algorithm::kSGBM || left_image.channels != 3 || right_image.channels != 3) {
      return std::unexpected{AnalysisError::kInvalidSettings};  // template / ORB matching are not on the GPU path
    }
    auto const& r = this->settings_.rectification;
    svgpu_ctx* c = this->ctx_.get();
    auto const W = r.width, H = r.height;
    auto const n = static_cast<std::size_t>(W) * H;

    // rescale on its own (the fused chain starts at calibration size), everything else in one svgpu_pipeline call
    auto scale = [&](Image const& in, std::vector<std::uint8_t>& store) -> Image {
      if (in.width == W) return in;
      store.resize(n * 3);
      check(c, svgpu_resize(c, in.data, in.width, in.height, in.stride, 3, store.data(), static_cast<std::size_t>(W) * 3, SV_MEM_HOST),
            "svgpu_resize");
      return Image{store.data(), W, H, 3, static_cast<std::size_t>(W) * 3};
    };
    std::vector<std::uint8_t> sl, sr;
    Image const l = scale(left_image, sl), rr = scale(right_image, sr);

    AnalysisResult out;
    out.left_image = {std::vector<std::uint8_t>(n * 3), W, H, 3};
    out.right_image = {std::vector<std::uint8_t>(n * 3), W, H, 3};
    out.disparity = {std::vector<std::int16_t>(n), W, H, 1};
    Buffer<float> xyz{std::vector<float>(n * 3), W, H, 3};
    Buffer<float> depth{std::vector<float>(n), W, H, 1};

    svgpu_pipeline_io io{};
    io.left = l.data; io.right = rr.data; io.left_stride = l.stride; io.right_stride = rr.stride; io.channels = 3;
    io.left_rect = out.left_image.data.data(); io.right_rect = out.right_image.data.data(); io.rect_stride = static_cast<std::size_t>(W) * 3;
    io.disp = out.disparity.data.data(); io.disp_stride = static_cast<std::size_t>(W) * 2;
    io.xyz = xyz.data.data(); io.xyz_stride = static_cast<std::size_t>(W) * 12;
    io.depth = depth.data.data(); io.depth_stride = static_cast<std::size_t>(W) * 4;
    io.mem = SV_MEM_HOST;
    int const rc = svgpu_pipeline(c, &io, SV_STAGE_REMAP | SV_STAGE_PREFILTER | SV_STAGE_SGBM | SV_STAGE_POSTFILTER | SV_STAGE_REPROJECT | SV_STAGE_DEPTH);
    if (rc == SV_INVALID_IMAGE) return std::unexpected{AnalysisError::kInvalidImage};
    if (rc != SV_OK) return std::unexpected{AnalysisError::kInvalidSettings};

    auto at = [&](Point p) -> Vec3f {  // src/lib.cpp:155: reprojected.at<cv::Vec3f>(point)
      if (p.x < 0 || p.y < 0 || p.x >= W || p.y >= H) return {0.0f, 0.0f, 0.0f};
      float const* v = xyz.row(p.y) + 3 * p.x;
      return {v[0], v[1], v[2]};
    };
    out.points_3d = {at(search_points.first), at(search_points.second)};
    out.depth_map = std::move(depth);
    return out;
  }

  auto StereoVision::AnalyzeAndAnnotateImages(std::span<Image const> left_images, std::span<Image const> right_images,
                                              std::span<std::pair<Point, Point> const> search_points, int max_in_flight) const
      -> std::vector<std::expected<AnalysisResult, AnalysisError>> {
    std::size_t const count = std::min(left_images.size(), right_images.size());
    std::vector<std::expected<AnalysisResult, AnalysisError>> results;
    results.reserve(count);
    for (std::size_t i = 0; i < count; i++) results.emplace_back(std::unexpected{AnalysisError::kInvalidImage});
    auto const& r = this->settings_.rectification;
    svgpu_ctx* c = this->ctx_.get();
    auto const W = r.width, H = r.height;
    auto const n = static_cast<std::size_t>(W) * H;
    // frames that can go through the device: non-empty BGR pairs (the others keep their error, src/lib.cpp:54-56, 71-73)
    std::vector<std::size_t> todo;
    for (std::size_t i = 0; i < count; i++) {
      if (left_images[i].empty() || right_images[i].empty()) continue;
      if (this->settings_.algorithm != Settings::Algorithm::kSGBM || left_images[i].channels != 3 || right_images[i].channels != 3) {
        results[i] = std::unexpected{AnalysisError::kInvalidSettings};
        continue;
      }
      todo.push_back(i);
    }
    int const chunk = std::max(1, max_in_flight);
    check(c, svgpu_reserve(c, chunk), "svgpu_reserve");
    std::vector<std::uint8_t> in_l, in_r, rect_l, rect_r;
    std::vector<std::int16_t> disp;
    std::vector<float> xyz, depth;
    for (std::size_t b0 = 0; b0 < todo.size(); b0 += static_cast<std::size_t>(chunk)) {
      std::size_t const m = std::min<std::size_t>(chunk, todo.size() - b0);
      in_l.resize(m * n * 3); in_r.resize(m * n * 3); rect_l.resize(m * n * 3); rect_r.resize(m * n * 3);
      disp.resize(m * n); xyz.resize(m * n * 3); depth.resize(m * n);
      // gather the frames tightly packed at calibration size (RescaleImage, src/lib.cpp:96-104, where the width differs)
      for (std::size_t k = 0; k < m; k++) {
        auto pack = [&](Image const& in, std::uint8_t* dst) {
          if (in.width != W) {
            check(c, svgpu_resize(c, in.data, in.width, in.height, in.stride, 3, dst, static_cast<std::size_t>(W) * 3, SV_MEM_HOST), "svgpu_resize");
          } else {
            for (int y = 0; y < H; y++)
              std::copy_n(in.data + static_cast<std::size_t>(y) * in.stride, static_cast<std::size_t>(W) * 3, dst + static_cast<std::size_t>(y) * W * 3);
          }
        };
        pack(left_images[todo[b0 + k]], in_l.data() + k * n * 3);
        pack(right_images[todo[b0 + k]], in_r.data() + k * n * 3);
      }
      svgpu_pipeline_batch_io io{};
      io.left = in_l.data(); io.right = in_r.data(); io.channels = 3;
      io.left_rect = rect_l.data(); io.right_rect = rect_r.data();
      io.disp = disp.data(); io.xyz = xyz.data(); io.depth = depth.data();
      io.mem = SV_MEM_HOST;
      int const rc = svgpu_pipeline_batch(c, static_cast<int>(m), &io,
                                          SV_STAGE_REMAP | SV_STAGE_PREFILTER | SV_STAGE_SGBM | SV_STAGE_POSTFILTER | SV_STAGE_REPROJECT | SV_STAGE_DEPTH);
      for (std::size_t k = 0; k < m; k++) {
        std::size_t const i = todo[b0 + k];
        if (rc == SV_INVALID_IMAGE) { results[i] = std::unexpected{AnalysisError::kInvalidImage}; continue; }
        if (rc != SV_OK) { results[i] = std::unexpected{AnalysisError::kInvalidSettings}; continue; }
        AnalysisResult out;
        out.left_image = {std::vector<std::uint8_t>(rect_l.begin() + k * n * 3, rect_l.begin() + (k + 1) * n * 3), W, H, 3};
        out.right_image = {std::vector<std::uint8_t>(rect_r.begin() + k * n * 3, rect_r.begin() + (k + 1) * n * 3), W, H, 3};
        out.disparity = {std::vector<std::int16_t>(disp.begin() + k * n, disp.begin() + (k + 1) * n), W, H, 1};
        float const* px = xyz.data() + k * n * 3;
        auto at = [&](Point p) -> Vec3f {
          if (p.x < 0 || p.y < 0 || p.x >= W || p.y >= H) return {0.0f, 0.0f, 0.0f};
          float const* v = px + (static_cast<std::size_t>(p.y) * W + p.x) * 3;
          return {v[0], v[1], v[2]};
        };
        auto const& sp = search_points[std::min(i, search_points.size() - 1)];
        out.points_3d = {at(sp.first), at(sp.second)};
        out.depth_map = Buffer<float>{std::vector<float>(depth.begin() + k * n, depth.begin() + (k + 1) * n), W, H, 1};
        results[i] = std::move(out);
      }
    }
    return results;
  }

}  // namespace stereo_vision
