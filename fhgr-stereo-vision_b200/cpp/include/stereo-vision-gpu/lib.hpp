// stereo-vision-gpu/lib.hpp — the public surface of stereo-vision-lib, re-declared on top of the svgpu C ABI.
//
// Mirrors /root/reference/include/stereo-vision-lib/lib.hpp:21-112 (AnalysisError, AnalysisResult,
// StereoVision::Settings, StereoVision::RescaleAndRectifyImages / AnalyzeAndAnnotateImage) for the kSGBM path that
// src/lib.cpp:53-85 dispatches to.  OpenCV's C++ headers are not part of this build, so images travel as plain
// views (`Image`: pointer, size, byte stride, channels — the fields of a cv::Mat) and the one-time calibration
// products that src/lib.cpp:29-47 obtains from cv::stereoRectify / cv::initUndistortRectifyMap are handed in by the
// caller (`Rectification`).  INTEGRATION.md shows the cv::Mat adapter a maintainer of the reference would write.
#ifndef STEREO_VISION_GPU_LIB_HPP
#define STEREO_VISION_GPU_LIB_HPP

#include <array>
#include <cstdint>
#include <expected>
#include <memory>
#include <optional>
#include <ostream>
#include <span>
#include <string>
#include <utility>
#include <vector>

#include <stereo-vision-gpu/StereoCameraInfo.hpp>

struct svgpu_ctx;

namespace stereo_vision {

  /// lib.hpp:21-47 — same values, same strings (src/lib.cpp:7-20).
  class AnalysisError {
  public:
    enum Value {
      kInvalidImage,
      kInvalidSettings,
    };
    AnalysisError() = default;
    constexpr AnalysisError(Value const value) : value_(value) {}
    constexpr explicit operator Value() const { return value_; }
    bool operator==(Value const value) const { return this->value_ == value; }
    explicit operator bool() const = delete;
    [[nodiscard]] auto ToString() const -> std::string;
    explicit operator std::string() const;

  private:
    Value value_;
  };

  /// The fields of a cv::Mat that the path uses; non-owning.
  struct Image {
    std::uint8_t const* data = nullptr;
    int width = 0, height = 0, channels = 0;  // 8UC1 or 8UC3 (BGR)
    std::size_t stride = 0;                    // bytes per row (cv::Mat::step)
    [[nodiscard]] bool empty() const { return data == nullptr || width <= 0 || height <= 0; }
  };
  struct Point { int x = 0, y = 0; };
  using Vec3f = std::array<float, 3>;

  /// Owning row-major buffer for results (what a cv::Mat return value is in the reference).
  template <typename T>
  struct Buffer {
    std::vector<T> data;
    int width = 0, height = 0, channels = 1;
    [[nodiscard]] T const* row(int y) const { return data.data() + static_cast<std::size_t>(y) * width * channels; }
  };

  /// lib.hpp:52-61
  struct AnalysisResult {
    Buffer<std::uint8_t> left_image;    ///< rectified left image
    Buffer<std::uint8_t> right_image;   ///< rectified right image
    std::pair<Vec3f, Vec3f> points_3d;  ///< 3D points of the two search points
    std::optional<Buffer<float>> depth_map;
    Buffer<std::int16_t> disparity;     ///< CV_16S disparity after the post filter (not exposed by the reference)
  };
  std::ostream& operator<<(std::ostream& o, AnalysisResult const& analysis_result);

  /// What StereoVision::StereoVision computes once on the host (src/lib.cpp:32-46) and keeps as members
  /// (lib.hpp:118-124): the two CV_16SC2 map pairs and Q_, plus the two scalars of CalculateDepthMapSimple
  /// (src/lib.cpp:283-284).
  struct Rectification {
    int width = 0, height = 0;                       ///< StereoCameraInfo::image_size
    std::vector<std::int16_t> map1_left, map1_right;  ///< H*W*2, (x, y) integer parts
    std::vector<std::uint16_t> map2_left, map2_right; ///< H*W, fy*32+fx
    std::array<double, 16> Q{};                       ///< row-major 4x4
    float baseline = 0.0f;                            ///< |T[0]|
    float focal_length = 0.0f;                        ///< camera_matrix_left(0,0)
  };

  class StereoVision {
  public:
    struct Settings {
      enum class Algorithm {
        kMatchTemplate,
        kORB,
        kSGBM,
      };
      /// Either the calibration itself: the constructor then does what src/lib.cpp:32-46 does (stereoRectify on the
      /// host, both initUndistortRectifyMap on the device) ...
      std::optional<StereoCameraInfo> stereo_camera_info;
      /// ... or the products of those calls, computed elsewhere (e.g. by OpenCV in the reference's constructor).
      Rectification rectification;
      /// lib.hpp:80 — same default as the reference.  Only kSGBM (the dense path, src/lib.cpp:67-69) has a GPU implementation;
      /// kMatchTemplate and kORB (src/lib.cpp:61-66, per-point CPU heuristics outside the hot path) yield kInvalidSettings.
      Algorithm algorithm = Algorithm::kMatchTemplate;
      bool show_debug_info = false;
      int device = 0;                                 ///< CUDA device of this instance
    };

    /// Throws std::runtime_error when the device context cannot be created or the settings are unusable
    /// (as the reference's constructor path does, StereoCameraInfo.hpp:69,75).
    explicit StereoVision(Settings settings);

    /// lib.hpp:96 — cv::resize to the calibration size if the width differs (src/lib.cpp:96-104), then cv::remap
    /// with the rectification maps (src/lib.cpp:87-94).
    [[nodiscard]] auto RescaleAndRectifyImages(Image const& left_image, Image const& right_image) const
        -> std::pair<Buffer<std::uint8_t>, Buffer<std::uint8_t>>;

    /// lib.hpp:108 — src/lib.cpp:53-85 with Algorithm::kSGBM: rescale, rectify, blur/gray/sharpen, SGBM
    /// (MODE_HH4, 224 disparities, src/lib.cpp:248-259), filterSpeckles + close, reprojectImageTo3D, depth map.
    [[nodiscard]] auto AnalyzeAndAnnotateImage(Image const& left_image, Image const& right_image,
                                               std::pair<Point, Point> const& search_points) const
        -> std::expected<AnalysisResult, AnalysisError>;

    /// The same for n frame pairs in one call (svgpu_pipeline_batch): what the folder loop (apps/stereo-vision/
    /// static-image-main.cpp:107-127) and the livecam loop (livecam-main.cpp:70-106) do frame by frame.  One result per pair, in
    /// order; a pair with an empty image yields kInvalidImage without affecting the others.  `max_in_flight` pairs are processed
    /// per device call (scratch memory: about 1 GB per pair at 1080p / 128 disparities).
    [[nodiscard]] auto AnalyzeAndAnnotateImages(std::span<Image const> left_images, std::span<Image const> right_images,
                                                std::span<std::pair<Point, Point> const> search_points, int max_in_flight = 16) const
        -> std::vector<std::expected<AnalysisResult, AnalysisError>>;

    /// lib.hpp:111-112 — returns a copy, like the reference.
    [[nodiscard]] Settings settings() const { return this->settings_; }

  private:
    Settings settings_;
    std::shared_ptr<svgpu_ctx> ctx_;  // shared by copies, as the reference's const copies share their maps (Viewer.hpp:43)
  };

}  // namespace stereo_vision

#endif  // STEREO_VISION_GPU_LIB_HPP
