// stereo-vision-gpu/StereoCameraInfo.hpp — the calibration record of the reference and its YAML (de)serialisation,
// without OpenCV.
//
// Mirrors /root/reference/include/stereo-vision-lib/StereoCameraInfo.hpp:11-103: the same fields (cv::Mat -> fixed-size
// row-major arrays, cv::Size -> width / height), `write` / `read` for the `calibration:` node that cv::FileStorage
// produces (file_version, image_size, eight !!opencv-matrix entries, reprojection_error), `LoadStereoCameraInfo(path)`
// and the stream operator.  The reader accepts what cv::FileStorage writes (YAML 1.0 with `!!opencv-matrix` maps and
// flow sequences that wrap over several lines); the writer emits the same layout, so files round-trip with the
// reference's calibration application (apps/calibrate) and with cv::FileStorage.
#ifndef STEREO_VISION_GPU_STEREOCAMERAINFO_HPP
#define STEREO_VISION_GPU_STEREOCAMERAINFO_HPP

#include <array>
#include <cstdint>
#include <filesystem>
#include <iosfwd>
#include <string>

namespace stereo_vision {
  /// StereoCameraInfo.hpp:12 — version of the file format; checked when a file is read.
  constexpr std::uint8_t kStereoCameraInfoVersion = 1;

  /// StereoCameraInfo.hpp:15-45 (row-major; the distortion vectors hold k1 k2 p1 p2 k3 as cv::stereoCalibrate writes them).
  struct StereoCameraInfo {
    int width = 0, height = 0;                        ///< image_size
    std::array<double, 9> camera_matrix_left{}, camera_matrix_right{};
    std::array<double, 5> dist_coeffs_left{}, dist_coeffs_right{};
    std::array<double, 9> R{};
    std::array<double, 3> T{};
    std::array<double, 9> E{}, F{};
    double reprojection_error = 0.0;
  };

  /// StereoCameraInfo.hpp:51-65 — the `calibration` node as YAML text (a whole document when `document` is true).
  [[nodiscard]] auto write(StereoCameraInfo const& x, bool document = true) -> std::string;
  /// StereoCameraInfo.hpp:67-89 — parses a document holding a `calibration:` node.  Throws std::runtime_error with the
  /// reference's messages: "missing camera info in config file!" / "invalid config file version! Expected 1 but found N! ..."
  [[nodiscard]] auto read(std::string const& yaml_text) -> StereoCameraInfo;
  /// StereoCameraInfo.hpp:91-97
  [[nodiscard]] auto LoadStereoCameraInfo(std::filesystem::path const& path) -> StereoCameraInfo;
  void SaveStereoCameraInfo(std::filesystem::path const& path, StereoCameraInfo const& x);
}  // namespace stereo_vision

/// StereoCameraInfo.hpp:100-114
std::ostream& operator<<(std::ostream& os, stereo_vision::StereoCameraInfo const& stereo_camera_info);

#endif  // STEREO_VISION_GPU_STEREOCAMERAINFO_HPP
