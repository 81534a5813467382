// camera_info_test — the calibration-file code of the facade (stereo-vision-gpu/StereoCameraInfo.hpp) from the command line:
//   camera_info_test dump <in.yml>             prints every field with 17 significant digits (one "key v v v ..." line each)
//   camera_info_test copy <in.yml> <out.yml>   LoadStereoCameraInfo + SaveStereoCameraInfo
//   camera_info_test print <in.yml>            the stream operator
// Exit code 3 + the exception text on stderr when the file is rejected (missing node, wrong file_version).
#include <stereo-vision-gpu/StereoCameraInfo.hpp>

#include <cstdio>
#include <cstring>
#include <iostream>

using namespace stereo_vision;

template <std::size_t N>
static void line(char const* key, std::array<double, N> const& a) {
  std::printf("%s", key);
  for (double v : a) std::printf(" %.17g", v);
  std::printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: camera_info_test dump|copy|print <in.yml> [out.yml]\n"); return 2; }
  try {
    StereoCameraInfo const x = LoadStereoCameraInfo(argv[2]);
    if (!std::strcmp(argv[1], "copy") && argc > 3) {
      SaveStereoCameraInfo(argv[3], x);
    } else if (!std::strcmp(argv[1], "print")) {
      std::cout << x;
    } else {
      std::printf("image_size %d %d\n", x.width, x.height);
      line("camera_matrix_left", x.camera_matrix_left);
      line("camera_matrix_right", x.camera_matrix_right);
      line("dist_coeffs_left", x.dist_coeffs_left);
      line("dist_coeffs_right", x.dist_coeffs_right);
      line("R", x.R);
      line("T", x.T);
      line("E", x.E);
      line("F", x.F);
      std::printf("reprojection_error %.17g\n", x.reprojection_error);
    }
  } catch (std::exception const& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  return 0;
}
