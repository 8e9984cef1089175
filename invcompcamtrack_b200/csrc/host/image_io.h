// image_io.h — grayscale image loading for the drivers.  The reference uses cv::imread(GRAYSCALE)
// (run_io_reprojection_test.cpp:135-136, run_track_nposes.cpp:164); OpenCV is not part of this build, so the
// drivers read binary PGM (P5, maxval 255) — convert other formats with any image tool.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace ictio {

inline bool read_pgm(const std::string& name, std::vector<unsigned char>& pix, int& w, int& h) {
  FILE* f = std::fopen(name.c_str(), "rb");
  if (!f) return false;
  auto token = [&](char* buf, int cap) -> bool {
    int c = std::fgetc(f);
    for (;;) {
      while (c == ' ' || c == '\n' || c == '\r' || c == '\t') c = std::fgetc(f);
      if (c != '#') break;
      while (c != '\n' && c != EOF) c = std::fgetc(f);
    }
    int n = 0;
    while (c != EOF && c != ' ' && c != '\n' && c != '\r' && c != '\t' && n < cap - 1) { buf[n++] = (char)c; c = std::fgetc(f); }
    buf[n] = 0;
    return n > 0;
  };
  char t[32];
  bool ok = token(t, 32) && std::strcmp(t, "P5") == 0;
  int maxv = 0;
  if (ok) ok = token(t, 32) && (w = std::atoi(t)) > 0;
  if (ok) ok = token(t, 32) && (h = std::atoi(t)) > 0;
  if (ok) ok = token(t, 32) && (maxv = std::atoi(t)) == 255;
  if (ok) {
    pix.resize((size_t)w * h);
    ok = std::fread(pix.data(), 1, pix.size(), f) == pix.size();
  }
  std::fclose(f);
  return ok;
}

}  // namespace ictio
