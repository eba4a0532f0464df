// Stub for the one OpenCV use reachable from the reference's CUDA translation units:
// PinholeCamera<T>::FromFile (common/pinhole_camera_impl.h:144-160), which the parity
// oracle never calls.  Only the names that function mentions are declared.
#pragma once
#include <stdexcept>
#include <string>
namespace cv
{
struct Mat
{
  template <typename T>
  T at(int, int) const { throw std::runtime_error("opencv stub"); }
  Mat clone() const { return *this; } // Frame's copy constructor (core/mapping/frame.h:37), reached by check_mapper_header.cpp
};
struct FileNode
{
  template <typename T>
  void operator>>(T &) const { throw std::runtime_error("opencv stub"); }
};
struct FileStorage
{
  enum { READ = 0 };
  FileStorage(const std::string &, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const char *) const { return FileNode(); }
};
} // namespace cv
