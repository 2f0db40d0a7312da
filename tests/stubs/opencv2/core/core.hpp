// Minimal stand-in for cv::Mat as the SemanticPlane interface touches it (8-bit single-channel label image). OpenCV's C++
// headers are absent from this image (SURVEY.md 8b "Compile-compat"); with the real headers on the include path this
// directory is simply not used.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#ifndef CV_8UC1
#define CV_8UC1 0
#endif

namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), d_(std::make_shared<std::vector<uchar>>((size_t)r * (size_t)c)) {}
    Mat(int r, int c, int /*type*/, void* data) : rows(r), cols(c), d_(std::make_shared<std::vector<uchar>>((size_t)r * (size_t)c)) {
        std::memcpy(d_->data(), data, d_->size());
    }
    int type() const { return CV_8UC1; }
    int channels() const { return 1; }
    bool empty() const { return rows == 0 || cols == 0; }
    template <typename T> T* ptr(int y) { return reinterpret_cast<T*>(d_->data() + (size_t)y * (size_t)cols); }
    template <typename T> const T* ptr(int y) const { return reinterpret_cast<const T*>(d_->data() + (size_t)y * (size_t)cols); }

private:
    std::shared_ptr<std::vector<uchar>> d_;
};
}  // namespace cv
