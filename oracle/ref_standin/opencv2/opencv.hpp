// Stand-in for the few OpenCV types monolidar_fusion touches (cv::FileStorage for parameters.yaml, cv::Mat /
// cv::Point for SemanticPlane). TEST INFRASTRUCTURE for oracle/_ref only; OpenCV's C++ headers are absent here.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;  // OpenCV's cvdef.h declares it at global scope

namespace cv {
typedef ::uchar uchar;

template <class T>
struct Point_ {
    T x = 0, y = 0;
    Point_() {}
    Point_(T x_, T y_) : x(x_), y(y_) {}  // double arguments convert implicitly (truncation for int)
};
typedef Point_<int> Point;

// single-channel 8-bit image, row-major, shared data like cv::Mat's ref-counted header copy
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, const uchar* src) : rows(r), cols(c), data_(std::make_shared<std::vector<uchar>>(src, src + size_t(r) * size_t(c))) {}
    // OpenCV does not bounds-check at(); the reference's SemanticPlane reads x == cols / y == rows (its validity
    // test uses '>', RansacPlane.cpp:205-206), which is undefined behaviour there. Here such reads return label 255.
    template <class T> T& at(const Point& p) {
        static T outside; outside = T(255);
        if (!inside(p)) return outside;
        return reinterpret_cast<T*>(data_->data())[size_t(p.y) * size_t(cols) + size_t(p.x)];
    }
    bool inside(const Point& p) const { return p.x >= 0 && p.x < cols && p.y >= 0 && p.y < rows; }

private:
    std::shared_ptr<std::vector<uchar>> data_;
};

// flat "key: value  # comment" YAML, which is all parameters.yaml contains; absent keys read as 0 like
// cv::FileNode's conversion of an empty node
class FileNode {
public:
    FileNode() {}
    explicit FileNode(const std::string& v) : v_(v), ok_(true) {}
    operator int() const { return ok_ ? int(std::strtod(v_.c_str(), nullptr)) : 0; }
    operator double() const { return ok_ ? std::strtod(v_.c_str(), nullptr) : 0.0; }
    operator float() const { return float(double(*this)); }
    operator std::string() const { return v_; }
    bool empty() const { return !ok_; }

private:
    std::string v_;
    bool ok_ = false;
};
class FileStorage {
public:
    enum Mode { READ = 0, WRITE = 1 };
    FileStorage(const std::string& path, int) {
        std::ifstream f(path);
        open_ = bool(f);
        std::string line;
        while (std::getline(f, line)) {
            size_t hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            size_t colon = line.find(':');
            if (colon == std::string::npos || line[0] == '%') continue;
            auto trim = [](std::string s) { size_t a = s.find_first_not_of(" \t\r\""), b = s.find_last_not_of(" \t\r\""); return a == std::string::npos ? std::string() : s.substr(a, b - a + 1); };
            std::string k = trim(line.substr(0, colon)), v = trim(line.substr(colon + 1));
            if (!k.empty() && !v.empty()) kv_[k] = v;
        }
    }
    bool isOpened() const { return open_; }
    FileNode operator[](const std::string& k) const { auto it = kv_.find(k); return it == kv_.end() ? FileNode() : FileNode(it->second); }
    FileNode operator[](const char* k) const { return (*this)[std::string(k)]; }
    void release() {}

private:
    bool open_ = false;
    std::map<std::string, std::string> kv_;
};
}  // namespace cv
