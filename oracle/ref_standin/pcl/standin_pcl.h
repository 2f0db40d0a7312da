// Stand-in for the subset of PCL 1.8 that monolidar_fusion's hot path calls. TEST INFRASTRUCTURE for
// oracle/_ref only: PCL is an un-vendored, absent dependency of the reference; this header restates the
// published algorithms of the classes used (PassThrough, RandomSample, SampleConsensusModelPlane,
// SampleConsensusModelPerpendicularPlane, RandomSampleConsensus, transformPointCloud,
// pointToPlaneDistance) so that the reference's own RansacPlane.cpp / DepthEstimator.cpp compile and run
// unmodified. Arithmetic is float where PCL's is float. The random streams are rand()-based like PCL's,
// but seeded through pcl::standin::seed (PCL seeds RandomSample with time(), which is irreproducible).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include <Eigen/Eigen>

namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
}  // namespace boost

namespace pcl {

namespace standin {
inline unsigned& seed() { static unsigned s = 12345u; return s; }
}  // namespace standin

struct PointXYZ {
    float x = 0, y = 0, z = 0, pad = 1.f;
    PointXYZ() {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) PointXYZI {  // 32 bytes: x y z pad | intensity pad pad pad (PCL's SSE-padded layout)
    float x = 0, y = 0, z = 0, pad0 = 1.f;
    float intensity = 0, pad1 = 0, pad2 = 0, pad3 = 0;
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");

struct PCLHeader {
    uint32_t seq = 0;
    uint64_t stamp = 0;
    std::string frame_id;
};

template <class PointT>
class PointCloud {
public:
    typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
    typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
    PCLHeader header;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    void clear() { points.clear(); width = 0; height = 0; }
    size_t size() const { return points.size(); }
    void push_back(const PointT& p) { points.push_back(p); width = uint32_t(points.size()); height = 1; }
    // Eigen::Map of the first 4 floats of every point (dim 4 x n, stride sizeof(PointT))
    struct MapXf {
        const PointCloud* c;
        template <class U> struct Cast {
            const PointCloud* c;
            template <int N> Eigen::Matrix<U, N, Eigen::Dynamic> topRows() const {
                Eigen::Matrix<U, N, Eigen::Dynamic> m; m.resize(N, int(c->points.size()));
                for (size_t j = 0; j < c->points.size(); j++) { const float* f = reinterpret_cast<const float*>(&c->points[j]); for (int i = 0; i < N; i++) m(i, int(j)) = U(f[i]); }
                return m;
            }
        };
        template <class U> Cast<U> cast() const { return Cast<U>{c}; }
    };
    MapXf getMatrixXfMap() const { return MapXf{this}; }
};

// pcl::pointToPlaneDistance (sample_consensus/sac_model_plane.h): float expression, left to right
template <class Point>
inline double pointToPlaneDistance(const Point& p, const Eigen::Vector4f& c) {
    return std::fabs(c[0] * p.x + c[1] * p.y + c[2] * p.z + c[3]);
}

// pcl::transformPointCloud(cloud_in, cloud_out, Transform<Scalar,3,Affine>) (PCL 1.8 common/impl/transforms.hpp): per
// point and coordinate  static_cast<float>(t(i,0) * x + t(i,1) * y + t(i,2) * z + t(i,3))  evaluated in the transform's
// Scalar (double for the Affine3d the reference passes), left to right; non-finite points of a non-dense cloud are
// copied unchanged (they stay non-finite either way).
template <class PointT, class Scalar>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Transform<Scalar, 3, Eigen::Affine>& tf) {
    out = in;
    const auto& m = tf.matrix();
    for (size_t k = 0; k < in.points.size(); k++) {
        const PointT& p = in.points[k];
        const Scalar x = p.x, y = p.y, z = p.z;
        out.points[k].x = static_cast<float>(m(0, 0) * x + m(0, 1) * y + m(0, 2) * z + m(0, 3));
        out.points[k].y = static_cast<float>(m(1, 0) * x + m(1, 1) * y + m(1, 2) * z + m(1, 3));
        out.points[k].z = static_cast<float>(m(2, 0) * x + m(2, 1) * y + m(2, 2) * z + m(2, 3));
    }
}

// ---- filters ----
template <class PointT>
class PCLBase {
public:
    typedef typename PointCloud<PointT>::ConstPtr CloudConstPtr;
    void setInputCloud(const CloudConstPtr& c) { input_ = c; }
    void setIndices(const boost::shared_ptr<std::vector<int>>& idx) { indices_ = idx; }

protected:
    CloudConstPtr input_;
    boost::shared_ptr<std::vector<int>> indices_;
    const std::vector<int>& indicesOrAll() {
        if (!indices_) { indices_ = boost::make_shared<std::vector<int>>(input_->points.size()); for (size_t i = 0; i < indices_->size(); i++) (*indices_)[i] = int(i); }
        return *indices_;
    }
};

// PassThrough on one field: keeps finite values with min <= v <= max (filter_limit_negative_ = false)
template <class PointT>
class PassThrough : public PCLBase<PointT> {
public:
    void setFilterFieldName(const std::string& f) { field_ = f; }
    void setFilterLimits(double lo, double hi) { lo_ = float(lo); hi_ = float(hi); }
    void filter(std::vector<int>& out) {
        const std::vector<int>& idx = this->indicesOrAll();
        std::vector<int> keep;
        for (int i : idx) {
            const PointT& p = this->input_->points[size_t(i)];
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
            float v = field_ == "x" ? p.x : field_ == "y" ? p.y : p.z;
            if (!std::isfinite(v)) continue;
            if (v > hi_ || v < lo_) continue;
            keep.push_back(i);
        }
        out.swap(keep);
    }

private:
    std::string field_ = "z";
    float lo_ = -3.4e38f, hi_ = 3.4e38f;
};

// RandomSample: Vitter's algorithm S over the index list (order preserving), rand()-driven
template <class PointT>
class RandomSample : public PCLBase<PointT> {
public:
    void setSample(unsigned s) { sample_ = s; }
    void setSeed(unsigned s) { seed_ = s; }
    void filter(std::vector<int>& out) {
        const std::vector<int>& src = this->indicesOrAll();
        unsigned N = unsigned(src.size());
        std::vector<int> res;
        if (sample_ >= N) { res = src; }
        else {
            res.resize(sample_);
            std::srand(seed_);
            unsigned top = N - sample_, i = 0, index = 0;
            for (size_t n = sample_; n >= 2; n--) {
                float V = unifRand();
                unsigned S = 0;
                float quot = float(top) / float(N);
                while (quot > V) { S++; top--; N--; quot = quot * float(top) / float(N); }
                index += S;
                res[i++] = src[index++];
                N--;
            }
            index += N * unsigned(unifRand());
            res[i++] = src[index++];
        }
        out.swap(res);
    }

private:
    static float unifRand() { return float(std::rand()) / float(RAND_MAX); }
    unsigned sample_ = UINT_MAX;
    unsigned seed_ = standin::seed();
};
template <class PointT> class ApproximateVoxelGrid {};

// ---- sample consensus ----
template <class PointT>
class SampleConsensusModelPlane {
public:
    typedef boost::shared_ptr<SampleConsensusModelPlane> Ptr;
    typedef typename PointCloud<PointT>::ConstPtr CloudConstPtr;
    explicit SampleConsensusModelPlane(const CloudConstPtr& c) : input_(c) {
        indices_.resize(c->points.size());
        for (size_t i = 0; i < indices_.size(); i++) indices_[i] = int(i);
        shuffled_ = indices_;
        std::srand(standin::seed());
    }
    virtual ~SampleConsensusModelPlane() {}
    void setIndices(const std::vector<int>& idx) { indices_ = idx; shuffled_ = idx; }
    const std::vector<int>& getIndices() const { return indices_; }
    unsigned getSampleSize() const { return 3; }

    // SampleConsensusModel::getSamples / drawIndexSample / isSampleGood
    void getSamples(int&, std::vector<int>& samples) {
        if (indices_.size() < 3) { samples.clear(); return; }
        samples.resize(3);
        for (unsigned iter = 0; iter < 1000; ++iter) {
            size_t n = shuffled_.size();
            for (unsigned i = 0; i < 3; ++i) std::swap(shuffled_[i], shuffled_[i + (size_t(std::rand()) % (n - i))]);
            std::copy(shuffled_.begin(), shuffled_.begin() + 3, samples.begin());
            if (isSampleGood(samples)) return;
        }
        samples.clear();
    }
    bool isSampleGood(const std::vector<int>& s) const {
        const PointT &a = input_->points[size_t(s[0])], &b = input_->points[size_t(s[1])], &c = input_->points[size_t(s[2])];
        float p1x = b.x - a.x, p1y = b.y - a.y, p1z = b.z - a.z, p2x = c.x - a.x, p2y = c.y - a.y, p2z = c.z - a.z;
        float d0 = p1x / p2x, d1 = p1y / p2y, d2 = p1z / p2z;
        return (d0 != d1) || (d2 != d1);
    }
    bool computeModelCoefficients(const std::vector<int>& s, Eigen::VectorXf& mc) const {
        if (s.size() != 3) return false;
        const PointT &a = input_->points[size_t(s[0])], &b = input_->points[size_t(s[1])], &c = input_->points[size_t(s[2])];
        float p1x = b.x - a.x, p1y = b.y - a.y, p1z = b.z - a.z, p2x = c.x - a.x, p2y = c.y - a.y, p2z = c.z - a.z;
        float d0 = p1x / p2x, d1 = p1y / p2y, d2 = p1z / p2z;
        if ((d0 == d1) && (d2 == d1)) return false;  // collinear
        mc.resize(4);
        mc[0] = p1y * p2z - p1z * p2y;
        mc[1] = p1z * p2x - p1x * p2z;
        mc[2] = p1x * p2y - p1y * p2x;
        float nrm = std::sqrt(mc[0] * mc[0] + mc[1] * mc[1] + mc[2] * mc[2]);
        mc[0] /= nrm; mc[1] /= nrm; mc[2] /= nrm;
        mc[3] = -1.f * (mc[0] * a.x + mc[1] * a.y + mc[2] * a.z);
        return true;
    }
    float dist(const Eigen::VectorXf& mc, int i) const {
        const PointT& p = input_->points[size_t(i)];
        return std::fabs(mc[0] * p.x + mc[1] * p.y + mc[2] * p.z + mc[3]);  // pt = (x,y,z,1), dot with coefficients
    }
    virtual bool isModelValid(const Eigen::VectorXf& mc) const { return mc.size() == 4; }
    virtual int countWithinDistance(const Eigen::VectorXf& mc, double thr) const {
        if (!isModelValid(mc)) return 0;
        int n = 0;
        for (int i : indices_) if (double(dist(mc, i)) < thr) n++;
        return n;
    }
    virtual void selectWithinDistance(const Eigen::VectorXf& mc, double thr, std::vector<int>& inliers) const {
        std::vector<int> out;
        if (isModelValid(mc)) for (int i : indices_) if (double(dist(mc, i)) < thr) out.push_back(i);
        inliers.swap(out);
    }
    // least-squares refit: centroid + eigenvector of the smallest eigenvalue of the covariance (float
    // accumulation like computeMeanAndCovarianceMatrix<float>); falls back to the input when invalid
    void optimizeModelCoefficients(const std::vector<int>& inliers, const Eigen::VectorXf& mc, Eigen::VectorXf& out) const {
        out = mc;
        if (mc.size() != 4 || inliers.size() <= 3) return;
        float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        size_t cnt = 0;
        for (int i : inliers) {
            const PointT& p = input_->points[size_t(i)];
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
            acc[0] += p.x * p.x; acc[1] += p.x * p.y; acc[2] += p.x * p.z; acc[3] += p.y * p.y; acc[4] += p.y * p.z; acc[5] += p.z * p.z;
            acc[6] += p.x; acc[7] += p.y; acc[8] += p.z;
            cnt++;
        }
        if (!cnt) return;
        for (float& a : acc) a /= float(cnt);
        float cx = acc[6], cy = acc[7], cz = acc[8];
        Eigen::MatrixXd cov(3, 3);
        cov(0, 0) = acc[0] - cx * cx; cov(1, 1) = acc[3] - cy * cy; cov(2, 2) = acc[5] - cz * cz;
        cov(0, 1) = cov(1, 0) = acc[1] - cx * cy; cov(0, 2) = cov(2, 0) = acc[2] - cx * cz; cov(1, 2) = cov(2, 1) = acc[4] - cy * cz;
        Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> es(cov);
        Eigen::VectorXf o(4);
        for (int i = 0; i < 3; i++) o[i] = float(es.eigenvectors()(i, 0));
        o[3] = -1.f * (o[0] * cx + o[1] * cy + o[2] * cz);
        if (isModelValid(o)) out = o;
    }

protected:
    CloudConstPtr input_;
    std::vector<int> indices_, shuffled_;
};

template <class PointT>
class SampleConsensusModelPerpendicularPlane : public SampleConsensusModelPlane<PointT> {
public:
    typedef boost::shared_ptr<SampleConsensusModelPerpendicularPlane> Ptr;
    using SampleConsensusModelPlane<PointT>::SampleConsensusModelPlane;
    void setAxis(const Eigen::Vector3f& a) { axis_ = a; }
    void setEpsAngle(double e) { eps_ = e; }
    // valid iff the plane normal is within eps of the axis (either orientation)
    bool isModelValid(const Eigen::VectorXf& mc) const override {
        if (mc.size() != 4) return false;
        if (eps_ > 0.0) {
            double d = double(axis_[0] * mc[0] + axis_[1] * mc[1] + axis_[2] * mc[2]);  // both unit vectors
            double rad = d < -1.0 ? -1.0 : (d > 1.0 ? 1.0 : d);
            double angle = std::fabs(std::acos(rad));
            angle = std::min(angle, M_PI - angle);
            if (angle > eps_) return false;
        }
        return true;
    }

private:
    Eigen::Vector3f axis_ = Eigen::Vector3f(0.f, 0.f, 1.f);
    double eps_ = 0.0;
};

template <class PointT>
class RandomSampleConsensus {
public:
    explicit RandomSampleConsensus(const boost::shared_ptr<SampleConsensusModelPlane<PointT>>& m) : model_(m) {}
    template <class M> explicit RandomSampleConsensus(const boost::shared_ptr<M>& m) : model_(m) {}
    void setDistanceThreshold(double t) { threshold_ = t; }
    void setMaxIterations(int n) { max_iterations_ = n; }
    void setProbability(double p) { probability_ = p; }
    bool computeModel(int = 0) {
        iterations_ = 0;
        int n_best = -INT_MAX;
        double k = 1.0;
        std::vector<int> selection;
        Eigen::VectorXf mc;
        const double log_probability = std::log(1.0 - probability_);
        const double one_over_indices = 1.0 / double(model_->getIndices().size());
        unsigned skipped = 0;
        const unsigned max_skip = unsigned(max_iterations_) * 10u;
        while (iterations_ < k && skipped < max_skip) {
            model_->getSamples(iterations_, selection);
            if (selection.empty()) break;
            if (!model_->computeModelCoefficients(selection, mc)) { ++skipped; continue; }
            int n = model_->countWithinDistance(mc, threshold_);
            if (n > n_best) {
                n_best = n;
                best_selection_ = selection;
                coeffs_ = mc;
                double w = double(n_best) * one_over_indices;
                double p_no_outliers = 1.0 - std::pow(w, double(selection.size()));
                p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
                p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
                k = log_probability / std::log(p_no_outliers);
            }
            ++iterations_;
            if (iterations_ > max_iterations_) break;
        }
        if (best_selection_.empty()) { inliers_.clear(); return false; }
        model_->selectWithinDistance(coeffs_, threshold_, inliers_);
        return true;
    }
    void getInliers(std::vector<int>& out) const { out = inliers_; }
    void getModelCoefficients(Eigen::VectorXf& out) const { out = coeffs_; }
    int iterations() const { return iterations_; }

private:
    boost::shared_ptr<SampleConsensusModelPlane<PointT>> model_;
    double threshold_ = 0, probability_ = 0.99;
    int max_iterations_ = 1000, iterations_ = 0;
    std::vector<int> best_selection_, inliers_;
    Eigen::VectorXf coeffs_;
};

struct ModelCoefficients { std::vector<float> values; };

}  // namespace pcl
