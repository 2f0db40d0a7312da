// Exercises the source-compatible shim the way tracklets_depth does
// (/root/reference/tracklets_depth/src/tracklet_depth_module.cpp:63-117): InitConfig from a parameter
// object, Initialize with a CameraPinhole and the lidar->camera transform, then the 5-argument
// CalculateDepth(cloud, features, depths, resultType, plane). Inputs are read from / results written
// to raw binary files so that tests/test_shim_cpp.py can diff them against the oracle.
//   shim_selftest <points.f32 (n x 4)> <features.f64 (F x 2)> <out_depth.f64> <out_status.i32> <use_plane 0|1> <out_plane.bin>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "monolidar_fusion/DepthEstimator.h"

template <typename T>
static std::vector<T> read_all(const char* path) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) {
        std::cerr << "cannot read " << path << std::endl;
        std::exit(2);
    }
    size_t bytes = (size_t)in.tellg();
    in.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    auto pts = read_all<float>(argv[1]);
    auto uv = read_all<double>(argv[2]);
    const bool use_plane = std::atoi(argv[5]) != 0;
    using namespace Mono_Lidar;
    try {
        DepthEstimator est;
        // misuse is reported like the reference does (throw const char*)
        bool threw = false;
        try {
            est.Initialize(std::make_shared<CameraPinhole>(1241, 376, 718.856, 607.1928, 185.2157), Eigen::Affine3d());
        } catch (const char*) {
            threw = true;
        }
        if (!threw) return 3;

        auto params = std::make_shared<DepthEstimatorParameters>();
        params->pixelarea_search_witdh = 6;
        params->pixelarea_search_height = 9;
        params->radiusSearch_count_min = 4;  // overwritten after InitConfig, see below
        params->histogram_segmentation_bin_witdh = 0.3;
        params->pca_treshold_2_1_rel_min = 1.5;
        params->ransac_plane_distance_treshold = 0.3;
        params->do_use_ransac_plane = use_plane ? 1 : 0;
        est.InitConfig(params, false);
        // changed through the shared pointer AFTER InitConfig: the reference reads the block when Initialize builds its modules
        params->viewray_plane_orthoganality_treshold = 0.03;
        params->radiusSearch_count_min = 1;

        Eigen::Affine3d T;
        const double Tm[12] = {7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03, 1.480249e-02, 7.280733e-04,
                               -9.998902e-01, -7.631618e-02, 9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01};
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) T.matrix()(r, c) = Tm[r * 4 + c];
        est.Initialize(std::make_shared<CameraPinhole>(1241, 376, 718.856, 607.1928, 185.2157), T);

        auto cloud = std::make_shared<DepthEstimator::Cloud>();
        cloud->points.resize(pts.size() / 4);
        for (size_t i = 0; i < cloud->points.size(); i++) {
            cloud->points[i].x = pts[i * 4];
            cloud->points[i].y = pts[i * 4 + 1];
            cloud->points[i].z = pts[i * 4 + 2];
            cloud->points[i].intensity = pts[i * 4 + 3];
        }
        const int F = (int)(uv.size() / 2);
        Eigen::Matrix2Xd feats(2, F);
        for (int i = 0; i < F; i++) {
            feats(0, i) = uv[(size_t)i * 2];
            feats(1, i) = uv[(size_t)i * 2 + 1];
        }
        Eigen::VectorXd depths;
        Eigen::VectorXi types;
        GroundPlane::Ptr plane;  // nullptr: created and fitted by setInputCloud when do_use_ransac_plane
        est.setRansacSeed(99);
        DepthEstimator::Cloud::ConstPtr ccloud = cloud;
        est.CalculateDepth(ccloud, feats, depths, types, plane);
        if (use_plane && (plane == nullptr || !plane->isSegmented())) return 4;
        if (!use_plane && plane != nullptr) return 5;

        // statistics and debug views of DepthEstimator.h:116-164
        {
            auto stats = est.getDepthCalcStats();
            int n_success = 0, n_reached = 0;
            for (int i = 0; i < F; i++) {
                n_success += types(i) == Success ? 1 : 0;
                // statuses that are decided after the corner selection
                n_reached += (types(i) == Success || types(i) == TriangleNotPlanar || types(i) == PlaneViewrayNotOrthogonal ||
                              (types(i) >= TresholdDepthGlobalGreaterMax && types(i) <= TresholdDepthLocalSmallerMin) || types(i) == CornerBehindCamera)
                                 ? 1 : 0;
            }
            if (stats.getPointCount() != F || stats.getSuccess() != n_success) return 20;
            int total = 0;
            for (int t = 0; t < DepthCalculationStatistics::kTypes; t++) total += stats.count((DepthResultType)t);
            if (total != F) return 21;
            Eigen::Matrix2Xd vis;
            est.getPointsCloudImageCs(vis);
            auto camc = std::make_shared<DepthEstimator::Cloud>();
            DepthEstimator::Cloud::Ptr camp = camc;
            est.getCloudCameraCs(camp);
            if (camp->points.size() != cloud->points.size() || vis.cols() < 100) return 22;
            // visible point 0: its camera-frame depth re-projects its image coordinates' scale (z > 0 is not required upstream)
            const double z0 = est.getPointDepthCamVisible(0), zl = est.getPointDepthCamVisible(vis.cols() - 1);
            if (!(z0 == z0) || !(zl == zl)) return 23;
            auto tri = std::make_shared<DepthEstimator::Cloud>();
            DepthEstimator::Cloud::Ptr trip = tri;
            est.getCloudTriangleCorners(trip);
            if (trip->points.size() % 3 != 0 || (int)trip->points.size() < 3 * n_success) return 24;
            if (!use_plane && (int)trip->points.size() != 3 * n_reached) return 25;  // road-path triangles are not part of this view
            auto gpc = std::make_shared<DepthEstimator::Cloud>();
            DepthEstimator::Cloud::Ptr gpp = gpc;
            est.getCloudRansacPlane(gpp);
            if (use_plane && gpp->points.size() != plane->getInlinersIndex().size()) return 26;
            if (!use_plane && !gpp->points.empty()) return 27;
            auto e1 = std::make_shared<DepthEstimator::Cloud>();
            DepthEstimator::Cloud::Ptr e1p = e1;
            est.getCloudInterpolated(e1p);
            est.getCloudInterpolatedPlane(e1p);
            est.getCloudNeighbors(e1p);
            if (!e1p->points.empty()) return 28;
        }

        // the 4-argument overload the real caller uses discards the status vector
        Eigen::VectorXd depths2;
        est.CalculateDepth(feats, depths2, plane);
        for (int i = 0; i < F; i++)
            if (!(depths2(i) == depths(i)) && !(depths2(i) != depths2(i) && depths(i) != depths(i))) return 6;
        // single-point overload
        Eigen::Vector2d one;
        one[0] = feats(0, 0);
        one[1] = feats(1, 0);
        auto pr = est.CalculateDepth(one, plane);
        if ((int)pr.first != types(0)) return 7;

        // batch adaptor: previous + current cloud in one call gives the same depths; statistics add up
        {
            Eigen::VectorXd dl, dc;
            GroundPlane::Ptr pl_last = plane, pl_cur = plane;
            est.CalculateDepthPair(ccloud, feats, dl, pl_last, ccloud, feats, dc, pl_cur);
            for (int i = 0; i < F; i++) {
                const bool same_l = (dl(i) == depths(i)) || (dl(i) != dl(i) && depths(i) != depths(i));
                const bool same_c = (dc(i) == depths(i)) || (dc(i) != dc(i) && depths(i) != depths(i));
                if (!same_l || !same_c) return 8;
            }
            long long counters[21];
            est.getDepthCalcStats(types, counters);
            long long total = 0;
            for (int i = 0; i < 21; i++) total += counters[i];
            if (total != F) return 9;
        }

        // use_plane == 2: the production configuration (tracklets_depth/src/tracklet_depth_module.cpp:269-284) -- a SemanticPlane
        // built from a label image (argv[7]: 376 x 1241 uint8, argv[8]: inlier threshold) is handed to CalculateDepth
        if (std::atoi(argv[5]) == 2 && argc >= 9) {
            auto lab = read_all<unsigned char>(argv[7]);
            cv::Mat img(376, 1241, CV_8UC1, lab.data());
            SemanticPlane::Camera cam;
            cam.f = 718.856;
            cam.cu = 607.1928;
            cam.cv = 185.2157;
            cam.transform_cam_lidar = T;
            plane = std::make_shared<SemanticPlane>(img, cam, std::set<int>{6, 7, 8, 9}, std::atof(argv[8]));
            est.CalculateDepth(ccloud, feats, depths, types, plane);  // setInputCloud segments the plane (DepthEstimator.cpp:281-283)
            if (!plane->isSegmented()) return 13;
        }

        std::ofstream(argv[3], std::ios::binary).write(reinterpret_cast<const char*>(depths.data()), (std::streamsize)(F * sizeof(double)));
        std::ofstream(argv[4], std::ios::binary).write(reinterpret_cast<const char*>(types.data()), (std::streamsize)(F * sizeof(int)));
        std::ofstream po(argv[6], std::ios::binary);
        if (plane != nullptr) {
            po.write(reinterpret_cast<const char*>(plane->getModelCoeffs().data()), 4 * sizeof(float));
            const auto& inl = plane->getInlinersIndex();
            po.write(reinterpret_cast<const char*>(inl.data()), (std::streamsize)(inl.size() * sizeof(int)));
        }
        std::cout << "shim ok F=" << F << " n=" << cloud->points.size() << std::endl;
    } catch (const char* e) {
        std::cerr << "const char*: " << e << std::endl;
        return 10;
    } catch (const std::string& e) {
        std::cerr << "std::string: " << e << std::endl;
        return 11;
    } catch (const std::exception& e) {
        std::cerr << "std::exception: " << e.what() << std::endl;
        return 12;
    }
    return 0;
}
